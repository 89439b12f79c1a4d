"""Top source lines of an .ncu-rep by stall samples / instructions executed (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = {}; fname = ''
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
        si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed')
    elif hdr and r[0].isdigit():
        try:
            key = (fname, int(r[0]), r[1].strip()[:100])
            s, i, t = int(r[si] or 0), int(r[ii] or 0), int(r[ti] or 0)
            o = data.get(key, (0, 0, 0))
            data[key] = (o[0] + s, o[1] + i, o[2] + t)
        except (ValueError, IndexError):
            pass
ts = sum(d[0] for d in data.values()) or 1; tins = sum(d[1] for d in data.values()) or 1
print(f'total samples {ts} warp-instr {tins}')
for key, d in sorted(data.items(), key=lambda x: -x[1][0])[:topn]:
    print(f'{key[0][:14]:14s}:{key[1]:4d} smp {100*d[0]/ts:5.1f}%  ins {100*d[1]/tins:5.1f}%  thr/ins {d[2]/max(d[1],1):4.1f}  {key[2]}')
if len(sys.argv) > 3:
    # region sums: pairs lo-hi
    for rg in sys.argv[3:]:
        lo, hi = map(int, rg.split('-'))
        ss = sum(d[0] for k_, d in data.items() if k_[0].startswith('block_attn') and lo <= k_[1] <= hi)
        ii = sum(d[1] for k_, d in data.items() if k_[0].startswith('block_attn') and lo <= k_[1] <= hi)
        print(f'lines {lo}-{hi}: samples {100*ss/ts:.1f}%  instr {100*ii/tins:.1f}%')
    ss = sum(d[0] for k_, d in data.items() if not k_[0].startswith('block_attn'))
    print(f'other files: samples {100*ss/ts:.1f}%')
