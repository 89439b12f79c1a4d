"""SEA_ATTN_TRACE=1 python scripts/attn_trace.py: where the softmax / MMA warps of the tcgen05 attention kernel spend their cycles
(per-warp counters of csrc/block_attn_umma.cu, averaged over the CTAs of each row block)."""
import os
os.environ['SEA_ATTN_TRACE'] = '1'
import ctypes, importlib, sys
import numpy as np
import torch
sys.argv = [sys.argv[0]] + sys.argv[1:]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(__file__), 'run_attn.py')).read())         # runs the kernel on the real north-star mask
sea = importlib.import_module('sea-attention_b200')
lib = sea._lib.load()
buf = np.zeros((4096, 20, 16), dtype=np.uint32)
n = int(lib.sea_debug_attn_trace_read(buf.ctypes.data_as(ctypes.c_void_p), buf.size))
print('traced CTAs', n)
buf = buf[:n].astype(np.float64)
NH = 32
nblk = n // NH
names_s = ['wait S', 'tmem ld', 'exp/pack', 'wait Pbuf', 'st+handoff', 'boundary', 'mask word+union', 'loop total']
names_m = ['wait Sbuf', 'issue S', 'wait V', 'wait P', 'issue PV', 'wait K', '-', 'loop total']
for blk_pos in (0, nblk // 2, nblk - 1):          # blockIdx.x / NH: 0 = heaviest (last) row block
    ctas = buf[blk_pos * NH:(blk_pos + 1) * NH]
    sm = ctas[:, 4:20].mean(axis=(0, 1))
    mm = ctas[:, 1:3].mean(axis=(0, 1))
    print(f'-- row-block position {blk_pos} (0 = heaviest): softmax warps, cycles (share of loop)')
    print('   ' + '  '.join(f'{a}: {b:.0f} ({100 * b / max(sm[7], 1):.0f}%)' for a, b in zip(names_s, sm)))
    print('   MMA warps: ' + '  '.join(f'{a}: {b:.0f} ({100 * b / max(mm[7], 1):.0f}%)' for a, b in zip(names_m, mm) if a != '-'))
    # spread between the softmax warps of a CTA (who is the slowest?)
    per_warp = ctas[:, 4:20, 2].mean(axis=0)
    print('   exp/pack cycles per softmax warp:', ' '.join(f'{x:.0f}' for x in per_warp))

# CTA-level timeline (slot of warp 3): SM id, start / end (globaltimer, ns), set-up / loop end / epilogue end / exit (cycles from entry)
cta = buf[:, 3, :]
t0 = cta[:, 1].min()
start, end, sm = (cta[:, 1] - t0) % 2 ** 32, (cta[:, 2] - t0) % 2 ** 32, cta[:, 0].astype(int)
print(f'kernel span (first CTA start -> last CTA end): {end.max() / 1e3:.1f} us;  CTAs per SM: min {np.bincount(sm).min()} max {np.bincount(sm).max()}')
for blk_pos in (0, 4, nblk // 2, nblk - 4, nblk - 3, nblk - 2, nblk - 1):
    c = cta[blk_pos * NH:(blk_pos + 1) * NH]
    print(f'-- row-block position {blk_pos}: cycles from entry: bits+barriers+tmem {c[:, 8].mean():.0f}, masks+cstar {c[:, 9].mean():.0f}, q.k {c[:, 10].mean():.0f}, set-up {c[:, 4].mean():.0f}, loop end {c[:, 5].mean():.0f}, l exchanged {c[:, 11].mean():.0f}, O loaded {c[:, 12].mean():.0f}, epilogue end {c[:, 6].mean():.0f}, exit {c[:, 7].mean():.0f};'
          f'  wall {((c[:, 2] - c[:, 1]) % 2 ** 32).mean() / 1e3:.2f} us')
busy = np.zeros(sm.max() + 1); last_end = np.zeros(sm.max() + 1); gaps = []
order = np.argsort(start)
for i in order:
    if last_end[sm[i]] > 0:
        gaps.append(start[i] - last_end[sm[i]])
    last_end[sm[i]] = end[i]; busy[sm[i]] += end[i] - start[i]
print(f'per SM: busy {busy.mean() / 1e3:.1f} us avg ({busy.min() / 1e3:.1f} .. {busy.max() / 1e3:.1f}), last end {last_end.mean() / 1e3:.1f} us avg ({last_end.min() / 1e3:.1f} .. {last_end.max() / 1e3:.1f});'
      f' gap between consecutive CTAs on an SM: {np.mean(gaps) / 1e3:.2f} us avg, {np.max(gaps) / 1e3:.2f} max')
