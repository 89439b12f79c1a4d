"""Headline benchmark: SEA attention layer forward, tokens/s at the OPT-1.3B shape (32 heads, d=64,
seq 4096, k=64, predictor length 256, nbf=8), bf16, 1..8 B200 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one PerlinAttention forward over one synthetic batch item per GPU (weak scaling: the path
shards by batch with no collective).  Timing: CUDA events around every step on the launch
stream, L2 flushed (256 MiB memset) between steps, barrier + synchronize on both sides, max over ranks.
`value`  : inputs resident in HBM.            `e2e` : pinned-host q,k,v -> device, forward, context -> host.
`roofline`: dominant kernel, algorithmic bytes (SURVEY 8d stage model) / its mean launch time.
`cpu_baseline` / `--impl reference`: the reference's dense torch path restated in oracle/ (the reference
is python and cannot travel to the GPU box), timed on the host cores on a bounded sample.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'SEA attn fwd tokens/sec @OPT-1.3B 4k ctx'
UNIT = 'tokens/s'
NS = dict(H=32, d=64, T=4096, P=256, k=64, nbf=8)


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d.get('bf16_tflops_sustained', d.get('bf16_tflops')), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every ~2 ms (the timed loops last tens of
    milliseconds), nvidia-smi as the fallback when pynvml is unavailable."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread, self.how = index, [], False, None, 'nvidia-smi'

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
        try:
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id
            dom = torch.cuda.get_device_properties(self.index).pci_domain_id
            dev = torch.cuda.get_device_properties(self.index).pci_device_id
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f'{dom:08x}:{bus:02x}:{dev:02x}.0')
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = [(getattr(pynvml, 'nvmlClocksThrottleReasonHwSlowdown', 0x8), 0), (getattr(pynvml, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40), 1),
                (getattr(pynvml, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20), 2), (getattr(pynvml, 'nvmlClocksThrottleReasonSwPowerCap', 0x4), 3)]
        self.how = 'nvml'
        while not self.stop_flag:
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            row = [str(sm), str(mx)] + ['Active' if r & b else 'Not Active' for b, _ in bits]
            self.samples.append(row)
            time.sleep(0.002)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            self.how = 'nvidia-smi'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = sorted({self.NAMES[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples), 'source': self.how}


def stage_bytes(N, H, d, T, P, k, F, b, Z):
    """ALGORITHMIC bytes per stage: every stage reads its inputs once and writes its outputs once."""
    NHTd = N * H * T * d
    NHT = N * H * T
    return {
        'performer': 3 * NHTd * b + 2 * NHTd * b + NHTd * b + F * d * 4 + T * d * 4,          # q,k,v in; ctx (2d) + cumavg out
        'mlp': 2 * NHTd * b + NHTd * b + NHT * (P // 2) * b + NHT * 8,                         # ctx, v in; cnn_in, scales out
        'conv': 2 * NHT * (P // 2) * b,                                                        # per conv: in + out
        'tail': NHT * (P // 2) * b + NHT * P * 4,                                              # conv out in; probs fp32 out
        'topk': NHT * P * 4 + N * T * H * P // 8,                                              # probs in; bitmask out
        'csr': N * T * H * P // 8 * 2 + N * (T + 1) * 4 + Z * 4,                               # bitmask in (count+fill); crow, col (int32) out
        'attn': 3 * NHTd * b + Z * 4 + NHT * 8 + NHTd * b + NHTd * b,                          # q,k,v, col, scales, cumavg in; ctx out
    }


def stage_flops(N, H, d, T, P, k, F, Z):
    C = 2 * H
    return {
        'performer': 4 * N * H * T * d * F + 8 * N * H * T * F * d,
        'mlp': 2 * N * H * T * (3 * d * 2 * d + 2 * d * (P // 2) + 2 * d * 2),
        'conv': 2 * N * T * (P // 4) * C * C * 9,
        'tail': 2 * N * T * (P // 4) * C * H,
        'attn': 4 * Z * d,
    }


def _bind_to_gpu_numa_node(index):
    """N > 1: the ranks of a box share the host's memory / PCIe paths.  Bind this rank's threads to the CPU set NVML reports as
    local to its GPU, so that its pinned staging buffers are first-touched on that GPU's host bridge.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(index)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f'{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0')
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f'{len(cpus)} cpus local to the GPU (NVML)'
    except Exception as e:
        return f'not bound ({type(e).__name__})'
    return 'not bound'


_STDOUT_FD = None


def _claim_stdout():
    """Library chatter (NCCL's version banner is a bare printf) must not share stdout with the ONE JSON line: everything written
    to fd 1 from here on goes to stderr, the JSON line is written to the saved descriptor."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, (json.dumps(line) + '\n').encode())


WORKLOAD = 'SEA attention layer fwd, OPT-1.3B shape: 32 heads d=64 seq 4096 k=64 predictor_length=256 nbf=8, causal, N=1/GPU'
# identical in both arms (the driver compares the arms' `config`)
CONFIG = {'workload': WORKLOAD,
          'sharding': 'batch (one item per GPU, no collective on the hot path)',
          'weights': 'random-init (seed 42)',
          'l2': 'GPU arm: flushed between steps (256 MiB memset), per-step CUDA events; CPU reference arm: not applicable'}


def _reference_step(H, d, T, P, k, nbf, causal=True, k_flatten_dim=None):
    """-> (step callable, kind, note).  kind 'reference' = the UNMODIFIED reference module (oracle/_ref, the verbatim copy
    `make -C oracle` makes in the build container; /root/reference itself does not exist on the GPU box) running its own
    dense torch CPU path (benchmarking=False, fp32, no_grad: what the reference runs on a CPU);
    kind 'port' = the oracle restatement, only when no copy of the reference travelled."""
    from oracle import ref_harness as rh
    q = torch.randn(1, H, T, d) * (d ** -0.5 if causal else 1.0)
    kk = torch.randn(1, H, T, d)
    v = torch.randn(1, H, T, d)
    if rh.reference_available():
        m = rh.build_reference_attention(H, d, T, k, P, nbf, causal, k_flatten_dim=k_flatten_dim)
        m.benchmarking = False
        mask = rh.causal_additive_mask(T, torch.float32) if causal else torch.zeros(1, 1, 1, T)

        def step():
            with torch.no_grad():
                return m(q, kk, v, q, kk, v, q, kk, mask, None, None)
        return step, 'reference', f'unmodified reference PerlinAttention ({os.path.relpath(rh.REFERENCE_ROOT, ROOT)}), benchmarking=False dense torch path'
    from oracle import sea_oracle as so
    import transformers
    sea = importlib.import_module('sea-attention_b200')
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=causal)).eval()
    sd = {k_: v_.detach().float() for k_, v_ in mod.state_dict().items() if 'enc_per_layer' not in k_}
    fwd = so.perlin_forward_causal if causal else so.perlin_forward_noncausal

    def step():
        with torch.no_grad():
            return fwd(sd, q, kk, v, k_top=k, P=P, sparse=False)
    return step, 'port', 'oracle restatement of the dense torch path (no copy of the reference on this box)'


def _time_cpu(step, warmup, steps):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def cpu_config1_row():
    """BASELINE configs[0] (the reference's own CPU-runnable case): BERT-base PerlinAttention layer fwd, seq 512, k=64,
    predictor length 128, nbf=1, non-causal, k_flatten_dim='batch' -- the mandatory CPU row of BASELINE.md section 3."""
    torch.manual_seed(42)
    step, kind, note = _reference_step(12, 64, 512, 128, 64, 1, causal=False, k_flatten_dim='batch')
    dt = _time_cpu(step, 3, 5)
    return {'workload': 'BERT-base PerlinAttention layer fwd, seq 512, k=64, predictor-length 128, nbf=1, fp32 torch CPU path',
            'value': 512 / dt, 'unit': UNIT, 'ms_per_step': dt * 1e3, 'cores': torch.get_num_threads(), 'kind': kind, 'steps': 5, 'warmup': 3}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, at the arm's own
    config (full T = 4096), with exactly --warmup / --steps iterations.  Rank 0 only under torchrun."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    H, d, T, P, k, nbf = (NS[x] for x in ('H', 'd', 'T', 'P', 'k', 'nbf'))
    torch.manual_seed(42)
    step, kind, note = _reference_step(H, d, T, P, k, nbf)
    # exactly --warmup / --steps iterations, unless that would run past the wall-clock budget (then fewer, and the line says so)
    budget = float(os.environ.get('SEA_REF_BUDGET_S', '240'))
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    warmup = max(1, min(args.warmup, int(budget * 0.2 / first)))
    steps = max(1, min(args.steps, int(budget * 0.8 / first)))
    dt = _time_cpu(step, warmup - 1, steps)
    args.steps, args.warmup = steps, warmup
    val = T / dt
    sample = f'the full workload (all {T} tokens), fp32, {note}; {args.warmup} warm-up + {args.steps} timed steps'
    line = {
        'metric': METRIC, 'value': val, 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(CONFIG),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind, 'sample': sample,
                         'cpu_count': os.cpu_count()},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if not args.no_cpu_baseline:
        try:
            line['cpu_config1'] = cpu_config1_row()
        except Exception as e:          # the BERT row is extra information; never lose the headline line over it
            line['cpu_config1'] = {'error': repr(e)[:200]}
    _emit(line)


def run_c5(args):
    """BASELINE configs[4]: SEA attention layer alone, 32 heads, sequence sweep, query-block sharded over the ranks (strong scaling:
    the sequence is fixed, every rank holds all of q / k / v and produces its own query blocks; no collective on the data path).
    One JSON line: tokens/s per sequence length, device-timed, max over ranks."""
    import transformers
    sea = importlib.import_module('sea-attention_b200')
    par = importlib.import_module('sea-attention_b200.parallel')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    H, P, nbf = 32, 256, 8
    d, k = (128, 128) if args.workload == 'c5' else (64, 64)
    dt = torch.bfloat16
    seqs = [int(x) for x in args.seqs.split(',')]
    max_T = max(seqs)
    torch.manual_seed(42)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=max_T)
    mod = sea.PerlinAttention(cfg, sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)).eval().to(dev)
    mod.benchmarking = True
    mod.check_padding = False
    mod.freeze_packed_weights()
    rows_per_block = 16384                                     # a rank walks its rows in blocks of at most this many (bounds the fp32 probabilities)
    sweep = []
    for T in seqs:
        gen = torch.Generator().manual_seed(T)
        q = (torch.randn(1, H, T, d, generator=gen) * d ** -0.5).to(dt).to(dev)
        kk = torch.randn(1, H, T, d, generator=gen).to(dt).to(dev)
        v = torch.randn(1, H, T, d, generator=gen).to(dt).to(dev)
        n_blocks = max(world, -(-T // rows_per_block))
        n_blocks = -(-n_blocks // world) * world
        mine = [par.query_block_bounds(T, n_blocks, b) for b in range(rank, n_blocks, world)]       # round-robin: late (longer) blocks spread evenly

        exchange = world > 1 and not args.no_exchange and sea.ops.performer_range_supported(q, mod.performer.projection_matrix)

        def step():
            outs = None
            if exchange:
                # contiguous rows per rank (the O(T k) gather attention costs the same for every row); the linear-attention stage runs over
                # the rank's own rows only, started from the state sums the ranks exchange (one NCCL all-gather of ~3 MB per rank)
                perf, (r0, r1) = par.performer_exchanged(mod, q, kk, v, world, rank)
                for b0 in range(r0, r1, rows_per_block):
                    outs = mod.forward_query_block(q, kk, v, b0, min(b0 + rows_per_block, r1), performer=perf).context_layer
                return outs
            live = [(t0, t1) for t0, t1 in mine if t1 > t0]
            # a rank that walks several blocks runs the linear-attention stage once over its longest prefix (no exchange with other ranks)
            perf = mod.performer_prefix(q, kk, v, max(t1 for _, t1 in live)) if len(live) > 1 else None
            for t0, t1 in live:
                outs = mod.forward_query_block(q, kk, v, t0, t1, performer=perf).context_layer
            return outs

        with torch.no_grad():
            for _ in range(max(1, min(args.warmup, 2))):
                step()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            steps = max(1, min(args.steps, 3))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        sweep.append({'T': T, 'ms': ms, 'tokens_per_s': T / ms * 1e3, 'blocks': n_blocks})
        del q, kk, v
        torch.cuda.empty_cache()
    if rank == 0:
        _emit({'metric': 'SEA attn layer fwd tokens/sec, long-context sweep (BASELINE configs[4])', 'unit': 'tokens/s', 'n_gpus': world,
               'value': sweep[-1]['tokens_per_s'], 'higher_is_better': True, 'scaling': 'strong', 'dtype': 'bf16', 'data': 'synthetic',
               'config': {'workload': f'SEA attention layer alone, 32 heads d={d}, k={k}, predictor_length={P}, nbf={nbf}, causal, N=1, seq sweep',
                          'sharding': ('contiguous query rows per rank, K/V replicated, 8-row CNN halo; one collective: all-gather of the Performer state sums '
                                       '(exclusive scan across ranks)') if world > 1 and not args.no_exchange else
                                      'query blocks (K/V replicated, Performer prefix computed once per rank, 8-row CNN halo); no collective'},
               'sweep': sweep})
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='sea')
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the kernels eagerly instead of replaying one CUDA graph per step')
    ap.add_argument('--workload', default='ns', choices=['ns', 'c5', 'c5-d64'],
                    help="ns: the north-star layer forward (default, the contract line); c5 / c5-d64: BASELINE configs[4], the layer alone over a "
                         "sequence sweep, query-block sharded over the ranks (d=128 k=128 / d=64 k=64)")
    ap.add_argument('--seqs', default='4096,8192,16384,32768,65536,131072')
    ap.add_argument('--no-exchange', action='store_true', help='c5 at N > 1: every rank recomputes the Performer prefix instead of exchanging state sums')
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload != 'ns':
        return run_c5(args)

    import transformers
    sea = importlib.import_module('sea-attention_b200')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    affinity = _bind_to_gpu_numa_node(local_rank) if world > 1 else None    # before the pinned buffers are allocated (first touch)
    dt = {'bf16': torch.bfloat16, 'fp32': torch.float32, 'fp16': torch.float16}[args.dtype]
    H, d, T, P, k, nbf = (NS[x] for x in ('H', 'd', 'T', 'P', 'k', 'nbf'))
    N = 1                                                     # batch items per GPU (weak scaling)
    torch.manual_seed(42 + rank)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T)
    pc = sea.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=True)
    mod = sea.PerlinAttention(cfg, pc).eval().to(dev)
    mod.benchmarking = True
    mod.check_padding = False                                 # synthetic data has no padding; keeps the step free of host syncs
    mod.freeze_packed_weights()                               # inference: weights are constant, bf16 packings made once
    F = mod.performer_nb_features
    gen = torch.Generator().manual_seed(42 + rank)
    hq = (torch.randn(N, H, T, d, generator=gen) * d ** -0.5).to(dt).pin_memory()
    hk = torch.randn(N, H, T, d, generator=gen).to(dt).pin_memory()
    hv = torch.randn(N, H, T, d, generator=gen).to(dt).pin_memory()
    hout = torch.empty(N, T, H * d, dtype=dt).pin_memory()
    q, kk, v = hq.to(dev), hk.to(dev), hv.to(dev)
    fp_min = torch.finfo(torch.float16).min / 2 if dt != torch.float32 else torch.finfo(torch.float32).min / 2
    mask = ((torch.arange(T, device=dev).view(1, T) > torch.arange(T, device=dev).view(T, 1)) * fp_min).to(dt).view(1, 1, T, T).expand(N, 1, T, T)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        return mod(q, kk, v, q, kk, v, q, kk, mask, None, None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def capture(fn):
        """The forward has no host synchronisation, so one step = one CUDA graph launch."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        return g, out

    use_graph = not args.no_graph
    lib = sea._lib
    for _ in range(2):
        step_device()
    lib.LAUNCH_COUNT = 0
    step_device()
    launches_per_step = lib.LAUNCH_COUNT
    torch.cuda.synchronize()
    if use_graph:
        graph, graph_out = capture(step_device)
        run_step = graph.replay
    else:
        run_step = step_device

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps)

    # ---- e2e: pinned host q,k,v -> device, forward, context -> pinned host; double-buffered so that the copies of
    # step i+1 overlap the kernels of step i (three streams); every step moves its own inputs and its own result.
    dbuf = [[torch.empty_like(q), torch.empty_like(kk), torch.empty_like(v)] for _ in range(2)]
    if use_graph:
        e2e_graphs = [capture(lambda b=b: mod(dbuf[b][0], dbuf[b][1], dbuf[b][2], dbuf[b][0], dbuf[b][1], dbuf[b][2], dbuf[b][0], dbuf[b][1], mask, None, None))
                      for b in range(2)]
    s_h2d, s_d2h, s_h2d2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    split_h2d = os.environ.get('SEA_BENCH_SPLIT_H2D', '0') == '1'

    def timed_e2e(steps, warmup):
        main = torch.cuda.current_stream()
        comp_done = [torch.cuda.Event() for _ in range(2)]
        d2h_done = [torch.cuda.Event() for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        copied2 = [torch.cuda.Event() for _ in range(2)]
        outs = [None, None]

        def one(i):
            b = i & 1
            with torch.cuda.stream(s_h2d):
                s_h2d.wait_event(comp_done[b])          # the forward that last read this buffer set has finished
                dbuf[b][0].copy_(hq, non_blocking=True)
                dbuf[b][1][:, : H // 2].copy_(hk[:, : H // 2], non_blocking=True) if split_h2d else dbuf[b][1].copy_(hk, non_blocking=True)
                if not split_h2d:
                    dbuf[b][2].copy_(hv, non_blocking=True)
                copied[b].record(s_h2d)
            if split_h2d:                               # second copy stream: keeps two DMA engines busy on the host->device direction
                with torch.cuda.stream(s_h2d2):
                    s_h2d2.wait_event(comp_done[b])
                    dbuf[b][1][:, H // 2:].copy_(hk[:, H // 2:], non_blocking=True)
                    dbuf[b][2].copy_(hv, non_blocking=True)
                    copied2[b].record(s_h2d2)
                main.wait_event(copied2[b])
            main.wait_event(copied[b])
            main.wait_event(d2h_done[b])                # the result buffer of this set has been drained
            if use_graph:
                e2e_graphs[b][0].replay()
                outs[b] = e2e_graphs[b][1]
            else:
                outs[b] = mod(dbuf[b][0], dbuf[b][1], dbuf[b][2], dbuf[b][0], dbuf[b][1], dbuf[b][2], dbuf[b][0], dbuf[b][1], mask, None, None)
            comp_done[b].record(main)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(comp_done[b])
                hout.copy_(outs[b].context_layer, non_blocking=True)
                d2h_done[b].record(s_d2h)

        for i in range(warmup):
            one(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        main.wait_stream(s_h2d)
        main.wait_stream(s_h2d2)
        main.wait_stream(s_d2h)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev = timed(run_step, args.steps, warm)
    launches = launches_per_step * args.steps             # kernels of libsea_b200.so launched inside the timed region
    ms_e2e = timed_e2e(args.steps, warm)
    clocks = sampler.stop() if sampler else None
    # beside the headline: the same step launched eagerly through mod() (python + 9 launches per step), and with the
    # estimated_attention_probs output not materialised (what the reference's OPT caller keeps, perlin_opt.py:477)
    ms_eager = timed(step_device, min(args.steps, 50), 3) if use_graph else ms_dev
    mod.keep_estimated_probs = False
    for _ in range(2):
        step_device()
    torch.cuda.synchronize()
    if use_graph:
        graph_np, _ = capture(step_device)
        ms_noprobs = timed(graph_np.replay, min(args.steps, 50), 3)
    else:
        ms_noprobs = timed(step_device, min(args.steps, 50), 3)
    mod.keep_estimated_probs = True

    # per-kernel breakdown (instrumented pass: events around every C-ABI call) -> dominant kernel + roofline
    lib.TRACE = []
    for _ in range(3):
        flush.zero_()
        step_device()
    torch.cuda.synchronize()
    per = {}
    for name, e0, e1 in lib.TRACE:
        per.setdefault(name, []).append(e0.elapsed_time(e1))
    lib.TRACE = None
    out = step_device()
    torch.cuda.synchronize()

    if rank == 0:
        hbm, tf, which = _peaks()
        crow = None
        mod.output_attentions = True
        o2 = step_device()
        Z = int(o2.partial_attention_mask.crow_indices()[0, -1].item())
        mod.output_attentions = False
        b = 2 if dt != torch.float32 else 4
        sb = stage_bytes(N, H, d, T, P, k, F, b, Z)
        if 'sea_sparse_attention_bits_fwd' in per or 'sea_block_attention_fwd' in per:
            # attention driven by the bit mask: no CSR tensors on the hot path (a8 is folded into the attention kernel)
            sb['csr'] = 0
            sb['attn'] = sb['attn'] - Z * 4 + N * T * H * P // 8
        sf = stage_flops(N, H, d, T, P, k, F, Z)
        entry_stage = {'sea_performer_causal_fwd': 'performer', 'sea_predictor_mlp_fwd': 'mlp', 'sea_causal_conv3x3_dil2_relu': 'conv',
                       'sea_causal_conv3x3_dil2_relu_umma': 'conv', 'sea_conv1x1_umma': 'tail', 'sea_causal_conv3x3_dil2_relu_conv1x1_umma': 'conv', 'sea_predictor_tail_topk_fwd': 'tail', 'sea_predictor_tail_topk_expand_fwd': 'tail',
                       'sea_predictor_mlp_umma_fwd': 'mlp', 'sea_predictor_mlp_umma_fwd_ex': 'mlp', 'sea_performer_causal_mma_fwd': 'performer',
                       'sea_predictor_tail_fwd': 'tail', 'sea_topk_mask_bits': 'topk', 'sea_csr_count': 'csr', 'sea_csr_fill': 'csr', 'sea_crow_scan': 'csr',
                       'sea_sparse_attention_fwd': 'attn', 'sea_sparse_attention_bits_fwd': 'attn', 'sea_block_attention_fwd': 'attn'}
        kernels = {}
        for name, times in per.items():
            st = entry_stage.get(name, name)
            calls_per_step = len(times) / 3
            kernels.setdefault(st, {'ms_per_step': 0.0, 'calls_per_step': 0})
            kernels[st]['ms_per_step'] += sum(times) / 3
            kernels[st]['calls_per_step'] += calls_per_step
        dom = max(kernels, key=lambda s: kernels[s]['ms_per_step'])
        dom_ms_launch = kernels[dom]['ms_per_step'] / max(1.0, kernels[dom]['calls_per_step'] if dom == 'conv' else 1.0)
        tensor_bound = dom in ('conv', 'mlp')
        if tensor_bound:
            ach = sf[dom] / (dom_ms_launch * 1e-3) / 1e12
            roof = {'kernel': dom, 'bound': 'tensor', 'achieved': ach, 'peak': tf, 'unit': 'TFLOP/s', 'frac': ach / tf, 'traffic': None,
                    'peak_source': which + ' (sustained)'}
        else:
            ach = sb[dom] / (dom_ms_launch * 1e-3) / 1e9
            roof = {'kernel': dom, 'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm, 'traffic': None,
                    'peak_source': which}
        try:        # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture (profiles/)
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(dom)
            if tr:
                roof['traffic'] = tr['dram_bytes_per_launch']
                roof['traffic_source'] = tr['source']
        except Exception:
            pass
        if dom == 'attn' and 'sea_block_attention_fwd' in per:
            # tensor view of the same kernel: dense FLOPs it EXECUTES (every 128 x 64 tile at or below the diagonal of every head:
            # S = Q.K^T and O += P.V, 4*128*64*64 flop per tile step; tiles without an alive element are skipped, so this is an upper bound)
            tiles = sum(-(-min(T, (rb + 1) * 128) // 64) for rb in range(-(-T // 128)))
            fl = N * H * tiles * 4 * 128 * 64 * d
            roof['tensor_view'] = {'flops_executed_upper_bound': fl, 'achieved': fl / (dom_ms_launch * 1e-3) / 1e12, 'peak': tf, 'unit': 'TFLOP/s',
                                   'frac': fl / (dom_ms_launch * 1e-3) / 1e12 / tf, 'algorithmic_flops': sf['attn']}
        total_bytes = sb['performer'] + sb['mlp'] + 2 * sb['conv'] + sb['tail'] + sb['topk'] + sb['csr'] + sb['attn']
        tokens = N * T * world
        line = {
            'metric': METRIC, 'value': tokens / (ms_dev * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
            'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': dict(CONFIG),
            'notes': {'launch': 'one CUDA graph replay per step' if use_graph else 'eager kernel launches',
                      'outputs': 'context_layer + estimated_attention_probs (output_attentions=False)', 'nnz': Z,
                      'eager_ms_per_step': ms_eager, 'eager_note': 'the same step through mod(...) without a CUDA graph (python dispatch + 9 launches)',
                      'ms_per_step_without_probs_output': ms_noprobs,
                      'without_probs_note': 'mod.keep_estimated_probs = False: estimated_attention_probs (134 MB fp32, dropped by the OPT caller, '
                                            'perlin_opt.py:477) is not written; NOT the headline value'},
            'e2e': {'value': tokens / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': 3 * q.numel() * q.element_size(),
                    'd2h_bytes_per_step': hout.numel() * hout.element_size(), 'ms_per_step': ms_e2e,
                    'h2d_gbs_aggregate': world * 3 * q.numel() * q.element_size() / (ms_e2e * 1e-3) / 1e9, 'cpu_affinity': affinity,
                    'note': 'causal additive mask is a shape constant kept on the device; H2D / forward / D2H on three streams, double-buffered, '
                            'every step copies its own inputs and result; no L2 flush in this loop'},
            'gpu_launches': launches,
            'roofline': roof,
            'layer_hbm': {'algorithmic_bytes': total_bytes, 'achieved_gbs': total_bytes / (ms_dev * 1e-3) / 1e9, 'peak_gbs': hbm,
                          'frac': total_bytes / (ms_dev * 1e-3) / 1e9 / hbm},
            'kernels_ms_per_step': {s: round(kv['ms_per_step'], 4) for s, kv in kernels.items()},
            'clocks': clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            # bounded sample of the SAME workload on the host cores: 1 warm-up + 2 timed steps of the full T = 4096 layer
            torch.set_num_threads(os.cpu_count())
            torch.manual_seed(42)
            step, kind, note = _reference_step(H, d, T, P, k, nbf)
            cdt = _time_cpu(step, 1, 2)
            line['cpu_baseline'] = {'value': T / cdt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind, 'cpu_count': os.cpu_count(),
                                    'ms_per_step': cdt * 1e3,
                                    'sample': f'the full workload (all {T} tokens), fp32, {note}; 1 warm-up + 2 timed steps'}
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
