#!/usr/bin/env python
"""bench.py -- ROI-crops/s through iDispNet (cost volume + 28-layer 3-D stack + soft-argmin) on B200.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`; under torchrun one rank
per GPU.  A step = one pass of the hot path over one batch of synthetic ROI feature pairs
(BASELINE.json configs[1]: 32 ROI pairs of 112x112x32-ch features, D=48 -> 448x448 disparity, per
GPU; weak scaling).  Prints ONE JSON line on rank 0.

  value ....... whole-job ROIs/s with the inputs already resident in HBM (device-timed, max over ranks)
  e2e ......... same metric through the C-ABI call with HOST (pinned) buffers, H2D + D2H inside the timed region
  roofline .... the 3-D conv launches: algorithmic FLOPs (644544*V - (64-2C)*1728*V per ROI)
                / their summed device time (CUDA events between launches on the launching stream),
                against MEASURED_PEAKS.json's sustained bf16 tensor peak
  cpu_baseline  the oracle (CPU port of the reference's PyTorch path) on the host cores, bounded sample
`--impl reference` times that CPU path alone with the same metric/config keys.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, HF, WF, MIND, MAXD, B_PER_GPU = 32, 112, 112, -96, 96, 32
GATHER_CHUNKS = int(os.environ.get('IDISP_GATHER_CHUNKS', '1'))
D = (MAXD - MIND) // 4
V = D * HF * WF
H, W = 4 * HF, 4 * WF
FLOP_PER_ROI = 644544 * V - (64 - 2 * C) * 1728 * V  # 388.09 GFLOP
WORKLOAD = (f'configs[1]: batch={B_PER_GPU} ROI pairs/GPU, {HF}x{WF}x{C}ch features, D={D} '
            f'(mindisp {MIND}, maxdisp {MAXD}) -> {H}x{W} disparity')


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get('bf16_tflops_sustained', 1400.0), j.get('bf16_tflops', 1590.0), j.get('hbm_gbs', 6650.0), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.idx}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def usable_cpus():
    """Hardware threads this process may actually use: affinity mask and cgroup CPU quota, not just os.cpu_count()."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        pass
    return n


def cpu_port_rois_per_s(model_sd, n_steps, n_warm, budget_s=150.0):
    """Time the oracle (CPU restatement of stackhourglass.py:115-174 on torch CPU ops, all host threads).

    Each step = a bounded sample of the config-2 workload: one ROI pair at full depth/width and `rows`
    feature rows (work is linear in V, so ROIs/s = (rows/HF) / t); rows shrinks until K+W steps fit the budget.
    """
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import idispnet_oracle as O  # the checker, used here ONLY as the reported CPU baseline
    cores = usable_cpus()
    torch.set_num_threads(cores)
    sd = {k: v.detach().cpu().float() for k, v in model_sd.items()}
    g = torch.Generator().manual_seed(0)

    def run(rows):
        L = torch.randn(1, C, rows, WF, generator=g).relu()
        R = torch.randn(1, C, rows, WF, generator=g).relu()
        t0 = time.perf_counter()
        with torch.no_grad():
            O.idispnet_from_features(L, R, sd, MIND, MAXD)
        return time.perf_counter() - t0

    run(28)               # first call pays oneDNN primitive creation
    t_probe = run(28)     # a quarter ROI; larger slices parallelise better on many cores, so scaling up is conservative
    # "all the host threads it can use": oneDNN's 3-D convs do not always get faster with every hardware thread of a big box
    # (with all 128 threads the GPU box measured 0.06-0.09 ROIs/s, an 8-core container 0.7-0.8) -- keep the fastest thread count
    best = cores
    for cand in sorted({max(1, cores // 2), 64, 32, 16}):
        if cand >= cores:
            continue
        torch.set_num_threads(cand)
        run(28)
        t = run(28)
        if t < t_probe:
            t_probe, best = t, cand
    cores = best
    torch.set_num_threads(cores)
    rows = HF
    while rows > 28 and t_probe * (rows / 28.0) * (n_steps + n_warm) > budget_s:
        rows //= 2
    for _ in range(n_warm):
        run(rows)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        run(rows)
    dt = (time.perf_counter() - t0) / max(n_steps, 1)
    return (rows / HF) / dt, dt, cores, f'{n_steps} step(s) of 1 ROI pair x {rows}/{HF} feature rows (D={D}, W={WF}, C={C}), {n_warm} warm-up'


def bench_roi_align(dev, hbm_gbs):
    """The ROIAlign that feeds iDispNet, reported separately: the live variant -- 8 synthetic 3x375x1242
    images, 32 integer-cornered boxes, 224x224 crops with the ImageNet normalisation fused (disprcnn3d.py:44-50)."""
    import torch
    from disprcnn_b200.layers.roi_align import crop_and_transform_roi_img
    g = torch.Generator().manual_seed(0)
    im = torch.rand(8, 3, 375, 1242, generator=g).to(dev)
    x1 = torch.randint(0, 800, (32,), generator=g).float()
    y1 = torch.randint(0, 150, (32,), generator=g).float()
    w = torch.randint(60, 400, (32,), generator=g).float()
    h = torch.randint(60, 200, (32,), generator=g).float()
    rois = torch.stack([torch.arange(32).float() % 8, x1, y1, x1 + w, y1 + h], 1).to(dev)
    for _ in range(3):
        out = crop_and_transform_roi_img(im, rois, 224)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 20
    for _ in range(n):
        out = crop_and_transform_roi_img(im, rois, 224)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out_bytes = out.numel() * 4
    return {'rois_per_s': 32 / (ms / 1e3), 'ms_per_call': ms, 'workload': '32 ROIs -> 3x224x224 from 8 images 3x375x1242, fused normalise',
            'algorithmic_bytes': out_bytes, 'achieved_gbs': out_bytes / (ms / 1e3) / 1e9, 'hbm_peak_gbs': hbm_gbs,
            'note': 'output-write bytes only; the gather reads hit L2 (3.7 MB image set), launch-latency dominated at this size'}


def bench_cost_volume(dev, hbm_gbs, L, R):
    """The standalone correlation kernel (stackhourglass.py:115-128 in the reference's NCDHW layout, C-ABI idisp_cost_volume):
    8 ROI pairs of the benchmark shape -> 1.23 GB written per call.  (The tensor-core modes never materialise this volume: the
    first conv's TMA loader assembles it; this entry exists for callers that want the tensor itself.)"""
    import torch
    from disprcnn_b200 import _lib
    lib = _lib.load()
    n = 8
    D = (MAXD - MIND) // 4
    cost = torch.empty(n, 2 * C, D, HF, WF, device=dev)
    Ln, Rn = L[:n].contiguous(), R[:n].contiguous()

    def run():
        _lib.check(lib.idisp_cost_volume(_lib.ptr(Ln), _lib.ptr(Rn), n, C, HF, WF, MIND, MAXD, _lib.ptr(cost), _lib.stream_ptr()))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    k = 10
    for _ in range(k):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    nbytes = cost.numel() * 4
    return {'ms_per_call': ms, 'workload': f'{n} ROI pairs 112x112x32ch, D=48 -> [{n},64,48,112,112] f32', 'algorithmic_bytes': nbytes,
            'achieved_gbs': nbytes / (ms / 1e3) / 1e9, 'hbm_peak_gbs': hbm_gbs, 'frac': nbytes / (ms / 1e3) / 1e9 / hbm_gbs,
            'rois_per_s': n / (ms / 1e3)}


def bench_live(dev):
    """The shape tools/test_net.py really runs (KITTI: R = 1..15 ROI pairs per image, 224x224 crops -> 56x56x32ch, D=24):
    latency per call through the public API, (a) 3-D stack from features, (b) whole PSMNet.forward from image crops
    (feature_extraction + stack).  Host-timed with a stream sync per call = what the caller of one image waits for."""
    import torch
    import torch.nn as nn
    from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet
    torch.manual_seed(0)
    m = PSMNet(48, -48)   # precision='auto' -> fp16x2
    with torch.no_grad():
        for c in (m.classif1, m.classif2, m.classif3):
            c[2].weight.mul_(0.1)
    m = m.to(dev)
    # random-init BatchNorm statistics are (0, 1): calibrate the extractor's on synthetic crops (two train-mode passes through the
    # torch modules, as tests/golden does) and scale its output to unit variance, so that activations stay in the range trained
    # weights produce -- otherwise the fp16 words of the split-precision mode overflow and 'auto' (correctly) reruns in fp32
    gcal = torch.Generator().manual_seed(7)
    fe = m.feature_extraction
    for mod in fe.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.momentum = None
    fe.train()
    with torch.no_grad():
        for _ in range(2):
            f = fe(torch.randn(4, 3, 224, 224, generator=gcal).to(dev))
        fe.lastconv[2].weight.mul_(1.0 / float(f.std()))
    m = m.eval()
    import warnings
    out = {'workload': 'R ROI pairs, 224x224 crops -> 56x56x32ch features, D=24 (mindisp -48, maxdisp 48) -> 224x224; precision auto (fp16x2)',
           'stack_ms': {}, 'psmnet_ms': {}, 'extractor_ms': {}, 'psmnet_rois_per_s': {}}
    g = torch.Generator().manual_seed(5)
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False   # whatever still runs through torch must be fp32-grade
    with torch.no_grad(), warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        for R in (1, 4, 8, 15):
            fl = torch.randn(R, 32, 56, 56, generator=g).to(dev)
            fr = torch.randn(R, 32, 56, 56, generator=g).to(dev)
            il = torch.randn(R, 3, 224, 224, generator=g).to(dev)
            ir = torch.randn(R, 3, 224, 224, generator=g).to(dev)
            both = torch.cat([il, ir])
            for name, fn in (('stack_ms', lambda: m.forward_features(fl, fr)), ('psmnet_ms', lambda: m({'left': il, 'right': ir})),
                             ('extractor_ms', lambda: m.feature_extraction(both))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(10):
                    t0 = time.perf_counter()
                    fn()
                    torch.cuda.synchronize()
                    ts.append((time.perf_counter() - t0) * 1e3)
                ts.sort()
                out[name][str(R)] = round(ts[len(ts) // 2], 4)
            out['psmnet_rois_per_s'][str(R)] = round(R / (out['psmnet_ms'][str(R)] / 1e3), 1)
    torch.backends.cudnn.allow_tf32 = old_tf32
    out['fp32_fallbacks'] = sum('fp16 range' in str(w.message) for w in caught)   # must be 0 for the numbers above to be the fp16x2 path
    out['extractor'] = type(m.feature_extraction).__module__ + ('.native' if getattr(m.feature_extraction, 'native', False) else ' (torch modules)')
    return out


def make_model(precision, device):
    import torch
    import torch.nn as nn
    from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet
    torch.manual_seed(0)
    m = PSMNet(MAXD, MIND, precision=precision)  # random init of the reference architecture
    m.feature_extraction = nn.Identity()
    # benign BN statistics + damped classifier so logits stay finite through 28 random layers
    with torch.no_grad():
        for c in (m.classif1, m.classif2, m.classif3):
            c[2].weight.mul_(0.1)
    return m.to(device).eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('IDISP_BENCH_PRECISION', 'fp16x2'), choices=['fp32', 'bf16', 'fp16', 'fp16x2'],
                    help='fp16x2 (default): split-precision tensor-core mode, meets the 1e-3 parity bar; fp16 / bf16: one-word '
                         'tensor-core modes (faster, 0.07 / 0.4 px from the reference); fp32: SIMT parity mode')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    warm = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    tens_sus, tens_burst, hbm, peak_src = peaks()

    if args.impl == 'reference':
        if rank != 0:
            return
        import torch
        import torch.nn as nn
        from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet
        torch.manual_seed(0)
        m = PSMNet(MAXD, MIND)
        with torch.no_grad():
            for c in (m.classif1, m.classif2, m.classif3):
                c[2].weight.mul_(0.1)
        sd = {k: v for k, v in m.state_dict().items() if not k.startswith('feature_extraction')}
        val, dt, cores, sample = cpu_port_rois_per_s(sd, args.steps, args.warmup)
        print(json.dumps({
            'impl': 'reference', 'metric': 'idispnet_roi_crops_per_s', 'value': val, 'unit': 'ROIs/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': B_PER_GPU * args.gpus, 'parallelism': f'dp{args.gpus}'},
            'cpu_baseline': {'value': val, 'unit': 'ROIs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'ROIs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
        }))
        return

    import torch
    import torch.distributed as dist
    from disprcnn_b200 import _lib
    from disprcnn_b200.parallel import gather_disparity, sharded_forward_async
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    lib = _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    Bg = B_PER_GPU * world
    m = make_model(args.precision, dev)
    g = torch.Generator().manual_seed(1234 + rank)
    L_host = torch.randn(B_PER_GPU, C, HF, WF, generator=g).relu().pin_memory()
    R_host = torch.randn(B_PER_GPU, C, HF, WF, generator=g).relu().pin_memory()
    out_host = torch.empty(B_PER_GPU, H, W).pin_memory()
    L, R = L_host.to(dev), R_host.to(dev)
    stream = torch.cuda.current_stream()

    pending = []  # N > 1: all-gathers in flight on NCCL's own stream (each overlaps the next step's kernels)

    def step_device():
        if world == 1:
            return m.forward_features(L, R)
        pending.append(sharded_forward_async(m, L, R, presharded=True, chunks=GATHER_CHUNKS))
        if len(pending) > 2:  # at most two result buffers (2 x Bg x H x W f32) alive: step i-2 must have landed
            return pending.pop(0).wait()
        return None

    def drain():
        out = None
        while pending:
            out = pending.pop(0).wait()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return ms.item()

    gather_info = None
    with torch.no_grad():
        for _ in range(warm):
            out = step_device()
        if world > 1:
            out = drain()
            # sharded + gathered == unsharded, checked once before timing: this rank recomputes the NEXT rank's shard from that
            # rank's seeded inputs and compares it bit for bit with what the gather delivered (ROIs are independent, kernels
            # deterministic, so the per-shard result is the unsharded result restricted to the shard)
            nxt = (rank + 1) % world
            gn = torch.Generator().manual_seed(1234 + nxt)
            Ln = torch.randn(B_PER_GPU, C, HF, WF, generator=gn).relu().to(dev)
            Rn = torch.randn(B_PER_GPU, C, HF, WF, generator=gn).relu().to(dev)
            mine = m.forward_features(Ln, Rn)
            same = torch.equal(mine, out[nxt * B_PER_GPU:(nxt + 1) * B_PER_GPU])
            flag = torch.tensor([0 if same else 1], device=dev)
            dist.all_reduce(flag)
            assert flag.item() == 0, f'rank {rank}: gathered shard of rank {nxt} differs from its recomputation'
            del Ln, Rn, mine
            # the collective alone: one all_gather_into_tensor of [B_PER_GPU, H, W] f32 per rank, CUDA events on this stream
            local = out[rank * B_PER_GPU:(rank + 1) * B_PER_GPU].clone()
            buf = torch.empty_like(out)
            for _ in range(2):
                dist.all_gather_into_tensor(buf, local)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for _ in range(5):
                dist.all_gather_into_tensor(buf, local)
            g1.record(stream)
            torch.cuda.synchronize()
            gms = torch.tensor([g0.elapsed_time(g1) / 5], device=dev)
            dist.all_reduce(gms, op=dist.ReduceOp.MAX)
            nbytes = local.numel() * 4
            gather_info = {'allgather_ms': gms.item(), 'bytes_per_rank': nbytes, 'algbw_gbs': nbytes * world / gms.item() / 1e6,
                           'busbw_gbs': nbytes * (world - 1) / gms.item() / 1e6, 'sharded_equals_unsharded': True,
                           'overlap': f'async_op all-gather on NCCL\'s stream in {GATHER_CHUNKS} ROI sub-chunk(s); the gather of step i '
                                      'overlaps the kernels of step i+1; every gather completes inside the timed region'}
            del buf, local
        assert os.environ.get('IDISP_TC_DBG') or torch.isfinite(out).all(), 'non-finite disparity in warm-up'
        plan = m._plan
        # ---- value: device-resident inputs; K steps enqueued back to back, no host sync inside the timed region ----
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        drain()  # N > 1: the current stream waits for every gather still in flight
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = t.item()
        launches_per_step = lib.idisp_plan_launches_per_forward(plan) + (GATHER_CHUNKS if world > 1 else 0)
        # ---- roofline leg: the same K steps again with CUDA events between the launches (on the launching stream; the plan
        # then launches eagerly instead of replaying its CUDA graph).  Only the per-kernel durations come from here. ----
        _lib.check(lib.idisp_plan_enable_timing(plan, 1))
        per_step = []
        for _ in range(args.steps):
            m.forward_features(L, R)
            n = lib.idisp_plan_launches_per_forward(plan)
            ms = (ctypes.c_float * n)()
            ly = (ctypes.c_int * n)(*([-100] * n))
            _lib.check(lib.idisp_plan_get_timing(plan, ms, ly, n))
            per_step.append((list(ms), list(ly)))
        torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        _lib.check(lib.idisp_plan_enable_timing(plan, 0))
        conv_ms = sum(v for msl, lyl in per_step for v, l in zip(msl, lyl) if 0 <= l <= 27)
        other_ms = sum(v for msl, lyl in per_step for v, l in zip(msl, lyl) if -100 < l < 0)
        by_layer = {}
        for msl, lyl in per_step:
            for v, l in zip(msl, lyl):
                if l > -100:
                    by_layer[l] = by_layer.get(l, 0.0) + v / args.steps

        # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----
        # A stream of batches through idisp_plan_forward_host_async: every step copies ITS inputs from pinned host memory and ITS
        # result back to pinned host memory; the library overlaps the H2D of step i with the kernels of step i-1 and the D2H of
        # step i with the kernels of step i+1 (double-buffered staging), and the user reads result i-1 (host_wait) while step i
        # runs -- the loader/consumer pattern of engine/inference.py:24-50.  The timed region closes only after the host has seen
        # the LAST result land.
        out_hosts = [out_host, torch.empty(B_PER_GPU, H, W).pin_memory()]
        state = {'i': 0, 'prev': None}

        def consume(prev):
            _lib.check(lib.idisp_plan_host_wait(plan, prev[0]))   # result of the previous step is in its pinned buffer now
            if world > 1:
                # the gathered result is what a multi-GPU user reads; gather from the device copy of the output
                gather_disparity(prev[1].to(dev, non_blocking=True), Bg)

        def step_host():
            buf = out_hosts[state['i'] & 1]
            t = ctypes.c_ulonglong()
            _lib.check(lib.idisp_plan_forward_host_async(plan, _lib.ptr(L_host), _lib.ptr(R_host), B_PER_GPU, HF, WF, H, W,
                                                         _lib.ptr(buf), _lib.stream_ptr(), ctypes.byref(t)))
            if state['prev'] is not None:
                consume(state['prev'])
            state['prev'] = (t.value, buf)
            state['i'] += 1

        def finish_host():
            if state['prev'] is not None:
                consume(state['prev'])
                state['prev'] = None

        for _ in range(2):
            step_host()
        finish_host()
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        for _ in range(args.steps):
            step_host()
        finish_host()          # host-side wait for the last result copy
        h1.record(stream)
        torch.cuda.synchronize()
        ms_t = torch.tensor([h0.elapsed_time(h1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            dist.barrier()
        ms_e2e = ms_t.item()
        # the same call, one batch at a time (copy in -> kernels -> copy out -> host reads): the latency form
        def step_host_sync():
            _lib.check(lib.idisp_plan_forward_host(plan, _lib.ptr(L_host), _lib.ptr(R_host), B_PER_GPU, HF, WF, H, W,
                                                   _lib.ptr(out_host), _lib.stream_ptr()))
            torch.cuda.current_stream().synchronize()
        step_host_sync()
        ms_e2e_sync = timed(step_host_sync, args.steps)

    value = Bg * args.steps / (ms_total / 1e3)
    e2e = Bg * args.steps / (ms_e2e / 1e3)
    conv_flops = FLOP_PER_ROI * B_PER_GPU * args.steps  # this rank's conv launches
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0

    traffic = None
    tpath = os.path.join(ROOT, 'profiles', f'r02_traffic_{args.precision}.json')
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, 'profiles', f'r01_traffic_{args.precision}.json')
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    if args.precision != 'fp32' and os.path.exists(tpath):
        tj = json.load(open(tpath))  # DRAM bytes of the 28 conv launches of one step, from the committed ncu --set full capture
        traffic = {'dram_gb_per_step': tj['dram_read_gb_per_step'] + tj['dram_write_gb_per_step'], 'source': tj['source']}
    if rank == 0:
        result = {
            'metric': 'idispnet_roi_crops_per_s', 'value': value, 'unit': 'ROIs/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warm, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'fp16x2': 'fp16x2 (operands and activations as hi+lo IEEE-half word pairs, fp32 accumulate: fp32-grade)'}.get(
                args.precision, f'{args.precision} (fp32 accumulate)'), 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': Bg, 'parallelism': f'dp{world} (ROI shards, one all-gather of disparity maps' + (', overlapped with the next step on NCCL\'s stream)' if world > 1 else ')'),
                       'precision_mode': args.precision,
                       'l2': 'no explicit flush: each step streams >10 GB of activations per GPU, far beyond the 126 MB L2'},
            'e2e': {'value': e2e, 'unit': 'ROIs/s', 'h2d_bytes_per_step': 2 * B_PER_GPU * C * HF * WF * 4 * world,
                    'd2h_bytes_per_step': B_PER_GPU * H * W * 4 * world, 'ms_per_step': ms_e2e / args.steps,
                    'api': 'idisp_plan_forward_host_async + idisp_plan_host_wait: per step pinned-host inputs in, pinned-host result out; '
                           'the copies of step i overlap the kernels of steps i-1 / i+1 (double-buffered staging); the timed region ends '
                           'after the host has waited for the last result',
                    'one_batch_at_a_time': {'value': Bg * args.steps / (ms_e2e_sync / 1e3), 'ms_per_step': ms_e2e_sync / args.steps,
                                            'api': 'idisp_plan_forward_host, host synchronises after every step (latency form)'}},
            'gpu_launches': launches_per_step * args.steps,
            'roofline': {'bound': 'tensor', 'kernel': '3-D conv launches (28 layers/step)', 'achieved': achieved,
                         'peak': tens_sus, 'unit': 'TFLOP/s', 'frac': achieved / tens_sus, 'traffic': traffic,
                         'peak_source': f'{peak_src} bf16 sustained (burst {tens_burst})',
                         'flop_per_roi': FLOP_PER_ROI, 'conv_ms_per_step': conv_ms / args.steps,
                         'other_ms_per_step': other_ms / args.steps,
                         'whole_step_frac': value / world * FLOP_PER_ROI / 1e12 / tens_sus,
                         # split precision issues three half-precision MMAs per algorithmic product (x_hi*w_hi, x_lo*w_hi,
                         # x_hi*w_lo): `achieved` counts ALGORITHMIC flops, the tensor pipe executes mma_passes times as many
                         'mma_passes': 3 if args.precision == 'fp16x2' else 1,
                         'executed_frac': achieved / tens_sus * (3 if args.precision == 'fp16x2' else 1)},
            'ms_by_layer': {str(k): round(v, 4) for k, v in sorted(by_layer.items())},
            'clocks': clocks,
        }
        if gather_info:
            result['allgather'] = gather_info
        if world == 1 and args.precision != 'fp32' and not os.environ.get('IDISP_TC_DBG'):
            # the parity (fp32 FFMA) mode on the same inputs and weights: its throughput, and how far the bf16 tensor-core
            # mode's disparities sit from it (fp32 mode itself is 2-7e-5 px from the reference, tests/test_gpu_parity.py)
            m32 = make_model('fp32', dev)
            m32.load_state_dict(m.state_dict())
            with torch.no_grad():
                d16 = m.forward_features(L[:4], R[:4])
                d32 = m32.forward_features(L[:4], R[:4])
                diff = (d16 - d32).abs()
                for _ in range(2):
                    m32.forward_features(L, R)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                m32.forward_features(L, R)
                e1.record(stream)
                torch.cuda.synchronize()
            result['fp32_parity_mode'] = {'value': B_PER_GPU / (e0.elapsed_time(e1) / 1e3), 'unit': 'ROIs/s',
                                          f'{args.precision}_vs_fp32_disparity_px': {'max': diff.max().item(), 'mean': diff.mean().item()}}
            if args.precision == 'fp16x2':
                # the one-word tensor-core mode on the same inputs: 3.5x the throughput, but 0.07 px from the reference
                # (tests/test_gpu_parity.py) -- outside north_star's 1e-3 bar, so it is reported beside the headline, not as it
                m16 = make_model('fp16', dev)
                m16.load_state_dict(m.state_dict())
                with torch.no_grad():
                    dh = (m16.forward_features(L[:4], R[:4]) - d32).abs()
                    for _ in range(3):
                        m16.forward_features(L, R)
                    torch.cuda.synchronize()
                    e0.record(stream)
                    for _ in range(3):
                        m16.forward_features(L, R)
                    e1.record(stream)
                    torch.cuda.synchronize()
                result['fp16_one_word_mode'] = {'value': 3 * B_PER_GPU / (e0.elapsed_time(e1) / 1e3), 'unit': 'ROIs/s',
                                                'fp16_vs_fp32_disparity_px': {'max': dh.max().item(), 'mean': dh.mean().item()},
                                                'note': 'fails the 1e-3 parity bar; not the headline'}
                del m16
            del m32
        model_sd = None
        if world == 1:
            result['roi_align'] = bench_roi_align(dev, hbm)
            result['cost_volume'] = bench_cost_volume(dev, hbm, L, R)
            if not os.environ.get('IDISP_BENCH_SKIP_LIVE'):
                result['live_shape'] = bench_live(dev)
        if world == 1 and not os.environ.get('IDISP_BENCH_SKIP_REFGPU') and not os.environ.get('IDISP_TC_DBG'):
            # north_star's ">= 4x over the reference GPU path": the same stack in eager PyTorch + cuDNN on THIS GPU, outside the timed
            # region, with its own clock record (tools/ref_gpu_timing.py: torch.nn restatement of SURVEY.md Appendix A, random weights)
            model_sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.startswith('feature_extraction')}
            del m
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import ref_gpu_timing
            rs = ClockSampler(local_rank)
            rs.start()
            best = None
            for chunk in (8, 32):   # give the reference its best batch split
                try:
                    r = ref_gpu_timing.measure(B_PER_GPU, chunk, 2, as_written=False, device=dev)
                except RuntimeError as e:   # (out of memory at the larger split)
                    r = None
                    torch.cuda.empty_cache()
                if r and (best is None or r['tf32_rois_per_s'] > best['tf32_rois_per_s']):
                    best = r
            rc = rs.stop()
            if best:
                result['reference_gpu'] = {
                    'what': 'eager PyTorch + cuDNN restatement of the 3-D stack (incl. device-side cost volume, upsample, softmax, regression) on this GPU',
                    'workload': best['workload'], 'tf32_rois_per_s': best['tf32_rois_per_s'], 'fp32_rois_per_s': best['fp32_rois_per_s'],
                    'tf32_note': 'allow_tf32=True is PyTorch\'s default for convolutions, i.e. what tools/test_net.py executes',
                    'tf32_disparity_vs_fp32_px': best['variants']['tf32 convs (PyTorch default allow_tf32=True), device-side cost volume']['disparity_vs_fp32_px'],
                    'speedup_vs_tf32': value / best['tf32_rois_per_s'], 'speedup_vs_fp32': value / best['fp32_rois_per_s'],
                    'e2e_speedup_vs_tf32': e2e / best['tf32_rois_per_s'], 'clocks': rc}
        if world == 1 and not args.no_cpu_baseline:
            if model_sd is None:
                model_sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.startswith('feature_extraction')}
            val, dt, cores, sample = cpu_port_rois_per_s(model_sd, 1, 1, budget_s=45.0)  # 1 warm-up: oneDNN primitive creation is per shape
            result['cpu_baseline'] = {'value': val, 'unit': 'ROIs/s', 'cores': cores, 'kind': 'port', 'sample': sample}
        print(json.dumps(result))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
