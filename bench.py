"""EMSANet R34-NBt1D 640x480 bf16 forward+backward throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path (N>1: launched by torchrun)
    python bench.py --impl reference [...]                       the reference algorithm (CPU oracle port) on host cores

A "step" is one forward+backward pass of the full RGB-D model (all tasks, train mode, Dropout2d active) over one
batch of 32 synthetic 640x480 images per GPU; loss = sum over outputs of mean(o^2).
  value : images/s with the batch resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e   : the same through the nn.Module API (`EMSANetB200.forward` / autograd backward) with the batch in pinned
          host memory: H2D copy of the inputs and D2H read of the loss inside the timed region
  roofline : tensor-core roofline of the dominant kernel (conv_tc_kernel), timed per launch with CUDA events
  cpu_baseline : the oracle port (same algorithm, torch fp32 on the host cores) on a bounded 2-image sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP_PER_IMG = 121.62          # SURVEY.md §8(d), config 2, 2*MAC over conv+linear, train mode
FWDBWD_GFLOP_PER_IMG = 362.93
FWDBWD_MB_PER_IMG = 1475.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='images per GPU')
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--backbone', default='resnet34')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, height, width, sample_n=2, backbone='resnet34'):
    """the reference's algorithm (oracle port: same ATen fp32 ops as the reference nn.Modules) on the host cores"""
    import torch
    from oracle import emsanet_oracle as O
    cfg = O.OracleConfig(backbone=backbone)
    sd = O.make_state_dict(cfg, seed=0)
    rgb, depth = O.make_inputs(sample_n, height, width, seed=1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward_backward(sd, cfg, rgb, depth)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return sample_n / mean, mean, torch.get_num_threads()


METRIC = 'images/sec EMSANet R34-NBt1D 640x480 bf16 fwd+bwd'   # BASELINE.json's metric; both arms print the same string


def reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(a.steps, 4))
    warmup = max(1, min(a.warmup, 1))
    ips, mean, cores = cpu_reference_run(steps, warmup, a.height, a.width, backbone=a.backbone)
    sample = f'{steps} timed fwd+bwd passes over 2 images {a.width}x{a.height} (fp32, NCHW, all host threads)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/s',
        'n_gpus': a.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': mean * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # our arm's config (same workload); what was actually timed on the host is in cpu_baseline.sample
        'config': {**workload_config(a, a.batch, f'dp{max(1, a.gpus)}'), 'reference_sample_batch': 2,
                   'reference_runs_on': 'host CPU of rank 0'},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(a, batch, par):
    return {'workload': f'full EMSANet RGB-D {a.backbone}-NBt1D, tasks semantic+scene+instance+orientation (panoptic), '
                        f'train-mode forward+backward, batch {batch} per GPU, {a.width}x{a.height}',
            'global_batch': batch * max(1, a.gpus), 'parallelism': par,
            'l2_policy': 'per-step working set (>10 GB of activations) exceeds the 126 MB L2; no explicit flush',
            'weights_repacked_each_step': True,
            'cuda_graph': os.environ.get('EB200_NO_GRAPH', '0') in ('', '0')}


class ClockSampler:
    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.idx),
                 '--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        # under load = top half of the samples (the sampler also sees idle gaps around the timed region)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {'sm_mhz': med, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons), 'samples': len(sm)}


def main():
    a = parse()
    if a.impl == 'reference':
        reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the EMSANet path has no CPU fallback); '
                         'use --impl reference for the host-CPU reference arm')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from emsanet_b200 import _lib, build as _build
    if not os.path.exists(_lib.LIB_PATH):
        if rank == 0:
            _build.build()
        if world > 1:
            dist.barrier()
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    from emsanet_b200.patch import _engine_for
    from emsanet_b200 import ops

    dev = torch.device('cuda', local)
    torch.manual_seed(0)
    margs = default_args(input_height=a.height, input_width=a.width)
    for m in ('rgb', 'depth', 'rgbd'):
        setattr(margs, f'{m}_encoder_backbone', a.backbone)
    model = EMSANetB200(margs, simple_dataset_config())
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():   # randomise BN tensors incl. the zero-initialised decoder norm2 gains
        for k, p in model.named_parameters():
            if 'norm' in k or 'downsample.1' in k:
                if k.endswith('weight'):
                    p.copy_((0.5 + torch.rand(p.shape, generator=g)) * (0.15 if k.endswith('norm2.weight') else 1.0))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
    model.to(dev).train()
    eng = _engine_for(model)
    N, H, W = a.batch, a.height, a.width
    gi = torch.Generator().manual_seed(1 + rank)
    rgb_h = torch.randn(N, 3, H, W, generator=gi).pin_memory()
    depth_h = torch.randn(N, 1, H, W, generator=gi).pin_memory()
    rgb_d, depth_d = rgb_h.to(dev), depth_h.to(dev)

    from emsanet_b200.ddp import GradAllReducer
    reducer = GradAllReducer(eng) if world > 1 else None   # bucketed NCCL mean all-reduce, overlapped with backward
    eng.force_repack = True   # a training step changes every weight: each timed step pays for the re-layout

    def step_eager():
        res = eng.forward(rgb_d, depth_d, True)
        gouts = {t: [o * (2.0 / o.numel()) for o in outs] for t, outs in res.items()}
        eng.backward(gouts)
        if reducer is not None:
            reducer.finish()

    # The whole step (weight re-layout, dropout masks, forward, loss gradient, backward) is recorded once into a CUDA
    # graph and replayed: ~1400 launches per step would otherwise be bound by the host's launch rate.
    use_graph = os.environ.get('EB200_NO_GRAPH', '0') in ('', '0')
    graph, graph_launches = None, 0
    if use_graph:
        saved_cb, eng.on_grads_ready = eng.on_grads_ready, None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_eager()                      # lazy one-time initialisation must not happen inside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            res = eng.forward(rgb_d, depth_d, True)
            gouts = {t: [o * (2.0 / o.numel()) for o in outs] for t, outs in res.items()}
            eng.backward(gouts)
            flat_static = eng.flat_grad
        graph_launches = _lib.launch_count() - l0
        del res, gouts
        eng.on_grads_ready = saved_cb

    def step_resident():
        if graph is None:
            return step_eager()
        graph.replay()
        if reducer is not None:               # data parallel: mean all-reduce of the flat gradient buffer (NCCL)
            reducer.on_grads_ready(flat_static, 0, flat_static.numel())
            reducer.finish()

    # e2e input pipeline: the usual pinned-memory prefetcher — the H2D copy of step i+1 is issued on a copy stream
    # while step i computes; every step still pays for one full copy of its own inputs inside the timed region.
    copy_stream = torch.cuda.Stream()
    staged = {}

    def stage_next():
        with torch.cuda.stream(copy_stream):                       # fresh device tensors (record_stream below guards reuse)
            staged['batch'] = {'rgb': rgb_h.to(dev, non_blocking=True), 'depth': depth_h.to(dev, non_blocking=True)}
            staged['event'] = torch.cuda.Event()
            staged['event'].record(copy_stream)

    def step_e2e():
        if 'batch' not in staged:
            stage_next()
        torch.cuda.current_stream().wait_event(staged['event'])
        batch = staged.pop('batch')
        for t in batch.values():
            t.record_stream(torch.cuda.current_stream())
        out = model(batch)
        stage_next()                                               # overlaps this step's backward
        loss = sum((o.float() ** 2).mean() for o in flatten(out))
        for p in model.parameters():
            p.grad = None
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return float(loss.item())

    def flatten(o):
        if o is None:
            return []
        if isinstance(o, (list, tuple)):
            return [t for x in o for t in flatten(x)]
        return [o]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launched = _lib.launch_count() - l0
        if graph is not None and fn is step_resident:
            launched = graph_launches * steps     # replayed launches are not seen by the C-ABI counter
        return ms, launched

    warm = max(3, a.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, a.steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    ips = N * world * a.steps / (ms / 1e3)
    ms_per_step = ms / a.steps

    e2e = None
    if not a.no_e2e:
        ms_e, _ = timed(step_e2e, a.steps, warm)      # same warm-up rule (>= 3) as the resident number
        e2e = {'value': N * world * a.steps / (ms_e / 1e3), 'unit': 'images/s',
               'h2d_bytes_per_step': int(rgb_h.numel() * 4 + depth_h.numel() * 4), 'd2h_bytes_per_step': 4,
               'ms_per_step': ms_e / a.steps}

    roofline = None
    if not a.no_roofline:
        roofline = conv_roofline(eng, ops, step_eager)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        cips, cmean, cores = cpu_reference_run(2, 1, H, W, backbone=a.backbone)
        cpu_baseline = {'value': cips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                        'sample': f'2 timed fwd+bwd passes over 2 images {W}x{H} (oracle port, fp32 NCHW, all host threads)'}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    line = {
        'metric': METRIC, 'value': ips, 'unit': 'images/s',
        'n_gpus': world, 'steps': a.steps, 'warmup': warm, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': workload_config(a, N, f'dp{world}'),
        'gpu_launches': int(launches), 'clocks': clocks, 'e2e': e2e, 'roofline': roofline,
        'cpu_baseline': cpu_baseline,
        'model_roofline': {
            'tensor_frac_of_measured': FWDBWD_GFLOP_PER_IMG * 1e9 * ips / world / (tf_peak * 1e12),
            'hbm_frac_of_measured': FWDBWD_MB_PER_IMG * 1e6 * ips / world / (hbm_peak * 1e9),
            'peaks': 'MEASURED_PEAKS.json' if peaks else 'fallback (B200_PROFILING.md)'},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def conv_roofline(eng, ops, step_fn):
    """Dominant kernel = conv3_tc_kernel: every 3-tap stride-1 convolution and data gradient of the step (the 3x1 / 1x3
    filters of the NBt1D blocks, ~77 % of the model's FLOPs).  Each launch is timed with CUDA events on the launching
    stream while the GPU is kept backlogged (so the events see device time, not host enqueue time) against its
    algorithmic FLOPs: 2 * pixels * Cout * Cin * 3, true channel counts.  `traffic` = DRAM bytes per launch from the
    committed ncu capture (profiles/r1_conv3_traffic.json), launch-weighted over the layer classes."""
    import torch
    records = []
    orig = ops.conv2d_raw

    def traced(views, tap_view, tap_dy, tap_dx, tap_w, weight, cin, cout, out_ptr, out_ext, out_strides, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(views, tap_view, tap_dy, tap_dx, tap_w, weight, cin, cout, out_ptr, out_ext, out_strides, **kw)
        e1.record()
        n, h, w = out_ext
        halo = len(tap_view) == 3 and all(v == 0 for v in tap_view) and cout % 64 == 0 and cin % 64 == 0
        records.append((e0, e1, 2.0 * n * h * w * cout * cin * len(tap_view), halo, cin))
    ops.conv2d_raw = traced
    saved_pair, eng.pair_siblings = eng.pair_siblings, False     # one descriptor per launch while tracing
    try:
        # CUDA events measure the launch itself only while the GPU is backlogged (otherwise they also see the host's
        # enqueue time): park the stream behind a ~0.3 s spin while the host queues up the step.
        torch.cuda._sleep(int(0.3 * 1.9e9))
        step_fn()
        torch.cuda.synchronize()
    finally:
        ops.conv2d_raw = orig
        eng.pair_siblings = saved_pair
    dom = [r for r in records if r[3]]
    tot_ms = sum(e0.elapsed_time(e1) for e0, e1, *_ in dom)
    tot_fl = sum(r[2] for r in dom)
    all_ms = sum(e0.elapsed_time(e1) for e0, e1, *_ in records)
    all_fl = sum(r[2] for r in records)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('bf16_tflops_sustained', 1400.0)
    ach = tot_fl / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r1_conv3_traffic.json')))['by_cin']
        vals = [tr[str(r[4])] for r in dom if str(r[4]) in tr]
        if vals:
            traffic = {'value': sum(vals) / len(vals), 'unit': 'MB per launch (dram read + write, launch-weighted)',
                       'source': 'profiles/r1_ncu_full_conv3_wgrad3.md'}
    except Exception:
        pass
    return {'bound': 'tensor', 'kernel': 'conv3_tc_kernel (3-tap halo implicit GEMM, tcgen05)', 'achieved': ach, 'peak': peak,
            'unit': 'TFLOP/s', 'frac': ach / peak, 'traffic': traffic, 'launches_per_step': len(dom),
            'avg_launch_us': tot_ms * 1e3 / max(1, len(dom)),
            'algorithmic_gflop_per_launch': tot_fl / 1e9 / max(1, len(dom)),
            'all_conv_launches': {'launches_per_step': len(records), 'achieved': all_fl / (all_ms / 1e3) / 1e12
                                  if all_ms > 0 else 0.0, 'unit': 'TFLOP/s'},
            'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)'
            if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'}


if __name__ == '__main__':
    main()
