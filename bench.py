"""EMSANet R34-NBt1D 640x480 bf16 forward+backward throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|4|5]    our CUDA path (N>1: launched by torchrun)
    python bench.py --impl reference [...]                                     the reference algorithm on the host cores

Default (= config 2, the configuration the metric is quoted on): a "step" is one forward+backward pass of the full
RGB-D model (all tasks, train mode, Dropout2d active) over one batch of 32 synthetic 640x480 images per GPU; loss = sum
over outputs of mean(o^2).
  value        images/s with the batch resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e          the same through the nn.Module API (`EMSANetB200.forward` / autograd backward) with the batch in pinned
               host memory: H2D copy of the inputs and D2H read of the loss inside the timed region
  roofline     tensor-core roofline of the dominant kernel (conv3_tc_kernel): every distinct launch of the step is
               re-issued as a run of identical launches inside ONE CUDA-event pair (operands rotated over buffers
               larger than L2), weighted by how often the step issues it
  cpu_baseline the oracle port (same ATen fp32 ops as the reference's nn.Modules) on a bounded sample, all host cores
  stock_torch_gpu   context: the same port executed by stock PyTorch/cuDNN on this GPU (fp32 NCHW as the reference
               runs it, and bf16 autocast + channels_last) — what the reference would do on this box
Other configurations of BASELINE.json: --config 1 (RGB-only semantic, batch 1, eval forward), 4 (ResNet101 1024x768
batch 8 fwd+bwd), 5 (full model, batch 1, eval forward = inference latency incl. H2D/D2H in e2e).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md §8(d) / BASELINE.md §2: algorithmic FLOP (2*MAC, conv+linear) and bytes per image
MODEL_COST = {   # config -> (GFLOP/img, MB/img)
    1: (59.33, 633.0 / 2),        # fwd only, eval (bytes: (78.50+79.73) M elems x 2 B bf16)
    2: (362.93, 1475.0),
    4: (1421.9, 4850.0),
    5: (119.67, 583.0),
}
PRESETS = {
    1: dict(backbone='resnet34', modalities=('rgb',), tasks=('semantic',), panoptic=False, batch=1, height=480,
            width=640, train=False),
    2: dict(backbone='resnet34', modalities=('rgb', 'depth'), tasks=('semantic', 'scene', 'instance', 'orientation'),
            panoptic=True, batch=32, height=480, width=640, train=True),
    4: dict(backbone='resnet101', modalities=('rgb', 'depth'), tasks=('semantic', 'scene', 'instance', 'orientation'),
            panoptic=True, batch=8, height=768, width=1024, train=True),
    5: dict(backbone='resnet34', modalities=('rgb', 'depth'), tasks=('semantic', 'scene', 'instance', 'orientation'),
            panoptic=True, batch=1, height=480, width=640, train=False),
}
METRIC = 'images/sec EMSANet R34-NBt1D 640x480 bf16 fwd+bwd'   # BASELINE.json's metric; both arms print the same string
METRICS = {1: 'images/sec EMSANet R34-NBt1D RGB-only semantic 640x480 eval forward (config 1)', 2: METRIC,
           4: 'images/sec EMSANet R101-NBt1D 1024x768 bf16 fwd+bwd (config 4)',
           5: 'images/sec EMSANet R34-NBt1D 640x480 bf16 batch-1 eval forward (config 5)'}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=sorted(PRESETS))
    ap.add_argument('--batch', type=int, default=None, help='images per GPU (default: the configuration\'s)')
    ap.add_argument('--height', type=int, default=None)
    ap.add_argument('--width', type=int, default=None)
    ap.add_argument('--backbone', default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-stock-torch', action='store_true')
    a = ap.parse_args()
    p = dict(PRESETS[a.config])
    for k in ('batch', 'height', 'width', 'backbone'):
        if getattr(a, k) is None:
            setattr(a, k, p[k])
    a.modalities, a.tasks, a.panoptic, a.train = p['modalities'], p['tasks'], p['panoptic'], p['train']
    return a


def workload_config(a, par):
    tasks = '+'.join(a.tasks) + (' (panoptic)' if a.panoptic else '')
    mode = 'train-mode forward+backward' if a.train else 'eval-mode forward'
    return {'workload': f'config {a.config}: EMSANet {"RGB-D" if len(a.modalities) == 2 else a.modalities[0]} '
                        f'{a.backbone}-NBt1D, tasks {tasks}, {mode}, batch {a.batch} per GPU, {a.width}x{a.height}',
            'global_batch': a.batch * max(1, a.gpus), 'parallelism': par,
            'l2_policy': ('per-step working set (>10 GB of activations) exceeds the 126 MB L2; no explicit flush'
                          if a.train else 'L2 flushed between timed iterations (256 MB memset)'),
            'weights_repacked_each_step': bool(a.train),
            'cuda_graph': os.environ.get('EB200_NO_GRAPH', '0') in ('', '0')}


# ------------------------------------------------------------------------------------------------ host-CPU reference
def _oracle_cfg(a):
    from oracle import emsanet_oracle as O
    return O, O.OracleConfig(backbone=a.backbone, modalities=a.modalities, tasks=a.tasks, enable_panoptic=a.panoptic)


def cpu_reference_run(a, steps, warmup, sample_n):
    """The reference's algorithm on the host cores: the oracle port issues the same ATen fp32 NCHW ops as the
    reference's nn.Modules (bit-identical to them, oracle/make_golden.py).  The reference itself is pure Python and
    does not exist on the GPU box (/root/reference is only in the build container; it is not pip-installable:
    the top level has no setup.py / pyproject.toml) -> kind "port".  All host threads: under torchrun OMP_NUM_THREADS
    is forced to 1, so the thread count is set explicitly."""
    import torch
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    O, cfg = _oracle_cfg(a)
    sd = O.make_state_dict(cfg, seed=0)
    rgb, depth = O.make_inputs(sample_n, a.height, a.width, seed=1)
    rgb = rgb if 'rgb' in a.modalities else None
    depth = depth if 'depth' in a.modalities else None
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if a.train:
            O.forward_backward(sd, cfg, rgb, depth)
        else:
            with torch.no_grad():
                O.forward(sd, cfg, rgb, depth, False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return sample_n / mean, mean, torch.get_num_threads(), min(times)


def reference_model_run(a, steps, warmup, sample_n):
    """The UNMODIFIED reference (`emsanet.model.EMSANet`, built by its own argument parser and dataset config) on the
    host cores — available where scripts/install_reference.sh has put it into baseline/_ref (or /root/reference
    exists).  Returns None when there is no reference install (-> the bit-identical oracle port is timed instead)."""
    import torch
    from emsanet_b200 import run as launcher
    try:
        root = launcher.find_reference()
    except FileNotFoundError:
        return None
    launcher.setup_paths(root)
    from emsanet.args import ArgParserEMSANet
    from emsanet.data import get_dataset
    from emsanet.model import EMSANet
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    argv = ['--dataset', 'nyuv2', '--tasks', *a.tasks, '--input-modalities', *a.modalities,
            '--rgb-encoder-backbone', a.backbone, '--depth-encoder-backbone', a.backbone,
            '--input-height', str(a.height), '--input-width', str(a.width), '--no-pretrained-backbone',
            '--wandb-mode', 'disabled', '--dropout-p', '0.1']
    if a.panoptic:
        argv.append('--enable-panoptic')
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):       # the reference prints while it builds; stdout carries ONE JSON line
        args = ArgParserEMSANet().parse_args(argv, verbose=False)
        torch.manual_seed(0)
        model = EMSANet(args, get_dataset(args, split=args.validation_split).config)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():   # same randomisation of the BatchNorm tensors as our arm
        for k, p in model.named_parameters():
            if 'norm' in k or 'downsample.1' in k:
                if k.endswith('weight'):
                    p.copy_((0.5 + torch.rand(p.shape, generator=g)) * (0.15 if k.endswith('norm2.weight') else 1.0))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
    model.train(a.train)
    gi = torch.Generator().manual_seed(1)
    batch = {}
    if 'rgb' in a.modalities:
        batch['rgb'] = torch.randn(sample_n, 3, a.height, a.width, generator=gi)
    if 'depth' in a.modalities:
        batch['depth'] = torch.randn(sample_n, 1, a.height, a.width, generator=gi)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if a.train:
            out = model(batch)
            loss = sum((o.float() ** 2).mean() for o in flatten(out))
            model.zero_grad(set_to_none=True)
            loss.backward()
        else:
            with torch.no_grad():
                model(batch)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return sample_n / mean, mean, torch.get_num_threads(), min(times), root


def _sample_batch(a):
    return max(1, min(a.batch, 2))


def reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    n = _sample_batch(a)
    kind, where = 'port', 'oracle port (bit-identical to the reference modules)'
    ref = None if os.environ.get('EB200_BENCH_PORT_ONLY') else reference_model_run(a, steps, warmup, n)
    if ref is not None:
        ips, mean, cores, best, root = ref
        kind, where = 'reference', f'unmodified reference EMSANet from {os.path.relpath(root, ROOT)}'
    else:
        ips, mean, cores, best = cpu_reference_run(a, steps, warmup, n)
    what = 'fwd+bwd passes' if a.train else 'eval forward passes'
    sample = (f'each step = one {what[:-2]} over {n} image(s) {a.width}x{a.height} of the workload (fp32, NCHW, '
              f'{cores} host threads); {steps} timed after {warmup} warm-up')
    line = {
        'impl': 'reference', 'metric': METRICS[a.config], 'value': ips, 'unit': 'images/s',
        'n_gpus': a.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': mean * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(a, f'dp{max(1, a.gpus)}'),      # identical to our arm's
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': kind, 'sample': sample,
                         'sample_batch': n, 'best_step_ms': best * 1e3, 'runs_on': 'host CPU of rank 0', 'what': where},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def stock_torch_gpu(a, dev, steps=5, warmup=3):
    """Context line: the port's ATen ops executed by stock PyTorch / cuDNN on THIS GPU, same batch —
    what the reference's nn.Modules would do here.  (a) fp32 NCHW, cudnn.benchmark as main.py:23-24 sets it;
    (b) bf16 autocast + channels_last.  Not part of the product path; measured after our numbers."""
    import torch
    O, cfg = _oracle_cfg(a)
    out = {}
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev) for k, v in O.make_state_dict(cfg, seed=0).items()}
    rgb, depth = (t.to(dev) for t in O.make_inputs(a.batch, a.height, a.width, seed=1))
    rgb = rgb if 'rgb' in a.modalities else None
    depth = depth if 'depth' in a.modalities else None
    for name in ('fp32_nchw', 'bf16_autocast_channels_last'):
        try:
            s, r, d = sd, rgb, depth
            if name != 'fp32_nchw':
                s = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
                r = r.contiguous(memory_format=torch.channels_last) if r is not None else None
                d = d.contiguous(memory_format=torch.channels_last) if d is not None else None

            def step():
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=name != 'fp32_nchw'):
                    if a.train:
                        O.forward_backward(s, cfg, r, d)
                    else:
                        with torch.no_grad():
                            O.forward(s, cfg, r, d, False)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'value': a.batch / (ms / 1e3), 'unit': 'images/s', 'ms_per_step': ms, 'steps': steps,
                         'warmup': warmup}
        except Exception as e:      # context only: never take the bench line down
            out[name] = {'error': f'{type(e).__name__}: {e}'[:200]}
        torch.cuda.empty_cache()
    out['what'] = ('oracle port (the reference modules\' ATen ops) run by stock PyTorch ' + torch.__version__ +
                   ' / cuDNN on this GPU, same batch, inputs resident, eager; context, not the reference arm')
    return out


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


def flatten(o):
    if o is None:
        return []
    if isinstance(o, (list, tuple)):
        return [t for x in o for t in flatten(x)]
    return [o]


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
    margs = default_args(input_height=a.height, input_width=a.width, input_modalities=a.modalities, tasks=a.tasks,
                         enable_panoptic=a.panoptic)
    for m in ('rgb', 'depth', 'rgbd'):
        setattr(margs, f'{m}_encoder_backbone', a.backbone)
    if len(a.modalities) == 1:
        margs.encoder_fusion = 'none'
    model = EMSANetB200(margs, simple_dataset_config())
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():   # randomise BN tensors incl. the zero-initialised decoder norm2 gains
        for k, p in model.named_parameters():
            if 'norm' in k or 'downsample.1' in k:
                if k.endswith('weight'):
                    p.copy_((0.5 + torch.rand(p.shape, generator=g)) * (0.15 if k.endswith('norm2.weight') else 1.0))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
    model.to(dev).train(a.train)
    eng = _engine_for(model)
    N, H, W = a.batch, a.height, a.width
    gi = torch.Generator().manual_seed(1 + rank)
    rgb_h = torch.randn(N, 3, H, W, generator=gi).pin_memory() if 'rgb' in a.modalities else None
    depth_h = torch.randn(N, 1, H, W, generator=gi).pin_memory() if 'depth' in a.modalities else None
    rgb_d = rgb_h.to(dev) if rgb_h is not None else None
    depth_d = depth_h.to(dev) if depth_h is not None else None
    h2d_bytes = sum(int(t.numel() * 4) for t in (rgb_h, depth_h) if t is not None)

    from emsanet_b200.ddp import GradAllReducer
    reducer = GradAllReducer(eng) if (world > 1 and a.train) else None   # NCCL mean all-reduce of the flat gradient
    eng.force_repack = a.train   # a training step changes every weight: each timed step pays for the re-layout
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if not a.train else None

    def step_eager():
        if not a.train:
            with torch.no_grad():
                return eng.forward(rgb_d, depth_d, False)
        res = eng.forward(rgb_d, depth_d, True)
        gouts = {t: [o * (2.0 / o.numel()) for o in outs] for t, outs in res.items()}
        eng.backward(gouts)
        if reducer is not None:
            reducer.finish()

    # The whole step (weight re-layout, dropout masks, forward, loss gradient, backward) is recorded once into CUDA
    # graphs and replayed: ~1100 launches per step would otherwise be bound by the host's launch rate.  Data parallel:
    # the step is split into THREE graphs where a bucket of the flat gradient buffer becomes final ([decoders + context] at
    # the encoder boundary, [encoder stages 3-4] after them, the rest at the end); each bucket is all-reduced on NCCL's
    # stream while the next graph runs.
    use_graph = os.environ.get('EB200_NO_GRAPH', '0') in ('', '0')
    graphs, graph_launches, flat_static = [], 0, None
    if use_graph:
        saved_cb, eng.on_grads_ready = eng.on_grads_ready, None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_eager()                      # lazy one-time initialisation must not happen inside the capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        pool = torch.cuda.graph_pool_handle()
        g1 = torch.cuda.CUDAGraph()
        from emsanet_b200.graphs import _NoGC
        nogc = _NoGC()
        nogc.__enter__()      # no cyclic GC (it may destroy other graphs / pools) while the captures below run
        if not a.train:
            with torch.no_grad(), torch.cuda.graph(g1, pool=pool):
                res = eng.forward(rgb_d, depth_d, False)
            graphs = [g1]
        elif reducer is None:
            with torch.cuda.graph(g1, pool=pool):
                res = eng.forward(rgb_d, depth_d, True)
                gouts = {t: [o * (2.0 / o.numel()) for o in outs] for t, outs in res.items()}
                eng.backward(gouts)
                flat_static = eng.flat_grad
            graphs = [g1]
        else:
            flat_static = torch.zeros(eng.param_grad_floats(), dtype=torch.float32, device=dev)
            with torch.cuda.graph(g1, pool=pool):
                res = eng.forward(rgb_d, depth_d, True)
                gouts = {t: [o * (2.0 / o.numel()) for o in outs] for t, outs in res.items()}
                eng.begin_backward(gouts, flat=flat_static)
                marker = eng.run_tape(stop_at_marker=True)
            graphs, bucket_ranges = [g1], [eng.ready_range(marker)]
            while eng.tape:             # one more graph per gradient bucket (encoder stages 3-4, then the rest)
                gn = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gn, pool=pool):
                    marker = eng.run_tape(stop_at_marker=True)
                graphs.append(gn)
                bucket_ranges.append(eng.ready_range(marker))
            eng.grads = None
        nogc.__exit__()
        graph_launches = _lib.launch_count() - l0
        eng.on_grads_ready = saved_cb

    def step_resident():
        if not graphs:
            return step_eager()
        if flush is not None:
            flush.zero_()                     # L2 flush between timed inference iterations (inside the timed region)
        if reducer is None:
            graphs[0].replay()
            return
        for g, (lo, hi) in zip(graphs, bucket_ranges):
            g.replay()
            reducer.on_grads_ready(flat_static, lo, hi)     # all-reduce of this bucket overlaps the next graph
        reducer.finish()

    # e2e input pipeline: the usual pinned-memory prefetcher — the H2D copy of step i+1 is issued on a copy stream
    # while step i computes; every step still pays for one full copy of its own inputs inside the timed region.
    copy_stream = torch.cuda.Stream()
    staged = {}

    def stage_next():
        with torch.cuda.stream(copy_stream):                       # fresh device tensors (record_stream below guards reuse)
            staged['batch'] = {k: t.to(dev, non_blocking=True) for k, t in (('rgb', rgb_h), ('depth', depth_h))
                               if t is not None}
            staged['event'] = torch.cuda.Event()
            staged['event'].record(copy_stream)

    d2h = {'bytes': 4}

    class _MeanSquares(torch.autograd.Function):
        """the stand-in loss sum_i mean(o_i^2) over the 17 outputs (1.8 GB fp32 at batch 32) in 2 + 1 passes: a norm
        reduction per output forward, ONE scaled copy per output backward.  Spelled `(o ** 2).mean()` autograd runs
        ~7 elementwise passes over the 1.8 GB (4.0 ms of torch kernels per step in `scripts/e2e_breakdown.py`'s
        profile, against 1.2 ms here) — time that belongs to the loss spelling, not to the path under test."""
        @staticmethod
        def forward(ctx, *outs):
            ctx.save_for_backward(*outs)
            return sum(torch.linalg.vector_norm(o).square() / o.numel() for o in outs)

        @staticmethod
        def backward(ctx, g):
            return tuple(o * (g * (2.0 / o.numel())) for o in ctx.saved_tensors)

    mean_squares = _MeanSquares.apply

    def step_e2e():
        if a.train:
            if 'batch' not in staged:
                stage_next()
            torch.cuda.current_stream().wait_event(staged['event'])
            batch = staged.pop('batch')
            for t in batch.values():
                t.record_stream(torch.cuda.current_stream())
            out = model(batch)
            stage_next()                                               # overlaps this step's backward
            loss = mean_squares(*flatten(out))
            for p in model.parameters():
                p.grad = None
            loss.backward()
            if reducer is not None:
                reducer.finish()
            return float(loss.item())
        # inference (time_inference_pytorch protocol, inference_time_whole_model.py:297-347): host inputs -> device,
        # forward, all outputs back on the host, one image at a time, nothing overlapped
        # (the reference calls `.cpu()`, i.e. a fresh pageable allocation per output and step — 49 MB of page faults for
        # the semantic logits, 5-25 ms of host time that has nothing to do with the path measured; results land in
        # pinned buffers allocated once, as a serving loop would hold them)
        with torch.no_grad():
            batch = {k: t.to(dev, non_blocking=True) for k, t in (('rgb', rgb_h), ('depth', depth_h)) if t is not None}
            res = flatten(model(batch))
            if 'host' not in d2h:
                d2h['host'] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in res]
            for hbuf, o in zip(d2h['host'], res):
                hbuf.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        d2h['bytes'] = sum(int(o.numel() * o.element_size()) for o in res)
        return d2h['host']

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
        if graphs and fn is step_resident:
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
               'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h['bytes'], 'ms_per_step': ms_e / a.steps}

    roofline = None
    if not a.no_roofline and a.train:
        roofline = conv_roofline(eng, ops, _lib, step_eager)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = None
    if world == 1 and not a.no_cpu_baseline:
        n = _sample_batch(a)
        cips, cmean, cores, best = cpu_reference_run(a, 2, 1, n)
        what = 'fwd+bwd passes' if a.train else 'eval forward passes'
        cpu_baseline = {'value': cips, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                        'sample': f'2 timed {what} over {n} image(s) {W}x{H} (oracle port, fp32 NCHW, all host threads)'}
    stock = None
    if world == 1 and not a.no_stock_torch:
        del graphs[:]
        torch.cuda.empty_cache()
        stock = stock_torch_gpu(a, dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    gflop, mb = MODEL_COST[a.config]
    line = {
        'metric': METRICS[a.config], 'value': ips, 'unit': 'images/s',
        'n_gpus': world, 'steps': a.steps, 'warmup': warm, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': workload_config(a, f'dp{world}'),
        'gpu_launches': int(launches), 'clocks': clocks, 'e2e': e2e, 'roofline': roofline,
        'cpu_baseline': cpu_baseline, 'stock_torch_gpu': stock,
        'model_roofline': {
            'tensor_frac_of_measured': gflop * 1e9 * ips / world / (tf_peak * 1e12),
            'hbm_frac_of_measured': mb * 1e6 * ips / world / (hbm_peak * 1e9),
            'gflop_per_image': gflop, 'mb_per_image': mb,
            'peaks': 'MEASURED_PEAKS.json' if peaks else 'fallback (B200_PROFILING.md)'},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ kernel roofline
def conv_roofline(eng, ops, _lib, step_fn, l2_bytes=126 << 20):
    """Dominant kernel = conv3_tc_kernel: every 3-tap stride-1 convolution and data gradient of the step (the 3x1 / 1x3
    filters of the NBt1D blocks, ~77 % of the model's FLOPs).

    1. One eager step is traced at the C-ABI boundary: every eb200_conv2d / eb200_conv2d_pair call and its descriptors
       (sibling pair launches stay pair launches, exactly as the graph-replayed step issues them).
    2. Every DISTINCT halo-kernel launch (shape, filter orientation, epilogue flags, single / pair) is re-issued as a
       run of R identical launches inside ONE CUDA-event pair on the launching stream (after 3 warm-up launches), with
       its activation operands rotated over enough private buffers that the run's working set exceeds 2x L2 — each
       launch reads cold operands, as in the step; programmatic dependent launch overlaps neighbours, as in the step.
    3. achieved = sum_launches(algorithmic FLOPs) / sum_launches(duration of its class); FLOPs = 2 * pixels * Cout * Cin * 3
       with true channel counts.  `traffic` = DRAM bytes per launch from the committed ncu capture."""
    import torch
    from emsanet_b200._lib import ConvDesc
    calls = []
    orig_call = _lib.call

    def traced(name, *args):
        if name in ('eb200_conv2d', 'eb200_conv2d_pair'):
            nd = 1 if name == 'eb200_conv2d' else 2
            calls.append((name, [ConvDesc.from_buffer_copy(args[i]._obj) for i in range(nd)]))
        return orig_call(name, *args)
    _lib.call = traced
    try:
        step_fn()
        torch.cuda.synchronize()
    finally:
        _lib.call = orig_call

    def is_halo(d):
        return (d.taps == 3 and all(d.tap_view[i] == 0 for i in range(3)) and d.cout % 64 == 0 and d.cin % 64 == 0
                and d.inp[0].c % 64 == 0)

    def flops(d):
        return 2.0 * d.n * d.h * d.w * d.cout * d.cin * d.taps

    def key(name, ds):
        return (name,) + tuple((d.n, d.h, d.w, d.cin, d.cout, tuple(d.tap_dy[:3]), tuple(d.tap_dx[:3]), d.flags,
                                d.inp[0].sn, d.inp[0].sh, d.inp[0].sw, d.out_sn, d.out_sh, d.out_sw) for d in ds)
    classes = {}
    for name, ds in calls:
        if not all(is_halo(d) for d in ds):
            continue
        c = classes.setdefault(key(name, ds), {'name': name, 'descs': ds, 'count': 0})
        c['count'] += 1
    stream = torch.cuda.current_stream().cuda_stream
    dev = eng.dev
    results = []
    for c in classes.values():
        per_launch_bytes = 0
        for d in c['descs']:
            per_launch_bytes += 2 * (d.inp[0].n * d.inp[0].sn + d.n * d.out_sn + (d.n * d.aux_sn if d.aux else 0))
        k_sets = int(min(24, max(2, (2 * l2_bytes) // max(1, per_launch_bytes) + 1)))
        sets, keep = [], []
        for _ in range(k_sets):
            ds = []
            for d in c['descs']:
                d2 = ConvDesc.from_buffer_copy(d)
                x = torch.randn(d.inp[0].n * d.inp[0].sn, device=dev, dtype=torch.bfloat16)
                y = torch.empty(d.n * d.out_sn, device=dev, dtype=torch.bfloat16)
                d2.inp[0].ptr, d2.out = x.data_ptr(), y.data_ptr()
                keep += [x, y]
                if d.aux:
                    aux = y if d.aux == d.out else torch.randn(d.n * d.aux_sn, device=dev, dtype=torch.bfloat16)
                    d2.aux = aux.data_ptr()
                    keep.append(aux)
                if d.stats:
                    st = torch.zeros(2 * max(d.cout_pad, d.cout) + 64, device=dev, dtype=torch.float32)
                    d2.stats = st.data_ptr()
                    keep.append(st)
                ds.append(d2)
            sets.append(ds)

        def launch(i):
            ds = sets[i % k_sets]
            orig_call(c['name'], *[ctypes.byref(d) for d in ds], stream)
        reps = max(24, 2 * k_sets)
        for i in range(3):
            launch(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            launch(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        fl = sum(flops(d) for d in c['descs'])
        d0 = c['descs'][0]
        results.append({'launch': c['name'].replace('eb200_', ''), 'n': d0.n, 'h': d0.h, 'w': d0.w, 'c': d0.cin,
                        'flags': d0.flags, 'per_step': c['count'], 'us': us, 'tflops': fl / us / 1e6,
                        'rotating_operand_sets': k_sets, 'launches_timed': reps})
        del sets, keep
    torch.cuda.empty_cache()
    tot_us = sum(r['us'] * r['per_step'] for r in results)
    tot_fl = sum(r['tflops'] * 1e6 * r['us'] * r['per_step'] for r in results)
    n_launch = sum(r['per_step'] for r in results)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('bf16_tflops_sustained', 1400.0)
    ach = tot_fl / tot_us / 1e6 if tot_us > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r2_conv3_traffic.json')))['by_cin']
        vals = [tr[str(r['c'])] for r in results for _ in range(r['per_step']) if str(r['c']) in tr]
        if vals:
            traffic = {'value': sum(vals) / len(vals), 'unit': 'MB per layer (dram read + write, launch-weighted)',
                       'source': 'profiles/r2_ncu_full_conv3_wgrad3.md'}
    except Exception:
        pass
    results.sort(key=lambda r: -r['us'] * r['per_step'])
    return {'bound': 'tensor', 'kernel': 'conv3_tc_kernel (3-tap halo implicit GEMM, tcgen05)', 'achieved': ach,
            'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak, 'traffic': traffic, 'launches_per_step': n_launch,
            'avg_launch_us': tot_us / max(1, n_launch), 'ms_per_step_in_this_kernel': tot_us / 1e3,
            'algorithmic_gflop_per_launch': tot_fl / 1e9 / max(1, n_launch),
            'method': 'per distinct launch: run of identical launches in one CUDA-event pair, operands rotated over '
                      '> 2x L2 of buffers; weighted by launches per step',
            'classes': results[:12],
            'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long run)'
            if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'}


if __name__ == '__main__':
    main()
