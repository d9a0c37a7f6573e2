"""CPU tests of the host-side logic: C-ABI symbol table, module mirror / state_dict contract, tap tables,
boundary re-nesting, config validation, the data-parallel gradient reducer over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='session')
def built_lib():
    from emsanet_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    """the C-ABI library loads without a GPU and exports every function include/emsanet_b200.h declares"""
    from emsanet_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'emsanet_b200.h')).read()
    declared = set(re.findall(r'\b(eb200_\w+)\s*\(', header))
    declared -= {'eb200_view', 'eb200_conv_desc', 'eb200_wgrad_desc', 'eb200_pack_entry'}
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS)
    assert declared == bound, (declared - bound, bound - declared)
    handle = ctypes.CDLL(built_lib)
    for name in declared:
        assert getattr(handle, name) is not None
    handle.eb200_version.restype = ctypes.c_int
    assert handle.eb200_version() >= 100


def test_struct_layouts_match_header(built_lib):
    """ctypes mirrors of the descriptor structs have the sizes nvcc computes for the header's structs"""
    from emsanet_b200 import _lib
    src = '#include "emsanet_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(eb200_view),' \
          'sizeof(eb200_conv_desc), sizeof(eb200_wgrad_desc), sizeof(eb200_pack_entry), sizeof(eb200_optim_entry),' \
          'sizeof(eb200_optim_hyper));return 0;}'
    exe = os.path.join(ROOT, 'emsanet_b200', 'lib', 'sizeof_check')
    subprocess.run(['gcc', '-x', 'c', '-', '-I', os.path.join(ROOT, 'include'), '-o', exe], input=src.encode(), check=True)
    sizes = list(map(int, subprocess.run([exe], capture_output=True, check=True).stdout.split()))
    assert sizes == [ctypes.sizeof(_lib.View), ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.WgradDesc),
                     ctypes.sizeof(_lib.PackEntry), ctypes.sizeof(_lib.OptimEntry), ctypes.sizeof(_lib.OptimHyper)]


def test_missing_library_fails_loudly(tmp_path):
    from emsanet_b200 import _lib
    saved = _lib._lib
    _lib._lib = None
    try:
        with pytest.raises(_lib.EB200Error, match='no CPU / PyTorch fallback'):
            _lib.load(str(tmp_path / 'nope.so'))
    finally:
        _lib._lib = saved


def test_module_mirror_state_dict_contract():
    """EMSANetB200 has the reference's keys/shapes/dtypes (checked against the oracle inventory, itself pinned to the
    reference by load_state_dict(strict=True) in oracle/make_golden.py) and reference-style init"""
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    from oracle import emsanet_oracle as O
    for kw_o, kw_a in [(dict(), dict()),
                       (dict(modalities=('rgb',), tasks=('semantic',), enable_panoptic=False),
                        dict(input_modalities=('rgb',), tasks=('semantic',), enable_panoptic=False)),
                       (dict(backbone='resnet101'), dict(rgb_encoder_backbone='resnet101', depth_encoder_backbone='resnet101'))]:
        m = EMSANetB200(default_args(**kw_a), simple_dataset_config())
        osd = O.make_state_dict(O.OracleConfig(**kw_o))
        sd = m.state_dict()
        assert list(sd.keys()) == list(osd.keys())
        assert all(sd[k].shape == osd[k].shape and sd[k].dtype == osd[k].dtype for k in sd)
        m.load_state_dict(osd, strict=True)
    m = EMSANetB200(default_args(), simple_dataset_config())
    sd = m.state_dict()
    assert len(sd) == 1056 and sum(p.numel() for p in m.parameters()) == 64246338
    # zero_residual_initialization (MT/model/initialization.py:69-81): decoder norm2 gains are 0, encoder's are 1
    assert sd['decoders.panoptic_helper.semantic_decoder.decoder_modules.0.blocks.0.norm2.weight'].abs().sum() == 0
    assert sd['encoder.backbone_rgb.layer1.0.norm2.weight'].min() == 1
    up = sd['decoders.panoptic_helper.semantic_decoder._task_head.upsample_0.conv.weight']
    assert torch.allclose(up[3, 0], torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16)
    assert list(m.decoders.keys()) == ['panoptic_helper', 'scene_decoder']
    assert m.decoders['panoptic_helper'].side_output_downscales == (16, 8, 4)


def test_unsupported_variants_raise():
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    for bad in (dict(activation='swish'), dict(context_module='appm-1-2-4-8'), dict(encoder_fusion='add-uni-rgb'),
                dict(rgb_encoder_backbone='resnet50', depth_encoder_backbone='resnet50'),
                dict(upsampling_prediction='bilinear'), dict(tasks=('semantic', 'normal')),
                dict(rgb_encoder_backbone_resnet_block='basicblock')):
        with pytest.raises(NotImplementedError):
            EMSANetB200(default_args(**bad), simple_dataset_config())


def test_cpu_model_refuses_to_run():
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    m = EMSANetB200(default_args(input_height=64, input_width=64), simple_dataset_config())
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m({'rgb': torch.zeros(2, 3, 64, 64), 'depth': torch.zeros(2, 1, 64, 64)})


def test_tap_tables_reproduce_strided_convolutions():
    """forward_taps (parity views) == torch conv2d index arithmetic for every filter/stride the path uses"""
    from emsanet_b200.ops import forward_taps
    import torch.nn.functional as F
    for kh, kw, sh, sw in [(3, 1, 1, 1), (1, 3, 1, 1), (3, 3, 1, 1), (1, 1, 1, 1), (3, 1, 2, 1), (1, 3, 1, 2), (1, 1, 2, 2)]:
        tt = forward_taps(kh, kw, sh, sw)
        assert len(tt.tap_view) == kh * kw and len(tt.views) <= 2
        x = torch.arange(2 * 1 * 8 * 10, dtype=torch.float64).reshape(2, 1, 8, 10)
        w = torch.randn(1, 1, kh, kw, dtype=torch.float64)
        ref = F.conv2d(x, w, None, (sh, sw), (kh // 2, kw // 2))
        ho, wo = ref.shape[2:]
        out = torch.zeros_like(ref)
        for t, (v, dy, dx) in enumerate(zip(tt.tap_view, tt.tap_dy, tt.tap_dx)):
            rp, cp = tt.views[v]
            view = x[:, :, (rp if rp is not None else 0)::(2 if rp is not None else 1),
                     (cp if cp is not None else 0)::(2 if cp is not None else 1)]
            vh, vw = view.shape[2:]
            for h in range(ho):
                for ww in range(wo):
                    hh, wx = h + dy, ww + dx
                    if 0 <= hh < vh and 0 <= wx < vw:
                        out[:, 0, h, ww] += w[0, 0, t // kw, t % kw] * view[:, 0, hh, wx]
        assert torch.allclose(out, ref)


def test_boundary_renesting_matches_reference_structure():
    """assemble_outputs reproduces SURVEY.md App. A for train and eval"""
    from emsanet_b200 import patch
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    m = EMSANetB200(default_args(), simple_dataset_config())
    m._eb200_engine = type('E', (), {'cfg': m._eb200_cfg})()
    t = lambda i: torch.full((1,), float(i))
    res = {'semantic': [t(0), t(1), t(2), t(3)], 'instance': [t(10 + i) for i in range(12)], 'scene': [t(99)]}
    m.train()
    out = patch.assemble_outputs(m, res, {}, False)
    (s, i), (ss, is_) = out[0]
    assert s is res['semantic'][0] and tuple(i) == tuple(res['instance'][:3])
    assert tuple(ss) == tuple(res['semantic'][1:]) and len(is_) == 3 and tuple(is_[1]) == tuple(res['instance'][6:9])
    assert out[1][0] is res['scene'][0] and out[1][1] is None
    m.eval()
    out = patch.assemble_outputs(m, {'semantic': [t(0)], 'instance': [t(1), t(2), t(3)], 'scene': [t(4)]}, {}, False)
    assert out[0][1] == ((None, None, None), (None, None, None))


def test_dropout_sites_and_config_roundtrip():
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    from oracle import emsanet_oracle as O
    m = EMSANetB200(default_args(), simple_dataset_config())
    cfg = m._eb200_cfg
    assert [(p, c) for p, c, _ in cfg.dropout_sites()] == [(p, c) for p, c, _ in O.dropout_sites(O.OracleConfig())]
    assert cfg.dropout_p_encoder == 0.1 and cfg.dropout_p_decoder == 0.2


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from emsanet_b200.ddp import GradAllReducer, shard_batch
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        red = GradAllReducer(engine=None)
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        enc_end = 600
        red.on_grads_ready(flat, enc_end, 1000)     # decoder + context bucket, mid-backward
        flat[:enc_end] += 1.0                        # "encoder backward" keeps writing its own slice meanwhile
        red.on_grads_ready(flat, 0, enc_end)
        red.finish()
        expect = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        expect[:enc_end] += 1.0
        lo, hi = shard_batch(257, rank, world)
        q.put((rank, bool(torch.allclose(flat, expect)), lo, hi))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """the bucketed mean all-reduce of the flat gradient buffer, world_size 2 over gloo on CPU"""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in results] == [True, True]
    assert (results[0][2], results[0][3], results[1][2], results[1][3]) == (0, 129, 129, 257)


class _StubEngine:
    """host-side stand-in for engine.Engine (which needs a GPU): same attributes the autograd node and the reducer use;
    backward() deposits rank-dependent gradients in a NEW flat buffer and reports the two buckets like the engine"""

    def __init__(self, params, rank):
        self.P = params
        self.grad_keys = list(params)
        self.rank = rank
        self.taps = None
        self.on_grads_ready = None
        self.calls = 0

    def forward(self, rgb, depth, training, track):
        return {'semantic': [rgb * 2.0]}

    def backward(self, by_task):
        self.calls += 1
        sizes = [self.P[k].numel() for k in self.grad_keys]
        self.flat_grad = torch.zeros(sum(sizes))
        self.grad_slices, off = [], 0
        for k, n in zip(self.grad_keys, sizes):
            self.grad_slices.append((off, n, tuple(self.P[k].shape)))
            self.flat_grad[off:off + n] = (self.rank + 1) * self.calls * torch.arange(1, n + 1, dtype=torch.float32)
            off += n
        enc_end = sizes[0]
        if self.on_grads_ready is not None:
            self.on_grads_ready(self.flat_grad, enc_end, off)
            self.on_grads_ready(self.flat_grad, 0, enc_end)


def _ddp_autograd_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['EB200_NO_GRAPH'] = '1'
    from emsanet_b200.ddp import GradAllReducer
    from emsanet_b200.patch import _EMSANetFunction
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        params = {'encoder.w': torch.nn.Parameter(torch.zeros(7, 3)), 'decoders.w': torch.nn.Parameter(torch.zeros(5))}
        eng = _StubEngine(params, rank)
        GradAllReducer(eng)
        mean_factor = sum(range(1, world + 1)) / world
        ok = True

        def step():
            out = _EMSANetFunction.apply(eng, [], torch.ones(2, 3), None, True, True, *params.values())
            out[0].sum().backward()

        def same_on_all_ranks(expect_scale):
            good = True
            for k, p_ in params.items():
                want = expect_scale * torch.arange(1, p_.numel() + 1, dtype=torch.float32).view(p_.shape)
                gathered = [torch.empty_like(p_.grad) for _ in range(world)]
                dist.all_gather(gathered, p_.grad.detach().clone())
                good &= all(torch.equal(g, gathered[0]) for g in gathered) and bool(torch.allclose(p_.grad, want))
            return good

        step()                                            # 1: .grad is None -> adopted
        ok &= same_on_all_ranks(mean_factor * 1)
        lo = eng.flat_grad.data_ptr()
        adopted = all(lo <= p_.grad.data_ptr() < lo + 4 * eng.flat_grad.numel() for p_ in params.values())
        step()                                            # 2: accumulation without zero_grad
        ok &= same_on_all_ranks(mean_factor * (1 + 2))
        for p_ in params.values():                        # 3: zero_grad(set_to_none=False)
            p_.grad.zero_()
        step()
        ok &= same_on_all_ranks(mean_factor * 3)
        q.put((rank, bool(ok), bool(adopted)))
    finally:
        dist.destroy_process_group()


def test_autograd_node_returns_reduced_gradients_two_ranks_gloo():
    """ADVICE r1 (patch.py:117): through the autograd node every rank must end up with the SAME, averaged `.grad` —
    with `.grad` adopted, accumulated into an existing `.grad`, and after zero_grad(set_to_none=False)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_ddp_autograd_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in results] == [True, True]
    assert [r[2] for r in results] == [True, True], 'gradient views were cloned instead of adopted'


REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout only exists in the build container')
def test_patch_on_the_real_reference_module():
    """patch() on an UNMODIFIED reference EMSANet instance: config is derived from its args, parameters stay the
    reference's objects, the forward is swapped per instance and restored by unpatch(); on CPU it refuses to run."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import make_golden as mg
    from oracle import emsanet_oracle as O
    mg.install_reference_shim()
    from emsanet.model import EMSANet
    from emsanet_b200 import patch as P
    cfg = O.OracleConfig()
    model = EMSANet(mg.make_args(cfg, 96, 128, dropout=0.1), mg.make_dataset_config(cfg))
    keys_before = list(model.state_dict().keys())
    params_before = [id(p) for p in model.parameters()]
    stock = model.forward
    P.patch(model)
    ecfg = P.config_from_model(model)
    assert ecfg.backbone == 'resnet34' and ecfg.modalities == ('rgb', 'depth') and ecfg.enable_panoptic
    assert ecfg.semantic_n_classes == 40 and ecfg.dropout_p_encoder == 0.1
    assert list(model.state_dict().keys()) == keys_before and [id(p) for p in model.parameters()] == params_before
    from emsanet_b200.module import param_inventory
    assert [k for k, _, _ in param_inventory(ecfg)] == keys_before
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model({'rgb': torch.zeros(2, 3, 96, 128), 'depth': torch.zeros(2, 1, 96, 128)})
    P.unpatch(model)
    assert model.forward == stock
    # unsupported variant of the real reference is rejected, not mis-handled
    bad = EMSANet(mg.make_args(O.OracleConfig(), 96, 128), mg.make_dataset_config(cfg))
    bad.args.activation = 'swish'
    with pytest.raises(NotImplementedError):
        P.patch(bad)


def test_lockstep_driver_pairs_and_survives_divergence():
    """Engine._drive: two generators advance in lock step (their requests are executed together), results are routed
    back to the right generator, and when one sibling finishes early the other is driven alone."""
    import torch
    from emsanet_b200.engine import Engine, EngineConfig
    eng = Engine(EngineConfig(), {'w': torch.zeros(1)})
    calls = []

    def fake_exec(reqs):
        calls.append(tuple(r[0] + str(r[1]) for r in reqs))
        return [f'out:{r[1]}' for r in reqs]
    eng._exec_requests = fake_exec

    def gen(tag, n):
        got = []
        for i in range(n):
            got.append((yield ('conv', f'{tag}{i}', {})))
        return got

    ra, rb = eng._drive([gen('a', 3), gen('b', 2)])
    assert ra == ['out:a0', 'out:a1', 'out:a2'] and rb == ['out:b0', 'out:b1']
    assert calls == [('conva0', 'convb0'), ('conva1', 'convb1'), ('conva2',)]
    calls.clear()
    (r,) = eng._drive([gen('s', 1)])
    assert r == ['out:s0'] and calls == [('convs0',)]


def test_lockstep_requests_are_not_paired_across_kinds():
    import torch
    from emsanet_b200 import ops
    from emsanet_b200.engine import Engine, EngineConfig
    eng = Engine(EngineConfig(), {'w': torch.zeros(1)})
    seen = []
    orig = (ops.conv2d, ops.conv2d_wgrad)
    ops.conv2d = lambda *a, **k: seen.append(('conv', a)) or 'c'
    ops.conv2d_wgrad = lambda *a, **k: seen.append(('wgrad', a)) or 'w'
    try:
        out = eng._exec_requests([('conv', (1,), {}), ('wgrad', (2,), {})])   # different kinds: executed one by one
    finally:
        ops.conv2d, ops.conv2d_wgrad = orig
    assert out == ['c', 'w'] and [s[0] for s in seen] == ['conv', 'wgrad']
    assert ops._defer_conv is None and ops._defer_wgrad is None


def test_bench_reference_arm_contract_and_gpu_arm_refuses_cpu():
    """`bench.py --impl reference` prints ONE JSON line with our arm's metric / unit / config and the contract's keys;
    without a CUDA device our own arm exits loudly instead of measuring a fallback."""
    import json
    import subprocess
    bench = os.path.join(ROOT, 'bench.py')
    out = subprocess.run([sys.executable, bench, '--impl', 'reference', '--steps', '1', '--warmup', '1', '--height', '96',
                          '--width', '128', '--backbone', 'resnet18'], capture_output=True, text=True, check=True)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'images/s' and d['higher_is_better'] is True
    assert d['metric'] == 'images/sec EMSANet R34-NBt1D 640x480 bf16 fwd+bwd'
    # the unmodified reference where scripts/install_reference.sh installed it (baseline/_ref), else the oracle port
    have_ref = os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'main.py')) or os.path.isdir('/root/reference')
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert len([ln for ln in out.stdout.splitlines() if ln.strip()]) == 1, 'stdout must carry the JSON line only'
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('config 2: EMSANet RGB-D') and d['gpu_launches'] == 0
    # the arm keeps our arm's config keys and the steps / warm-up it was given; what was sampled is under cpu_baseline
    assert set(d['config']) == {'workload', 'global_batch', 'parallelism', 'l2_policy', 'weights_repacked_each_step',
                                'cuda_graph'}
    assert d['steps'] == 1 and d['warmup'] == 1 and d['cpu_baseline']['sample_batch'] == 2
    # under torchrun OMP_NUM_THREADS=1 is inherited: the arm must still use every host core it may run on
    one = subprocess.run([sys.executable, bench, '--impl', 'reference', '--steps', '1', '--warmup', '0', '--height', '64',
                          '--width', '96', '--backbone', 'resnet18', '--config', '1'], capture_output=True, text=True,
                         check=True, env={**os.environ, 'OMP_NUM_THREADS': '1', 'EB200_BENCH_PORT_ONLY': '1'})
    d1 = json.loads([ln for ln in one.stdout.splitlines() if ln.startswith('{')][0])
    assert d1['cpu_baseline']['kind'] == 'port'
    assert d1['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))
    assert 'config 1' in d1['config']['workload'] and 'eval-mode forward' in d1['config']['workload']
    # a rank other than 0 does no work and prints nothing
    quiet = subprocess.run([sys.executable, bench, '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                           capture_output=True, text=True, check=True, env={**os.environ, 'RANK': '1', 'WORLD_SIZE': '2'})
    assert quiet.stdout.strip() == ''
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, bench, '--steps', '1', '--warmup', '0'], capture_output=True, text=True)
        assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_synthetic_nyuv2_layout(tmp_path):
    """the on-disk layout DS/datasets/nyuv2/dataset.py:60-178 reads (checked here without the reference; with it in
    tests/test_pipeline_reference.py): file lists, directories, dtypes, value ranges, determinism"""
    import json
    import cv2
    import numpy as np
    from emsanet_b200 import synthetic_nyuv2 as S
    root = str(tmp_path / 'ds')
    assert S.write_dataset(root, n_train=3, n_test=2, height=96, width=128, seed=7, with_normal=True) == (3, 2)
    for split, n in (('train', 3), ('test', 2)):
        names = open(os.path.join(root, f'{split}.txt')).read().split()
        assert names == [f'{i:04d}' for i in range(n)]
        for name in names:
            base = os.path.join(root, split)
            rgb = cv2.imread(os.path.join(base, 'rgb', name + '.png'), cv2.IMREAD_UNCHANGED)
            depth = cv2.imread(os.path.join(base, 'depth', name + '.png'), cv2.IMREAD_UNCHANGED)
            raw = cv2.imread(os.path.join(base, 'depth_raw', name + '.png'), cv2.IMREAD_UNCHANGED)
            sem = cv2.imread(os.path.join(base, 'semantic_40', name + '.png'), cv2.IMREAD_UNCHANGED)
            ins = cv2.imread(os.path.join(base, 'instance', name + '.png'), cv2.IMREAD_UNCHANGED)
            nrm = cv2.imread(os.path.join(base, 'normal', name + '.png'), cv2.IMREAD_UNCHANGED)
            assert rgb.shape == (96, 128, 3) and rgb.dtype == np.uint8 and nrm.shape == (96, 128, 3)
            assert depth.shape == (96, 128) and depth.dtype == np.uint16 and raw.dtype == np.uint16
            assert 713 <= depth.min() and depth.max() <= 9995 and (raw == 0).any()
            assert sem.dtype == np.uint8 and sem.max() <= 40 and ins.dtype == np.uint16
            assert ((ins > 0) <= (sem > 0)).all()                         # instances only on labelled pixels
            ori = json.load(open(os.path.join(base, 'orientations', name + '.json')))
            assert all(int(k) in np.unique(ins) and 0 <= v < 2 * np.pi + 1e-6 for k, v in ori.items())
            assert open(os.path.join(base, 'scene_class', name + '.txt')).read() in S.SCENES
    again = str(tmp_path / 'ds2')
    S.write_dataset(again, n_train=3, n_test=2, height=96, width=128, seed=7, with_normal=True)
    a = cv2.imread(os.path.join(root, 'train', 'semantic_40', '0001.png'), cv2.IMREAD_UNCHANGED)
    b = cv2.imread(os.path.join(again, 'train', 'semantic_40', '0001.png'), cv2.IMREAD_UNCHANGED)
    assert (a == b).all()


def test_gradient_buckets_follow_backward_completion_order():
    """data parallel: the flat gradient buffer is laid out as [encoder stem + stages 1-2 | encoder
    stages 3-4 | context module + decoders]; every parameter belongs to exactly one section and the section that can be
    all-reduced in the middle of the encoder's backward holds the bulk of the encoder"""
    from emsanet_b200.engine import Engine
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    m = EMSANetB200(default_args(), simple_dataset_config())
    numel = {0: 0, 1: 0, 2: 0}
    for k, p in m.named_parameters():
        sec = Engine._grad_section(k)
        numel[sec] += p.numel()
        if sec == 2:
            assert not k.startswith('encoder.')
        elif sec == 1:
            assert any(t in k for t in ('.layer3.', '.layer4.', 'fusions.3.', 'fusions.4.')), k
        else:
            assert k.startswith('encoder.') and not any(t in k for t in ('.layer3.', '.layer4.')), k
    assert sum(numel.values()) == 64246338
    assert numel[1] / (numel[0] + numel[1]) > 0.9          # 27.8 M of the encoder's 29.6 M parameters
    assert numel[0] * 4 < 8e6                               # what waits for the end of the step: 7.5 MB
