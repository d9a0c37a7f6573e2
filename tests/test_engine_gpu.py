"""End-to-end parity (GPU): the CUDA engine against the CPU oracle on the same seeded weights and inputs.

Tolerance: the reference computes in fp32; this path stores every activation as bf16 (2^-8 relative rounding per
stored tensor) with fp32 accumulation, so end-to-end deviations are rounding noise accumulated over ~100 layers.
Outputs and parameter gradients are compared by relative L2 error per tensor; class maps by arg-max agreement on
pixels whose oracle top-2 margin exceeds the observed output error bound (near-ties legitimately flip in bf16)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

FACTOR = 4.0           # allowed multiple of the network's own bf16 sensitivity (see _budget)
OUT_FLOOR = 1e-2       # rel-L2 floor for outputs (one bf16 rounding is 4e-3)
GRAD_FLOOR = 3e-2      # rel-L2 floor for parameter gradients
STAT_FLOOR = 5e-3


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _bf16_round(sd):
    return {k: (v.to(torch.bfloat16).float() if (v.is_floating_point() and v.dim() == 4) else v)
            for k, v in sd.items()}


def _setup(kw, n, h, w, dropout=False):
    """Seeded weights/inputs on both sides.  The residual-branch BN gains (norm2.weight) are scaled to
    U(0.075, 0.225): with gains ~1 a *random* 16-block residual BN network amplifies any perturbation ~150x
    end to end in train mode (measured on the fp32 oracle itself), which would turn bf16 rounding into O(1)
    output differences and make an end-to-end comparison meaningless; trained / zero-init-residual networks
    (the reference zero-initialises the decoder norm2 gains, emsanet/model.py:189-190) are not in that regime."""
    from oracle import emsanet_oracle as O
    from emsanet_b200.engine import Engine, EngineConfig
    ocfg = O.OracleConfig(**kw)
    sd = O.make_state_dict(ocfg, seed=0)
    for k in sd:
        if k.endswith('norm2.weight'):
            sd[k] = sd[k] * 0.15
    rgb, depth = O.make_inputs(n, h, w, seed=1)
    if 'rgb' not in ocfg.modalities:
        rgb = None
    if 'depth' not in ocfg.modalities:
        depth = None
    ecfg = EngineConfig(backbone=ocfg.backbone, modalities=ocfg.modalities, tasks=ocfg.tasks,
                        enable_panoptic=ocfg.enable_panoptic, semantic_n_classes=ocfg.semantic_n_classes,
                        scene_n_classes=ocfg.scene_n_classes,
                        dropout_p_encoder=0.1 if dropout else 0.0, dropout_p_decoder=0.2 if dropout else 0.0)
    params = {k: v.cuda() for k, v in sd.items()}
    eng = Engine(ecfg, params)
    return O, ocfg, sd, rgb, depth, eng


def _r(t):
    return None if t is None else t.to(torch.bfloat16).float()


def _flat_engine(res):
    """engine result -> the oracle's depth-first flat order"""
    flat = []
    if 'semantic' in res and 'instance' in res:
        s, i = res['semantic'], res['instance']
        nt = 3
        flat += [s[0]] + list(i[:nt]) + list(s[1:]) + list(i[nt:])
    else:
        for t in ('semantic', 'instance'):
            if t in res:
                flat += list(res[t])
    if 'scene' in res:
        flat += list(res['scene'])
    return flat


CASES = {
    'full_rgbd_r34': (dict(), 4, 96, 128),
    'rgb_semantic_r34': (dict(modalities=('rgb',), tasks=('semantic',), enable_panoptic=False), 4, 64, 96),
    'full_rgbd_r18_ragged': (dict(backbone='resnet18'), 5, 96, 160),
}


@pytest.mark.parametrize('name', list(CASES))
def test_eval_forward_matches_oracle(name):
    kw, n, h, w = CASES[name]
    O, ocfg, sd, rgb, depth, eng = _setup(kw, n, h, w)
    with torch.no_grad():
        # calibrate the running statistics (a trained checkpoint's buffers describe its activations; random ones
        # make activations explode by 1e5 in eval mode): one oracle train pass with momentum 1
        cal = O.OracleConfig(**{**kw, 'bn_momentum': 1.0})
        _, stats = O.forward(sd, cal, rgb, depth, True)
        sd.update({k: v for k, v in stats.items() if 'num_batches' not in k})
        for k, v in stats.items():
            if 'num_batches' not in k:
                eng.P[k].copy_(v)
        ref = O.flatten_outputs(O.forward(sd, ocfg, rgb, depth, False)[0])
        bud = O.flatten_outputs(O.forward(_bf16_round(sd), ocfg, _r(rgb), _r(depth), False)[0])
        res = eng.forward(rgb.cuda() if rgb is not None else None, depth.cuda() if depth is not None else None, False)
    got = _flat_engine(res)
    assert len(got) == len(ref)
    report, fails = {}, {}
    for i, (g, r, b) in enumerate(zip(got, ref, bud)):
        assert tuple(g.shape) == tuple(r.shape)
        e, lim = rel_l2(g, r), max(OUT_FLOOR, FACTOR * rel_l2(b, r))
        report[f'out{i}'] = (e, lim)
        if e > lim:
            fails[f'out{i}'] = (e, lim)
    _dump(f'eval_{name}', report)
    assert not fails, fails
    if 'semantic' in ocfg.tasks:   # arg-max exactness where the oracle's decision margin is above the error bound
        r, g = ref[0], got[0].cpu()
        top2 = r.topk(2, dim=1).values
        margin = top2[:, 0] - top2[:, 1]
        bound = 2 * (g - r).abs().max().item()
        sure = margin > bound
        agree = (g.argmax(1) == r.argmax(1))
        assert bool(agree[sure].all()), f'argmax flips on confident pixels: {(~agree[sure]).sum().item()}'
        assert agree.float().mean() > 0.9, agree.float().mean()


@pytest.mark.parametrize('name', list(CASES))
def test_train_forward_backward_matches_oracle(name):
    kw, n, h, w = CASES[name]
    O, ocfg, sd, rgb, depth, eng = _setup(kw, n, h, w)
    res = eng.forward(rgb.cuda() if rgb is not None else None, depth.cuda() if depth is not None else None, True)
    got = _flat_engine(res)
    # fixed random cotangents (the same on both sides), scaled like d(mean)/do
    gg = torch.Generator().manual_seed(7)
    cot = [torch.randn(o.shape, generator=gg) / o.numel() ** 0.5 for o in got]
    ref_out, ref_grads, ref_stats = O.forward_backward(sd, ocfg, rgb, depth, grad_outputs=cot)
    bud_out, bud_grads, bud_stats = O.forward_backward(_bf16_round(sd), ocfg, _r(rgb), _r(depth), grad_outputs=cot)
    ref, bud = O.flatten_outputs(ref_out), O.flatten_outputs(bud_out)
    report, fails = {}, {}

    def check(key, g, r, b, floor):
        budget = rel_l2(b, r)
        # tensors whose fp32 oracle value moves by > 25 % under bf16 rounding of inputs/weights are noise
        # dominated (near-zero true gradients): hold them to a multiple of that noise only
        e, lim = rel_l2(g, r), max(floor, (2 * FACTOR if budget > 0.25 else FACTOR) * budget)
        report[key] = (e, lim)
        if e > lim:
            fails[key] = (e, lim)
    for i, (g, r, b) in enumerate(zip(got, ref, bud)):
        check(f'out{i}', g, r, b, OUT_FLOOR)
    it = iter(cot)
    gouts = {}
    for t in ('semantic', 'instance', 'scene'):
        pass
    # map the flat cotangents back to the engine's per-task output lists
    flat_keys = []
    if 'semantic' in res and 'instance' in res:
        ns, ni = len(res['semantic']), len(res['instance'])
        flat_keys = [('semantic', 0)] + [('instance', j) for j in range(3)] + [('semantic', j) for j in range(1, ns)] \
            + [('instance', j) for j in range(3, ni)]
    else:
        for t in ('semantic', 'instance'):
            if t in res:
                flat_keys += [(t, j) for j in range(len(res[t]))]
    if 'scene' in res:
        flat_keys += [('scene', 0)]
    gouts = {t: [None] * len(outs) for t, outs in res.items()}
    for (t, j), c in zip(flat_keys, cot):
        gouts[t][j] = c.cuda()
    grads = eng.backward(gouts)
    torch.cuda.synchronize()
    assert set(grads.keys()) == set(ref_grads.keys())
    for k, rg in ref_grads.items():
        check('grad:' + k, grads[k], rg, bud_grads[k], GRAD_FLOOR)
    for k, v in ref_stats.items():
        if 'num_batches' in k:
            assert int(eng.P[k].item()) == int(v.item())
        else:
            check('stat:' + k, eng.P[k], v, bud_stats[k], STAT_FLOOR)
    _dump(f'train_{name}', report)
    assert not fails, dict(sorted(fails.items(), key=lambda kv: -kv[1][0] / kv[1][1])[:25])


def test_dropout_masks_are_applied():
    """same masks on both sides -> parity holds with Dropout2d active"""
    kw, n, h, w = CASES['full_rgbd_r18_ragged']
    O, ocfg, sd, rgb, depth, eng = _setup(kw, n, h, w, dropout=True)
    masks = eng.make_dropout_masks(n)
    assert len(masks) == len(O.dropout_sites(ocfg))
    ref_out, _ = O.forward(sd, ocfg, rgb, depth, True, dropout_masks={k: v.cpu() for k, v in masks.items()})
    res = eng.forward(rgb.cuda(), depth.cuda(), True, dropout_masks=masks)
    errs = [rel_l2(g, r) for g, r in zip(_flat_engine(res), O.flatten_outputs(ref_out))]
    assert max(errs) < 0.1, errs
    vals = torch.cat([m.flatten() for m in masks.values()]).unique().cpu().tolist()
    assert len(vals) <= 3 and 0.0 in vals


def _dump(name, report):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f'parity_{name}.json'), 'w') as f:
        json.dump(dict(sorted(report.items(), key=lambda kv: -kv[1][0] / kv[1][1])), f, indent=1)
