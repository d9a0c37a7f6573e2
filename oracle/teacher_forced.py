"""Teacher-forced whole-network parity of the CUDA engine against the CPU oracle.  TEST INFRASTRUCTURE ONLY
(imported by tests/ and __graft_entry__.smoke(); nothing in emsanet_b200/ imports it).

Why.  A free-running comparison of two bf16-storage implementations of this network cannot be tight: one bf16
rounding that falls differently (fp32 summation order) perturbs everything downstream, the train-mode BatchNorms
amplify it, ~0.5 % of the ReLU masks of every later layer flip, and each flipped mask element is a 100 % change of a
gradient element.  `scripts/parity_noise_floor.py` measures it on the oracle alone: the bf16-storage oracle with
fp32 accumulation and the SAME oracle with fp64 accumulation (identical rounding points) end up 20 % apart in the
parameter gradients (median rel-L2; outputs 1-8 %).  Any bound wide enough for that passes an all-zero gradient.

How.  The engine runs once with every stored activation exposed (Engine.taps, 285 tensors for resnet18, 545 for
resnet34).  The oracle then recomputes the network with `teacher=` those tensors (oracle/emsanet_oracle.py `_Ctx.q`):
each layer's value is computed from the ENGINE's stored inputs, compared, and replaced by the engine's stored output
(straight-through for autograd).  Forward: every layer is checked on identical inputs, at the real shapes, through the
real launch configuration (pair launches, CTA pairs, strided parity views).  Backward: autograd runs on exactly the
activations, batch statistics and ReLU masks the engine used, so the backward pass is the same LINEAR map on both
sides and all parameter gradients must agree to rounding (bf16 storage of the activation gradients, fp32 summation
order) — a wrong, missing or mis-wired data/weight gradient anywhere in the network shows up at full size.
The rounding level is measured, per gradient tensor, on the oracle itself: the same autograd graph is differentiated
twice, exactly and with the gradient at every storage point rounded to bf16 (`emsanet_oracle.ROUND_GRADS`); the
engine's gradient may deviate from the exact one by a small multiple of that yardstick (tests/test_engine_gpu.py).

The cotangents are those of the coherent bench loss sum_i mean(o_i^2) plus a fixed random probe
sum_i mean(o_i * r_i): the probe keeps the gradient of the unit-length orientation outputs (whose mean(o^2) is
constant) away from 0/0.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import torch

from . import emsanet_oracle as O


def rel_l2(a, b) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b) -> float:
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def flat_engine_outputs(res: Dict[str, List[torch.Tensor]], nt: int = 3) -> List[torch.Tensor]:
    """engine result {'semantic': [...], 'instance': [...], 'scene': [...]} -> the oracle's depth-first flat order"""
    flat: List[torch.Tensor] = []
    if 'semantic' in res and 'instance' in res:
        s, i = res['semantic'], res['instance']
        flat += [s[0]] + list(i[:nt]) + list(s[1:]) + list(i[nt:])
    else:
        for t in ('semantic', 'instance'):
            if t in res:
                flat += list(res[t])
    if 'scene' in res:
        flat += list(res['scene'])
    return flat


def flat_output_keys(res: Dict[str, List[torch.Tensor]], nt: int = 3):
    keys = []
    if 'semantic' in res and 'instance' in res:
        ns, ni = len(res['semantic']), len(res['instance'])
        keys = [('semantic', 0)] + [('instance', j) for j in range(nt)] + [('semantic', j) for j in range(1, ns)] \
            + [('instance', j) for j in range(nt, ni)]
    else:
        for t in ('semantic', 'instance'):
            if t in res:
                keys += [(t, j) for j in range(len(res[t]))]
    if 'scene' in res:
        keys += [('scene', 0)]
    return keys


def cotangents(outputs: List[torch.Tensor], seed: int = 7) -> List[torch.Tensor]:
    """d/do of  sum_i mean(o_i^2) + sum_i mean(o_i * r_i),  r_i ~ N(0,1) seeded: 2 o / numel + r / numel"""
    g = torch.Generator().manual_seed(seed)
    cot = []
    for o in outputs:
        o = o.detach().float().cpu()
        cot.append((2.0 * o + torch.randn(o.shape, generator=g)) / o.numel())
    return cot


def nchw(t: torch.Tensor) -> torch.Tensor:
    """engine activation (NHWC bf16, maybe a channel slice, maybe [N, C]) -> fp32 NCHW on the CPU"""
    t = t.detach().float().cpu()
    return t.permute(0, 3, 1, 2).contiguous() if t.dim() == 4 else t


def run(eng, sd: Dict[str, torch.Tensor], ocfg: O.OracleConfig, rgb: Optional[torch.Tensor],
        depth: Optional[torch.Tensor], dropout_masks: Optional[Dict[str, torch.Tensor]] = None,
        training: bool = True, mutate=None) -> Dict:
    """One teacher-forced comparison.  `eng`: emsanet_b200.engine.Engine holding `sd` on the GPU.  Returns a report:
      forward[name]  rel-L2 of the oracle's value of a stored activation (from the engine's stored inputs) vs the engine's
      outputs[i]     rel-L2 of the fp32 NCHW network outputs
      grads[key]     (rel-L2, cosine, oracle norm, yardstick) of every parameter gradient, yardstick = rel-L2 between
                     the oracle's gradient with bf16-rounded activation gradients and its exact one   (training only)
      stats[key]     rel-L2 of the updated running statistics                      (training only)
    `mutate(eng)` may sabotage the engine between its forward and backward pass (mutation tests)."""
    dev = eng.dev
    eng.taps = {}
    try:
        with torch.no_grad():
            res = eng.forward(rgb.to(dev) if rgb is not None else None, depth.to(dev) if depth is not None else None,
                              training, dropout_masks=dropout_masks)
        taps = dict(eng.taps)
    finally:
        eng.taps = None
    nt = 3 if ocfg.with_orientation else 2
    got = flat_engine_outputs(res, nt)
    teacher = {k: nchw(v) for k, v in taps.items() if v.dtype == torch.bfloat16}
    computed: Dict[str, torch.Tensor] = {}
    emu = dataclasses.replace(ocfg, emulate_bf16_storage=True)
    masks_cpu = {k: v.cpu() for k, v in dropout_masks.items()} if dropout_masks else None
    report: Dict = {'n_storage_points': len(teacher)}
    if not training:
        with torch.no_grad():
            out, _ = O.forward(sd, emu, rgb, depth, False, teacher=teacher, computed=computed)
        ref = O.flatten_outputs(out)
        grads_ref = stats_ref = None
    else:
        cot = cotangents(got)
        leaves = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v)
                  for k, v in sd.items()}
        out, stats_ref = O.forward(leaves, emu, rgb, depth, True, masks_cpu, teacher=teacher, computed=computed)
        ref = O.flatten_outputs(out)
        keys = [k for k, v in leaves.items() if v.is_floating_point() and v.requires_grad]
        # the same graph differentiated twice: exact fp32 gradients, and with every stored activation gradient rounded
        # to bf16 (the yardstick: what bf16 gradient storage alone does to each parameter gradient)
        g_exact = torch.autograd.grad(ref, [leaves[k] for k in keys], cot, retain_graph=True)
        O.ROUND_GRADS[0] = True
        try:
            g_round = torch.autograd.grad(ref, [leaves[k] for k in keys], cot)
        finally:
            O.ROUND_GRADS[0] = False
        grads_ref = dict(zip(keys, g_exact))
        grads_rounded = dict(zip(keys, g_round))
    missing = sorted(set(computed) - set(teacher))
    assert not missing, f'storage points of the oracle that the engine does not expose: {missing[:5]}'
    report['forward'] = {k: rel_l2(teacher[k][:, :computed[k].shape[1]], computed[k]) for k in computed}
    report['outputs'] = [rel_l2(g, r) for g, r in zip(got, ref)]
    report['engine_outputs'], report['oracle_outputs'] = [g.detach().cpu() for g in got], [r.detach() for r in ref]
    if training:
        if mutate is not None:
            mutate(eng)
        gouts = {t: [None] * len(outs) for t, outs in res.items()}
        for (t, j), c in zip(flat_output_keys(res, nt), cot):
            gouts[t][j] = c.to(dev)
        grads = eng.backward(gouts)
        torch.cuda.synchronize()
        assert set(grads) == set(grads_ref)
        report['grads'] = {k: (rel_l2(grads[k], grads_ref[k]), cosine(grads[k], grads_ref[k]),
                               float(grads_ref[k].double().norm()), rel_l2(grads_rounded[k], grads_ref[k]))
                           for k in grads_ref}
        report['stats'] = {}
        for k, v in stats_ref.items():
            if 'num_batches' in k:
                assert int(eng.P[k].item()) == int(v.item()), k
            else:
                report['stats'][k] = rel_l2(eng.P[k], v)
    return report


def summarize(report: Dict) -> Dict:
    """the numbers the tests assert on and dump under gpurun_out/ -> profiles/"""
    fw = sorted(report['forward'].items(), key=lambda kv: -kv[1])
    out = {'n_storage_points': report['n_storage_points'], 'forward_max': fw[0][1] if fw else 0.0,
           'forward_worst': fw[:5], 'outputs_max': max(report['outputs']), 'outputs': report['outputs']}
    if 'grads' in report:
        g = report['grads']
        rels = sorted(v[0] for v in g.values())
        out.update(
            n_grads=len(g), grad_rel_median=rels[len(rels) // 2], grad_rel_p99=rels[int(0.99 * (len(rels) - 1))],
            grad_rel_max=rels[-1], grad_cos_min=min(v[1] for v in g.values()),
            grad_worst=sorted(((k, v[0], v[1], v[3]) for k, v in g.items()), key=lambda x: -x[1])[:8],
            grad_yardstick_median=sorted(v[3] for v in g.values())[len(g) // 2],
            grad_over_yardstick_max=max(v[0] / max(v[3], 3e-3) for v in g.values()),
            stats_max=max(report['stats'].values()) if report['stats'] else 0.0)
    return out
