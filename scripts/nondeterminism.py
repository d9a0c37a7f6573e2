"""Run-to-run spread of the engine (same weights, same inputs, same process): fp32 atomics (BatchNorm statistics, split-K
weight gradients, bias / squeeze-excite sums) fix no summation order, a sum that lands on the other side of a bf16
rounding boundary changes one stored activation by one ulp, and the train-mode network amplifies that chaotically
(oracle/teacher_forced.py).  Prints what that amounts to: rel-L2 between two runs for outputs and parameter gradients."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import emsanet_oracle as O
from emsanet_b200.engine import Engine, EngineConfig


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(backbone, n, h, w):
    cfg = O.OracleConfig(backbone=backbone)
    sd = O.make_state_dict(cfg, seed=0)
    for k in sd:
        if k.endswith('norm2.weight'):
            sd[k] = sd[k] * 0.15
    eng = Engine(EngineConfig(backbone=backbone, dropout_p_encoder=0.0, dropout_p_decoder=0.0),
                 {k: v.cuda() for k, v in sd.items()})
    rgb, depth = (t.cuda() for t in O.make_inputs(n, h, w, seed=1))
    runs = []
    for _ in range(3):
        res = eng.forward(rgb, depth, True, track_running_stats=False)
        outs = [o.clone() for t in res.values() for o in t]
        gouts = {t: [o * (2.0 / o.numel()) for o in v] for t, v in res.items()}
        grads = {k: g.clone() for k, g in eng.backward(gouts).items()}
        runs.append((outs, grads))
    (o0, g0) = runs[0]
    rep = {'backbone': backbone, 'shape': [n, h, w]}
    for i, (o, g) in enumerate(runs[1:], 1):
        gr = sorted(rel(g[k], g0[k]) for k in g0 if float(g0[k].norm()) > 0)
        rep[f'run{i}_vs_run0'] = {'outputs_max': max(rel(a, b) for a, b in zip(o, o0)),
                                   'outputs_bit_identical': all(torch.equal(a, b) for a, b in zip(o, o0)),
                                   'grad_rel_median': gr[len(gr) // 2], 'grad_rel_p90': gr[int(0.9 * len(gr))],
                                   'grad_rel_max': gr[-1], 'grads_bit_identical': sum(int(torch.equal(g[k], g0[k])) for k in g0)}
    return rep


if __name__ == '__main__':
    print(json.dumps([run('resnet18', 4, 64, 96), run('resnet34', 8, 192, 256), run('resnet34', 4, 480, 640)], indent=1))
