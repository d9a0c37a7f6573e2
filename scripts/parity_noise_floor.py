"""How far apart are two CORRECT implementations with identical bf16 rounding points?  (CPU, oracle only.)

The bf16-storage oracle is run twice — fp32 accumulation and fp64 accumulation — on the same seeded network with the
coherent bench loss.  Their distance is the floor any engine-vs-oracle gradient comparison can resolve; the vehicles
compared here decide which network the whole-network parity tests run on (tests/test_engine_gpu.py)."""
import dataclasses
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import emsanet_oracle as O   # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def med(v):
    v = sorted(v)
    return v[len(v) // 2]


def vehicle(name, cfg, seed=0):
    sd = O.make_state_dict(cfg, seed=seed)
    if name == 'random_bn_015':
        for k in sd:
            if k.endswith('norm2.weight'):
                sd[k] = sd[k] * 0.15
    elif name == 'trained_like':
        sd = O.make_trained_like_state_dict(cfg, seed=seed)
    return sd


def main():
    n, h, w = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    backbone = sys.argv[4] if len(sys.argv) > 4 else 'resnet18'
    cfg = O.OracleConfig(backbone=backbone)
    emu = dataclasses.replace(cfg, emulate_bf16_storage=True)
    rgb, depth = O.make_inputs(n, h, w, seed=1)
    for name in sys.argv[5:] or ['random_bn_015', 'trained_like']:
        sd = vehicle(name, cfg)
        o32, g32, _ = O.forward_backward(sd, cfg, rgb, depth)
        oe, ge, _ = O.forward_backward(sd, emu, rgb, depth)
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        od, gd, _ = O.forward_backward(sd64, emu, rgb.double(), depth.double())
        f32, fe, fd = (O.flatten_outputs(x) for x in (o32, oe, od))
        keys = [k for k in g32 if float(g32[k].norm()) > 0]
        rep = {
            'vehicle': name, 'shape': [n, h, w], 'backbone': backbone,
            'out_emu_vs_fp32': [rel(a, b) for a, b in zip(fe, f32)],
            'out_emu64_vs_emu32': [rel(a, b) for a, b in zip(fd, fe)],
            'grad_emu_vs_fp32_median': med(rel(ge[k], g32[k]) for k in keys),
            'grad_emu_vs_fp32_max': max(rel(ge[k], g32[k]) for k in keys),
            'grad_emu64_vs_emu32_median': med(rel(gd[k], ge[k]) for k in keys),
            'grad_emu64_vs_emu32_p90': sorted(rel(gd[k], ge[k]) for k in keys)[int(0.9 * len(keys))],
            'grad_emu64_vs_emu32_max': max(rel(gd[k], ge[k]) for k in keys),
            'cos_emu64_vs_emu32_min': min(cos(gd[k], ge[k]) for k in keys),
            'cos_emu64_vs_emu32_median': med(cos(gd[k], ge[k]) for k in keys),
            'cos_emu_vs_fp32_min': min(cos(ge[k], g32[k]) for k in keys),
            'cos_emu_vs_fp32_median': med(cos(ge[k], g32[k]) for k in keys),
        }
        print(json.dumps(rep))


if __name__ == '__main__':
    main()
