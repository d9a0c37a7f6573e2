"""Fused semantic cross-entropy beside torch's own at the bench batch.

    python scripts/bench_loss.py [--batch 32] [--iters 10]

Times forward + backward of `CrossEntropyLossSemanticB200` and of `torch.nn.CrossEntropyLoss(weight, reduction='sum',
ignore_index=-1, label_smoothing)` (what MT/loss/ce.py:31-36 builds) on N x 40 x 480 x 640 fp32 logits with CUDA
events, checks that the two agree, and reports the fused path against the HBM roofline (algorithmic bytes:
forward 4C + 1 per pixel, backward 8C + 1).  The kernels had not run on a B200 when this script was written.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emsanet_b200 import losses   # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--smoothing', type=float, default=0.1)
    a = ap.parse_args()
    n, c, h, w = a.batch, 40, 480, 640
    g = torch.Generator(device='cuda').manual_seed(0)
    logits = (torch.randn(n, c, h, w, device='cuda', generator=g) * 3).requires_grad_(True)
    target = torch.randint(0, c + 1, (n, h, w), device='cuda', generator=g).to(torch.uint8)
    weights = 0.5 + 2 * torch.rand(c, device='cuda', generator=g)
    fused = losses.CrossEntropyLossSemanticB200(weights=weights, label_smoothing=a.smoothing)
    ref = torch.nn.CrossEntropyLoss(weight=weights, reduction='sum', ignore_index=-1, label_smoothing=a.smoothing)

    def run_fused():
        logits.grad = None
        (loss, _n), = fused([logits], [target])
        (loss * 0.5).backward()
        return loss

    def run_ref():
        logits.grad = None
        loss = ref(logits, target.long() - 1)
        _n = torch.sum(target > 0).item()            # the reference synchronises for the count (ce.py:50)
        (loss * 0.5).backward()
        return loss

    lf = run_fused()
    gf = logits.grad.clone()
    lr = run_ref()
    gr = logits.grad.clone()
    res = {'batch': n, 'shape': [c, h, w], 'label_smoothing': a.smoothing,
           'loss_rel_dev': abs(lf.item() - lr.item()) / abs(lr.item()), 'grad_max_dev': (gf - gr).abs().max().item()}
    ms_f, ms_r = timed(run_fused, a.iters), timed(run_ref, a.iters)
    px = n * h * w
    bytes_ = px * (4 * c + 1) + px * (8 * c + 1)
    peak, src = 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        src = 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        pass
    res.update({'fused_fwd_bwd_ms': ms_f, 'torch_fwd_bwd_ms': ms_r, 'algorithmic_GB': bytes_ / 1e9,
                'achieved_GBps': bytes_ / ms_f / 1e6, 'peak_GBps': peak, 'peak_source': src,
                'frac': bytes_ / ms_f / 1e6 / peak})
    print(json.dumps(res))


if __name__ == '__main__':
    main()
