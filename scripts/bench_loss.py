"""Fused semantic cross-entropy beside torch's own at the bench batch.

    python scripts/bench_loss.py [--batch 32] [--iters 10]

Times forward + backward of `CrossEntropyLossSemanticB200` and of `torch.nn.CrossEntropyLoss(weight, reduction='sum',
ignore_index=-1, label_smoothing)` (what MT/loss/ce.py:31-36 builds) on N x 40 x 480 x 640 fp32 logits with CUDA
events, checks that the two agree, and reports the fused path against the HBM roofline (algorithmic bytes:
forward 4C + 1 per pixel, backward 8C + 1).  Second line: the instance / orientation losses of one scale (masked MSE,
masked L1, von Mises) fused and sync-free (`losses.instance_losses`) beside the reference's composition of torch ops
with its three `.item()` synchronisations (MT/task_helper/instance.py:118-207).
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
    del logits, gf, gr
    torch.cuda.empty_cache()
    print(json.dumps(instance_line(n, h, w, a.iters, peak, src)))


def instance_line(n, h, w, iters, peak, src):
    """masked MSE (centers) + masked L1 (offsets) + von Mises (orientations) at the main scale, forward + backward"""
    g = torch.Generator(device='cuda').manual_seed(1)
    center = torch.rand(n, 1, h, w, device='cuda', generator=g).requires_grad_(True)
    offset = torch.tanh(torch.randn(n, 2, h, w, device='cuda', generator=g)).requires_grad_(True)
    o = torch.randn(n, 2, h, w, device='cuda', generator=g)
    orient = (o / o.norm(dim=1, keepdim=True)).requires_grad_(True)
    fg = torch.rand(n, h, w, device='cuda', generator=g) < 0.4
    cmask = torch.rand(n, h, w, device='cuda', generator=g) < 0.9
    t_center = torch.rand(n, h, w, device='cuda', generator=g) * fg
    t_offset = torch.tanh(torch.randn(n, 2, h, w, device='cuda', generator=g)) * fg[:, None]
    ang = torch.rand(n, h, w, device='cuda', generator=g) * 6.2831853
    t_orient = torch.stack([torch.cos(ang), torch.sin(ang)], 1)
    ofg = fg & (torch.rand(n, h, w, device='cuda', generator=g) < 0.5)
    leaves = (center, offset, orient)

    def fused():
        for t in leaves:
            t.grad = None
        s0, c0 = losses.fused_masked_loss(losses.MSE, center[:, 0], t_center, cmask, None)
        s1, c1 = losses.fused_masked_loss(losses.L1, offset, t_offset, fg, 1)
        s2, c2 = losses.fused_masked_loss(losses.VONMISES, orient, t_orient, ofg, 1)
        total = s0 / c0.float().reshape(()) + s1 / c1.float().reshape(()) + s2 / c2.clamp(min=1).float().reshape(())
        total.backward()
        return total

    def reference():                                     # instance.py:118-207, op for op
        for t in leaves:
            t.grad = None
        l0 = torch.nn.functional.mse_loss(center[:, 0] * cmask, t_center, reduction='none').sum()
        n0 = cmask.sum().cpu().detach().item()
        m = fg.unsqueeze(1).expand_as(offset)
        l1 = torch.nn.functional.l1_loss(offset * m, t_offset, reduction='none').mean(dim=1).sum()
        n1 = fg.sum().cpu().detach().item()
        p = orient.contiguous().permute((0, 2, 3, 1)).reshape(-1, 2)
        t = t_orient.permute((0, 2, 3, 1)).reshape(-1, 2)
        mm = ofg.flatten()
        n2 = max(mm.sum().cpu().detach().item(), 1)
        cos = (p[mm, :] * t[mm, :]).sum(dim=1, keepdim=True)
        l2 = (1 - torch.exp(1.0 * (cos - 1))).sum()
        total = l0 / n0 + l1 / n1 + l2 / n2
        total.backward()
        return total

    lf, lr = float(fused()), float(reference())
    gdev = max(float((a.grad - b).abs().max()) for a, b in zip(leaves, [t.grad.clone() for t in leaves]))
    ms_f, ms_r = timed(fused, iters), timed(reference, iters)
    px = n * h * w
    # fwd: pred + target + mask per loss; bwd: the same + the gradient written
    bytes_ = px * ((4 + 4 + 1) + (8 + 8 + 1) + (8 + 8 + 1)) * 2 + px * (4 + 8 + 8)
    return {'what': 'instance / orientation losses, main scale', 'batch': n, 'loss_rel_dev': abs(lf - lr) / abs(lr),
            'fused_fwd_bwd_ms': ms_f, 'torch_fwd_bwd_ms': ms_r, 'host_syncs_fused': 0, 'host_syncs_reference': 3,
            'algorithmic_GB': bytes_ / 1e9, 'achieved_GBps': bytes_ / ms_f / 1e6, 'peak_GBps': peak, 'peak_source': src,
            'frac': bytes_ / ms_f / 1e6 / peak}


if __name__ == '__main__':
    main()
