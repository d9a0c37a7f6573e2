"""CPU oracle of the semantic cross-entropy loss that consumes the network's largest output
(SURVEY.md §8(f) row 2).  TEST INFRASTRUCTURE ONLY: imported by tests/, never by emsanet_b200/.

Restates  MT/loss/ce.py:13-68 (CrossEntropyLossSemantic, weighted_reduction=False) =
torch.nn.CrossEntropyLoss(weight=w, reduction='sum', ignore_index=-1, label_smoothing=eps) on
`target.long() - 1` (MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/), per pixel:

    logp   = x - max(x) - log(sum(exp(x - max(x))))
    loss_i = (1 - eps) * w[t] * (-logp[t]) + eps / C * sum_c w[c] * (-logp[c])        (t = target - 1 >= 0)
    loss   = sum over the non-void pixels;   n_elements = number of non-void pixels
    dloss/dx[k] = -a[k] + (sum_c a[c]) * softmax(x)[k],   a[c] = (1 - eps) * w[t] * [c == t] + eps / C * w[c]

Accumulated in float64 (the reference sums in fp32 in ATen's order; it agrees to ~1e-6 relative).
Pinned: oracle/make_golden_loss.py runs the UNMODIFIED reference class (loss and autograd gradient) on seeded inputs;
fixtures in tests/golden/loss/, re-checked by tests/test_loss.py.
"""
from typing import Optional, Tuple

import numpy as np
import torch


def make_inputs(n: int, c: int, h: int, w: int, seed: int, void_fraction: float = 0.2, dtype=torch.uint8):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, c, h, w, generator=g) * 3.0
    target = torch.randint(1, c + 1, (n, h, w), generator=g)
    target[torch.rand(n, h, w, generator=g) < void_fraction] = 0          # 0 = void (not predicted, ce.py:46)
    weights = 0.5 + torch.rand(c, generator=g) * 2.0
    return logits, target.to(dtype), weights


def cross_entropy_semantic(logits: torch.Tensor, target: torch.Tensor, weights: Optional[torch.Tensor],
                           label_smoothing: float = 0.0) -> Tuple[float, int, np.ndarray]:
    """-> (loss, n_elements, dloss/dlogits as float64 [N,C,H,W])"""
    x = logits.detach().double().numpy()
    n, c, h, w = x.shape
    t = target.long().numpy() - 1
    valid = t >= 0
    wv = np.ones(c) if weights is None else weights.double().numpy()
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    s = e.sum(axis=1, keepdims=True)
    logp = x - m - np.log(s)
    p = e / s
    tc = np.where(valid, t, 0)
    onehot = np.zeros_like(x)
    np.put_along_axis(onehot, tc[:, None], 1.0, axis=1)
    a = (1.0 - label_smoothing) * wv[tc][:, None] * onehot + (label_smoothing / c) * wv[None, :, None, None]
    a = a * valid[:, None]
    loss = float(-(a * logp).sum())
    grad = -a + a.sum(axis=1, keepdims=True) * p
    return loss, int(valid.sum()), grad
