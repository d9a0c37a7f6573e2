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


# ----------------------------------------------------------------------------------------------------------------
# Regression losses of the instance / orientation task (MT/loss/mse.py:23-41, l1.py:23-41, vonmises.py:29-51) as
# MT/task_helper/instance.py:118-207 composes them with its masks; float64.
#   kind 0: sum_p 1/C sum_c (x_pc * m_p - t_pc)^2     kind 1: ... |x_pc * m_p - t_pc|
#   kind 2: sum_{p: m_p} 1 - exp(kappa * (sum_c x_pc t_pc - 1))
# -> (loss, count = sum_p m_p, dloss/dpred as float64 in pred's shape).  channel_dim: axis of c (None = no channel axis).
def make_instance_inputs(n: int, h: int, w: int, seed: int, fg_fraction: float = 0.4):
    g = torch.Generator().manual_seed(seed)
    center = torch.rand(n, 1, h, w, generator=g)
    offset = torch.tanh(torch.randn(n, 2, h, w, generator=g))
    o = torch.randn(n, 2, h, w, generator=g)
    orient = o / (o.norm(dim=1, keepdim=True) + 1e-7)
    fg = torch.rand(n, h, w, generator=g) < fg_fraction
    center_mask = torch.rand(n, h, w, generator=g) < 0.9
    t_center = torch.rand(n, h, w, generator=g) * fg
    t_offset = torch.tanh(torch.randn(n, 2, h, w, generator=g)) * fg[:, None]
    a = torch.rand(n, h, w, generator=g) * 6.2831853
    t_orient = torch.stack([torch.cos(a), torch.sin(a)], 1)
    ofg = fg & (torch.rand(n, h, w, generator=g) < 0.5)
    return dict(center=center, offset=offset, orientation=orient, center_mask=center_mask, fg=fg, ofg=ofg,
                t_center=t_center, t_offset=t_offset, t_orientation=t_orient)


def masked_loss(kind: int, pred: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor],
                channel_dim: Optional[int], kappa: float = 1.0) -> Tuple[float, int, np.ndarray]:
    x = pred.detach().double().numpy()
    t = target.detach().double().numpy()
    if channel_dim is None:
        x, t = x[None], t[None]
        cd = 0
    else:
        cd = channel_dim % x.ndim
    c = x.shape[cd]
    pix_shape = tuple(s for i, s in enumerate(x.shape) if i != cd)
    m = np.ones(pix_shape) if mask is None else (mask.detach().numpy() != 0).astype(np.float64).reshape(pix_shape)
    me = np.expand_dims(m, cd)
    if kind == 2:
        dot = (x * t).sum(axis=cd, keepdims=True)
        e = np.exp(kappa * (dot - 1.0))
        loss = float(((1.0 - e) * me).sum())
        grad = -kappa * e * t * me
    else:
        d = x * me - t
        loss = float(((d * d if kind == 0 else np.abs(d)).sum(axis=cd) / c).sum())
        grad = (2.0 * d if kind == 0 else np.sign(d)) * me / c
    if channel_dim is None:
        grad = grad[0]
    return loss, int(m.sum()), grad
