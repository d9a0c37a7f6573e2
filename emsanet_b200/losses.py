"""Fused semantic cross-entropy behind the reference's loss interface.

`CrossEntropyLossSemanticB200` mirrors `CrossEntropyLossSemantic` (MT/loss/ce.py:13-68, MT/loss/base.py:11-33;
MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/): same constructor, same
`loss(input_tensors, target_tensors) -> ((loss, n_elements), ...)` contract, `loss` an autograd-connected fp32 scalar.
Forward = ONE pass over the logits (`eb200_ce_loss_fwd`: loss and the non-void count accumulated on the device),
backward = ONE pass that writes the gradient already scaled by the upstream gradient (`eb200_ce_loss_bwd`), instead
of log_softmax / nll_loss / smoothing / their autograd backward as separate full-tensor passes.

`install(task_helper)` swaps a reference SemanticTaskHelper's `_loss` (MT/task_helper/semantic.py:40-49).

STATUS: the kernels have NOT run on a B200 yet (round 1 ended without GPU time).  The oracle is pinned against the
reference (oracle/loss_oracle.py), the host side below is tested on CPU with the two C-ABI calls replaced by the
oracle; the GPU parity tests exist but are skipped unless EB200_RUN_UNVERIFIED=1 (tests/test_loss.py).
No CPU fallback: non-CUDA inputs raise.
"""
import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.EB200Error(f'{what} must be a CUDA tensor: emsanet_b200 losses have no CPU path')
    return t.contiguous()


def ce_forward(logits: torch.Tensor, target: torch.Tensor, weights: Optional[torch.Tensor], eps: float
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """eb200_ce_loss_fwd -> (loss fp64 [1], count int64 [1]) on the device"""
    n, c, h, w = logits.shape
    loss = torch.empty(1, dtype=torch.float64, device=logits.device)
    count = torch.empty(1, dtype=torch.int64, device=logits.device)
    _lib.call('eb200_ce_loss_fwd', _p(logits), _p(target), target.element_size(), _p(weights), float(eps), n, c, h, w,
              _p(loss), _p(count), _stream())
    return loss, count


def ce_backward(logits: torch.Tensor, target: torch.Tensor, weights: Optional[torch.Tensor], eps: float,
                grad_out: torch.Tensor) -> torch.Tensor:
    """eb200_ce_loss_bwd -> dlogits fp32 [N,C,H,W]"""
    n, c, h, w = logits.shape
    d = torch.empty_like(logits)
    _lib.call('eb200_ce_loss_bwd', _p(logits), _p(target), target.element_size(), _p(weights), float(eps), _p(grad_out),
              n, c, h, w, _p(d), _stream())
    return d


class _FusedCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weights, eps):
        loss64, count = ce_forward(logits, target, weights, eps)
        ctx.save_for_backward(logits, target, weights if weights is not None else logits.new_empty(0))
        ctx.eps = eps
        ctx.has_weights = weights is not None
        ctx.mark_non_differentiable(count)
        return loss64.to(torch.float32).reshape(()), count

    @staticmethod
    def backward(ctx, grad_loss, _grad_count):
        logits, target, weights = ctx.saved_tensors
        g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        d = ce_backward(logits, target, weights if ctx.has_weights else None, ctx.eps, g)
        return d, None, None, None


class CrossEntropyLossSemanticB200(torch.nn.Module):
    def __init__(self, weights: Optional[torch.Tensor] = None, label_smoothing: float = 0.0,
                 weighted_reduction: bool = False) -> None:
        super().__init__()
        if weighted_reduction:
            raise NotImplementedError('weighted_reduction=True (the ESANet reduction, ce.py:24-29,55-68) is not covered')
        self._weights = weights
        self._label_smoothing = float(label_smoothing)
        self._weighted_reduction = False

    def _compute_loss(self, input_: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, int]:
        x = _check(input_, 'logits')
        if x.dtype != torch.float32:
            x = x.float()
        t = _check(target, 'target')
        if t.dtype not in (torch.uint8, torch.int32, torch.int64):
            t = t.long()
        w = None
        if self._weights is not None:
            w = self._weights.to(device=x.device, dtype=torch.float32).contiguous()
            if w.numel() != x.shape[1]:
                raise ValueError(f'{w.numel()} class weights for {x.shape[1]} classes')
        loss, count = _FusedCrossEntropy.apply(x, t, w, self._label_smoothing)
        return loss, int(count.item())          # the reference synchronises here as well (ce.py:50)

    def forward(self, input_tensors: Sequence[torch.Tensor], target_tensors: Sequence[torch.Tensor]):
        return tuple(self._compute_loss(i, t) for i, t in zip(input_tensors, target_tensors))   # base.py:23-33


def install(task_helper):
    """swap the `_loss` of a reference SemanticTaskHelper (after `initialize(device)`) for the fused mirror"""
    old = task_helper._loss
    if type(old).__name__ != 'CrossEntropyLossSemantic':
        raise NotImplementedError(f'no fused mirror for {type(old).__name__}')
    task_helper._loss = CrossEntropyLossSemanticB200(weights=old._weights, label_smoothing=old._loss.label_smoothing,
                                                    weighted_reduction=old._weighted_reduction)
    return task_helper
