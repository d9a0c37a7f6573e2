"""Fused loss layer behind the reference's loss interfaces.

MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/.  Mirrors, same constructors and the same
`loss(input_tensors, target_tensors) -> ((loss, n_elements), ...)` contract, `loss` an autograd-connected fp32 scalar:

  CrossEntropyLossSemanticB200   MT/loss/ce.py:13-68 (weighted_reduction=False)        eb200_ce_loss_fwd / _bwd
  MSELossB200                    MT/loss/mse.py:13-41  (reduction='sum')                eb200_masked_loss_fwd / _bwd kind 0
  L1LossB200                     MT/loss/l1.py:13-41   (reduction='sum')                kind 1
  VonMisesLossBiternionB200      MT/loss/vonmises.py:18-51 (reduction='sum')            kind 2

Each forward is ONE pass over the prediction (loss and element count accumulated on the device in fp64 / int64), each
backward ONE pass that writes the gradient already scaled by the upstream gradient — instead of the elementwise loss
tensor, its channel mean, the sum and their autograd backward as separate full-tensor passes.

`instance_losses()` is `InstanceTaskHelper._compute_losses` (MT/task_helper/instance.py:92-262) without its host
synchronisations: the reference multiplies predictions by masks, calls the three losses and fetches every valid-pixel
count with `.sum().cpu().detach().item()` (three syncs per scale, twelve per step: SURVEY.md P9); here mask, loss and
count are one kernel and the normalisation `loss / n` stays on the device.  `install(task_helper)` swaps the loss objects
of a reference SemanticTaskHelper / InstanceTaskHelper (after `initialize(device)`) and, for the instance helper, its
`_compute_losses`.

All kernels verified on a B200 against oracle/loss_oracle.py, which is pinned against the unmodified reference classes
(tests/test_loss.py; first CE run: profiles/r2_unverified_kernels_first_run.log).  No CPU fallback: non-CUDA inputs raise.
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
        n = int(count.item())                   # the reference synchronises here as well (ce.py:50)
        if n >> 62:                             # out-of-range flag set by the kernel (torch device-asserts on this)
            raise _lib.EB200Error(f'cross entropy: target labels above the number of classes ({x.shape[1]}) — dataset / '
                                  'n_classes mismatch')
        return loss, n

    def forward(self, input_tensors: Sequence[torch.Tensor], target_tensors: Sequence[torch.Tensor]):
        return tuple(self._compute_loss(i, t) for i, t in zip(input_tensors, target_tensors))   # base.py:23-33


# ------------------------------------------------------------------------------------------------ regression losses
MSE, L1, VONMISES = 0, 1, 2


def _layout(x: torch.Tensor, channel_dim: Optional[int]):
    """(N, C, P, sn, sc, sp) of a contiguous tensor whose channel axis is `channel_dim` (None: no channel axis)"""
    if channel_dim is None:
        return 1, 1, x.numel(), x.numel(), 0, 1
    if channel_dim == 1:
        n, c = x.shape[0], x.shape[1]
        p = x.numel() // (n * c)
        return n, c, p, c * p, p, 1
    assert channel_dim == x.dim() - 1
    c = x.shape[-1]
    return 1, c, x.numel() // c, x.numel(), 1, c


def masked_loss_forward(kind: int, pred, target, mask, channel_dim, kappa: float = 1.0):
    n, c, p, sn, sc, sp = _layout(pred, channel_dim)
    loss = torch.empty(1, dtype=torch.float64, device=pred.device)
    count = torch.empty(1, dtype=torch.int64, device=pred.device)
    _lib.call('eb200_masked_loss_fwd', kind, _p(pred), _p(target), _p(mask), n, c, p, sn, sc, sp, float(kappa), _p(loss),
              _p(count), _stream())
    return loss, count


def masked_loss_backward(kind: int, pred, target, mask, channel_dim, kappa, grad_out):
    n, c, p, sn, sc, sp = _layout(pred, channel_dim)
    d = torch.empty_like(pred)
    _lib.call('eb200_masked_loss_bwd', kind, _p(pred), _p(target), _p(mask), n, c, p, sn, sc, sp, float(kappa),
              _p(grad_out), _p(d), _stream())
    return d


class _FusedMaskedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, kind, channel_dim, kappa):
        loss64, count = masked_loss_forward(kind, pred, target, mask, channel_dim, kappa)
        ctx.save_for_backward(pred, target, mask if mask is not None else pred.new_empty(0))
        ctx.args = (kind, channel_dim, kappa, mask is not None)
        ctx.mark_non_differentiable(count)
        return loss64.to(torch.float32).reshape(()), count

    @staticmethod
    def backward(ctx, grad_loss, _grad_count):
        pred, target, mask = ctx.saved_tensors
        kind, channel_dim, kappa, has_mask = ctx.args
        g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        d = masked_loss_backward(kind, pred, target, mask if has_mask else None, channel_dim, kappa, g)
        return d, None, None, None, None, None


def fused_masked_loss(kind: int, pred: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor],
                      channel_dim: Optional[int], kappa: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (loss fp32 scalar with autograd, valid-pixel count int64 [1] on the device)"""
    x = _check(pred, 'prediction')
    if x.dtype != torch.float32:
        x = x.float()
    t = _check(target, 'target').to(torch.float32)
    if t.shape != x.shape:
        raise ValueError(f'prediction {tuple(x.shape)} and target {tuple(t.shape)} differ in shape')
    m = None
    if mask is not None:
        m = _check(mask, 'mask')
        m = (m != 0).to(torch.uint8) if m.dtype not in (torch.bool, torch.uint8) else m.contiguous()
        if m.dtype == torch.bool:
            m = m.view(torch.uint8)
        n, c, p, *_ = _layout(x, channel_dim)
        if m.numel() != n * p:
            raise ValueError(f'mask with {m.numel()} elements for {n * p} pixels')
    return _FusedMaskedLoss.apply(x, t, m, kind, channel_dim, float(kappa))


class _RegressionLossB200(torch.nn.Module):
    KIND = MSE

    def __init__(self, reduction: str = 'sum') -> None:
        super().__init__()
        if reduction != 'sum':
            raise NotImplementedError(f"reduction='{reduction}' is not covered (the task helpers use 'sum')")
        self._reduction = reduction

    def _compute_loss(self, input_: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, int]:
        channel_dim = 1 if input_.ndim in (2, 4) else None          # mse.py:31-34 / l1.py:31-34: mean over dim 1
        loss, _ = fused_masked_loss(self.KIND, input_, target, None, channel_dim)
        n = input_.numel() // (input_.shape[1] if channel_dim == 1 else 1)
        return loss, n

    def forward(self, input_tensors: Sequence[torch.Tensor], target_tensors: Sequence[torch.Tensor]):
        return tuple(self._compute_loss(i, t) for i, t in zip(input_tensors, target_tensors))   # base.py:23-33


class MSELossB200(_RegressionLossB200):
    KIND = MSE


class L1LossB200(_RegressionLossB200):
    KIND = L1


class VonMisesLossBiternionB200(torch.nn.Module):
    def __init__(self, reduction: str = 'sum', kappa: float = 1.0) -> None:
        super().__init__()
        if reduction != 'sum':
            raise NotImplementedError("reduction='none' is not covered")
        self._kappa, self._reduction = float(kappa), reduction

    def _compute_loss(self, input_: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, int]:
        if input_.ndim != 2 or target.ndim != 2:                    # vonmises.py:34-42
            raise ValueError('VonMisesLossBiternion does only support 2d inputs with shape (n, 2)')
        if input_.shape[0] == 0:                                    # no oriented pixel in the batch: loss 0 (:47-49)
            return input_.sum() * 0.0, 0
        loss, _ = fused_masked_loss(VONMISES, input_, target, None, input_.dim() - 1, self._kappa)
        return loss, input_.shape[0]

    def forward(self, input_tensors: Sequence[torch.Tensor], target_tensors: Sequence[torch.Tensor]):
        return tuple(self._compute_loss(i, t) for i, t in zip(input_tensors, target_tensors))


def instance_losses(helper, batch, batch_idx, predictions_post):
    """`InstanceTaskHelper._compute_losses` (MT/task_helper/instance.py:92-262) with mask, loss and valid-pixel count of
    every scale fused into one kernel each and NO host synchronisation: the per-scale losses are `sum / count` with the
    count kept on the device (orientation: max(count, 1), instance.py:199-201), the totals `sum(sums) / sum(counts)`
    (base.py:161-182).  Same keys and values as the reference's dictionary."""
    no_multiscale = helper._disable_multiscale_supervision
    preds, keys, downscales = helper.collect_predictions_for_loss(
        predictions_post=predictions_post, predictions_post_key='instance_output',
        side_outputs_key=None if no_multiscale else 'instance_side_outputs')
    with_orientation = len(preds[0]) == 3
    helper._with_orientation = with_orientation

    def targets(key):
        return helper.collect_targets_for_loss(batch, batch_key=key, downscales=downscales)
    t_cmask, t_center = targets('instance_center_mask'), targets('instance_center')
    t_fg, t_offset = targets('instance_foreground'), targets('instance_offset')
    kinds = {'mse': MSE, 'l1': L1}
    center_kind = kinds[getattr(helper, '_loss_name_instance_center', 'mse')]
    sums = {'center': [], 'offset': [], 'orientation': []}
    counts = {'center': [], 'offset': [], 'orientation': []}
    for i, pred in enumerate(preds):
        s, c = fused_masked_loss(center_kind, pred[0][:, 0], t_center[i], t_cmask[i], None)     # instance.py:118-139
        sums['center'].append(s), counts['center'].append(c)
        s, c = fused_masked_loss(L1, pred[1], t_offset[i], t_fg[i], 1)                          # :141-167
        sums['offset'].append(s), counts['offset'].append(c)
    if with_orientation:
        t_ofg, t_orient = targets('orientation_foreground'), targets('orientation')
        kappa = getattr(helper._loss_orientation, '_kappa', 1.0)
        for i, pred in enumerate(preds):
            s, c = fused_masked_loss(VONMISES, pred[2], t_orient[i], t_ofg[i], 1, kappa)        # :169-207
            sums['orientation'].append(s), counts['orientation'].append(c.clamp(min=1))
    loss_dict = {}
    names = ['center', 'offset'] + (['orientation'] if with_orientation else [])
    for name in names:
        for key, s, c in zip(keys, sums[name], counts[name]):
            loss_dict[f'instance_{name}_loss_{key}'] = s / c.to(torch.float32).reshape(())
    for name in names:                                                                          # :240-262
        total_n = torch.stack(counts[name]).sum().to(torch.float32)
        total = torch.stack(sums[name]).sum()
        loss_dict[helper.mark_as_total(f'instance_{name}')] = torch.where(total_n > 0, total / total_n.clamp(min=1),
                                                                          total)
    return loss_dict


def install(task_helper):
    """swap the loss objects of a reference task helper (after `initialize(device)`) for the fused mirrors:
    SemanticTaskHelper._loss; InstanceTaskHelper._loss_center / _loss_offset / _loss_orientation and its
    `_compute_losses` (the sync-free `instance_losses`)."""
    import types
    if hasattr(task_helper, '_loss_center'):
        old_c = task_helper._loss_center
        task_helper._loss_center = {'MSELoss': MSELossB200, 'L1Loss': L1LossB200}[type(old_c).__name__]('sum')
        task_helper._loss_offset = L1LossB200('sum')
        task_helper._loss_orientation = VonMisesLossBiternionB200(kappa=task_helper._loss_orientation._kappa)
        task_helper._compute_losses = types.MethodType(instance_losses, task_helper)
        return task_helper
    old = task_helper._loss
    if type(old).__name__ != 'CrossEntropyLossSemantic':
        raise NotImplementedError(f'no fused mirror for {type(old).__name__}')
    task_helper._loss = CrossEntropyLossSemanticB200(weights=old._weights, label_smoothing=old._loss.label_smoothing,
                                                    weighted_reduction=old._weighted_reduction)
    return task_helper
