"""Fused optimizer step + conv-weight re-layout.

`get_optimizer(args, model)` mirrors the reference's `get_optimizer(args, parameters)` (emsanet/optimizer.py:29-59:
SGD with momentum and nesterov=True, Adam, AdamW — same hyper-parameters from the same `args` fields) but returns
optimizers whose `step()` is ONE kernel launch (csrc/optim.cu, `eb200_optim_step`) over all 675 parameter tensors:
it reads each gradient where autograd left it (normally a view of the engine's flat fp32 gradient buffer, i.e. the
reduced bucket in data-parallel runs), updates the fp32 master parameters and the optimizer state in place with
torch.optim's arithmetic, and writes the bf16 tensor-core layouts of every conv weight from the updated values — so
the engine's per-step weight re-layout launch (`eb200_pack_conv_weights_batched`) and its extra pass over all weights
are gone.  RAdam is not covered (raises).

They are `torch.optim.Optimizer` subclasses: `param_groups`, `zero_grad()`, lr schedulers (the reference steps
OneCycleLR per epoch, emsanet/lr_scheduler.py:23-31), `state_dict()` / `load_state_dict()` (same keys as
torch.optim.SGD / AdamW: 'momentum_buffer' / 'step', 'exp_avg', 'exp_avg_sq') work as with the stock classes, so
main.py's checkpointing (main.py:459-466) is unaffected.  The state tensors are views of flat buffers.

No CPU path: parameters must be CUDA tensors of a model that runs on the emsanet_b200 engine.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import OptimEntry, OptimHyper

SGD, ADAM, ADAMW = 0, 1, 2


class _FusedOptimizer(torch.optim.Optimizer):
    KIND = SGD
    STATE_KEYS = ('momentum_buffer',)

    def __init__(self, model, defaults: Dict):
        from .patch import _engine_for
        self.engine = _engine_for(model)          # raises for CPU models: there is no CPU path
        eng = self.engine
        self.keys: List[str] = list(eng.grad_keys)
        params = [eng.P[k] for k in self.keys]
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise NotImplementedError('one parameter group (the reference passes model.parameters())')
        self.dev = eng.dev
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        self._offsets = [0]
        for s in sizes:
            self._offsets.append(self._offsets[-1] + s)
        self._flat_state = {name: torch.zeros(self._offsets[-1], dtype=torch.float32, device=self.dev)
                            for name in self.STATE_KEYS}
        self._step = 0
        self._tables = None
        self._grad_ptrs = None
        eng.weights_packed_by_optimizer = True    # graph capture leaves the re-layout launch out of the forward graph
        self.flags = 0

    # ------------------------------------------------------------------ state as torch.optim exposes it
    def _state_view(self, name: str, i: int, p: torch.Tensor) -> torch.Tensor:
        o = self._offsets[i]
        return self._flat_state[name][o:o + p.numel()].view(p.shape)

    def _publish_state(self) -> None:
        for i, p in enumerate(self.param_groups[0]['params']):
            st = self.state[p]
            for name in self.STATE_KEYS:
                if name not in st or st[name] is None or st[name].data_ptr() != self._state_view(name, i, p).data_ptr():
                    st[name] = self._state_view(name, i, p)
            if self.KIND != SGD:
                st['step'] = torch.tensor(float(self._step))

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        steps = []
        for i, p in enumerate(self.param_groups[0]['params']):
            st = self.state.get(p, {})
            for name in self.STATE_KEYS:
                if st.get(name) is not None:
                    self._state_view(name, i, p).copy_(st[name])
            if 'step' in st:
                steps.append(int(float(st['step'])))
        if self.KIND == SGD:
            self._step = 1 if any(self.state.get(p, {}).get('momentum_buffer') is not None
                                  for p in self.param_groups[0]['params']) else 0
        else:
            self._step = max(steps) if steps else 0
        self._publish_state()

    # ------------------------------------------------------------------ tables
    def _build_tables(self, grads: List[torch.Tensor]) -> None:
        eng = self.engine
        eng.refresh_weights()                     # makes sure the pack plan (and the packed buffers) exist
        owners = {k: i for i, k in enumerate(eng._pack_owner_keys)}
        chunk = int(_lib.load().eb200_optim_chunk())
        arr = (OptimEntry * len(self.keys))()
        block_entry, block_start = [], []
        params = self.param_groups[0]['params']
        for i, (k, p, g) in enumerate(zip(self.keys, params, grads)):
            e = arr[i]
            e.p, e.g, e.numel = p.data_ptr(), g.data_ptr(), p.numel()
            e.m = self._state_view(self.STATE_KEYS[0], i, p).data_ptr() if self._uses_first_state() else None
            e.v = self._state_view(self.STATE_KEYS[1], i, p).data_ptr() if len(self.STATE_KEYS) > 1 else None
            e.pack = owners.get(k, -1)
            if e.pack >= 0:
                if p.dim() != 4:
                    raise AssertionError(k)
                cout = p.shape[0]
                cin = p.shape[1] * 49 if p.shape[2] == 7 else p.shape[1]   # stem: packed as a 1x1 conv over K = cin * 49
                for ct in range((cout + 31) // 32):
                    for it in range((cin + 31) // 32):
                        block_entry.append(i)
                        block_start.append(ct | (it << 16))
            else:
                for c in range((p.numel() + chunk - 1) // chunk):
                    block_entry.append(i)
                    block_start.append(c)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        self._tables = (raw.to(self.dev), torch.tensor(block_entry, dtype=torch.int32, device=self.dev),
                        torch.tensor(block_start, dtype=torch.int32, device=self.dev), eng._pack_entries)
        self._grad_ptrs = tuple(g.data_ptr() for g in grads)
        self._param_ptrs = tuple(p.data_ptr() for p in params)

    def _uses_first_state(self) -> bool:
        return self.KIND != SGD or self.param_groups[0]['momentum'] != 0

    def _hyper(self) -> OptimHyper:
        raise NotImplementedError

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        params = self.param_groups[0]['params']
        grads = []
        for k, p in zip(self.keys, params):
            g = p.grad
            if g is None:
                raise RuntimeError(f'{k} has no gradient: the fused step updates all parameters of the network together '
                                   '(every parameter of EMSANet receives a gradient from loss.backward())')
            if not g.is_contiguous() or g.dtype != torch.float32:
                raise RuntimeError(f'{k}: gradient must be a contiguous fp32 tensor')
            grads.append(g)
        eng = self.engine
        if (self._tables is None or self._grad_ptrs != tuple(g.data_ptr() for g in grads)
                or self._param_ptrs != tuple(p.data_ptr() for p in params)
                or self._tables[3] is not eng._pack_entries):
            self._build_tables(grads)
        self._step += 1
        h = self._hyper()
        entries, block_entry, block_start, packs = self._tables
        _lib.call('eb200_optim_step', entries.data_ptr(), packs.data_ptr(), block_entry.data_ptr(),
                  block_start.data_ptr(), block_entry.numel(), C.byref(h),
                  torch.cuda.current_stream().cuda_stream)
        # the packed layouts are current: tell the engine (the kernel does not bump torch's version counters)
        eng._pack_versions = tuple(eng.P[k]._version for k in eng._pack_owner_keys)
        self._publish_state()
        return loss


class FusedSGD(_FusedOptimizer):
    """torch.optim.SGD(params, lr, momentum, weight_decay, nesterov) — bit-identical fp32 master parameters"""
    KIND = SGD
    STATE_KEYS = ('momentum_buffer',)

    def __init__(self, model, lr: float, momentum: float = 0.0, weight_decay: float = 0.0, nesterov: bool = False):
        if nesterov and momentum <= 0:
            raise ValueError('Nesterov momentum requires a momentum')
        super().__init__(model, dict(lr=lr, momentum=momentum, dampening=0.0, weight_decay=weight_decay,
                                     nesterov=nesterov, maximize=False, foreach=None, differentiable=False, fused=None))

    def _hyper(self) -> OptimHyper:
        g = self.param_groups[0]
        h = OptimHyper()
        h.kind, h.lr, h.momentum, h.weight_decay = SGD, float(g['lr']), float(g['momentum']), float(g['weight_decay'])
        h.nesterov, h.step, h.flags = int(bool(g['nesterov'])), self._step, self.flags
        h.bias_correction1 = h.bias_correction2_sqrt = 1.0
        return h


class _FusedAdamBase(_FusedOptimizer):
    STATE_KEYS = ('exp_avg', 'exp_avg_sq')

    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        super().__init__(model, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False,
                                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                                     decoupled_weight_decay=self.KIND == ADAMW))

    def _hyper(self) -> OptimHyper:
        g = self.param_groups[0]
        b1, b2 = g['betas']
        h = OptimHyper()
        h.kind, h.lr, h.momentum, h.beta2, h.eps = self.KIND, float(g['lr']), float(b1), float(b2), float(g['eps'])
        h.weight_decay, h.step, h.flags = float(g['weight_decay']), self._step, self.flags
        h.bias_correction1 = 1.0 - b1 ** self._step
        h.bias_correction2_sqrt = math.sqrt(1.0 - b2 ** self._step)
        # computed in Python doubles like torch.optim does, rounded to float once
        h.one_minus_beta1, h.one_minus_beta2 = 1.0 - b1, 1.0 - b2
        h.step_size = g['lr'] / (1.0 - b1 ** self._step)
        h.decay = 1.0 - g['lr'] * g['weight_decay']
        return h


class FusedAdam(_FusedAdamBase):
    KIND = ADAM


class FusedAdamW(_FusedAdamBase):
    KIND = ADAMW


def get_optimizer(args, model):
    """emsanet/optimizer.py:20-59 with the fused step: same `args.optimizer / learning_rate / weight_decay / momentum`"""
    name = args.optimizer.lower()
    if name == 'sgd':
        return FusedSGD(model, lr=args.learning_rate, weight_decay=args.weight_decay, momentum=args.momentum, nesterov=True)
    if name == 'adam':
        return FusedAdam(model, lr=args.learning_rate, weight_decay=args.weight_decay, betas=(0.9, 0.999))
    if name == 'adamw':
        return FusedAdamW(model, lr=args.learning_rate, weight_decay=args.weight_decay, betas=(0.9, 0.999))
    if name == 'radam':
        raise NotImplementedError("no fused step for 'radam': use the reference's get_optimizer for it")
    raise ValueError(f"Unknown optimizer: '{name}'")
