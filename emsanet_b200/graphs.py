"""CUDA-graph replay of the EMSANet kernel programs.

One training step is ~1400 launches of 20-80 us kernels; issuing them one by one from Python costs more host time
than the GPU needs to run them.  `GraphRunner` therefore records the engine's forward program (and, at the first
backward, its backward program) once per (input shape, mode) into CUDA graphs over static buffers and replays them:

  forward : static input buffers <- copy of the caller's rgb/depth ; replay ; then the OUTPUT BOUNDARY — the handful of
            kernels that write the fp32 NCHW outputs — runs eagerly into fresh tensors: nothing handed to the caller
            aliases a static buffer (predictions kept over a validation loop stay valid) and no copy-out is needed
  backward: the kernels that READ dL/d(outputs) run eagerly on the tensors autograd hands over (no copy-in), writing the
            activation gradients the recorded program starts from ; replay ; parameter gradients in the static flat buffer

Nothing numerical changes: the graphs contain exactly the launches of `Engine.forward` / `Engine.backward`
(reference: EMSANet.forward, emsanet/model.py:192-233, and its autograd backward entered at main.py:598), the weight
re-layout and the Dropout2d mask generation (torch's graph-safe Philox) included.  `EB200_NO_GRAPH=1` disables it.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch

from .engine import Engine, _Grads

MAX_ENTRIES = 6   # distinct (shape, mode) programs kept per engine; the least recently used one is evicted


class _NoGC:
    """No cyclic garbage collection while a stream capture is in progress: the engine's tape closures, runners and graph
    entries form reference cycles, so an EARLIER model's CUDA graphs and private pools are released by the cyclic
    collector at an arbitrary allocation — if that happens inside a capture, the cudaGraphExecDestroy / cudaFree it
    triggers invalidates the capture (cudaErrorStreamCaptureInvalidated)."""

    def __enter__(self):
        import gc
        gc.collect()
        self.was = gc.isenabled()
        gc.disable()

    def __exit__(self, *a):
        import gc
        if self.was:
            gc.enable()


def enabled() -> bool:
    return os.environ.get('EB200_NO_GRAPH', '0') in ('', '0')


class _Entry:
    """graphs + static buffers of one (shape, mode)"""

    def __init__(self):
        self.pool = torch.cuda.graph_pool_handle()
        self.rgb: Optional[torch.Tensor] = None
        self.depth: Optional[torch.Tensor] = None
        self.g_fwd: Optional[torch.cuda.CUDAGraph] = None
        self.res: Dict[str, List[torch.Tensor]] = {}
        self.tape: List = []
        self.slots: Dict[str, List] = {}
        self.bwd: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, Dict[str, List[Optional[torch.Tensor]]], torch.Tensor,
                                    Dict[str, torch.Tensor]]] = {}
        self.fwd_launches = 0
        self.bwd_launches = 0
        self.out_fns: List = []          # deferred output-boundary launches: (task, first index, count, make)


class GraphRunner:
    def __init__(self, eng: Engine):
        self.eng = eng
        self.entries: Dict[Tuple, _Entry] = {}
        self.current: Optional[_Entry] = None
        self.generation = 0          # bumped by every forward: a backward must belong to the latest one
        self._sig = None

    # ------------------------------------------------------------------ bookkeeping
    def _signature(self) -> Tuple:
        """addresses of every parameter / buffer: a change (e.g. .to(), load into new storage) invalidates the graphs"""
        return tuple(t.data_ptr() for t in self.eng.P.values())

    def _key(self, rgb, depth, training: bool, track: bool) -> Tuple:
        ref = rgb if rgb is not None else depth
        return (tuple(ref.shape), rgb is not None, depth is not None, bool(training), bool(track))

    def usable(self, rgb, depth, training, track) -> bool:
        if not enabled():
            return False
        if self.eng.taps is not None:     # debug taps need the eager program
            return False
        # the full address signature costs ~0.2 ms of host time (1056 tensors) on the critical path between two steps.
        # A device move builds a new engine (patch._engine_for) and with it a new runner; what is left are exotic
        # re-assignments (`p.data = ...`, load_state_dict(assign=True)): checked on the first call and every 32nd one
        self._calls = getattr(self, '_calls', 0) + 1
        if self._sig is None or self._calls % 32 == 0:
            sig = self._signature()
            if sig != self._sig:
                self.entries.clear()
                self.current = None
                self._sig = sig
        return True

    # ------------------------------------------------------------------ forward
    def forward(self, rgb, depth, training: bool, track: bool = True) -> Dict[str, List[torch.Tensor]]:
        key = self._key(rgb, depth, training, track)
        e = self.entries.pop(key, None)
        if e is None:
            while len(self.entries) >= MAX_ENTRIES:
                # least recently used program goes (ragged last batches, several validation shapes, main.py's sanity
                # check): its graphs and its private pool — a full activation set — are released, nothing falls back
                # to the slow eager path
                old_key = next(iter(self.entries))
                old = self.entries.pop(old_key)
                if old is self.current:
                    self.current = None
                del old
            e = self._capture_forward(rgb, depth, training, track)
        self.entries[key] = e          # dict order = recency
        eng = self.eng
        if not training or eng.weights_packed_by_optimizer:
            # inference: weights only change when somebody loads / trains in between -> re-lay-out eagerly, on demand.
            # Training with a fused optimizer (optim.py): its step kernel already wrote the bf16 layouts; only changes
            # made behind its back (load_state_dict, manual edits: they bump torch's version counters) re-pack here.
            eng.refresh_weights(force=False)
            if not training:
                eng.refresh_folded()
        if e.rgb is not None:
            e.rgb.copy_(rgb, non_blocking=True)
        if e.depth is not None:
            e.depth.copy_(depth, non_blocking=True)
        e.g_fwd.replay()
        self.current = e
        self.generation += 1
        # output boundary: the kernels that write the fp32 NCHW outputs run here, eagerly, into FRESH tensors — nothing
        # the caller receives aliases a static buffer (predictions collected over a validation loop stay valid) and no
        # copy out of static buffers is needed
        res = {t: list(outs) for t, outs in e.res.items()}
        for task, idx, count, make in e.out_fns:
            t = make()
            res[task][idx:idx + count] = list(t) if isinstance(t, (tuple, list)) else [t]
        return res

    def _capture_forward(self, rgb, depth, training, track) -> _Entry:
        from . import _lib
        eng = self.eng
        e = _Entry()
        e.rgb = torch.empty_like(rgb) if rgb is not None else None
        e.depth = torch.empty_like(depth) if depth is not None else None
        saved = (eng.on_grads_ready, eng.force_repack)
        eng.on_grads_ready = None
        # warm-up on a side stream: lazy one-time initialisation (kernel attributes, scratch buffers, pack plan) must not
        # happen inside a capture.  track_running_stats=False: the warm-up must not touch the running statistics.
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            if e.rgb is not None:
                e.rgb.copy_(rgb)
            if e.depth is not None:
                e.depth.copy_(depth)
            res = eng.forward(e.rgb, e.depth, training, False)
            if training:
                eng.backward({t: [torch.zeros_like(o) for o in outs] for t, outs in res.items()})
            del res
        torch.cuda.current_stream().wait_stream(s)
        eng._eval_bn.clear()
        try:
            # a training step changes every weight: the re-layout is part of the graph — unless a fused optimizer does it
            eng.force_repack = bool(training) and not eng.weights_packed_by_optimizer
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            eng.boundary_fwd = []
            with _NoGC(), torch.no_grad(), torch.cuda.graph(g, pool=e.pool):
                e.res = eng.forward(e.rgb, e.depth, training, track)
            e.fwd_launches = _lib.launch_count() - l0
            e.g_fwd = g
            e.tape, e.slots = eng.tape, eng.grad_out_slots
            by_list = {id(outs): t for t, outs in e.res.items()}
            e.out_fns = [(by_list[id(outs)], idx, count, make) for outs, idx, count, make in eng.boundary_fwd]
            eng.tape, eng.grads = [], None
        finally:
            eng.boundary_fwd = None
            eng.on_grads_ready, eng.force_repack = saved
            eng._eval_bn.clear()   # affine tensors computed inside the capture live in the graph's pool
        return e

    # ------------------------------------------------------------------ backward
    def backward(self, grad_outputs: Dict[str, List[Optional[torch.Tensor]]]) -> Dict[str, torch.Tensor]:
        e = self.current
        if e is None or not e.tape:
            raise RuntimeError('emsanet_b200: backward without a preceding training-mode forward')
        pattern = (self.eng.on_grads_ready is not None,) + tuple((t, tuple(g is not None for g in gs))
                                                                for t, gs in sorted(grad_outputs.items()))
        hit = e.bwd.get(pattern)
        if hit is None:
            hit = self._capture_backward(e, grad_outputs)
            e.bwd[pattern] = hit
        parts, boundary, flat, G = hit
        self._rescue_aliased_grads(flat)
        eng = self.eng
        eng.flat_grad = flat
        # output boundary: the kernels that read dL/d(outputs) run here, eagerly, on the tensors autograd hands over (no
        # copy into static buffers); they write the activation gradients the recorded program starts from and already
        # accumulate parameter gradients, so the flat gradient buffer is zeroed here and not inside the graph
        flat.zero_()
        for task, indices, launch, dx in boundary:
            gs = [grad_outputs[task][i] for i in indices]
            launch(*[g.contiguous() if g is not None else None for g in gs], out=dx)
        # data parallel: the program is recorded in up to three graphs, split where a bucket of the flat gradient buffer
        # becomes final ([decoders + context] at the encoder boundary, [encoder stages 3-4] after them, the rest at the
        # end); each bucket is all-reduced on NCCL's stream while the next graph runs (same buckets as the eager path)
        for g, (lo, hi) in parts:
            g.replay()
            if eng.on_grads_ready is not None:
                eng.on_grads_ready(flat, lo, hi)
        return G

    def _rescue_aliased_grads(self, flat: torch.Tensor) -> None:
        """Gradient accumulation without zero_grad(): a `.grad` adopted from the previous replay aliases the static
        buffer this replay is about to overwrite — move it to private memory first."""
        lo = flat.data_ptr()
        hi = lo + flat.numel() * 4
        for p in self.eng.P.values():
            g = p.grad
            if g is not None and lo <= g.data_ptr() < hi:
                p.grad = g.clone()

    def _capture_backward(self, e: _Entry, grad_outputs):
        from . import _lib
        eng = self.eng
        saved = eng.on_grads_ready
        eng.on_grads_ready = None
        # the flat gradient buffer lives OUTSIDE the graph pool: `.grad`s adopted from it must not be scribbled over by
        # the next forward replay (pool memory is recycled between the two graphs), only by the next backward replay
        flat_static = torch.zeros(eng.param_grad_floats(), dtype=torch.float32, device=eng.dev)
        split = saved is not None      # a gradient reducer is attached: one graph per gradient bucket
        try:
            parts = []
            l0 = _lib.launch_count()
            eng.boundary_bwd = []
            first = True
            while first or eng.tape:
                g = torch.cuda.CUDAGraph()
                with _NoGC(), torch.no_grad(), torch.cuda.graph(g, pool=e.pool):
                    if first:
                        eng.tape = list(e.tape)
                        eng.grads = _Grads()
                        eng.grad_out_slots = e.slots
                        eng.training = True
                        # the output gradients only decide which boundary ops exist (None-ness) and their shapes here:
                        # the launches that read them are deferred
                        eng.begin_backward(grad_outputs, flat=flat_static, zero=False)
                    marker = eng.run_tape(stop_at_marker=split)
                first = False
                parts.append((g, eng.ready_range(marker)))
            if marker is not None:   # a marker was the last entry of the tape: the remaining range is final as well
                parts[-1] = (parts[-1][0], (0, parts[-1][1][1]))
            if not split:
                parts = [(parts[0][0], (0, flat_static.numel()))]
            eng.grads = None
            G = dict(eng.G)          # copy: the engine's dict is refilled by eager runs
            flat = eng.flat_grad
            e.bwd_launches = _lib.launch_count() - l0
            by_slot = {id(sl): t for t, sl in e.slots.items()}
            boundary = [(by_slot[id(sl)], indices, launch, dx) for sl, indices, launch, dx in eng.boundary_bwd]
            for sl in e.slots.values():          # do not keep the caller's gradient tensors alive
                sl.clear()
        finally:
            eng.boundary_bwd = None
            eng.on_grads_ready = saved
        return parts, boundary, flat, G
