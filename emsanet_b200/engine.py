"""The EMSANet forward/backward program over our CUDA kernels (host side).

Mirrors, op for op, what `EMSANet.forward` (emsanet/model.py:192-233) dispatches to — encoder stages with
SE-add RGB-D fusion (MT/model/encoder.py:220-261), pyramid pooling (MT/model/context_module/ppm.py:57-78),
the dense decoders with skip fusion (MT/model/decoder/dense_base.py:229-284), the task heads — and the
autograd backward of all of it (entered at main.py:598), as an explicit tape of fused kernels:

  * activations: NHWC bf16 in HBM; accumulators, statistics, parameters and parameter gradients fp32
  * conv + bias + ReLU, conv + BN statistics, dgrad + ReLU-mask + bias-gradient are single kernels
  * BN apply fuses dropout, residual / skip add, ReLU and the SE squeeze
  * stride-2 convs run as stride-1 convs over parity views; the stem runs as im2col + 1x1 tensor-core GEMM

Parameters are addressed by the reference's state_dict keys and stay the caller's
tensors; there is no fallback path: a missing CUDA library raises.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import ops
from .ops import BF16, PackedWeight

RESNET_LAYERS = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3), 'resnet101': (3, 4, 23, 3)}
STAGE_CHANNELS = (64, 64, 128, 256, 512)


@dataclasses.dataclass(frozen=True)
class EngineConfig:
    backbone: str = 'resnet34'
    modalities: Tuple[str, ...] = ('rgb', 'depth')
    tasks: Tuple[str, ...] = ('semantic', 'scene', 'instance', 'orientation')
    enable_panoptic: bool = True
    semantic_n_classes: int = 40
    scene_n_classes: int = 10
    decoder_n_channels: Tuple[int, ...] = (512, 256, 128)
    decoder_n_blocks: int = 3
    ppm_bins: Tuple[int, ...] = (1, 5)
    dropout_p_encoder: float = 0.1
    dropout_p_decoder: float = 0.2
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1

    @property
    def layers(self):
        return RESNET_LAYERS[self.backbone]

    def backbone_prefix(self, m: str) -> str:
        return f'encoder.backbone_{m}.' if len(self.modalities) == 2 else 'encoder.backbone.'

    @property
    def with_orientation(self) -> bool:
        return 'orientation' in self.tasks

    @property
    def decoder_prefixes(self) -> Dict[str, str]:
        pre = 'decoders.panoptic_helper.' if (self.enable_panoptic and 'semantic' in self.tasks
                                              and 'instance' in self.tasks) else 'decoders.'
        out = {}
        if 'semantic' in self.tasks:
            out['semantic'] = pre + 'semantic_decoder.'
        if 'instance' in self.tasks:
            out['instance'] = pre + 'instance_decoder.'
        if 'scene' in self.tasks:
            out['scene'] = 'decoders.scene_decoder.'
        return out

    def dropout_sites(self) -> List[Tuple[str, int, float]]:
        """(block prefix, channels, p) of every Dropout2d in reference execution order."""
        sites = []
        for li, (n, c) in enumerate(zip(self.layers, (64, 128, 256, 512)), start=1):
            for m in ('rgb', 'depth'):
                if m in self.modalities:
                    for b in range(n):
                        sites.append((f'{self.backbone_prefix(m)}layer{li}.{b}.', c, self.dropout_p_encoder))
        for task in ('semantic', 'instance'):
            if task in self.decoder_prefixes:
                for i, c in enumerate(self.decoder_n_channels):
                    for b in range(self.decoder_n_blocks):
                        sites.append((f'{self.decoder_prefixes[task]}decoder_modules.{i}.blocks.{b}.', c,
                                      self.dropout_p_decoder))
        return sites


class _Grads:
    """gradient slots of activations, keyed by tensor identity"""

    def __init__(self):
        self.slots: Dict[int, torch.Tensor] = {}
        self.keep: Dict[int, torch.Tensor] = {}

    def has(self, t) -> bool:
        return id(t) in self.slots

    def get(self, t) -> Optional[torch.Tensor]:
        return self.slots.get(id(t))

    def pop(self, t) -> Optional[torch.Tensor]:
        self.keep.pop(id(t), None)
        return self.slots.pop(id(t), None)

    def add(self, t, g: torch.Tensor) -> None:
        cur = self.slots.get(id(t))
        if cur is None:
            self.slots[id(t)] = g
            self.keep[id(t)] = t   # pin the key tensor so ids stay unique
        else:
            ops.add_inplace(cur, g)


class Engine:
    def __init__(self, cfg: EngineConfig, params: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.P = params
        self.dev = next(iter(params.values())).device
        self._packed: Dict[str, Tuple[int, PackedWeight]] = {}
        self._scratch: Dict[Tuple, torch.Tensor] = {}
        self._eval_bn: Dict[str, Tuple[Tuple[int, ...], ops.BNState]] = {}
        self.grad_keys = [k for k, v in params.items() if v.is_floating_point() and 'running_' not in k]
        self.tape: List[Callable[[], None]] = []
        self.G: Dict[str, torch.Tensor] = {}
        self.grads: Optional[_Grads] = None
        self.training = False
        self.track = True
        self.masks: Dict[str, torch.Tensor] = {}
        self._bn_touched: List[str] = []
        self.taps: Optional[Dict[str, torch.Tensor]] = None   # debug: name -> NHWC activation
        self.on_grads_ready = None  # data parallel: callable(flat_grad, start, end) when that slice is final
        self._join_at_markers = False
        self._mid_done = False
        self._enc_lo_end = self._enc_end = 0
        self.force_repack = False   # benchmarks: pay for the weight re-layout every step, as training does
        self.weights_packed_by_optimizer = False   # optim.FusedSGD & co. write the bf16 layouts in their step kernel
        # Output boundary (graph replay, graphs.py): the kernels that WRITE the fp32 NCHW outputs and the kernels that READ
        # their gradients can be left out of the recorded programs and launched eagerly around the replay — outputs then
        # land in fresh tensors (no copy out of static buffers), output gradients are read where autograd left them (no
        # copy into static buffers).  None = launch in place.
        self.boundary_fwd: Optional[List] = None      # [(outs list, first index, count, make() -> tensor | tuple)]
        self.boundary_bwd: Optional[List] = None      # [(slot list, indices, launch(*grads))]
        self._pack_entries = None
        # zero-initialised fp32 scratch handed out per step (BN statistics, backward sum replicas): two bump arenas,
        # each cleared by ONE memset — the forward one in begin(), the backward one when backward starts
        self._arena = {'fwd': [None, 0, 1 << 16], 'bwd': [None, 0, 1 << 18]}   # [buffer, offset, floats needed]
        self._arena_cur = 'fwd'
        import os
        self.fuse_bn_bwd = os.environ.get('EB200_NO_BN_FUSE', '0') in ('', '0')   # experiments: unfused norm1 backward
        self.pair_siblings = os.environ.get('EB200_NO_PAIR', '0') in ('', '0')    # one launch for sibling branches
        # Weight gradients of the NBt1D blocks run on a second stream: nothing in the backward chain waits for them, and
        # their CTAs take the SMs that the data-gradient kernels' short waves (320 tiles on 148 SMs: a third of the CTAs
        # exit one tile early) and the bandwidth-bound BatchNorm kernels leave idle.  In-flight depth bounds the number of
        # activation-gradient tensors kept alive for the side stream.
        self.overlap_wgrad = os.environ.get('EB200_NO_WGRAD_OVERLAP', '0') in ('', '0')
        self.overlap_wgrad_layers = os.environ.get('EB200_NO_WGRAD_OVERLAP_LAYERS', '0') in ('', '0')
        self.fuse_output_upsample = os.environ.get('EB200_NO_FUSED_OUTPUT', '0') in ('', '0')
        # Sibling branches (RGB / depth encoder, semantic / instance decoder) issue their bandwidth-bound kernels
        # (BatchNorm apply / backward) in lock step: the second one goes to a side stream, so the two 15-50 us kernels
        # overlap their ramps and tails instead of running back to back (the tensor-core kernels of siblings already
        # share one launch).
        self.parallel_siblings = os.environ.get('EB200_NO_PARALLEL_SIBLINGS', '0') in ('', '0')
        self._side2: Optional[torch.cuda.Stream] = None
        # inference: every conv -> BatchNorm pair runs as ONE conv whose bf16 weights carry the BatchNorm scale and whose
        # epilogue adds the shift (+ residual) and applies the ReLU — the 127 bn_apply passes of an eval forward are gone
        self.fold_eval_bn = os.environ.get('EB200_NO_BN_FOLD', '0') in ('', '0')
        self._folded: Dict[str, list] = {}      # wkey -> [versions, PackedWeight, shift, bn prefix, stem?]
        self._side: Optional[torch.cuda.Stream] = None
        self._side_pending: List = []
        self._side_depth = int(os.environ.get('EB200_WGRAD_OVERLAP_DEPTH', '6'))

    # ------------------------------------------------------------------ helpers
    def _tap(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """debug / parity: expose a stored activation (NHWC bf16) under the oracle's name of that storage point"""
        if self.taps is not None:
            self.taps[name] = t
        return t

    def _emit_outputs(self, outs: List, count: int, make) -> int:
        """append `count` network outputs produced by make() (a tensor or a tuple of tensors) to `outs`; returns the
        index of the first one.  With the boundary deferred, placeholders are appended and make() runs after the replay."""
        idx = len(outs)
        if self.boundary_fwd is None:
            t = make()
            outs.extend(t if isinstance(t, (tuple, list)) else [t])
            assert len(outs) == idx + count
        else:
            self.boundary_fwd.append((outs, idx, count, make))
            outs.extend([None] * count)
        return idx

    def _boundary_dx(self, x: torch.Tensor) -> Optional[torch.Tensor]:
        """Deferred boundary, training forward: reserve the gradient buffer of a boundary op's input NOW.  The deferred
        backward launch writes it BEFORE the recorded backward program runs; a buffer allocated inside the backward
        capture could share its memory with a temporary the program uses earlier (the allocator recycles within a
        capture in program order) and be clobbered.  Allocated here it outlives every backward temporary."""
        return torch.empty_like(x) if (self.boundary_fwd is not None and self.training) else None

    def _boundary_backward(self, slot: List, indices, x: torch.Tensor, launch, dx: Optional[torch.Tensor] = None) -> None:
        """backward of an output-boundary op: dx (a gradient slot of activation x) = launch(*output gradients, out=dx).
        With the boundary deferred only the bookkeeping happens here; the launch runs eagerly before the replay."""
        gs = [slot[i] for i in indices]
        if all(g is None for g in gs):
            return
        if dx is None:
            assert self.boundary_bwd is None, 'deferred boundary backward without a reserved gradient buffer'
            dx = torch.empty_like(x)
        if self.boundary_bwd is None:
            launch(*[g.contiguous() if g is not None else None for g in gs], out=dx)
        else:
            self.boundary_bwd.append((slot, tuple(indices), launch, dx))
        self.grads.add(x, dx)

    def _arena_reset(self, which: str) -> None:
        a = self._arena[which]
        if a[0] is None or a[0].numel() < a[2]:
            a[0] = torch.zeros(int(a[2] * 1.25) + 1024, dtype=torch.float32, device=self.dev)
        else:
            a[0].zero_()
        a[1] = 0
        self._arena_cur = which

    def arena_zeros(self, n: int) -> torch.Tensor:
        """n zeroed floats, valid for the current forward / backward program (16-byte aligned)"""
        a = self._arena[self._arena_cur]
        n4 = (n + 3) // 4 * 4
        off = a[1]
        a[1] = off + n4
        a[2] = max(a[2], a[1])
        if a[0] is None or a[1] > a[0].numel():   # first step of a bigger program: private buffer now, bigger arena next
            t = torch.empty(n, dtype=torch.float32, device=self.dev)
            # cleared on the stream its consumer is launched on (a sibling branch may be launching on the side stream:
            # a torch.zeros() would be ordered on the main stream only)
            ops._lib.call('eb200_memset_zero', t.data_ptr(), 4 * n, ops._stream())
            return t
        return a[0][off:off + n]

    def zeros(self, tag, n) -> torch.Tensor:
        """persistent zero-initialised fp32 scratch; the kernels that consume it zero it again"""
        key = (tag, n)
        t = self._scratch.get(key)
        if t is None:
            t = torch.zeros(n, dtype=torch.float32, device=self.dev)
            self._scratch[key] = t
        return t

    def weight(self, key: str, need_bwd: bool = True) -> PackedWeight:
        w = self.P[key]
        ver = w._version
        hit = self._packed.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        pw = ops.pack_weight(w.detach(), need_bwd, out=hit[1] if hit is not None else None)
        self._packed[key] = (ver, pw)
        return pw

    @property
    def folding(self) -> bool:
        """eval-mode BatchNorm folding is on (not while debug taps ask for the raw conv outputs)"""
        return (not self.training) and self.fold_eval_bn and self.taps is None

    def folded(self, wkey: str, bnp: str, stem: bool = False) -> Tuple[PackedWeight, torch.Tensor]:
        """eval: (bf16 layout of W * gamma / sqrt(running_var + eps), shift = beta - running_mean * scale) of a
        conv -> BatchNorm pair (nn.BatchNorm2d eval arithmetic).  Cached per parameter versions and
        refreshed IN PLACE (graph replay keeps the addresses)."""
        P = self.P
        vers = (P[wkey]._version, P[bnp + 'weight']._version, P[bnp + 'bias']._version,
                P[bnp + 'running_mean']._version, P[bnp + 'running_var']._version)
        hit = self._folded.get(wkey)
        if hit is not None and hit[0] == vers:
            return hit[1], hit[2]
        with torch.no_grad():
            scale = P[bnp + 'weight'] * torch.rsqrt(P[bnp + 'running_var'] + self.cfg.bn_eps)
            shift = (P[bnp + 'bias'] - P[bnp + 'running_mean'] * scale).contiguous()
            wf = P[wkey].detach() * scale.view(-1, 1, 1, 1)
            if stem:
                wf = wf.reshape(wf.shape[0], -1, 1, 1)
            pw = ops.pack_weight(wf.contiguous(), need_bwd=False, out=hit[1] if hit is not None else None)
        if hit is not None:
            hit[2].copy_(shift)
            hit[0] = vers
            return hit[1], hit[2]
        self._folded[wkey] = [vers, pw, shift, bnp, stem]
        return pw, shift

    def refresh_folded(self) -> None:
        """re-fold whatever changed since the last eval forward (called eagerly before an eval graph replay)"""
        for wkey, (vers, pw, shift, bnp, stem) in list(self._folded.items()):
            self.folded(wkey, bnp, stem)

    def block_diag_weight(self, key: str, wkeys: List[str], couts: List[int], cin_each: int) -> PackedWeight:
        """instance task convs (MT/model/decoder/instance.py:100-109): conv t reads channels [32t, 32t+32) and
        writes its own output channels -> one block-diagonal conv 32*T -> 8 (5 real) channels."""
        vers = tuple(self.P[k]._version for k in wkeys)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == vers:
            return hit[1]
        w0 = self.P[wkeys[0]]
        kh, kw = w0.shape[2], w0.shape[3]
        cin = cin_each * len(wkeys)
        if hit is None:
            fwd = torch.zeros(kh * kw, 16, ops.round_up(cin, 64), dtype=BF16, device=self.dev)
            bwd = torch.zeros(kh * kw, ops.pad_cout(cin), 64, dtype=BF16, device=self.dev)
            pw = PackedWeight(fwd, bwd, 8, cin, kh, kw)
        else:
            pw = hit[1]
        co = 0
        for t, (k, c) in enumerate(zip(wkeys, couts)):
            w = self.P[k].detach()
            ops._lib.call('eb200_pack_conv_weight', w.data_ptr(), c, cin_each, kh, kw, pw.fwd.data_ptr(),
                          pw.fwd.shape[1], pw.fwd.shape[2], 0, co, t * cin_each, ops._stream())
            ops._lib.call('eb200_pack_conv_weight', w.data_ptr(), c, cin_each, kh, kw, pw.bwd.data_ptr(),
                          pw.bwd.shape[1], pw.bwd.shape[2], 1, co, t * cin_each, ops._stream())
            co += c
        self._packed[key] = (vers, pw)
        return pw

    # ------------------------------------------------------------------ batched weight re-layout
    def _build_pack_plan(self) -> None:
        """One table for every tensor-core conv weight: regular convs, the stems (as 1x1 over the im2col K axis)
        and the block-diagonal instance task convs; refreshed by ONE kernel launch per step."""
        import ctypes as C
        from ._lib import PackEntry
        entries, owners = [], []
        groups: Dict[str, List[str]] = {}
        for key, w in self.P.items():
            if w.dim() != 4:
                continue
            if 'task_convs.' in key:
                groups.setdefault(key.split('task_convs.')[0] + 'task_convs', []).append(key)
                continue
            if w.shape[1] == 1 and w.shape[2:] == (3, 3) and 'upsampl' in key:
                continue                                  # depthwise upsampling weights stay fp32
            if key.startswith('encoder.fusions.'):
                continue                                  # SE squeeze MLP runs in fp32
            wd = w.detach()
            if w.shape[2] == 7:                           # stem
                wv = wd.reshape(w.shape[0], -1, 1, 1)
                cout, cin, kh, kw = wv.shape
                fwd = torch.zeros(1, ops.pad_cout(cout), ops.round_up(cin, 64), dtype=BF16, device=self.dev)
                pw = PackedWeight(fwd, None, cout, cin, 1, 1)
            else:
                cout, cin, kh, kw = w.shape
                fwd = torch.zeros(kh * kw, ops.pad_cout(cout), ops.round_up(cin, 64), dtype=BF16, device=self.dev)
                bwd = torch.zeros(kh * kw, ops.pad_cout(ops.round_up(cin, 8)), ops.round_up(cout, 64), dtype=BF16,
                                  device=self.dev)
                pw = PackedWeight(fwd, bwd, cout, cin, kh, kw)
            self._packed[key] = (w._version, pw)
            entries.append((wd, pw, cout, cin, pw.kh * pw.kw, 0, 0))
            owners.append(key)
        for gkey, wkeys in groups.items():
            wkeys = sorted(wkeys)
            w0 = self.P[wkeys[0]]
            kh, kw = w0.shape[2], w0.shape[3]
            cin_each = w0.shape[1]
            cin = cin_each * len(wkeys)
            fwd = torch.zeros(kh * kw, 16, ops.round_up(cin, 64), dtype=BF16, device=self.dev)
            bwd = torch.zeros(kh * kw, ops.pad_cout(cin), 64, dtype=BF16, device=self.dev)
            pw = PackedWeight(fwd, bwd, 8, cin, kh, kw)
            self._packed[gkey] = (tuple(self.P[k]._version for k in wkeys), pw)
            co = 0
            for t, k in enumerate(wkeys):
                w = self.P[k]
                entries.append((w.detach(), pw, w.shape[0], cin_each, kh * kw, co, t * cin_each))
                owners.append(k)
                co += w.shape[0]
        if not entries:
            self._pack_entries = torch.zeros(1, dtype=torch.uint8, device=self.dev)
            self._pack_block_entry = torch.zeros(0, dtype=torch.int32, device=self.dev)
            self._pack_block_start = self._pack_block_entry
            self._pack_owner_keys, self._pack_ptrs, self._pack_versions, self._pack_groups = [], (), None, {}
            return
        arr = (PackEntry * len(entries))()
        block_entry, block_start = [], []
        for i, (w, pw, cout, cin, taps, co_off, ci_off) in enumerate(entries):
            e = arr[i]
            e.w, e.fwd, e.bwd = w.data_ptr(), pw.fwd.data_ptr(), (pw.bwd.data_ptr() if pw.bwd is not None else None)
            e.cout, e.cin, e.taps = cout, cin, taps
            e.fwd_rows, e.fwd_cols = pw.fwd.shape[1], pw.fwd.shape[2]
            if pw.bwd is not None:
                e.bwd_rows, e.bwd_cols = pw.bwd.shape[1], pw.bwd.shape[2]
            e.co_off, e.ci_off = co_off, ci_off
            assert taps <= 9
            for ct in range((cout + 31) // 32):            # one block per 32 x 32 (co, ci) tile, all taps
                for it in range((cin + 31) // 32):
                    block_entry.append(i)
                    block_start.append(ct | (it << 16))
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        self._pack_entries = raw.to(self.dev)
        self._pack_block_entry = torch.tensor(block_entry, dtype=torch.int32, device=self.dev)
        self._pack_block_start = torch.tensor(block_start, dtype=torch.int32, device=self.dev)
        self._pack_owner_keys = owners
        self._pack_ptrs = tuple(self.P[k].data_ptr() for k in owners)
        self._pack_versions = None
        self._pack_groups = {g: sorted(ks) for g, ks in groups.items()}

    def refresh_weights(self, force: bool = False) -> None:
        """Re-lay-out the conv weights if any parameter changed since the last call (one launch)."""
        if getattr(self, '_pack_entries', None) is None or \
                self._pack_ptrs != tuple(self.P[k].data_ptr() for k in self._pack_owner_keys):
            self._packed.clear()
            self._build_pack_plan()
        vers = tuple(self.P[k]._version for k in self._pack_owner_keys)
        if (force or vers != self._pack_versions) and self._pack_owner_keys:
            ops._lib.call('eb200_pack_conv_weights_batched', self._pack_entries.data_ptr(),
                          self._pack_block_entry.data_ptr(), self._pack_block_start.data_ptr(),
                          self._pack_block_entry.numel(), ops._stream())
            self._pack_versions = vers
            for k in self._pack_owner_keys:
                if k in self._packed:
                    self._packed[k] = (self.P[k]._version, self._packed[k][1])
            for g, ks in self._pack_groups.items():
                self._packed[g] = (tuple(self.P[k]._version for k in ks), self._packed[g][1])

    def bn_state(self, x_raw_count: int, stats: Optional[torch.Tensor], p: str) -> ops.BNState:
        """train: finalize batch statistics (+ running update); eval: affine from the running buffers"""
        P = self.P
        if self.training:
            rm = P[p + 'running_mean'] if self.track else None
            rv = P[p + 'running_var'] if self.track else None
            if self.track:
                self._bn_touched.append(p + 'num_batches_tracked')
            return ops.bn_finalize(stats, x_raw_count, P[p + 'weight'], P[p + 'bias'], rm, rv, self.cfg.bn_eps,
                                   self.cfg.bn_momentum)
        vers = (P[p + 'weight']._version, P[p + 'bias']._version, P[p + 'running_mean']._version,
                P[p + 'running_var']._version)
        hit = self._eval_bn.get(p)
        if hit is not None and hit[0] == vers:
            return hit[1]
        with torch.no_grad():
            rstd = torch.rsqrt(P[p + 'running_var'] + self.cfg.bn_eps)
            scale = (P[p + 'weight'] * rstd).contiguous()
            shift = (P[p + 'bias'] - P[p + 'running_mean'] * scale).contiguous()
        st = ops.BNState(scale, shift, None, None, 0)
        self._eval_bn[p] = (vers, st)
        return st

    def wgrad_ws(self, pw: PackedWeight) -> Optional[torch.Tensor]:
        """persistent zeroed staging for the weight gradient of 3x3 filters (the kernel leaves it zeroed)"""
        if pw.kh * pw.kw != 9:
            return None
        return self.zeros('wgrad_ws', 9 * ops.round_up(pw.cin, 16) * ops.round_up(pw.cout, 8))

    def conv_stats(self, c: int) -> Optional[torch.Tensor]:
        return self.arena_zeros(2 * c) if self.training else None

    def bn_rep(self, c: int) -> torch.Tensor:
        return self.arena_zeros(ops.bn_rep_floats(c))

    # ------------------------------------------------------------------ layers
    def conv_bn_act(self, x: torch.Tensor, wkey: str, bnp: str, stride=(1, 1), *, relu=True, res_post=None,
                    gap=None, out=None, out_coff=0, need_dx=True, cin=None, kernel=None) -> torch.Tensor:
        """ConvNormAct (MT/model/utils.py:44-69), optionally followed by `+ res_post` (skip fusion add)."""
        if self.folding and res_post is None and out is None:
            pw, shift = self.folded(wkey, bnp)
            y = ops.conv2d(x, pw, stride, bias=shift, relu=relu, cin=cin)
            if gap is not None:
                ops.gap(y, gap)
            return y
        pw = self.weight(wkey, need_bwd=need_dx)
        cout = pw.cout
        stats = self.conv_stats(cout)
        c = ops.conv2d(x, pw, stride, stats=stats, cin=cin)
        n, h, w, _ = c.shape
        st = self.bn_state(n * h * w, stats, bnp)
        y = ops.bn_apply(c, st, relu=relu, res_post=res_post, gap=gap, out=out, out_coff=out_coff)
        if self.taps is not None:
            self._tap(bnp[:-len('norm.')] + 'c', c)
            self._tap(bnp[:-len('norm.')] + 'out', y)
        if self.training:
            sliced = out is not None and out.shape[3] != cout

            def bwd():
                dy = self.grads.pop(y)
                if dy is None:
                    return
                mode = 0 if not relu else (2 if (res_post is not None or sliced) else 1)
                dc, _ = ops.bn_backward(dy, c, st, self.P[bnp + 'weight'], rep=self.bn_rep(cout),
                                        relu_mode=mode, mask_src=y if mode == 1 else None, dy_coff=out_coff,
                                        dgamma=self.G[bnp + 'weight'], dbeta=self.G[bnp + 'bias'])
                if res_post is not None:
                    self.grads.add(res_post, dy)
                self._wgrad_side(dc, x, self.G[wkey], pw.kh, pw.kw, stride, cin=cin, ws=self.wgrad_ws(pw))
                if need_dx:
                    self.dgrad_to(x, dc, pw, stride)
            self.tape.append(bwd)
        return y

    def dgrad_to(self, x: torch.Tensor, dy: torch.Tensor, pw: PackedWeight, stride=(1, 1), *, aux=None,
                 aux_mode=None, stats=None) -> None:
        """accumulate conv data-gradient into the gradient slot of x"""
        cur = self.grads.get(x)
        shape = tuple(x.shape)
        one_by_one_strided = (stride != (1, 1)) and pw.kh == 1 and pw.kw == 1
        if cur is None:
            g = ops.conv2d_dgrad(dy, pw, shape, stride, aux=aux, aux_mode=aux_mode, stats=stats)
            self.grads.add(x, g)
        elif aux_mode is None and stats is None:
            ops.conv2d_dgrad(dy, pw, shape, stride, out=cur, accumulate_into_out=True)
        else:
            assert not one_by_one_strided
            g = ops.conv2d_dgrad(dy, pw, shape, stride, aux=aux, aux_mode=aux_mode, stats=stats)
            ops.add_inplace(cur, g)

    # ------------------------------------------------------------------ lock-step execution of sibling branches
    # The RGB / depth encoder branches and the semantic / instance decoders run the same layer shapes with different
    # weights.  Their NBt1D blocks are written once as generators that *yield* their tensor-core launches
    # ('conv' / 'dgrad' / 'wgrad' requests); `_drive` advances one generator (plain execution) or two in lock step, in
    # which case the two branches' descriptors go to eb200_conv2d_pair / eb200_conv2d_wgrad_pair: ONE launch where
    # the halo kernels allow it (wide layers: 2.16 rounds of tiles per launch become 4.3 rounds per double launch).
    def _side_launch(self, fn, keep) -> None:
        """run fn() (kernel launches) on the side stream, ordered after everything issued so far on the current one;
        `keep` = the tensors those kernels read: held until the current stream has waited for them (their memory may
        only be recycled for work that is ordered after the side kernels)"""
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            fn()
            done = torch.cuda.Event()
            done.record(self._side)
        self._side_pending.append((done, keep))
        if len(self._side_pending) > self._side_depth:
            old, _ = self._side_pending.pop(0)
            main.wait_event(old)

    def _wgrad_side(self, dy, x, *args, **kw) -> None:
        """weight gradient of a layer outside the lock-step generators (ConvNormAct, conv + bias, downsample): on the
        side stream like the NBt1D ones, so the data gradient that follows on the current stream does not wait for it"""
        if self._side is None or not self.overlap_wgrad_layers:
            ops.conv2d_wgrad(dy, x, *args, **kw)
            return
        self._side_launch(lambda: ops.conv2d_wgrad(dy, x, *args, **kw), [dy, x])

    def _side_join(self) -> None:
        """current stream waits for all side-stream work (the side stream is in-order: its newest event covers all)"""
        if self._side_pending:
            torch.cuda.current_stream().wait_event(self._side_pending[-1][0])
            self._side_pending.clear()

    def _run_parallel(self, fns):
        """run the launch closures of two sibling branches concurrently: fns[0] on the current stream, fns[1] on a side
        stream that is ordered after everything issued so far and joined again before this returns (in a recorded graph
        the two kernels become parallel branches)"""
        # Allocator discipline: every tensor is allocated under the MAIN stream (ops._stream_override only redirects
        # the launches), and the closures must not free anything they touch before the join — a block freed while a
        # kernel of the other stream still uses it could be handed to a tensor the other closure writes.  The BatchNorm
        # wrappers qualify: they allocate their outputs, launch, and return them.  (Whole layer calls with temporaries —
        # stem, max-pool, upsampling — were tried on this path and corrupted gradients at 640x480: the teacher-forced
        # parity test caught it.)
        if len(fns) != 2 or not self.parallel_siblings or self._side2 is None:
            return [fn() for fn in fns]
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        self._side2.wait_event(fork)
        r0 = fns[0]()
        ops._stream_override = self._side2.cuda_stream
        try:
            r1 = fns[1]()
        finally:
            ops._stream_override = None
        join = torch.cuda.Event()
        join.record(self._side2)
        main.wait_event(join)
        return [r0, r1]

    def _exec_requests(self, reqs):
        kind = reqs[0][0]
        if any(r[0] == 'par' for r in reqs):
            if len(reqs) == 2 and reqs[0][0] == 'par' and reqs[1][0] == 'par':
                return self._run_parallel([reqs[0][1], reqs[1][1]])
            return [r[1]() if r[0] == 'par' else self._exec_requests([r])[0] for r in reqs]
        if kind == 'wgrad' and self._side is not None and all(r[0] == 'wgrad' for r in reqs):
            self._side_launch(lambda: self._exec_requests_now(reqs), [r[1] for r in reqs])
            return [None] * len(reqs)
        return self._exec_requests_now(reqs)

    def _exec_requests_now(self, reqs):
        calls = {'conv': ops.conv2d, 'dgrad': ops.conv2d_dgrad, 'wgrad': ops.conv2d_wgrad}
        kind = reqs[0][0]
        if len(reqs) != 2 or not self.pair_siblings or reqs[1][0] != kind:
            return [calls[r[0]](*r[1], **r[2]) for r in reqs]
        outs, descs = [], []
        attr = '_defer_wgrad' if kind == 'wgrad' else '_defer_conv'
        try:
            for r in reqs:
                setattr(ops, attr, [])
                outs.append(calls[kind](*r[1], **r[2]))
                descs.append(getattr(ops, attr))
        finally:
            setattr(ops, attr, None)
        (ops.launch_wgrad_descs if kind == 'wgrad' else ops.launch_conv_descs)(descs[0], descs[1])
        return outs

    def _drive(self, gens):
        """run 1 or 2 generators to completion (in lock step while both are alive); returns their return values"""
        rets = [None] * len(gens)
        alive, reqs = [], []
        for i, g in enumerate(gens):
            try:
                reqs.append(next(g))
                alive.append(i)
            except StopIteration as e:
                rets[i] = e.value
        while alive:
            outs = self._exec_requests(reqs)
            nalive, nreqs = [], []
            for i, o in zip(alive, outs):
                try:
                    nreqs.append(gens[i].send(o))
                    nalive.append(i)
                except StopIteration as e:
                    rets[i] = e.value
            alive, reqs = nalive, nreqs
        return rets

    def nbt1d(self, x: torch.Tensor, p: str, stride: int, gap=None) -> torch.Tensor:
        """NonBottleneck1D.forward (MT/model/block.py:201-221)."""
        out, bwd_gen = self._drive([self._nbt1d_gen(x, p, stride, gap)])[0]
        if bwd_gen is not None:
            def bwd():
                g = bwd_gen()
                if g is not None:
                    self._drive([g])
            self.tape.append(bwd)
        return out

    def nbt1d_pair(self, xs, ps, stride: int, gaps=(None, None)):
        """two NBt1D blocks of identical shape (sibling branches) in lock step"""
        (o0, b0), (o1, b1) = self._drive([self._nbt1d_gen(xs[0], ps[0], stride, gaps[0]),
                                          self._nbt1d_gen(xs[1], ps[1], stride, gaps[1])])
        if b0 is not None:
            def bwd():
                g0, g1 = b0(), b1()
                live = [g for g in (g0, g1) if g is not None]
                if live:
                    self._drive(live)
            self.tape.append(bwd)
        return o0, o1

    def _nbt1d_gen(self, x: torch.Tensor, p: str, stride: int, gap=None):
        P, G = self.P, self.G
        has_ds = (p + 'downsample.0.weight') in P
        w11 = self.weight(p + 'conv1_1.weight')
        w12 = self.weight(p + 'conv1_2.weight')
        w21 = self.weight(p + 'conv2_1.weight')
        w22 = self.weight(p + 'conv2_2.weight')
        C = w11.cout
        s1, s2 = (stride, 1), (1, stride)
        if self.folding:
            # inference: norm1 / norm2 / the downsample BatchNorm live in the weights of the conv in front of them
            w12f, sh1 = self.folded(p + 'conv1_2.weight', p + 'norm1.')
            w22f, sh2 = self.folded(p + 'conv2_2.weight', p + 'norm2.')
            a11 = yield ('conv', (x, w11, s1), dict(bias=P[p + 'conv1_1.bias'], relu=True))
            a12 = yield ('conv', (a11, w12f, s2), dict(bias=sh1, relu=True))
            a21 = yield ('conv', (a12, w21), dict(bias=P[p + 'conv2_1.bias'], relu=True))
            if has_ds:
                wdsf, shd = self.folded(p + 'downsample.0.weight', p + 'downsample.1.')
                idt = yield ('conv', (x, wdsf, (stride, stride)), dict(bias=shd))
            else:
                idt = x
            out = yield ('conv', (a21, w22f), dict(bias=sh2, relu=True, aux=idt, aux_mode='add'))
            if gap is not None:
                ops.gap(out, gap)
            return out, None
        a11 = yield ('conv', (x, w11, s1), dict(bias=P[p + 'conv1_1.bias'], relu=True))
        st1_stats = self.conv_stats(C)
        c12 = yield ('conv', (a11, w12, s2), dict(stats=st1_stats))
        n, h, w, _ = c12.shape
        count = n * h * w
        st1 = self.bn_state(count, st1_stats, p + 'norm1.')
        a12 = yield ('par', lambda: ops.bn_apply(c12, st1, relu=True))
        a21 = yield ('conv', (a12, w21), dict(bias=P[p + 'conv2_1.bias'], relu=True))
        st2_stats = self.conv_stats(C)
        c22 = yield ('conv', (a21, w22), dict(stats=st2_stats))
        st2 = self.bn_state(count, st2_stats, p + 'norm2.')
        if has_ds:
            wds = self.weight(p + 'downsample.0.weight')
            ds_stats = self.conv_stats(C)
            cds = yield ('conv', (x, wds, (stride, stride)), dict(stats=ds_stats))
            std = self.bn_state(count, ds_stats, p + 'downsample.1.')
            idt = yield ('par', lambda: ops.bn_apply(cds, std, relu=False))
        else:
            idt = x
        drop = self.masks.get(p) if self.training else None
        out = yield ('par', lambda: ops.bn_apply(c22, st2, relu=True, drop=drop, res_pre=idt, gap=gap))
        if self.taps is not None:
            for nm, t in (('a11', a11), ('c12', c12), ('a12', a12), ('a21', a21), ('c22', c22), ('out', out)):
                self.taps[p + nm] = t
            if has_ds:
                self.taps[p + 'cds'], self.taps[p + 'idt'] = cds, idt
        if not self.training:
            return out, None

        def bwd_gen():
            dout = self.grads.pop(out)
            if dout is None:
                return None
            return bwd_body(dout)

        def bwd_body(dout):
            rep2 = self.bn_rep(C)
            dc22, dz = yield ('par', lambda: ops.bn_backward(
                dout, c22, st2, P[p + 'norm2.weight'], rep=rep2, relu_mode=1, mask_src=out, drop=drop, want_dres=True,
                dgamma=G[p + 'norm2.weight'], dbeta=G[p + 'norm2.bias']))
            yield ('wgrad', (dc22, a21, G[p + 'conv2_2.weight'], 1, 3), {})
            # bias gradients: the data-gradient epilogue sums its stored values straight into the bias .grad
            dc21 = yield ('dgrad', (dc22, w22, tuple(a21.shape)),
                          dict(aux=a21, aux_mode='mask', stats=G[p + 'conv2_1.bias'], stats_sum_only=True))
            yield ('wgrad', (dc21, a12, G[p + 'conv2_1.weight'], 3, 1), {})
            if self.fuse_bn_bwd:
                # conv2_1's data gradient also does the ReLU mask and the two sums of norm1's backward in its epilogue
                raw = self.arena_zeros(2 * C)
                g12 = yield ('dgrad', (dc21, w21, tuple(a12.shape)),
                             dict(aux=c12, aux_mode='mask', stats=raw, bn_bwd=(st1.scale, st1.shift)))
                dc12 = yield ('par', lambda: ops.bn_bwd_apply_raw(g12, c12, st1, P[p + 'norm1.weight'], raw,
                                                                  G[p + 'norm1.weight'], G[p + 'norm1.bias']))
            else:
                da12 = yield ('dgrad', (dc21, w21, tuple(a12.shape)), {})
                dc12, _ = ops.bn_backward(da12, c12, st1, P[p + 'norm1.weight'], rep=self.bn_rep(C), relu_mode=1,
                                          mask_src=a12, dgamma=G[p + 'norm1.weight'], dbeta=G[p + 'norm1.bias'])
            yield ('wgrad', (dc12, a11, G[p + 'conv1_2.weight'], 1, 3, s2), {})
            dc11 = yield ('dgrad', (dc12, w12, tuple(a11.shape), s2),
                          dict(aux=a11, aux_mode='mask', stats=G[p + 'conv1_1.bias'], stats_sum_only=True))
            yield ('wgrad', (dc11, x, G[p + 'conv1_1.weight'], 3, 1, s1), {})
            if has_ds:
                self.dgrad_to(x, dc11, w11, s1)
                dcds, _ = ops.bn_backward(dz, cds, std, P[p + 'downsample.1.weight'], rep=self.bn_rep(C), relu_mode=0,
                                          dgamma=G[p + 'downsample.1.weight'], dbeta=G[p + 'downsample.1.bias'])
                self._wgrad_side(dcds, x, G[p + 'downsample.0.weight'], 1, 1, (stride, stride))
                self.dgrad_to(x, dcds, wds, (stride, stride))
            elif self.grads.has(x):
                self.dgrad_to(x, dc11, w11, s1)
                ops.add_inplace(self.grads.get(x), dz)
            else:
                g = yield ('dgrad', (dc11, w11, tuple(x.shape), s1), dict(aux=dz, aux_mode='add'))
                self.grads.add(x, g)
        return out, bwd_gen

    def upsample(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """Upsampling 'learned-3x3-zeropad' (MT/model/upsampling.py:85-96)."""
        w, b = self.P[p + 'conv.weight'], self.P[p + 'conv.bias']
        y = self._tap(p + 'out', ops.upsample_dw_fwd(x, w, b))
        if self.training:
            def bwd():
                dy = self.grads.pop(y)
                if dy is None:
                    return
                dx = ops.upsample_dw_bwd(dy, x, w, self.G[p + 'conv.weight'], self.G[p + 'conv.bias'])
                self.grads.add(x, dx)
            self.tape.append(bwd)
        return y

    def upsample_output(self, x: torch.Tensor, p: str, outs: List, slot: List) -> None:
        """the LAST Upsampling of a head (MT/model/upsampling.py:85-96) fused with the fp32 NCHW output boundary: the
        upsampled bf16 map is never stored, the output gradient is read once for dx, dW and db"""
        w, b = self.P[p + 'conv.weight'], self.P[p + 'conv.bias']
        idx = self._emit_outputs(outs, 1, lambda: ops.upsample_dw_fwd_nchw(x, w, b))
        if self.taps is not None:      # the oracle's storage point: the values are bf16-representable by construction
            self.taps[p + 'out'] = outs[idx].permute(0, 2, 3, 1).to(BF16)
        if self.training:
            dxb = self._boundary_dx(x)

            def bwd():
                self._boundary_backward(slot, [idx], x, lambda g, out: ops.upsample_dw_bwd_nchw(
                    g, x, w, self.G[p + 'conv.weight'], self.G[p + 'conv.bias'], out=out), dxb)
            self.tape.append(bwd)

    # ------------------------------------------------------------------ encoder
    def stem(self, inp: torch.Tensor, bp: str, gap) -> torch.Tensor:
        """conv 7x7 s2 + BN + ReLU (MT/model/backbone/resnet.py:64-66) as im2col + 1x1 tensor-core GEMM"""
        cols = ops.im2col_stem(inp.contiguous())
        cin = inp.shape[1] * 49
        if self.folding:
            pw, shift = self.folded(bp + 'conv1.weight', bp + 'norm1.', stem=True)
            y = ops.conv2d(cols, pw, bias=shift, relu=True)
            if gap is not None:
                ops.gap(y, gap)
            return y
        pw = self.weight_stem(bp + 'conv1.weight')
        stats = self.conv_stats(64)
        c = ops.conv2d(cols, pw, stats=stats)
        n, h, w, _ = c.shape
        st = self.bn_state(n * h * w, stats, bp + 'norm1.')
        y = ops.bn_apply(c, st, relu=True, gap=gap)
        self._tap(bp + 'conv1.c', c)
        self._tap(bp + 'stem.out', y)
        if self.training:
            def bwd():
                dy = self.grads.pop(y)
                if dy is None:
                    return
                dc, _ = ops.bn_backward(dy, c, st, self.P[bp + 'norm1.weight'], rep=self.bn_rep(64),
                                        relu_mode=1, mask_src=y, dgamma=self.G[bp + 'norm1.weight'],
                                        dbeta=self.G[bp + 'norm1.bias'])
                g = self.G[bp + 'conv1.weight']
                ops.conv2d_wgrad(dc, cols, g.view(64, cin, 1, 1), 1, 1, cin=cin)
            self.tape.append(bwd)
        return y

    def weight_stem(self, key: str) -> PackedWeight:
        w = self.P[key]
        hit = self._packed.get(key)
        if hit is not None and hit[0] == w._version:
            return hit[1]
        w2 = w.detach().reshape(64, -1, 1, 1)
        pw = ops.pack_weight(w2, need_bwd=False, out=hit[1] if hit is not None else None)
        self._packed[key] = (w._version, pw)
        return pw

    def maxpool(self, x: torch.Tensor) -> torch.Tensor:
        y, idx = ops.maxpool_fwd(x)
        if self.training:
            def bwd():
                dy = self.grads.pop(y)
                if dy is None:
                    return
                self.grads.add(x, ops.maxpool_bwd(dy, idx, tuple(x.shape)))
            self.tape.append(bwd)
        return y

    def se_fuse(self, xr: torch.Tensor, xd: torch.Tensor, gr: torch.Tensor, gd: torch.Tensor, p: str):
        """EncoderRGBDFusionWeightedAdd 'se-add-uni-rgb' (MT/model/encoder_fusion.py:63-90)"""
        P, G = self.P, self.G
        n, h, w, c = xr.shape
        hw = h * w
        pr, pd = p + 'weighting_rgb.layers.', p + 'weighting_depth.layers.'
        sr = ops.se_mlp_fwd(gr, hw, P[pr + '0.weight'], P[pr + '0.bias'], P[pr + '2.weight'], P[pr + '2.bias'])
        sd = ops.se_mlp_fwd(gd, hw, P[pd + '0.weight'], P[pd + '0.bias'], P[pd + '2.weight'], P[pd + '2.bias'])
        fused = self._tap(p + 'out', ops.se_fuse_fwd(xr, xd, sr.wgt, sd.wgt))
        if self.training:
            def bwd():
                df = self.grads.pop(fused)
                if df is None:
                    return
                dwr = self.zeros(('dwgt', 'r'), n * c).view(n, c)
                dwd = self.zeros(('dwgt', 'd'), n * c).view(n, c)
                ops.se_fuse_bwd_reduce(df, xr, xd, dwr, dwd)
                dmr = ops.se_mlp_bwd(dwr, sr, hw, P[pr + '0.weight'], P[pr + '2.weight'], G[pr + '0.weight'],
                                     G[pr + '0.bias'], G[pr + '2.weight'], G[pr + '2.bias'])
                dmd = ops.se_mlp_bwd(dwd, sd, hw, P[pd + '0.weight'], P[pd + '2.weight'], G[pd + '0.weight'],
                                     G[pd + '0.bias'], G[pd + '2.weight'], G[pd + '2.bias'])
                prev = self.grads.pop(xd)
                da, db = ops.se_fuse_bwd_apply(df, sr.wgt, sd.wgt, dmr, dmd, prev)
                self.grads.add(xr, da)
                self.grads.add(xd, db)
            self.tape.append(bwd)
        return fused

    def encoder(self, rgb: Optional[torch.Tensor], depth: Optional[torch.Tensor]):
        cfg = self.cfg
        dual = len(cfg.modalities) == 2
        x: Dict[str, torch.Tensor] = {}
        skips: Dict[int, torch.Tensor] = {}
        n = (rgb if rgb is not None else depth).shape[0]
        for stage in range(5):
            c = STAGE_CHANNELS[stage]
            gaps = {}
            for m in cfg.modalities:
                gaps[m] = self.zeros(('gap', m), n * c).view(n, c) if dual else None
            for m in cfg.modalities:
                bp = cfg.backbone_prefix(m)
                g = gaps[m]
                if stage == 0:
                    x[m] = self.stem(rgb if m == 'rgb' else depth, bp, g)
                    continue
                t = x[m]
                if stage == 1:
                    t = self.maxpool(t)
                nb = cfg.layers[stage - 1]
                if dual:
                    x[m] = t
                    continue                       # the blocks of both modalities run in lock step below
                for b in range(nb):
                    t = self.nbt1d(t, f'{bp}layer{stage}.{b}.', 2 if (b == 0 and stage > 1) else 1,
                                   gap=g if b == nb - 1 else None)
                x[m] = t
            if dual and stage > 0:
                ma, mb = cfg.modalities
                nb = cfg.layers[stage - 1]
                for b in range(nb):
                    last = b == nb - 1
                    x[ma], x[mb] = self.nbt1d_pair(
                        (x[ma], x[mb]),
                        (f'{cfg.backbone_prefix(ma)}layer{stage}.{b}.', f'{cfg.backbone_prefix(mb)}layer{stage}.{b}.'),
                        2 if (b == 0 and stage > 1) else 1,
                        (gaps[ma] if last else None, gaps[mb] if last else None))
            if dual:
                x['rgb'] = self.se_fuse(x['rgb'], x['depth'], gaps['rgb'], gaps['depth'], f'encoder.fusions.{stage}.')
            key = 'rgb' if 'rgb' in x else 'depth'
            if self.taps is not None:
                self.taps[f'encoder.stage{stage}.{key}'] = x[key]
            if stage in (1, 2, 3):
                skips[4 * 2 ** (stage - 1)] = x[key]
            if stage == 2 and self.training:
                self.tape.append(self._encoder_mid_boundary)
        return x['rgb' if 'rgb' in x else 'depth'], skips

    # ------------------------------------------------------------------ context module
    def ppm(self, x: torch.Tensor):
        """PyramidPoolingModule.forward (MT/model/context_module/ppm.py:57-78)"""
        cfg = self.cfg
        n, h, w, c = x.shape
        cred = c // len(cfg.ppm_bins)
        ctot = c + cred * len(cfg.ppm_bins)
        cat = torch.empty(n, h, w, ctot, dtype=BF16, device=x.device)
        ops.copy_channels(x, cat, c, 0, 0, False)
        feats = []
        for i, b in enumerate(cfg.ppm_bins):
            pooled = self._tap(f'context_module.features.{i}.pooled', ops.adaptive_pool_fwd(x, b))
            f = self.conv_bn_act(pooled, f'context_module.features.{i}.1.conv.weight',
                                 f'context_module.features.{i}.1.norm.')
            feats.append(f)
            ops.bilinear_fwd(f, cat, c + i * cred)
            self._tap(f'context_module.features.{i}.up', cat[..., c + i * cred:c + (i + 1) * cred])
            if self.training:
                def bwd_pool(pooled=pooled):
                    dp = self.grads.pop(pooled)
                    if dp is None:
                        return
                    cur = self.grads.get(x)
                    if cur is None:
                        cur = torch.empty_like(x)
                        ops.adaptive_pool_bwd(dp, cur, False)
                        self.grads.add(x, cur)
                    else:
                        ops.adaptive_pool_bwd(dp, cur, True)
                # executes after the conv_bn_act backward (tape runs in reverse): insert *before* it
                self.tape.insert(len(self.tape) - 1, bwd_pool)
        y = self.conv_bn_act(cat, 'context_module.final_conv.conv.weight', 'context_module.final_conv.norm.')
        if self.training:
            # backward of the concat: runs right after final_conv's backward (i.e. placed before it on the tape)
            def bwd_cat():
                dcat = self.grads.pop(cat)
                if dcat is None:
                    return
                cur = self.grads.get(x)
                if cur is None:
                    cur = torch.empty_like(x)
                    ops.copy_channels(dcat, cur, c, 0, 0, False)
                    self.grads.add(x, cur)
                else:
                    ops.copy_channels(dcat, cur, c, 0, 0, True)
                for i, f in enumerate(feats):
                    self.grads.add(f, ops.bilinear_bwd(dcat, c + i * cred, tuple(f.shape)))
            self.tape.insert(len(self.tape) - 1, bwd_cat)
        return y, feats

    # ------------------------------------------------------------------ decoders
    def decoder_modules(self, x: torch.Tensor, skips: Dict[int, torch.Tensor], p: str):
        """DenseDecoderBase._forward_decoder_modules (MT/model/decoder/dense_base.py:229-259)"""
        sides = []
        for i in range(len(self.cfg.decoder_n_channels)):
            mp = f'{p}decoder_modules.{i}.'
            x = self.conv_bn_act(x, mp + 'conv.conv.weight', mp + 'conv.norm.')
            for b in range(self.cfg.decoder_n_blocks):
                x = self.nbt1d(x, f'{mp}blocks.{b}.', 1)
            sides.append(x if self.training else None)
            up = self.upsample(x, mp + 'upsample.')
            fp = f'{p}fusions.{i}.layer.'
            x = self.conv_bn_act(skips[16 // 2 ** i], fp + 'conv.weight', fp + 'norm.', res_post=up)
            if self.taps is not None:
                self.taps[mp + 'fused'] = x
        return x, sides

    def plain_conv(self, x: torch.Tensor, wkey: str, bkey: str) -> torch.Tensor:
        """nn.Conv2d with bias, no activation (task-head convs, MT/model/decoder/dense_utils.py:19-24)"""
        pw = self.weight(wkey)
        cpad = ops.round_up(pw.cout, 8)
        y = ops.conv2d(x, pw, bias=self.P[bkey], out=torch.empty(*x.shape[:3], cpad, dtype=BF16, device=x.device)
                       if cpad != pw.cout else None)
        self._tap(wkey[:-len('weight')] + 'out', y)
        if self.training:
            def bwd():
                dy = self.grads.pop(y)
                if dy is None:
                    return
                ops.colsum(dy, self.G[bkey], c=pw.cout)
                self._wgrad_side(dy, x, self.G[wkey], pw.kh, pw.kw, dy_c=pw.cout, ws=self.wgrad_ws(pw))
                self.dgrad_to(x, dy, pw)
            self.tape.append(bwd)
        return y

    def output_nchw(self, x: torch.Tensor, creal: int, outs: List, slot: List) -> None:
        idx = self._emit_outputs(outs, 1, lambda: ops.nhwc_to_nchw(x, creal))
        if self.training:
            dxb = self._boundary_dx(x)

            def bwd():
                self._boundary_backward(slot, [idx], x, lambda g, out: ops.nchw_grad_to_nhwc(g, tuple(x.shape), creal,
                                                                                          out=out), dxb)
            self.tape.append(bwd)

    def decoder_modules_pair(self, x: torch.Tensor, skips: Dict[int, torch.Tensor], ps):
        """the decoder modules of the semantic and the instance decoder (same shapes, different weights) with their
        NBt1D blocks in lock step (one launch per pair of tensor-core kernels where the halo kernels allow it)"""
        xs = [x, x]
        sides = ([], [])
        for i in range(len(self.cfg.decoder_n_channels)):
            mps = [f'{p}decoder_modules.{i}.' for p in ps]
            xs = [self.conv_bn_act(xs[k], mps[k] + 'conv.conv.weight', mps[k] + 'conv.norm.') for k in range(2)]
            for b in range(self.cfg.decoder_n_blocks):
                xs = list(self.nbt1d_pair(xs, [f'{mp}blocks.{b}.' for mp in mps], 1))
            for k in range(2):
                sides[k].append(xs[k] if self.training else None)
            for k in range(2):
                up = self.upsample(xs[k], mps[k] + 'upsample.')
                fp = f'{ps[k]}fusions.{i}.layer.'
                xs[k] = self.conv_bn_act(skips[16 // 2 ** i], fp + 'conv.weight', fp + 'norm.', res_post=up)
                if self.taps is not None:
                    self.taps[mps[k] + 'fused'] = xs[k]
        return (xs[0], sides[0]), (xs[1], sides[1])

    def semantic_decoder(self, x, skips, p: str, outs: List, slot: List, modules=None):
        """SemanticDecoder (MT/model/decoder/semantic.py:26-83)"""
        x, sides = modules if modules is not None else self.decoder_modules(x, skips, p)
        y = self.plain_conv(x, p + '_task_head.conv.weight', p + '_task_head.conv.bias')
        y = self.upsample(y, p + '_task_head.upsample_0.')
        if self.fuse_output_upsample:
            self.upsample_output(y, p + '_task_head.upsample_1.', outs, slot)
        else:
            y = self.upsample(y, p + '_task_head.upsample_1.')
            self.output_nchw(y, self.cfg.semantic_n_classes, outs, slot)
        for i, s in enumerate(sides):
            if s is not None:
                hp = p + f'_side_output_heads.{i}.conv.'
                self.output_nchw(self.plain_conv(s, hp + 'weight', hp + 'bias'), self.cfg.semantic_n_classes, outs,
                                 slot)

    def instance_head(self, x: torch.Tensor, p: str, k: int, n_up: int, outs: List, slot: List) -> None:
        """InstanceHead.forward (MT/model/decoder/instance.py:95-121)"""
        P, G = self.P, self.G
        nt = 3 if self.cfg.with_orientation else 2
        couts = [1, 2, 2][:nt]
        s = self.conv_bn_act(x, p + 'shared_conv.conv.weight', p + 'shared_conv.norm.')
        wkeys = [p + f'task_convs.{t}.weight' for t in range(nt)]
        bkeys = [p + f'task_convs.{t}.bias' for t in range(nt)]
        pw = self.block_diag_weight(p + 'task_convs', wkeys, couts, 32)
        bias = torch.zeros(8, dtype=torch.float32, device=self.dev)
        torch.cat([P[b].detach() for b in bkeys], out=bias[:sum(couts)])
        t8 = self._tap(p + 'task_convs.out', ops.conv2d(s, pw, bias=bias))
        if self.training:
            def bwd_task():
                dt = self.grads.pop(t8)
                if dt is None:
                    return
                db = torch.zeros(8, dtype=torch.float32, device=self.dev)
                ops.colsum(dt, db)
                taps = k * k
                dwd = torch.zeros(8, 32 * nt, taps, dtype=torch.float32, device=self.dev)
                ops.conv2d_wgrad(dt, s, dwd, k, k, dw_strides=(32 * nt * taps, taps, 1))
                co = 0
                for t, c in enumerate(couts):
                    G[bkeys[t]] += db[co:co + c]
                    G[wkeys[t]] += dwd[co:co + c, 32 * t:32 * t + 32].reshape(c, 32, k, k)
                    co += c
                self.dgrad_to(s, dt, pw)
            self.tape.append(bwd_task)
        y = t8
        ups = []
        for u in range(n_up):
            up_p = p + f'upsampling.{u}.'
            y = self.upsample(y, up_p)
            ups.append(y)
        if self.taps is not None:
            self.taps[p + 'shared'] = s
            self.taps[p + 'task8'] = t8
            self.taps[p + 'pre_act'] = y
        idx = self._emit_outputs(outs, nt, lambda: ops.instance_outputs(y, nt == 3))
        if self.training:
            dxb = self._boundary_dx(y)

            def bwd_out():
                def launch(*gs, out):
                    gs = list(gs) + [None] * (3 - nt)
                    return ops.instance_outputs_bwd(gs[0], gs[1], gs[2], y, out=out)
                self._boundary_backward(slot, [idx + j for j in range(nt)], y, launch, dxb)
            self.tape.append(bwd_out)

    def instance_decoder(self, x, skips, p: str, outs: List, slot: List, modules=None):
        x, sides = modules if modules is not None else self.decoder_modules(x, skips, p)
        self.instance_head(x, p + '_task_head.', 3, 2, outs, slot)
        for i, s in enumerate(sides):
            if s is not None:
                self.instance_head(s, p + f'_side_output_heads.{i}.', 1, 0, outs, slot)

    def scene_head(self, feat: torch.Tensor, p: str, outs: List, slot: List):
        """SceneClassificationDecoder (MT/model/decoder/scene.py:32-65): Linear on the PPM bin-1 feature"""
        w, b = self.P[p + '_task_head.weight'], self.P[p + '_task_head.bias']
        idx = self._emit_outputs(outs, 1, lambda: ops.linear_fwd(feat, w, b))
        if self.training:
            dxb = self._boundary_dx(feat)

            def bwd():
                self._boundary_backward(slot, [idx], feat, lambda g, out: ops.linear_bwd(
                    g, feat, w, self.G[p + '_task_head.weight'], self.G[p + '_task_head.bias'], out=out), dxb)
            self.tape.append(bwd)

    # ------------------------------------------------------------------ top level
    def make_dropout_masks(self, n: int) -> Dict[str, torch.Tensor]:
        """Dropout2d keep masks scaled by 1/(1-p), one [N, C] fp32 tensor per block (torch's CUDA generator)"""
        masks = {}
        sites = self.cfg.dropout_sites()
        for p_val in sorted({s[2] for s in sites}):
            if p_val <= 0:
                continue
            group = [s for s in sites if s[2] == p_val]
            total = sum(n * c for _, c, _ in group)
            keep = (torch.rand(total, device=self.dev) >= p_val).float() * (1.0 / (1.0 - p_val))
            off = 0
            for prefix, c, _ in group:
                masks[prefix] = keep[off:off + n * c].view(n, c)
                off += n * c
        return masks

    def begin(self, training: bool, track_running_stats: bool = True,
              dropout_masks: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """start a new forward program (also used by the block-level tests to drive single layers)"""
        self.training, self.track = training, track_running_stats
        self.tape, self.grads = [], _Grads()
        self._bn_touched = []
        self._arena_reset('bwd')
        self._arena_reset('fwd')
        self.refresh_weights(force=self.force_repack)
        self.masks = dropout_masks or {}
        if self.parallel_siblings and self._side2 is None and not torch.cuda.is_current_stream_capturing():
            self._side2 = torch.cuda.Stream(device=self.dev)

    def param_grad_floats(self) -> int:
        return sum((self.P[k].numel() + 3) // 4 * 4 for k in self.grad_keys)

    def alloc_param_grads(self, flat: Optional[torch.Tensor] = None, zero: bool = True) -> torch.Tensor:
        sizes = [self.P[k].numel() for k in self.grad_keys]
        padded = [(s + 3) // 4 * 4 for s in sizes]    # every slice starts 16-byte aligned (vector reductions)
        if flat is None:
            flat = torch.zeros(sum(padded), dtype=torch.float32, device=self.dev)
        else:                                         # caller-owned buffer (graph replay: static, outside the graph pool)
            assert flat.numel() == sum(padded)
            if zero:                                  # (deferred output boundary: the caller zeroes it before the eager
                flat.zero_()                          #  boundary launches, which already accumulate into it)
        self.flat_grad = flat   # one contiguous fp32 buffer: the unit of the data-parallel all-reduce
        self.G.clear()   # same dict object: the tape closures hold a reference to it
        # Layout = the order in which backward FINISHES the gradients, last first, so that each data-parallel bucket
        # is one contiguous range: [encoder stem + stages 1-2 | encoder stages 3-4 | context module + decoders].
        # (`grad_keys` / `grad_slices` stay in state_dict order: optimizers and `.grad` views go by key.)
        off = 0
        offsets = {}
        for section in (0, 1, 2):
            for k, sp in zip(self.grad_keys, padded):
                if self._grad_section(k) == section:
                    offsets[k] = off
                    off += sp
            if section == 0:
                self._enc_lo_end = off
            elif section == 1:
                self._enc_end = off
        self.grad_slices = []    # (offset, numel, shape) per grad key: lets callers make fresh views of `flat`
        for k, s in zip(self.grad_keys, sizes):
            o = offsets[k]
            self.grad_slices.append((o, s, tuple(self.P[k].shape)))
            self.G[k] = flat[o:o + s].view(self.P[k].shape)
        return flat

    @staticmethod
    def _grad_section(key: str) -> int:
        """0: encoder parameters whose gradient is final only when backward ends (stem, stages 1-2, fusions 0-2);
        1: encoder stages 3-4 and their fusions (final at the mid-encoder marker); 2: everything behind the encoder"""
        if not key.startswith('encoder.'):
            return 2
        late = any(f'.layer{st}.' in key or key.startswith(f'encoder.fusions.{st}.') for st in (3, 4))
        return 1 if late else 0

    def run_tape(self, stop_at_marker: bool = False) -> Optional[str]:
        """run the backward tape (last recorded first).  With stop_at_marker the run stops right after a gradient-ready
        marker — 'enc': every decoder / context-module gradient is final, 'mid': so are those of encoder stages 3-4 —
        and returns its name; the next call continues.  This is where the data-parallel path splits its CUDA graphs.
        Returns None when the tape is finished."""
        self._join_at_markers = stop_at_marker or self.on_grads_ready is not None
        while self.tape:
            fn = self.tape.pop()
            fn()
            if stop_at_marker:
                if fn == self._encoder_boundary:
                    return 'enc'
                if fn == self._encoder_mid_boundary:
                    return 'mid'
        self._side_join()
        return None

    def ready_range(self, marker: Optional[str]):
        """the range of the flat gradient buffer that became final when run_tape() returned `marker`"""
        if marker == 'enc':
            return self._enc_end, self.flat_grad.numel()
        if marker == 'mid':
            return self._enc_lo_end, self._enc_end
        return 0, (self._enc_lo_end if self._mid_done else self._enc_end)

    def forward(self, rgb: Optional[torch.Tensor], depth: Optional[torch.Tensor], training: bool,
                track_running_stats: bool = True, dropout_masks: Optional[Dict[str, torch.Tensor]] = None):
        """Returns the flat list of fp32 NCHW output tensors in the reference's depth-first output order
        (SURVEY.md App. A): semantic main, instance main (center, offset[, orientation]), semantic side outputs,
        instance side outputs, scene."""
        cfg = self.cfg
        n = (rgb if rgb is not None else depth).shape[0]
        self.begin(training, track_running_stats,
                   dropout_masks if (dropout_masks is not None or not training) else self.make_dropout_masks(n))
        enc, skips = self.encoder(rgb, depth)
        if training:
            # backward reaches this marker when every decoder / context-module gradient is final
            self.tape.append(self._encoder_boundary)
        ctx, feats = self.ppm(enc)
        if self.taps is not None:
            self.taps['context_module.out'] = ctx
        pre = cfg.decoder_prefixes
        self.grad_out_slots: Dict[str, List] = {}
        res: Dict[str, List[torch.Tensor]] = {}
        paired = {}
        if 'semantic' in pre and 'instance' in pre and self.pair_siblings:
            ms, mi = self.decoder_modules_pair(ctx, skips, (pre['semantic'], pre['instance']))
            paired = {'semantic': ms, 'instance': mi}
        for task, fn in (('semantic', self.semantic_decoder), ('instance', self.instance_decoder)):
            if task in pre:
                outs: List[torch.Tensor] = []
                slot: List = []
                self.grad_out_slots[task] = slot
                fn(ctx, skips, pre[task], outs, slot, modules=paired.get(task))
                res[task] = outs
        if 'scene' in pre:
            outs, slot = [], []
            self.grad_out_slots['scene'] = slot
            self.scene_head(feats[0], pre['scene'], outs, slot)
            res['scene'] = outs
        if training and self.track and self._bn_touched:
            torch._foreach_add_([self.P[k] for k in self._bn_touched], 1)
        return res

    def _encoder_boundary(self) -> None:
        self._side_join()       # every decoder / context-module weight gradient is final on the current stream
        if self.on_grads_ready is not None:
            self.on_grads_ready(self.flat_grad, self._enc_end, self.flat_grad.numel())

    def _encoder_mid_boundary(self) -> None:
        """recorded after encoder stage 2: backward reaches it when the gradients of stages 3-4 (94 % of the encoder's
        parameters) are final — their all-reduce hides behind the backward of stages 2, 1 and the stems"""
        if not self._join_at_markers:
            return              # single GPU, one graph: nobody consumes the bucket early, no stream join needed
        self._side_join()
        self._mid_done = True
        if self.on_grads_ready is not None:
            self.on_grads_ready(self.flat_grad, self._enc_lo_end, self._enc_end)

    def begin_backward(self, grad_outputs: Dict[str, List[Optional[torch.Tensor]]],
                       flat: Optional[torch.Tensor] = None, zero: bool = True) -> torch.Tensor:
        """first part of backward(): gradient buffer, scratch arena, output gradients in place; then run_tape()"""
        flat = self.alloc_param_grads(flat, zero)
        self._mid_done = False
        self._arena_reset('bwd')
        if self.overlap_wgrad and self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        if not self.overlap_wgrad:
            self._side = None
        for task, slot in self.grad_out_slots.items():
            slot.clear()
            slot.extend(grad_outputs.get(task, []))
        return flat

    def backward(self, grad_outputs: Dict[str, List[Optional[torch.Tensor]]],
                 flat: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """grad_outputs[task][i] = dL/d(output i of that task) (NCHW fp32) or None.  Returns fp32 parameter grads."""
        flat = self.begin_backward(grad_outputs, flat)
        self.run_tape()
        if self.on_grads_ready is not None:
            self.on_grads_ready(flat, *self.ready_range(None))
        self.grads = None
        return self.G
