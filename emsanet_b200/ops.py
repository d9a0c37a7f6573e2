"""Host-side wrappers around the C ABI: torch tensors in, kernels enqueued on torch's current stream.

torch is used for device memory and streams only; every op below is one of our CUDA kernels.
Activations are NHWC bf16 tensors of shape [N, H, W, C].
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import AUX_ADD, AUX_MASK, BIAS, BN_BWD, RELU, STATS, STATS_SUM_ONLY, ConvDesc, View, WgradDesc  # noqa: F401

BF16 = torch.bfloat16


# Launch-stream override (engine._run_parallel): kernels of the second sibling branch are enqueued on a side stream while
# every allocation still happens under torch's current stream — the caching allocator ties a block to the stream that
# was current when it was allocated, and these tensors are consumed on the main stream afterwards.
_stream_override: Optional[int] = None


def _stream() -> int:
    return _stream_override if _stream_override is not None else torch.cuda.current_stream().cuda_stream


# Lock-step execution of two sibling branches (engine._drive): while a list is installed here, eb200_conv2d /
# eb200_conv2d_wgrad descriptors are collected instead of launched, so that the driver can hand the two branches'
# descriptors to eb200_conv2d_pair / eb200_conv2d_wgrad_pair (ONE launch for both).
_defer_conv: Optional[list] = None
_defer_wgrad: Optional[list] = None


def launch_conv_descs(a: list, b: Optional[list] = None) -> None:
    """launch collected conv descriptors: pairwise with `b` where both lists line up, else one by one"""
    if b is not None and len(a) == len(b):
        for da, db in zip(a, b):
            _lib.call('eb200_conv2d_pair', C.byref(da), C.byref(db), _stream())
        return
    for d in a + (b or []):
        _lib.call('eb200_conv2d', C.byref(d), _stream())


def launch_wgrad_descs(a: list, b: Optional[list] = None) -> None:
    if b is not None and len(a) == len(b):
        for da, db in zip(a, b):
            _lib.call('eb200_conv2d_wgrad_pair', C.byref(da), C.byref(db), _stream())
        return
    for d in a + (b or []):
        _lib.call('eb200_conv2d_wgrad', C.byref(d), _stream())


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ----------------------------------------------------------------------------------------------
# views and tap tables
# ----------------------------------------------------------------------------------------------
def dense_view(x: torch.Tensor, c: Optional[int] = None) -> View:
    """View of a contiguous NHWC tensor (optionally only its first `c` channels)."""
    n, h, w, ct = x.shape
    assert x.is_contiguous() and x.dtype == BF16
    return View(x.data_ptr(), n, h, w, ct if c is None else c, h * w * ct, w * ct, ct)


def parity_view(x: torch.Tensor, row_par: Optional[int], col_par: Optional[int]) -> View:
    """Rows (cols) of parity row_par (col_par) of a contiguous NHWC tensor; None keeps the full axis."""
    n, h, w, c = x.shape
    ptr, vh, vw, sh, sw = x.data_ptr(), h, w, w * c, c
    if row_par is not None:
        ptr += row_par * w * c * 2
        vh = (h - row_par + 1) // 2
        sh = 2 * w * c
    if col_par is not None:
        ptr += col_par * c * 2
        vw = (w - col_par + 1) // 2
        sw = 2 * c
    return View(ptr, n, vh, vw, c, h * w * c, sh, sw)


@dataclass
class TapTable:
    """taps of a convolution expressed as stride-1 taps over <= 2 parity views of the input."""
    views: List[Tuple[Optional[int], Optional[int]]]   # (row parity, col parity) per view
    tap_view: List[int]
    tap_dy: List[int]
    tap_dx: List[int]


def forward_taps(kh: int, kw: int, sh: int, sw: int) -> TapTable:
    """out(h,w) = sum_{ky,kx} in(sh*h + ky - kh//2, sw*w + kx - kw//2); tap order = ky*kw + kx (torch's)."""
    views: List[Tuple[Optional[int], Optional[int]]] = []
    tv, tdy, tdx = [], [], []
    for ky in range(kh):
        for kx in range(kw):
            dyy, dxx = ky - kh // 2, kx - kw // 2
            rp = (dyy % 2) if sh == 2 else None
            cp = (dxx % 2) if sw == 2 else None
            if (rp, cp) not in views:
                views.append((rp, cp))
            tv.append(views.index((rp, cp)))
            tdy.append((dyy - rp) // 2 if sh == 2 else dyy)
            tdx.append((dxx - cp) // 2 if sw == 2 else dxx)
    assert len(views) <= 2, 'at most two input views per convolution'
    return TapTable(views, tv, tdy, tdx)


def _fill_taps(desc, tt_view, tt_dy, tt_dx):
    desc.taps = len(tt_view)
    for i, (v, dy, dx) in enumerate(zip(tt_view, tt_dy, tt_dx)):
        desc.tap_view[i], desc.tap_dy[i], desc.tap_dx[i] = v, dy, dx


# ----------------------------------------------------------------------------------------------
# packed weights
# ----------------------------------------------------------------------------------------------
@dataclass
class PackedWeight:
    fwd: torch.Tensor          # bf16 [taps][cout_pad][cin_pad]
    bwd: Optional[torch.Tensor]  # bf16 [taps][cin_pad'][cout_pad'] (transposed, for the data gradient)
    cout: int
    cin: int
    kh: int
    kw: int


def pad_cout(c: int) -> int:
    return round_up(c, 16) if c <= 256 else round_up(c, 128)


def pack_weight(w: torch.Tensor, need_bwd: bool = True, out: Optional[PackedWeight] = None) -> PackedWeight:
    """fp32 [Cout,Cin,kh,kw] -> PackedWeight (forward and, optionally, transposed for dgrad)."""
    cout, cin, kh, kw = w.shape
    assert w.dtype == torch.float32 and w.is_contiguous()
    taps = kh * kw
    if out is None:
        fwd = torch.zeros(taps, pad_cout(cout), round_up(cin, 64), dtype=BF16, device=w.device)
        bwd = torch.zeros(taps, pad_cout(round_up(cin, 8)), round_up(cout, 64), dtype=BF16, device=w.device) if need_bwd else None
        out = PackedWeight(fwd, bwd, cout, cin, kh, kw)
    _lib.call('eb200_pack_conv_weight', w.data_ptr(), cout, cin, kh, kw, out.fwd.data_ptr(), out.fwd.shape[1],
              out.fwd.shape[2], 0, 0, 0, _stream())
    if out.bwd is not None:
        _lib.call('eb200_pack_conv_weight', w.data_ptr(), cout, cin, kh, kw, out.bwd.data_ptr(), out.bwd.shape[1],
                  out.bwd.shape[2], 1, 0, 0, _stream())
    return out


# ----------------------------------------------------------------------------------------------
# convolution (tcgen05)
# ----------------------------------------------------------------------------------------------
def conv2d_raw(views: Sequence[View], tap_view, tap_dy, tap_dx, tap_w, weight: torch.Tensor, cin: int, cout: int,
               out_ptr: int, out_ext: Tuple[int, int, int], out_strides: Tuple[int, int, int], *,
               bias: Optional[torch.Tensor] = None, relu: bool = False, aux_ptr: Optional[int] = None,
               aux_strides: Optional[Tuple[int, int, int]] = None, aux_mode: Optional[str] = None,
               stats: Optional[torch.Tensor] = None, stats_sum_only: bool = False, bn_bwd=None) -> None:
    d = ConvDesc()
    for i, v in enumerate(views):
        d.inp[i] = v
    d.n, d.h, d.w = out_ext
    d.cin, d.cout = cin, cout
    _fill_taps(d, tap_view, tap_dy, tap_dx)
    for i, tw in enumerate(tap_w):
        d.tap_w[i] = tw
    d.weight_taps = weight.shape[0]
    d.weight = weight.data_ptr()
    d.cout_pad, d.cin_pad = weight.shape[1], weight.shape[2]
    d.out = out_ptr
    d.out_sn, d.out_sh, d.out_sw = out_strides
    flags = 0
    if bias is not None:
        flags |= BIAS
        d.bias = bias.data_ptr()
    if relu:
        flags |= RELU
    if aux_mode is not None:
        flags |= AUX_ADD if aux_mode == 'add' else AUX_MASK
        d.aux = aux_ptr
        d.aux_sn, d.aux_sh, d.aux_sw = aux_strides
    if stats is not None:
        flags |= STATS | (STATS_SUM_ONLY if stats_sum_only else 0)
        d.stats = stats.data_ptr()
    if bn_bwd is not None:     # (scale, shift) of the BatchNorm whose ReLU/statistics backward is fused (aux = its input)
        flags |= BN_BWD
        d.bn_scale, d.bn_shift = bn_bwd[0].data_ptr(), bn_bwd[1].data_ptr()
    d.flags = flags
    if _defer_conv is not None:
        d._keep = (views, weight, bias, stats, bn_bwd)      # the descriptor only holds raw pointers
        _defer_conv.append(d)
        return
    _lib.call('eb200_conv2d', C.byref(d), _stream())


def _dense_strides(t: torch.Tensor) -> Tuple[int, int, int]:
    n, h, w, c = t.shape
    return (h * w * c, w * c, c)


def conv2d(x: torch.Tensor, pw: PackedWeight, stride: Tuple[int, int] = (1, 1), *, out: Optional[torch.Tensor] = None,
           out_coff: int = 0, bias=None, relu=False, aux: Optional[torch.Tensor] = None, aux_mode=None,
           stats=None, cin: Optional[int] = None) -> torch.Tensor:
    """Forward convolution, padding k//2 (the only padding the reference uses for these layers)."""
    n, h, w, c = x.shape
    sh, sw = stride
    ho, wo = (h + sh - 1) // sh, (w + sw - 1) // sw
    tt = forward_taps(pw.kh, pw.kw, sh, sw)
    if sh == 1 and sw == 1:
        views = [dense_view(x, cin)]
    else:
        views = [parity_view(x, rp, cp) for rp, cp in tt.views]
    if out is None:
        out = torch.empty(n, ho, wo, pw.cout, dtype=BF16, device=x.device)
    ct = out.shape[3]
    conv2d_raw(views, tt.tap_view, tt.tap_dy, tt.tap_dx, list(range(pw.kh * pw.kw)), pw.fwd, pw.cin if cin is None else cin, pw.cout,
               out.data_ptr() + out_coff * 2, (n, ho, wo), (ho * wo * ct, wo * ct, ct), bias=bias, relu=relu,
               aux_ptr=_ptr(aux), aux_strides=_dense_strides(aux) if aux is not None else None, aux_mode=aux_mode,
               stats=stats)
    return out


def conv2d_dgrad(dy: torch.Tensor, pw: PackedWeight, in_shape: Tuple[int, int, int, int],
                 stride: Tuple[int, int] = (1, 1), *, out: Optional[torch.Tensor] = None,
                 aux: Optional[torch.Tensor] = None, aux_mode=None, stats=None, accumulate_into_out: bool = False,
                 dy_c: Optional[int] = None, stats_sum_only: bool = False, bn_bwd=None) -> torch.Tensor:
    """Data gradient: dx[n,r,s,ci] = sum_taps W[t][co][ci] * dy[...]; strided convs write per output parity.

    aux/aux_mode: 'mask' multiplies by (aux > 0) (ReLU backward of the producer), 'add' adds a tensor of dx's shape.
    accumulate_into_out: dx += result (read-modify-write through the aux-add path on the same addresses).
    stats: fp32 [2*cin]: sums of the stored dx per channel (bias gradient of the producer); with stats_sum_only it is
    fp32 [cin] and only the sums are accumulated (pass the bias parameter's .grad itself).
    """
    n, h, w, cin = in_shape
    sh, sw = stride
    kh, kw = pw.kh, pw.kw
    cin8 = round_up(cin, 8)
    if out is None:
        out = torch.empty(n, h, w, cin8, dtype=BF16, device=dy.device)
        if (sh == 2 and kh == 1) or (sw == 2 and kw == 1):
            out.zero_()   # 1x1 stride-2: odd rows/cols receive no gradient
    assert not (accumulate_into_out and aux_mode is not None)
    dyv = dense_view(dy, dy_c)
    ct = out.shape[3]
    row_pars = [0, 1] if (sh == 2 and kh > 1) else ([0] if sh == 2 else [None])
    col_pars = [0, 1] if (sw == 2 and kw > 1) else ([0] if sw == 2 else [None])
    for rp in row_pars:
        for cp in col_pars:
            tv, tdy, tdx = [], [], []
            for ky in range(kh):
                for kx in range(kw):
                    dyy, dxx = ky - kh // 2, kx - kw // 2
                    if rp is not None and (dyy % 2) != rp:
                        continue
                    if cp is not None and (dxx % 2) != cp:
                        continue
                    tv.append(ky * kw + kx)   # absolute tap id selects the weight slice
                    tdy.append(-((dyy - rp) // 2) if rp is not None else -dyy)
                    tdx.append(-((dxx - cp) // 2) if cp is not None else -dxx)
            if not tv:
                continue
            ptr = out.data_ptr()
            oh, ow, o_sh, o_sw = h, w, w * ct, ct
            if rp is not None:
                ptr += rp * w * ct * 2
                oh = (h - rp + 1) // 2
                o_sh = 2 * w * ct
            if cp is not None:
                ptr += cp * ct * 2
                ow = (w - cp + 1) // 2
                o_sw = 2 * ct
            strides = (h * w * ct, o_sh, o_sw)
            a_ptr, a_str, a_mode = None, None, aux_mode
            if accumulate_into_out:
                a_ptr, a_str, a_mode = ptr, strides, 'add'
            elif aux is not None:
                act = aux.shape[3]
                a_ptr = aux.data_ptr() + ((rp or 0) * w * act + (cp or 0) * act) * 2
                a_str = (h * w * act, (2 if rp is not None else 1) * w * act, (2 if cp is not None else 1) * act)
            conv2d_raw([dyv], [0] * len(tv), tdy, tdx, tv, pw.bwd, dyv.c, cin8, ptr, (n, oh, ow), strides,
                       aux_ptr=a_ptr, aux_strides=a_str, aux_mode=a_mode, stats=stats, stats_sum_only=stats_sum_only,
                       bn_bwd=bn_bwd)
    return out


def conv2d_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, kh: int, kw: int,
                 stride: Tuple[int, int] = (1, 1), *, cin: Optional[int] = None, dy_c: Optional[int] = None,
                 dw_strides: Optional[Tuple[int, int, int]] = None, ws: Optional[torch.Tensor] = None) -> None:
    """dw (fp32, reference layout [Cout,Cin,kh,kw] unless dw_strides given) += dy^T * x over all pixels.
    ws: optional ZEROED fp32 staging of >= cout*cin*9 floats for 3x3 filters (returned zeroed)."""
    sh, sw = stride
    tt = forward_taps(kh, kw, sh, sw)
    d = WgradDesc()
    d.dy = dense_view(dy, dy_c)
    if sh == 1 and sw == 1:
        d.x[0] = dense_view(x, cin)
    else:
        for i, (rp, cp) in enumerate(tt.views):
            d.x[i] = parity_view(x, rp, cp)
    _fill_taps(d, tt.tap_view, tt.tap_dy, tt.tap_dx)
    d.dw = dw.data_ptr()
    if dw_strides is None:
        cin_w = dw.shape[1]
        dw_strides = (cin_w * kh * kw, kh * kw, 1)
    d.dw_sco, d.dw_sci, d.dw_st = dw_strides
    if ws is not None:
        d.ws, d.ws_floats = ws.data_ptr(), ws.numel()
    if _defer_wgrad is not None:
        d._keep = (dy, x, dw, ws)
        _defer_wgrad.append(d)
        return
    _lib.call('eb200_conv2d_wgrad', C.byref(d), _stream())


# ----------------------------------------------------------------------------------------------
# batch norm
# ----------------------------------------------------------------------------------------------
def _f32(n, device):
    return torch.empty(n, dtype=torch.float32, device=device)


@dataclass
class BNState:
    """Per-call batch-norm affine + saved statistics (fp32 [C] each).  In train mode the finalize (raw sums -> affine,
    running-statistics update) is folded into the first `bn_apply`: until then `pending` holds its arguments."""
    scale: torch.Tensor
    shift: torch.Tensor
    mean: Optional[torch.Tensor]
    rstd: Optional[torch.Tensor]
    count: int
    pending: Optional[tuple] = None


def bn_finalize(stats: torch.Tensor, count: int, gamma, beta, running_mean, running_var, eps=1e-5,
                momentum=0.1) -> BNState:
    """stats: fp32 [2C] raw sum / sum of squares from the conv epilogue.  No launch here: the next bn_apply derives
    the affine from `stats` inside its own kernel (eb200_bn_apply_train) and publishes scale/shift/mean/rstd.
    `stats` is NOT zeroed any more: callers that re-use the buffer zero it themselves (the engine zeroes one arena
    per step)."""
    c = gamma.numel()
    buf = _f32(4 * c, gamma.device)
    return BNState(buf[0:c], buf[c:2 * c], buf[2 * c:3 * c], buf[3 * c:4 * c], count,
                   pending=(stats, gamma, beta, running_mean, running_var, eps, momentum))


def bn_apply(x: torch.Tensor, st: BNState, *, relu: bool, drop=None, res_pre=None, res_post=None, gap=None,
             out: Optional[torch.Tensor] = None, out_coff: int = 0) -> torch.Tensor:
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    if st.pending is not None:
        stats, gamma, beta, rm, rv, eps, momentum = st.pending
        st.pending = None
        _lib.call('eb200_bn_apply_train', x.data_ptr(), out.data_ptr(), stats.data_ptr(), st.count, gamma.data_ptr(),
                  beta.data_ptr(), eps, momentum, _ptr(rm), _ptr(rv), st.scale.data_ptr(), st.shift.data_ptr(),
                  st.mean.data_ptr(), st.rstd.data_ptr(), _ptr(drop), _ptr(res_pre), _ptr(res_post), _ptr(gap), n,
                  h * w, c, out.shape[3], out_coff, int(relu), _stream())
        return out
    _lib.call('eb200_bn_apply', x.data_ptr(), out.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(), _ptr(drop),
              _ptr(res_pre), _ptr(res_post), _ptr(gap), n, h * w, c, out.shape[3], out_coff, int(relu), _stream())
    return out


BN_REPLICAS = 16   # copies of the [2C] backward sums the reduce blocks spread their atomics over


def bn_rep_floats(c: int) -> int:
    """size of the zeroed scratch of bn_backward: replicas | folded sums | counter"""
    return (BN_REPLICAS + 1) * 2 * c + 4


def bn_backward(dy: torch.Tensor, x: torch.Tensor, st: BNState, gamma: torch.Tensor, sums: Optional[torch.Tensor] = None,
                *, relu_mode: int, mask_src=None, drop=None, want_dres: bool = False, dy_coff: int = 0,
                dgamma: torch.Tensor = None, dbeta: torch.Tensor = None, rep: Optional[torch.Tensor] = None):
    """Returns (dx, dres).  rep: ZEROED fp32 [bn_rep_floats(C)] scratch (allocated here if not given); dgamma / dbeta
    are accumulated.  Two launches: reduce (block sums -> replicas, last block folds them) and apply."""
    n, h, w, c = x.shape
    dy_cs = dy.shape[3]
    assert st.pending is None, 'bn_backward before the forward bn_apply of this BNState'
    if rep is None:
        rep = torch.zeros(bn_rep_floats(c), dtype=torch.float32, device=x.device)
    args = (dy.data_ptr(), x.data_ptr(), _ptr(mask_src), _ptr(drop), st.mean.data_ptr(), st.rstd.data_ptr(),
            st.scale.data_ptr(), st.shift.data_ptr())
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    # one launch: replica reduce, grid-wide barrier, apply (two launches inside when the grid is not fully resident)
    _lib.call('eb200_bn_bwd_fused', *args, gamma.data_ptr(), rep.data_ptr(), BN_REPLICAS, dgamma.data_ptr(),
              dbeta.data_ptr(), dx.data_ptr(), _ptr(dres), n, h * w, c, dy_cs, dy_coff, relu_mode, _stream())
    return dx, dres


def dgrad_with_bn_backward(dy: torch.Tensor, pw: PackedWeight, x_bn: torch.Tensor, st: BNState, gamma: torch.Tensor,
                           raw_sums: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor) -> torch.Tensor:
    """Backward through  x_bn --BN(st)--> ReLU --conv(pw, stride 1)--> . given dy of the conv output: returns d x_bn.
    The conv's data gradient applies the ReLU mask (recomputed from x_bn) and sums g and g*x_bn in its epilogue
    (raw_sums: ZEROED fp32 [2C]); one bandwidth pass then finishes the BatchNorm backward.  Replaces
    conv2d_dgrad + bn_backward(relu_mode=1) (three passes over the activation fewer)."""
    assert st.pending is None
    n, h, w, c = x_bn.shape
    g = conv2d_dgrad(dy, pw, (n, h, w, c), aux=x_bn, aux_mode='mask', stats=raw_sums, bn_bwd=(st.scale, st.shift))
    return bn_bwd_apply_raw(g, x_bn, st, gamma, raw_sums, dgamma, dbeta)


def bn_bwd_apply_raw(g: torch.Tensor, x_bn: torch.Tensor, st: BNState, gamma: torch.Tensor, raw_sums: torch.Tensor,
                     dgamma: torch.Tensor, dbeta: torch.Tensor) -> torch.Tensor:
    """second half of dgrad_with_bn_backward (g = masked data gradient, raw_sums = (sum g, sum g*x))"""
    n, h, w, c = x_bn.shape
    dx = torch.empty_like(x_bn)
    _lib.call('eb200_bn_bwd_apply_raw', g.data_ptr(), x_bn.data_ptr(), st.mean.data_ptr(), st.rstd.data_ptr(),
              gamma.data_ptr(), raw_sums.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), dx.data_ptr(), n, h * w, c,
              _stream())
    return dx


def sums_to_bias_grad(sums: torch.Tensor, dbias: torch.Tensor, scratch: torch.Tensor) -> None:
    """dbias += sums[0:C]; sums zeroed (second half goes to a scratch vector)."""
    _lib.call('eb200_bn_bwd_param', sums.data_ptr(), scratch.data_ptr(), dbias.data_ptr(), dbias.numel(), _stream())


def colsum(x: torch.Tensor, out: torch.Tensor, c: Optional[int] = None, coff: int = 0) -> None:
    cs = x.shape[-1]
    p = x.numel() // cs
    _lib.call('eb200_colsum', x.data_ptr(), out.data_ptr(), p, cs if c is None else c, cs, coff, _stream())


# ----------------------------------------------------------------------------------------------
# stem / pooling
# ----------------------------------------------------------------------------------------------
def im2col_stem(x_nchw: torch.Tensor) -> torch.Tensor:
    n, cin, h, w = x_nchw.shape
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    kpad = round_up(cin * 49, 64)
    out = torch.empty(n, (h + 1) // 2, (w + 1) // 2, kpad, dtype=BF16, device=x_nchw.device)
    _lib.call('eb200_im2col_stem', x_nchw.data_ptr(), out.data_ptr(), n, cin, h, w, kpad, _stream())
    return out


def maxpool_fwd(x: torch.Tensor):
    n, h, w, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty(n, ho, wo, c, dtype=BF16, device=x.device)
    idx = torch.empty(n, ho, wo, c, dtype=torch.uint8, device=x.device)
    _lib.call('eb200_maxpool_fwd', x.data_ptr(), y.data_ptr(), idx.data_ptr(), n, h, w, c, _stream())
    return y, idx


def maxpool_bwd(dy: torch.Tensor, idx: torch.Tensor, in_shape) -> torch.Tensor:
    n, h, w, c = in_shape
    dx = torch.empty(n, h, w, c, dtype=BF16, device=dy.device)
    _lib.call('eb200_maxpool_bwd', dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), n, h, w, c, _stream())
    return dx


# ----------------------------------------------------------------------------------------------
# SE fusion
# ----------------------------------------------------------------------------------------------
def gap(x: torch.Tensor, out: torch.Tensor) -> None:
    n, h, w, c = x.shape
    _lib.call('eb200_gap', x.data_ptr(), out.data_ptr(), n, h * w, c, _stream())


@dataclass
class SEState:
    mean: torch.Tensor
    hid: torch.Tensor
    wgt: torch.Tensor


def se_mlp_fwd(gap_sums: torch.Tensor, hw: int, w1, b1, w2, b2) -> SEState:
    n, c = gap_sums.shape
    cr = w1.shape[0]
    st = SEState(_f32(n * c, w1.device).view(n, c), _f32(n * cr, w1.device).view(n, cr),
                 _f32(n * c, w1.device).view(n, c))
    _lib.call('eb200_se_mlp_fwd', gap_sums.data_ptr(), hw, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
              st.mean.data_ptr(), st.hid.data_ptr(), st.wgt.data_ptr(), n, c, cr, _stream())
    return st


def se_mlp_bwd(dwgt: torch.Tensor, st: SEState, hw: int, w1, w2, dw1, db1, dw2, db2) -> torch.Tensor:
    n, c = dwgt.shape
    dmean = _f32(n * c, w1.device).view(n, c)
    _lib.call('eb200_se_mlp_bwd', dwgt.data_ptr(), st.wgt.data_ptr(), st.hid.data_ptr(), st.mean.data_ptr(), hw,
              w1.data_ptr(), w2.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(),
              dmean.data_ptr(), n, c, w1.shape[0], _stream())
    return dmean


def se_fuse_fwd(a: torch.Tensor, b: torch.Tensor, wa: torch.Tensor, wb: torch.Tensor) -> torch.Tensor:
    n, h, w, c = a.shape
    out = torch.empty_like(a)
    _lib.call('eb200_se_fuse_fwd', a.data_ptr(), b.data_ptr(), wa.data_ptr(), wb.data_ptr(), out.data_ptr(), n, h * w,
              c, _stream())
    return out


def se_fuse_bwd_reduce(dout, a, b, dwa: torch.Tensor, dwb: torch.Tensor) -> None:
    n, h, w, c = a.shape
    _lib.call('eb200_se_fuse_bwd_reduce', dout.data_ptr(), a.data_ptr(), b.data_ptr(), dwa.data_ptr(), dwb.data_ptr(),
              n, h * w, c, _stream())


def se_fuse_bwd_apply(dout, wa, wb, dmean_a, dmean_b, db_prev):
    n, h, w, c = dout.shape
    da, db = torch.empty_like(dout), torch.empty_like(dout)
    _lib.call('eb200_se_fuse_bwd_apply', dout.data_ptr(), wa.data_ptr(), wb.data_ptr(), dmean_a.data_ptr(),
              dmean_b.data_ptr(), _ptr(db_prev), da.data_ptr(), db.data_ptr(), n, h * w, c, _stream())
    return da, db


# ----------------------------------------------------------------------------------------------
# pyramid pooling
# ----------------------------------------------------------------------------------------------
def adaptive_pool_fwd(x: torch.Tensor, b: int) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty(n, b, b, c, dtype=BF16, device=x.device)
    _lib.call('eb200_adaptive_pool_fwd', x.data_ptr(), y.data_ptr(), n, h, w, c, b, _stream())
    return y


def adaptive_pool_bwd(dy: torch.Tensor, dx: torch.Tensor, accumulate: bool) -> None:
    n, h, w, c = dx.shape
    _lib.call('eb200_adaptive_pool_bwd', dy.data_ptr(), dx.data_ptr(), n, h, w, c, dy.shape[1], int(accumulate),
              _stream())


def bilinear_fwd(x: torch.Tensor, out: torch.Tensor, out_coff: int) -> None:
    n, hi, wi, c = x.shape
    _lib.call('eb200_bilinear_fwd', x.data_ptr(), out.data_ptr(), n, hi, wi, out.shape[1], out.shape[2], c,
              out.shape[3], out_coff, _stream())


def bilinear_bwd(dy: torch.Tensor, dy_coff: int, in_shape) -> torch.Tensor:
    n, hi, wi, c = in_shape
    dx = torch.empty(n, hi, wi, c, dtype=BF16, device=dy.device)
    _lib.call('eb200_bilinear_bwd', dy.data_ptr(), dx.data_ptr(), n, hi, wi, dy.shape[1], dy.shape[2], c, dy.shape[3],
              dy_coff, _stream())
    return dx


# ----------------------------------------------------------------------------------------------
# learned upsampling (nearest x2 + depthwise 3x3)
# ----------------------------------------------------------------------------------------------
def upsample_dw_fwd(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    n, h, wd, c = x.shape
    y = torch.empty(n, 2 * h, 2 * wd, c, dtype=BF16, device=x.device)
    _lib.call('eb200_upsample_dw_fwd', x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, wd, c,
              w.shape[0], _stream())
    return y


def upsample_dw_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, dw: torch.Tensor, db: torch.Tensor,
                    need_dx: bool = True) -> Optional[torch.Tensor]:
    n, h, wd, c = x.shape
    _lib.call('eb200_upsample_dw_bwd_weight', dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), n, h, wd, c,
              w.shape[0], _stream())
    if not need_dx:
        return None
    dx = torch.empty_like(x)
    _lib.call('eb200_upsample_dw_bwd_input', dy.data_ptr(), w.data_ptr(), dx.data_ptr(), n, h, wd, c, w.shape[0],
              _stream())
    return dx


def upsample_dw_fwd_nchw(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """last upsampling of a head + output boundary: NHWC bf16 [N,H,W,C] -> fp32 NCHW [N,Creal,2H,2W]"""
    n, h, wd, c = x.shape
    creal = w.shape[0]
    y = torch.empty(n, creal, 2 * h, 2 * wd, dtype=torch.float32, device=x.device)
    _lib.call('eb200_upsample_dw_fwd_nchw', x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, wd, c, creal,
              _stream())
    return y


def upsample_dw_bwd_nchw(g: torch.Tensor, x: torch.Tensor, w: torch.Tensor, dw: torch.Tensor, db: torch.Tensor,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """g: fp32 NCHW gradient of upsample_dw_fwd_nchw's output (read once) -> dx bf16 NHWC; dw / db accumulated"""
    n, h, wd, c = x.shape
    assert g.dtype == torch.float32 and g.is_contiguous() and tuple(g.shape) == (n, w.shape[0], 2 * h, 2 * wd)
    dx = torch.empty_like(x) if out is None else out
    _lib.call('eb200_upsample_dw_bwd_nchw', g.data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(),
              db.data_ptr(), n, h, wd, c, w.shape[0], _stream())
    return dx


# ----------------------------------------------------------------------------------------------
# output boundary, scene head, helpers
# ----------------------------------------------------------------------------------------------
def nhwc_to_nchw(x: torch.Tensor, creal: int) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty(n, creal, h, w, dtype=torch.float32, device=x.device)
    _lib.call('eb200_nhwc_to_nchw', x.data_ptr(), y.data_ptr(), None, None, n, h * w, c, creal, 0, _stream())
    return y


def instance_outputs(x: torch.Tensor, with_orientation: bool):
    n, h, w, c = x.shape
    assert c == 8
    y0 = torch.empty(n, 1, h, w, dtype=torch.float32, device=x.device)
    y1 = torch.empty(n, 2, h, w, dtype=torch.float32, device=x.device)
    y2 = torch.empty(n, 2, h, w, dtype=torch.float32, device=x.device) if with_orientation else None
    _lib.call('eb200_nhwc_to_nchw', x.data_ptr(), y0.data_ptr(), y1.data_ptr(), _ptr(y2), n, h * w, c, 5, 1, _stream())
    return (y0, y1, y2) if with_orientation else (y0, y1)


def nchw_grad_to_nhwc(g: Optional[torch.Tensor], shape, creal: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, h, w, c = shape
    dx = torch.empty(n, h, w, c, dtype=BF16, device=g.device) if out is None else out
    _lib.call('eb200_nchw_to_nhwc_grad', g.data_ptr(), None, None, None, dx.data_ptr(), n, h * w, c, creal, 0,
              _stream())
    return dx


def instance_outputs_bwd(g0, g1, g2, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, h, w, c = x.shape
    dx = torch.empty_like(x) if out is None else out
    _lib.call('eb200_nchw_to_nhwc_grad', _ptr(g0), _ptr(g1), _ptr(g2), x.data_ptr(), dx.data_ptr(), n, h * w, c, 5, 1,
              _stream())
    return dx


def linear_fwd(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    n, k = x.shape[0], x.shape[-1]
    m = w.shape[0]
    y = torch.empty(n, m, dtype=torch.float32, device=x.device)
    _lib.call('eb200_linear_fwd', x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, k, m, _stream())
    return y


def linear_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, dw: torch.Tensor, db: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n, k = x.shape[0], x.shape[-1]
    dx = torch.empty_like(x) if out is None else out
    _lib.call('eb200_linear_bwd', dy.data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), dw.data_ptr(),
              db.data_ptr(), n, k, w.shape[0], _stream())
    return dx


def add_inplace(a: torch.Tensor, b: torch.Tensor) -> None:
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    _lib.call('eb200_add_inplace', a.data_ptr(), b.data_ptr(), a.numel(), _stream())


def copy_channels(src: torch.Tensor, dst: torch.Tensor, c: int, scoff: int, dcoff: int, accumulate: bool) -> None:
    p = src.numel() // src.shape[-1]
    _lib.call('eb200_copy_channels', src.data_ptr(), dst.data_ptr(), p, c, src.shape[-1], scoff, dst.shape[-1], dcoff,
              int(accumulate), _stream())
