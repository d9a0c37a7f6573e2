"""Per-kernel parity (GPU): every C-ABI op against the same op in plain PyTorch fp32 on identical
(bf16-rounded) operands.  Tolerances: the accumulators are fp32, outputs are stored as bf16, so an
output may differ from the fp32 reference by one bf16 rounding (2^-8 relative) plus fp32 summation
noise; fp32 outputs (weight gradients, statistics) are held to 2e-3 relative to the tensor's max
(the bf16 rounding of the *inputs* is shared with the reference)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8


def _ops():
    from emsanet_b200 import ops
    return ops


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def rand_act(n, c, h, w, seed=0, relu=False):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(n, c, h, w, device='cuda', generator=g)
    if relu:
        x = x.clamp_min(0)
    return x.to(torch.bfloat16)


def assert_close_bf16(got, ref, what, extra=0.0):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    tol = (1.5 * BF16_EPS + extra) * scale
    assert err <= tol, f'{what}: max abs err {err:.4e} > tol {tol:.4e} (scale {scale:.3e})'


def assert_close_f32(got, ref, what, rtol=2e-3):
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, f'{what}: max abs err {err:.4e} > {rtol} * {scale:.3e}'


CONV_CASES = [
    # n, h, w, cin, cout, kh, kw, stride
    (2, 16, 24, 64, 64, 3, 1, (1, 1)),
    (2, 16, 24, 64, 64, 1, 3, (1, 1)),
    (3, 15, 20, 128, 128, 3, 1, (1, 1)),      # ragged tiles
    (2, 15, 20, 512, 512, 1, 3, (1, 1)),      # two N tiles, deep K
    (2, 30, 40, 256, 128, 3, 3, (1, 1)),
    (2, 12, 16, 128, 40, 3, 3, (1, 1)),       # semantic head: cout 40
    (2, 12, 16, 128, 96, 3, 3, (1, 1)),       # instance shared conv: cout 96
    (2, 12, 16, 96, 8, 3, 3, (1, 1)),         # instance task convs (cin not a multiple of 64, tiny cout)
    (2, 24, 32, 64, 128, 1, 1, (1, 1)),
    (2, 24, 32, 64, 128, 3, 1, (2, 1)),       # strided first block
    (2, 24, 32, 128, 128, 1, 3, (1, 2)),
    (2, 24, 32, 64, 128, 1, 1, (2, 2)),       # residual downsample
    (1, 5, 5, 512, 256, 1, 1, (1, 1)),        # PPM bin 5
    (4, 1, 1, 512, 256, 1, 1, (1, 1)),        # PPM bin 1
    (2, 15, 20, 1024, 512, 1, 1, (1, 1)),     # PPM final conv
    (2, 60, 80, 192, 64, 1, 1, (1, 1)),       # stem as 1x1 over im2col (K=192)
]


def _ref_conv(x, w, stride, kh, kw):
    return F.conv2d(x.float(), w.to(torch.bfloat16).float(), None, stride, (kh // 2, kw // 2))


@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'x'.join(map(str, c[:7])) + f's{c[7][0]}{c[7][1]}')
def test_conv2d_forward(case):
    ops = _ops()
    n, h, w, cin, cout, kh, kw, stride = case
    x = rand_act(n, cin, h, w, seed=1)
    g = torch.Generator(device='cuda').manual_seed(2)
    wt = torch.randn(cout, cin, kh, kw, device='cuda', generator=g) / math.sqrt(cin * kh * kw)
    pw = ops.pack_weight(wt)
    y = ops.conv2d(nhwc(x), pw, stride)
    torch.cuda.synchronize()
    ref = _ref_conv(x, wt, stride, kh, kw)
    assert y.shape == (n, ref.shape[2], ref.shape[3], cout)
    assert_close_bf16(nchw(y), ref, 'conv fwd')


def test_conv2d_epilogues():
    ops = _ops()
    n, h, w, cin, cout = 2, 15, 20, 128, 128
    x = rand_act(n, cin, h, w, seed=3)
    g = torch.Generator(device='cuda').manual_seed(4)
    wt = torch.randn(cout, cin, 1, 3, device='cuda', generator=g) / math.sqrt(cin * 3)
    bias = torch.randn(cout, device='cuda', generator=g)
    res = rand_act(n, cout, h, w, seed=5)
    pw = ops.pack_weight(wt)
    ref = _ref_conv(x, wt, 1, 1, 3)
    # bias + relu
    y = ops.conv2d(nhwc(x), pw, bias=bias, relu=True)
    assert_close_bf16(nchw(y), F.relu(ref + bias[None, :, None, None]), 'bias+relu')
    # bias + residual + relu (eval-mode block tail)
    y = ops.conv2d(nhwc(x), pw, bias=bias, relu=True, aux=nhwc(res), aux_mode='add')
    assert_close_bf16(nchw(y), F.relu(ref + bias[None, :, None, None] + res.float()), 'bias+res+relu', extra=BF16_EPS)
    # relu-backward mask
    y = ops.conv2d(nhwc(x), pw, aux=nhwc(res), aux_mode='mask')
    assert_close_bf16(nchw(y), ref * (res.float() > 0), 'mask')
    # statistics of the stored values
    stats = torch.zeros(2 * cout, device='cuda')
    y = ops.conv2d(nhwc(x), pw, stats=stats)
    torch.cuda.synchronize()
    yf = nchw(y).float()
    assert_close_f32(stats[:cout], yf.sum((0, 2, 3)), 'stats sum', 1e-3)
    assert_close_f32(stats[cout:], (yf * yf).sum((0, 2, 3)), 'stats sumsq', 1e-3)
    # channel-sliced output (concat fusion)
    wide = torch.zeros(n, h, w, 256, dtype=torch.bfloat16, device='cuda')
    ops.conv2d(nhwc(x), pw, out=wide, out_coff=64)
    assert_close_bf16(nchw(wide[..., 64:192]), ref, 'sliced out')
    assert wide[..., :64].abs().max().item() == 0 and wide[..., 192:].abs().max().item() == 0


@pytest.mark.parametrize('case', [c for c in CONV_CASES if c[4] % 8 == 0 and c[3] != 192],
                         ids=lambda c: 'x'.join(map(str, c[:7])) + f's{c[7][0]}{c[7][1]}')
def test_conv2d_dgrad(case):
    ops = _ops()
    n, h, w, cin, cout, kh, kw, stride = case
    g = torch.Generator(device='cuda').manual_seed(6)
    wt = torch.randn(cout, cin, kh, kw, device='cuda', generator=g) / math.sqrt(cout * kh * kw)
    ho, wo = (h + stride[0] - 1) // stride[0], (w + stride[1] - 1) // stride[1]
    dy = rand_act(n, cout, ho, wo, seed=7)
    pw = ops.pack_weight(wt)
    dx = ops.conv2d_dgrad(nhwc(dy), pw, (n, h, w, cin), stride)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt.to(torch.bfloat16).float(), dy.float(), stride,
                                     (kh // 2, kw // 2))
    assert_close_bf16(nchw(dx)[:, :cin], ref, 'dgrad')


@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'x'.join(map(str, c[:7])) + f's{c[7][0]}{c[7][1]}')
def test_conv2d_wgrad(case):
    ops = _ops()
    n, h, w, cin, cout, kh, kw, stride = case
    ho, wo = (h + stride[0] - 1) // stride[0], (w + stride[1] - 1) // stride[1]
    x = rand_act(n, cin, h, w, seed=8)
    dy = rand_act(n, cout, ho, wo, seed=9)
    dw = torch.zeros(cout, cin, kh, kw, device='cuda')
    # 3x3 filters: with a staging workspace the three tap groups leave as TMA bulk reductions (left zeroed afterwards)
    ws = torch.zeros(9 * cin * cout, device='cuda') if kh * kw == 9 else None
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), dw, kh, kw, stride, ws=ws)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, kh, kw), dy.float(), stride, (kh // 2, kw // 2))
    assert_close_f32(dw, ref, 'wgrad')
    # accumulation semantics
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), dw, kh, kw, stride, ws=ws)
    torch.cuda.synchronize()
    assert_close_f32(dw, 2 * ref, 'wgrad accumulate')
    if ws is not None:
        assert float(ws.abs().max()) == 0.0
        dw2 = torch.zeros_like(dw)
        ops.conv2d_wgrad(nhwc(dy), nhwc(x), dw2, kh, kw, stride)      # scalar-atomic path without the workspace
        torch.cuda.synchronize()
        assert_close_f32(dw2, ref, 'wgrad (no workspace)')


def test_conv2d_full_size_linearity():
    """Config-2 sized layer (32x120x160x64): conv(a+b) == conv(a)+conv(b) and a strided spot check vs torch."""
    ops = _ops()
    n, h, w, c = 32, 120, 160, 64
    a = rand_act(n, c, h, w, seed=10)
    g = torch.Generator(device='cuda').manual_seed(11)
    wt = torch.randn(c, c, 3, 1, device='cuda', generator=g) / math.sqrt(3 * c)
    pw = ops.pack_weight(wt)
    ya = ops.conv2d(nhwc(a), pw)
    ref = _ref_conv(a[:2], wt, 1, 3, 1)
    assert_close_bf16(nchw(ya[:2]), ref, 'full-size conv, first images')
    ref = _ref_conv(a[-1:], wt, 1, 3, 1)
    assert_close_bf16(nchw(ya[-1:]), ref, 'full-size conv, last image')
    y2 = ops.conv2d(nhwc((a.float() * 2).to(torch.bfloat16)), pw)
    assert_close_bf16(y2, ya.float() * 2, 'homogeneity', extra=BF16_EPS)


# ------------------------------------------------------------------------------------------------
# HBM-bound kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape', [(2, 64, 24, 32), (3, 96, 15, 20), (2, 512, 3, 4), (4, 256, 1, 1)])
def test_batchnorm_train_forward_backward(shape):
    ops = _ops()
    n, c, h, w = shape
    x = rand_act(n, c, h, w, seed=20) * 1.7 + 0.3
    x = x.to(torch.bfloat16)
    g = torch.Generator(device='cuda').manual_seed(21)
    gamma = torch.rand(c, device='cuda', generator=g) + 0.5
    beta = torch.randn(c, device='cuda', generator=g) * 0.1
    rm = torch.randn(c, device='cuda', generator=g) * 0.1
    rv = torch.rand(c, device='cuda', generator=g) + 0.5
    res = rand_act(n, c, h, w, seed=22)
    drop = (torch.rand(n, c, device='cuda', generator=g) > 0.3).float() / 0.7
    dy = rand_act(n, c, h, w, seed=23)
    # reference
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    yb = F.batch_norm(xr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
    out_ref = F.relu(yb * drop[:, :, None, None] + res.float())
    out_ref.backward(dy.float())
    # ours: statistics of x (as the conv epilogue would produce them)
    xf = x.float()
    stats = torch.cat([xf.sum((0, 2, 3)), (xf * xf).sum((0, 2, 3))]).contiguous()
    rm2, rv2 = rm.clone(), rv.clone()
    st = ops.bn_finalize(stats, n * h * w, gamma, beta, rm2, rv2)
    gap = torch.zeros(n, c, device='cuda')
    out = ops.bn_apply(nhwc(x), st, relu=True, drop=drop, res_pre=nhwc(res), gap=gap)
    torch.cuda.synchronize()
    assert st.pending is None          # the finalize ran inside this bn_apply (one launch)
    assert_close_bf16(nchw(out), out_ref, 'bn_apply')
    assert_close_f32(rm2, rm_ref, 'running_mean', 1e-4)
    assert_close_f32(rv2, rv_ref, 'running_var', 1e-3)
    assert_close_f32(gap, nchw(out).float().sum((2, 3)), 'gap', 1e-3)
    sums = torch.zeros(2 * c, device='cuda')
    dgamma, dbeta = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    dx, dres = ops.bn_backward(nhwc(dy), nhwc(x), st, gamma, sums, relu_mode=1, mask_src=out, drop=drop,
                               want_dres=True, dgamma=dgamma, dbeta=dbeta)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(dx), xr.grad, 'bn dx', extra=2 * BF16_EPS)
    assert_close_bf16(nchw(dres), dy.float() * (out_ref > 0), 'bn dres')
    assert_close_f32(dgamma, gr.grad, 'dgamma', 5e-3)
    assert_close_f32(dbeta, br.grad, 'dbeta', 5e-3)
    # relu_mode 2 (recompute mask), post-add variant
    out2 = ops.bn_apply(nhwc(x), st, relu=True, res_post=nhwc(res))
    ref2 = F.relu(yb.detach()) + res.float()
    assert_close_bf16(nchw(out2), ref2, 'bn_apply post-add')
    dx2, _ = ops.bn_backward(nhwc(dy), nhwc(x), st, gamma, sums, relu_mode=2, dgamma=dgamma, dbeta=dbeta)
    xr2 = x.float().requires_grad_(True)
    F.relu(F.batch_norm(xr2, None, None, gamma, beta, True, 0.1, 1e-5)).backward(dy.float())
    assert_close_bf16(nchw(dx2), xr2.grad, 'bn dx mode 2', extra=2 * BF16_EPS)


def test_conv_stats_feed_batchnorm():
    """conv epilogue statistics -> finalize -> apply == F.batch_norm(conv) for a multi-tile, multi-wave map"""
    ops = _ops()
    for (n, h, w, c) in [(2, 24, 32, 64), (2, 12, 16, 128), (2, 6, 8, 256), (2, 3, 4, 512), (8, 60, 80, 64)]:
        x = rand_act(n, c, h, w, seed=30, relu=True)
        g = torch.Generator(device='cuda').manual_seed(31)
        wt = torch.randn(c, c, 1, 3, device='cuda', generator=g) / math.sqrt(3 * c)
        gamma = torch.rand(c, device='cuda', generator=g) + 0.5
        beta = torch.randn(c, device='cuda', generator=g) * 0.1
        pw = ops.pack_weight(wt)
        stats = torch.zeros(2 * c, device='cuda')
        y = ops.conv2d(nhwc(x), pw, stats=stats)
        yf = nchw(y).float()
        assert_close_f32(stats[:c], yf.sum((0, 2, 3)), f'sum {c}', 1e-3)
        assert_close_f32(stats[c:], (yf * yf).sum((0, 2, 3)), f'sumsq {c}', 1e-3)
        st = ops.bn_finalize(stats, n * h * w, gamma, beta, None, None)
        out = ops.bn_apply(y, st, relu=True)
        ref = F.relu(F.batch_norm(yf, None, None, gamma, beta, True, 0.1, 1e-5))
        assert_close_bf16(nchw(out), ref, f'conv+bn {c}')


@pytest.mark.parametrize('shape', [(2, 40, 12, 16, 40), (2, 8, 24, 32, 5), (3, 128, 15, 20, 128), (1, 512, 3, 4, 512)])
def test_learned_upsampling_forward_backward(shape):
    ops = _ops()
    n, c, h, w, creal = shape
    x = rand_act(n, c, h, w, seed=40)
    if creal < c:
        x[:, creal:] = 0
    g = torch.Generator(device='cuda').manual_seed(41)
    wt = (torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device='cuda') / 16.).expand(creal, 1, 3, 3)
    wt = (wt * (1 + 0.3 * torch.randn(creal, 1, 3, 3, device='cuda', generator=g))).contiguous()
    b = torch.randn(creal, device='cuda', generator=g) * 0.1
    dy = rand_act(n, c, 2 * h, 2 * w, seed=42)
    if creal < c:
        dy[:, creal:] = 0
    xr = x[:, :creal].float().requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(F.interpolate(xr, scale_factor=2., mode='nearest'), wr, br, 1, 1, 1, creal)
    ref.backward(dy[:, :creal].float())
    y = ops.upsample_dw_fwd(nhwc(x), wt, b)
    assert_close_bf16(nchw(y)[:, :creal], ref, 'upsample fwd')
    if creal < c:
        assert nchw(y)[:, creal:].abs().max().item() == 0
    dw, db = torch.zeros_like(wt), torch.zeros_like(b)
    dx = ops.upsample_dw_bwd(nhwc(dy), nhwc(x), wt, dw, db)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(dx)[:, :creal], xr.grad, 'upsample dx')
    assert_close_f32(dw, wr.grad, 'upsample dw', 3e-3)
    assert_close_f32(db, br.grad, 'upsample db', 3e-3)


@pytest.mark.parametrize('shape', [(2, 40, 12, 16, 40), (2, 8, 24, 32, 5), (3, 48, 15, 20, 40), (1, 40, 37, 70, 40),
                                   (2, 40, 120, 160, 40)])
def test_learned_upsampling_fused_with_the_nchw_output_boundary(shape):
    """last upsampling of a head + fp32 NCHW boundary in one kernel each way (csrc/upsample_nchw.cu): forward values are
    the bf16-rounded ones of the unfused path (bit-equal to it), backward reads the fp32 gradient once"""
    ops = _ops()
    n, c, h, w, creal = shape
    x = rand_act(n, c, h, w, seed=43)
    if creal < c:
        x[:, creal:] = 0
    g = torch.Generator(device='cuda').manual_seed(44)
    wt = (torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device='cuda') / 16.).expand(creal, 1, 3, 3)
    wt = (wt * (1 + 0.3 * torch.randn(creal, 1, 3, 3, device='cuda', generator=g))).contiguous()
    b = torch.randn(creal, device='cuda', generator=g) * 0.1
    dy = torch.randn(n, creal, 2 * h, 2 * w, device='cuda', generator=g)
    xr = x[:, :creal].float().requires_grad_(True)
    wr, br = wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(F.interpolate(xr, scale_factor=2., mode='nearest'), wr, br, 1, 1, 1, creal)
    ref.backward(dy)
    y = ops.upsample_dw_fwd_nchw(nhwc(x), wt, b)
    assert y.dtype == torch.float32 and tuple(y.shape) == (n, creal, 2 * h, 2 * w)
    assert_close_bf16(y, ref, 'fused upsample fwd')
    unfused = ops.nhwc_to_nchw(ops.upsample_dw_fwd(nhwc(x), wt, b), creal)
    assert torch.equal(y, unfused), 'fused and unfused forward must agree bit for bit'
    dw, db = torch.zeros_like(wt), torch.zeros_like(b)
    dx = ops.upsample_dw_bwd_nchw(dy, nhwc(x), wt, dw, db)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(dx)[:, :creal], xr.grad, 'fused upsample dx')
    if creal < c:
        assert nchw(dx)[:, creal:].abs().max().item() == 0
    assert_close_f32(dw, wr.grad, 'fused upsample dw', 1e-3)
    assert_close_f32(db, br.grad, 'fused upsample db', 1e-3)


def test_maxpool_forward_backward():
    ops = _ops()
    x = rand_act(2, 64, 24, 32, seed=50)
    dy = rand_act(2, 64, 12, 16, seed=51)
    xr = x.float().requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    ref.backward(dy.float())
    y, idx = ops.maxpool_fwd(nhwc(x))
    assert torch.equal(nchw(y).float(), ref)
    dx = ops.maxpool_bwd(nhwc(dy), idx, (2, 24, 32, 64))
    assert_close_bf16(nchw(dx), xr.grad, 'maxpool bwd')


def test_se_fusion_forward_backward():
    ops = _ops()
    n, c, h, w = 3, 128, 12, 16
    cr = c // 16
    a, b = rand_act(n, c, h, w, seed=60, relu=True), rand_act(n, c, h, w, seed=61, relu=True)
    g = torch.Generator(device='cuda').manual_seed(62)
    P = {}
    for m in 'ab':
        P[m] = [torch.randn(cr, c, 1, 1, device='cuda', generator=g) * 0.2, torch.randn(cr, device='cuda', generator=g) * 0.1,
                torch.randn(c, cr, 1, 1, device='cuda', generator=g) * 0.5, torch.randn(c, device='cuda', generator=g) * 0.1]
    dy = rand_act(n, c, h, w, seed=63)
    # reference (MT/model/utils.py:91-95, encoder_fusion.py:84)
    leaves = {m: [t.clone().requires_grad_(True) for t in P[m]] for m in 'ab'}
    ar, br = a.float().requires_grad_(True), b.float().requires_grad_(True)

    def se(x, ps):
        s = F.adaptive_avg_pool2d(x, 1)
        s = torch.sigmoid(F.conv2d(F.relu(F.conv2d(s, ps[0], ps[1])), ps[2], ps[3]))
        return x * s
    ref = se(ar, leaves['a']) + se(br, leaves['b'])
    ref.backward(dy.float())
    # ours
    ga, gb = torch.zeros(n, c, device='cuda'), torch.zeros(n, c, device='cuda')
    ops.gap(nhwc(a), ga)
    ops.gap(nhwc(b), gb)
    sa = ops.se_mlp_fwd(ga, h * w, *P['a'])
    sb = ops.se_mlp_fwd(gb, h * w, *P['b'])
    assert ga.abs().max().item() == 0
    out = ops.se_fuse_fwd(nhwc(a), nhwc(b), sa.wgt, sb.wgt)
    assert_close_bf16(nchw(out), ref, 'se fuse fwd')
    dwa, dwb = torch.zeros(n, c, device='cuda'), torch.zeros(n, c, device='cuda')
    ops.se_fuse_bwd_reduce(nhwc(dy), nhwc(a), nhwc(b), dwa, dwb)
    G = {m: [torch.zeros_like(t) for t in P[m]] for m in 'ab'}
    dma = ops.se_mlp_bwd(dwa, sa, h * w, P['a'][0], P['a'][2], *G['a'])
    dmb = ops.se_mlp_bwd(dwb, sb, h * w, P['b'][0], P['b'][2], *G['b'])
    prev = rand_act(n, c, h, w, seed=64)
    da, db = ops.se_fuse_bwd_apply(nhwc(dy), sa.wgt, sb.wgt, dma, dmb, nhwc(prev))
    torch.cuda.synchronize()
    assert_close_bf16(nchw(da), ar.grad, 'se da', extra=BF16_EPS)
    assert_close_bf16(nchw(db), br.grad + prev.float(), 'se db (+prev)', extra=BF16_EPS)
    for m in 'ab':
        for got, leaf, nm in zip(G[m], leaves[m], ('w1', 'b1', 'w2', 'b2')):
            assert_close_f32(got, leaf.grad, f'se {m}.{nm}', 5e-3)


@pytest.mark.parametrize('hw', [(15, 20), (24, 32), (3, 4)])
def test_pyramid_pooling_ops(hw):
    ops = _ops()
    h, w = hw
    n, c = 2, 64
    x = rand_act(n, c, h, w, seed=70)
    for b in (1, 5):
        xr = x.float().requires_grad_(True)
        pooled = F.adaptive_avg_pool2d(xr, b)
        up = F.interpolate(pooled, (h, w), mode='bilinear', align_corners=False)
        dy = rand_act(n, c, h, w, seed=71)
        up.backward(dy.float())
        p = ops.adaptive_pool_fwd(nhwc(x), b)
        assert_close_bf16(nchw(p), pooled, f'adaptive pool {b}')
        wide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device='cuda')
        pr = pooled.detach().to(torch.bfloat16)
        ops.bilinear_fwd(nhwc(pr), wide, c)
        assert_close_bf16(nchw(wide[..., c:]), F.interpolate(pr.float(), (h, w), mode='bilinear', align_corners=False),
                          f'bilinear {b}')
        dwide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device='cuda')
        dwide[..., c:] = nhwc(dy)
        dp = ops.bilinear_bwd(dwide, c, (n, b, b, c))
        dx = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device='cuda')
        ops.adaptive_pool_bwd(dp, dx, False)
        assert_close_bf16(nchw(dx), xr.grad, f'pool+bilinear bwd {b}', extra=2 * BF16_EPS)


def test_output_boundary_and_scene_head():
    ops = _ops()
    n, h, w = 2, 12, 16
    x = rand_act(n, 40, h, w, seed=80)
    y = ops.nhwc_to_nchw(nhwc(x), 40)
    assert torch.equal(y, x.float())
    gr = torch.randn(n, 40, h, w, device='cuda')
    d = ops.nchw_grad_to_nhwc(gr, (n, h, w, 40), 40)
    assert torch.equal(nchw(d), gr.to(torch.bfloat16))
    # instance activations (MT/model/decoder/instance.py:113-119)
    t = rand_act(n, 8, h, w, seed=81) * 2
    t[:, 5:] = 0
    tr = t[:, :5].float().requires_grad_(True)
    r0, r1 = torch.sigmoid(tr[:, 0:1]), torch.tanh(tr[:, 1:3])
    o = tr[:, 3:5]
    r2 = o / (torch.sqrt(torch.sum(o * o, dim=1, keepdim=True)) + 1e-7)
    g0, g1, g2 = (torch.randn_like(r) for r in (r0, r1, r2))
    torch.autograd.backward([r0, r1, r2], [g0, g1, g2])
    y0, y1, y2 = ops.instance_outputs(nhwc(t), True)
    for got, ref, nm in ((y0, r0, 'sigmoid'), (y1, r1, 'tanh'), (y2, r2, 'unit')):
        assert_close_f32(got, ref.detach(), nm, 1e-4)
    dt = ops.instance_outputs_bwd(g0.contiguous(), g1.contiguous(), g2.contiguous(), nhwc(t))
    assert_close_bf16(nchw(dt)[:, :5], tr.grad, 'instance act bwd')
    # scene head
    f = rand_act(n, 256, 1, 1, seed=82)
    g = torch.Generator(device='cuda').manual_seed(83)
    wl, bl = torch.randn(10, 256, device='cuda', generator=g) * 0.1, torch.randn(10, device='cuda', generator=g)
    ys = ops.linear_fwd(nhwc(f), wl, bl)
    assert_close_f32(ys, F.linear(f.float().flatten(1), wl, bl), 'linear', 1e-4)
    dyl = torch.randn(n, 10, device='cuda', generator=g)
    dwl, dbl = torch.zeros_like(wl), torch.zeros_like(bl)
    dxl = ops.linear_bwd(dyl, nhwc(f), wl, dwl, dbl)
    assert_close_bf16(dxl.flatten(1), dyl @ wl, 'linear dx')
    assert_close_f32(dwl, dyl.t() @ f.float().flatten(1), 'linear dw', 1e-4)
    assert_close_f32(dbl, dyl.sum(0), 'linear db', 1e-4)


def test_stem_im2col_matches_conv7x7():
    ops = _ops()
    for cin in (3, 1):
        g = torch.Generator(device='cuda').manual_seed(90 + cin)
        x = torch.randn(2, cin, 32, 48, device='cuda', generator=g)
        wt = torch.randn(64, cin, 7, 7, device='cuda', generator=g) / math.sqrt(49 * cin)
        cols = ops.im2col_stem(x)
        pw = ops.pack_weight(wt.reshape(64, -1, 1, 1).contiguous(), need_bwd=False)
        y = ops.conv2d(cols, pw)
        ref = F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), None, 2, 3)
        assert_close_bf16(nchw(y), ref, f'stem cin={cin}')
        dy = rand_act(2, 64, 16, 24, seed=95)
        dw = torch.zeros(64, cin * 49, 1, 1, device='cuda')
        ops.conv2d_wgrad(nhwc(dy), cols, dw, 1, 1, cin=cin * 49)
        refw = torch.nn.grad.conv2d_weight(x.to(torch.bfloat16).float(), (64, cin, 7, 7), dy.float(), 2, 3)
        assert_close_f32(dw.view(64, cin, 7, 7), refw, f'stem wgrad cin={cin}')


# ------------------------------------------------------------------------------------------------
# the 3-tap "halo" kernel (conv3_tc.cuh) against the generic implicit-GEMM kernel on identical operands: the two
# differ only in fp32 summation order (taps inside / outside the channel loop), i.e. by isolated bf16 roundings
HALO_CASES = [
    # n, h, w, c  -> which conv3 instantiation it exercises
    (8, 120, 160, 64),     # BN=64, weights resident
    (8, 60, 80, 128),      # BN=128, weights resident
    (2, 60, 80, 128),      # BN=128, weight ring (few tiles)
    (4, 30, 40, 256),      # BN=256, one channel tile
    (4, 15, 20, 512),      # BN=256, two channel tiles, ragged image
    (3, 37, 53, 128),      # ragged in both axes
]


@pytest.mark.parametrize('case', HALO_CASES + [(4, 30, 40, 256, 'pair'), (4, 15, 20, 512, 'pair'), (5, 21, 40, 256, 'pair')],
                         ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('kh,kw', [(3, 1), (1, 3)])
def test_conv3_halo_kernel_matches_generic(case, kh, kw, monkeypatch):
    """'pair' cases force the cta_group::2 kernel (conv3_2cta.cuh); (5,21,40) has an odd tile count (dummy peer tile)"""
    ops = _ops()
    from emsanet_b200 import _lib
    if len(case) == 5:
        monkeypatch.setenv('EB200_CONV3_2CTA', '1')
        case = case[:4]
    n, h, w, c = case
    x = nhwc(rand_act(n, c, h, w, seed=1, relu=True))
    dy = nhwc(rand_act(n, c, h, w, seed=2))
    res = nhwc(rand_act(n, c, h, w, seed=3))
    g = torch.Generator(device='cuda').manual_seed(4)
    wt = torch.randn(c, c, kh, kw, device='cuda', generator=g) / math.sqrt(3 * c)
    bias = torch.randn(c, device='cuda', generator=g)
    pw = ops.pack_weight(wt)

    def run_all():
        out = {}
        out['plain'] = ops.conv2d(x, pw)
        out['bias_relu'] = ops.conv2d(x, pw, bias=bias, relu=True)
        st = torch.zeros(2 * c, device='cuda')
        out['stats'] = ops.conv2d(x, pw, stats=st)
        out['stats.sums'] = st
        st2 = torch.zeros(2 * c, device='cuda')
        out['dgrad_mask_stats'] = ops.conv2d_dgrad(dy, pw, tuple(x.shape), aux=x, aux_mode='mask', stats=st2)
        out['dgrad_mask_stats.sums'] = st2
        out['dgrad_add'] = ops.conv2d_dgrad(dy, pw, tuple(x.shape), aux=res, aux_mode='add')
        acc = res.clone()
        ops.conv2d_dgrad(dy, pw, tuple(x.shape), out=acc, accumulate_into_out=True)
        out['dgrad_accumulate'] = acc
        torch.cuda.synchronize()
        return out

    l0 = _lib.launch_count()
    new = run_all()
    assert _lib.launch_count() - l0 == 6
    monkeypatch.setenv('EB200_CONV3_DISABLE', '1')
    old = run_all()
    monkeypatch.delenv('EB200_CONV3_DISABLE')
    for k in new:
        a, b = new[k].float(), old[k].float()
        scale = b.abs().max().item() + 1e-12
        if k.endswith('.sums'):
            assert (a - b).abs().max().item() <= 2e-3 * scale, k
        else:
            diff = (a - b).abs()
            assert diff.max().item() <= 2 * BF16_EPS * scale, f'{k}: max diff {diff.max().item():.3e} (scale {scale:.3e})'
            frac = (diff > 0).float().mean().item()
            assert frac < 0.05, f'{k}: {100 * frac:.2f} % of the elements differ'


@pytest.mark.parametrize('case', HALO_CASES + [(2, 16, 24, 64), (2, 15, 20, 512)], ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('kh,kw', [(3, 1), (1, 3)])
@pytest.mark.parametrize('cluster', [False, True], ids=['single', 'dsmem-pair'])
def test_wgrad3_halo_kernel_matches_generic_and_torch(case, kh, kw, cluster, monkeypatch):
    """halo weight-gradient kernel (wgrad3_tc.cuh) vs the generic one (fp32 summation order only) and vs torch;
    'dsmem-pair': the opt-in variant whose CTA pairs pre-reduce their partial tiles through distributed shared memory"""
    ops = _ops()
    if cluster:
        monkeypatch.setenv('EB200_WGRAD_CLUSTER', '1')
    n, h, w, c = case
    x = rand_act(n, c, h, w, seed=5, relu=True)
    dy = rand_act(n, c, h, w, seed=6)
    new = torch.zeros(c, c, kh, kw, device='cuda')
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), new, kh, kw)
    monkeypatch.setenv('EB200_WGRAD3_DISABLE', '1')
    old = torch.zeros(c, c, kh, kw, device='cuda')
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), old, kh, kw)
    monkeypatch.delenv('EB200_WGRAD3_DISABLE')
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float(), (c, c, kh, kw), dy.float(), padding=(kh // 2, kw // 2))
    assert_close_f32(new, old, 'halo vs generic', rtol=1e-4)
    assert_close_f32(new, ref, 'halo vs torch')


@pytest.mark.parametrize('shape', [(8, 60, 80, 128), (4, 30, 40, 256), (3, 15, 20, 64), (2, 6, 5, 128)])
def test_dgrad_with_fused_batchnorm_backward(shape):
    """x --BN(train)--> ReLU --conv3x1--> y : the conv's data-gradient epilogue does the ReLU mask and both sums of the
    BatchNorm backward (EB200_BN_BWD) and eb200_bn_bwd_apply_raw finishes it; against torch autograd."""
    ops = _ops()
    n, h, w, c = shape
    x = (rand_act(n, c, h, w, seed=40).float() * 1.3 + 0.2).to(torch.bfloat16)
    dy = rand_act(n, c, h, w, seed=41)
    g = torch.Generator(device='cuda').manual_seed(42)
    wt = (torch.randn(c, c, 3, 1, device='cuda', generator=g) / math.sqrt(3 * c)).to(torch.bfloat16).float()
    gamma = torch.rand(c, device='cuda', generator=g) + 0.5
    beta = torch.randn(c, device='cuda', generator=g) * 0.3
    # reference (the ReLU output is stored as bf16 on our side: round it in the reference too)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a = F.relu(F.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5))
    y = F.conv2d(a, wt, padding=(1, 0))
    y.backward(dy.float())
    # ours
    xf = x.float()
    stats = torch.cat([xf.sum((0, 2, 3)), (xf * xf).sum((0, 2, 3))]).contiguous()
    st = ops.bn_finalize(stats, n * h * w, gamma, beta, None, None)
    ops.bn_apply(nhwc(x), st, relu=True)
    pw = ops.pack_weight(wt)
    dgamma, dbeta = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    raw = torch.zeros(2 * c, device='cuda')
    dx = ops.dgrad_with_bn_backward(nhwc(dy), pw, nhwc(x), st, gamma, raw, dgamma, dbeta)
    torch.cuda.synchronize()
    assert_close_bf16(nchw(dx), xr.grad, 'fused dgrad + bn backward dx', extra=3 * BF16_EPS)
    assert_close_f32(dgamma, gr.grad, 'dgamma', 1e-2)
    assert_close_f32(dbeta, br.grad, 'dbeta', 1e-2)


@pytest.mark.parametrize('case', [(8, 30, 40, 256), (8, 15, 20, 512), (3, 21, 40, 256)], ids=lambda c: 'x'.join(map(str, c)))
def test_sibling_pair_launch_matches_single_launches(case, monkeypatch):
    """eb200_conv2d_pair / eb200_conv2d_wgrad_pair (two problems of identical geometry in one launch, even / odd CTAs)
    against the same two problems launched one after the other — identical kernels, bit-identical conv results."""
    monkeypatch.setenv('EB200_WGRAD_DUAL', '1')      # the weight-gradient double launch is opt-in (no speed-up)
    ops = _ops()
    from emsanet_b200 import _lib
    n, h, w, c = case
    g = torch.Generator(device='cuda').manual_seed(50)
    xs = [nhwc(rand_act(n, c, h, w, seed=51 + i, relu=True)) for i in range(2)]
    dys = [nhwc(rand_act(n, c, h, w, seed=53 + i)) for i in range(2)]
    pws = [ops.pack_weight(torch.randn(c, c, 3, 1, device='cuda', generator=g) / math.sqrt(3 * c)) for _ in range(2)]
    single = [ops.conv2d(xs[i], pws[i]) for i in range(2)]
    single_dw = [torch.zeros(c, c, 3, 1, device='cuda') for _ in range(2)]
    for i in range(2):
        ops.conv2d_wgrad(dys[i], xs[i], single_dw[i], 3, 1)
    l0 = _lib.launch_count()
    descs, outs = [], []
    try:
        for i in range(2):
            ops._defer_conv = []
            outs.append(ops.conv2d(xs[i], pws[i]))
            descs.append(ops._defer_conv)
    finally:
        ops._defer_conv = None
    ops.launch_conv_descs(descs[0], descs[1])
    assert _lib.launch_count() - l0 == 1, 'the two convolutions should have shared one launch'
    pair_dw = [torch.zeros(c, c, 3, 1, device='cuda') for _ in range(2)]
    wd = []
    try:
        for i in range(2):
            ops._defer_wgrad = []
            ops.conv2d_wgrad(dys[i], xs[i], pair_dw[i], 3, 1)
            wd.append(ops._defer_wgrad)
    finally:
        ops._defer_wgrad = None
    l0 = _lib.launch_count()
    ops.launch_wgrad_descs(wd[0], wd[1])
    assert _lib.launch_count() - l0 == 1
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(outs[i], single[i]), f'conv {i}'
        assert_close_f32(pair_dw[i], single_dw[i], f'wgrad {i}', rtol=1e-4)
