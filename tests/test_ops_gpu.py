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
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), dw, kh, kw, stride)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, kh, kw), dy.float(), stride, (kh // 2, kw // 2))
    assert_close_f32(dw, ref, 'wgrad')
    # accumulation semantics
    ops.conv2d_wgrad(nhwc(dy), nhwc(x), dw, kh, kw, stride)
    torch.cuda.synchronize()
    assert_close_f32(dw, 2 * ref, 'wgrad accumulate')


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
    assert stats.abs().max().item() == 0
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
    assert sums.abs().max().item() == 0
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
