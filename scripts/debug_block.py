import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from emsanet_b200 import ops
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
nchw = lambda x: x.permute(0, 3, 1, 2).contiguous()
torch.manual_seed(0)
n, c, h, w = 4, 64, 24, 32
dev = 'cuda'
# --- dgrad alone, rel-L2
wt = (torch.randn(c, c, 1, 3, device=dev) / math.sqrt(3 * c)).to(torch.bfloat16).float()
dy = torch.randn(n, c, h, w, device=dev).to(torch.bfloat16)
pw = ops.pack_weight(wt)
dx = ops.conv2d_dgrad(nhwc(dy), pw, (n, h, w, c))
ref = torch.nn.grad.conv2d_input((n, c, h, w), wt, dy.float(), 1, (0, 1))
print('dgrad 1x3 rel', rel(nchw(dx).float(), ref))
mask = torch.randn(n, c, h, w, device=dev).clamp_min(0).to(torch.bfloat16)
st = torch.zeros(2 * c, device=dev)
dxm = ops.conv2d_dgrad(nhwc(dy), pw, (n, h, w, c), aux=nhwc(mask), aux_mode='mask', stats=st)
print('dgrad mask rel', rel(nchw(dxm).float(), ref * (mask > 0)), 'bias-sum rel', rel(st[:c], (ref * (mask > 0)).sum((0, 2, 3))))
# --- wgrad rel
x = torch.randn(n, c, h, w, device=dev).clamp_min(0).to(torch.bfloat16)
dw = torch.zeros(c, c, 1, 3, device=dev)
ops.conv2d_wgrad(nhwc(dy), nhwc(x), dw, 1, 3)
print('wgrad rel', rel(dw, torch.nn.grad.conv2d_weight(x.float(), (c, c, 1, 3), dy.float(), 1, (0, 1))))
# --- BN backward rel-L2
xb = (torch.randn(n, c, h, w, device=dev) * 1.3 + 0.4).to(torch.bfloat16)
gamma = torch.rand(c, device=dev) + 0.5; beta = torch.randn(c, device=dev) * 0.1
xr = xb.float().requires_grad_(True)
y = F.relu(F.batch_norm(xr, None, None, gamma, beta, True, 0.1, 1e-5))
y.backward(dy.float())
xf = xb.float()
stats = torch.cat([xf.sum((0, 2, 3)), (xf * xf).sum((0, 2, 3))]).contiguous()
bst = ops.bn_finalize(stats, n * h * w, gamma, beta, None, None)
out = ops.bn_apply(nhwc(xb), bst, relu=True)
print('bn fwd rel', rel(nchw(out).float(), y))
sums = torch.zeros(2 * c, device=dev); dg = torch.zeros(c, device=dev); db = torch.zeros(c, device=dev)
dxb, _ = ops.bn_backward(nhwc(dy), nhwc(xb), bst, gamma, sums, relu_mode=1, mask_src=out, dgamma=dg, dbeta=db)
print('bn bwd dx rel', rel(nchw(dxb).float(), xr.grad))
# manual formula with the same inputs
g = dy.float() * (y > 0)
mu = xf.mean((0, 2, 3), keepdim=True); var = xf.var((0, 2, 3), unbiased=False, keepdim=True); rstd = (var + 1e-5).rsqrt()
xh = (xf - mu) * rstd
man = gamma[None, :, None, None] * rstd * (g - g.mean((0, 2, 3), keepdim=True) - xh * (g * xh).mean((0, 2, 3), keepdim=True))
print('manual vs autograd', rel(man, xr.grad), ' ours vs manual', rel(nchw(dxb).float(), man))
print('sums rel', rel(sums[:c], g.sum((0, 2, 3))), rel(sums[c:], (g * xh).sum((0, 2, 3))), 'dgamma', rel(dg, (g * xh).sum((0, 2, 3))))
print('mean/rstd rel', rel(bst.mean, mu.flatten()), rel(bst.rstd, rstd.flatten()))
