import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from emsanet_b200 import ops
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
nchw = lambda x: x.permute(0, 3, 1, 2).contiguous()
bf = lambda x: x.to(torch.bfloat16).float()
torch.manual_seed(0)
n, cin, c, h, w = 4, 64, 128, 24, 32
x = bf(torch.randn(n, cin, h, w).clamp_min(0)).cuda()
for (kh, kw, s, ci) in [(3, 1, (2, 1), cin), (1, 3, (1, 2), c), (1, 1, (2, 2), cin), (3, 1, (1, 1), c)]:
    xin = x if ci == cin else bf(torch.randn(n, c, h // 2, w).clamp_min(0)).cuda()
    wt = bf(torch.randn(c, ci, kh, kw) * math.sqrt(2.0 / (ci * kh * kw))).cuda()
    pw = ops.pack_weight(wt)
    stats = torch.zeros(2 * c, device='cuda')
    y = ops.conv2d(nhwc(xin).to(torch.bfloat16), pw, s, stats=stats)
    ref = F.conv2d(xin, wt, None, s, (kh // 2, kw // 2))
    refq = bf(ref)
    yy = nchw(y).float()
    d = (yy - refq).abs()
    print(f'k{kh}x{kw} s{s}: rel vs fp32 {rel(yy, ref):.3e}  rel vs bf16(ref) {rel(yy, refq):.3e}  frac differing {float((d > 0).float().mean()):.4f}  max diff/|ref| {float((d / (refq.abs() + 1e-3)).max()):.3e}',
          ' stats sum rel', f'{rel(stats[:c], refq.sum((0, 2, 3))):.2e}', 'sumsq rel', f'{rel(stats[c:], (refq * refq).sum((0, 2, 3))):.2e}')
