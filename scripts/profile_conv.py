"""A few conv_tc_kernel / wgrad_tc_kernel launches at config-2 layer shapes for `ncu --set full`."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
torch.manual_seed(0)
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
for (n, h, w, c, kh, kw) in [(32, 30, 40, 256, 3, 1), (32, 60, 80, 128, 1, 3), (32, 120, 160, 64, 3, 1), (32, 15, 20, 512, 1, 3)]:
    x = torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16)
    dy = torch.randn(n, h, w, c, device='cuda').to(torch.bfloat16)
    wt = torch.randn(c, c, kh, kw, device='cuda') / math.sqrt(3 * c)
    bias = torch.randn(c, device='cuda')
    pw = ops.pack_weight(wt)
    stats = torch.zeros(2 * c, device='cuda')
    dw = torch.zeros_like(wt)
    for _ in range(2):
        ops.conv2d(x, pw, bias=bias, relu=True)
        ops.conv2d(x, pw, stats=stats)
        ops.conv2d_wgrad(dy, x, dw, kh, kw)
    torch.cuda.synchronize()
