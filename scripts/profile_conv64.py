"""C=64 120x160 N=32 3x1 conv (bias+ReLU), the layer class furthest from its roofline, for `ncu --set full --import-source on`."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
torch.manual_seed(0)
n, h, w, c = 32, 120, 160, int(sys.argv[1]) if len(sys.argv) > 1 else 64
if c == 128:
    h, w = 60, 80
x = torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16)
wt = torch.randn(c, c, 3, 1, device='cuda') / math.sqrt(3 * c)
bias = torch.randn(c, device='cuda')
pw = ops.pack_weight(wt)
for _ in range(3):
    ops.conv2d(x, pw, bias=bias, relu=True)
torch.cuda.synchronize()
