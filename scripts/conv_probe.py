import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
torch.manual_seed(0)
for (n, h, w, c, kh, kw) in [(32, 120, 160, 64, 3, 1), (32, 60, 80, 128, 1, 3), (32, 30, 40, 256, 3, 1)]:
    x = torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16)
    wt = torch.randn(c, c, kh, kw, device='cuda') / math.sqrt(3 * c)
    pw = ops.pack_weight(wt)
    for _ in range(3):
        ops.conv2d(x, pw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv2d(x, pw)
    e1.record(); torch.cuda.synchronize()
    print(f'C={c}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us', end='   ')
print()
