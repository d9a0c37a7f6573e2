import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
n, h, w, c, creal = 32, 240, 320, 40, 40
x = torch.randn(n, h, w, c, device='cuda').to(torch.bfloat16)
dy = torch.randn(n, 2 * h, 2 * w, c, device='cuda').to(torch.bfloat16)
wt = torch.randn(creal, 1, 3, 3, device='cuda'); b = torch.randn(creal, device='cuda')
dw, db = torch.zeros_like(wt), torch.zeros_like(b)
for _ in range(2):
    ops.upsample_dw_fwd(x, wt, b)
    ops.upsample_dw_bwd(dy, x, wt, dw, db)
torch.cuda.synchronize()
