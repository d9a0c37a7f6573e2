"""Conv main-loop ablation at config-2 layer shapes: time per launch with stores / MMA / TMA loads skipped
(needs a build with EB200_NVCC_EXTRA=-DEB200_CONV_PROBES=1; EB200_CONV_DEBUG is read once per process)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
torch.manual_seed(0)
shapes = [(32, 120, 160, 64, 3, 1), (32, 120, 160, 64, 1, 3), (32, 60, 80, 128, 1, 3), (32, 30, 40, 256, 3, 1), (32, 15, 20, 512, 1, 3)]
out = []
for (n, h, w, c, kh, kw) in shapes:
    x = torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16)
    wt = torch.randn(c, c, kh, kw, device='cuda') / math.sqrt(3 * c)
    bias = torch.randn(c, device='cuda')
    stats = torch.zeros(2 * c, device='cuda')
    pw = ops.pack_weight(wt)
    res = []
    for mode in ('plain', 'bias_relu', 'stats', 'mask_stats'):
        def run():
            if mode == 'plain': ops.conv2d(x, pw)
            elif mode == 'bias_relu': ops.conv2d(x, pw, bias=bias, relu=True)
            elif mode == 'stats': ops.conv2d(x, pw, stats=stats)
            else: ops.conv2d_dgrad(x, pw, tuple(x.shape), aux=x, aux_mode='mask', stats=stats)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        res.append(f'{mode} {e0.elapsed_time(e1) / 20 * 1e3:6.1f}')
    print(f'dbg={os.environ.get("EB200_CONV_DEBUG","0")} C={c:3d} {kh}x{kw} {h}x{w}: ' + '  '.join(res) + ' us', flush=True)
