"""last upsampling of the semantic head + NCHW boundary: unfused launches vs csrc/upsample_nchw.cu (CUDA events, L2-cold
by size: the tensors are 0.2-1.6 GB)"""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emsanet_b200 import ops

n, h, w, c = 32, 240, 320, 40
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(n, h, w, c, device='cuda', generator=g).to(torch.bfloat16)
wt = torch.randn(c, 1, 3, 3, device='cuda', generator=g)
b = torch.randn(c, device='cuda', generator=g)
gy = torch.randn(n, c, 2 * h, 2 * w, device='cuda', generator=g)
dw, db = torch.zeros_like(wt), torch.zeros_like(b)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def old_fwd():
    return ops.nhwc_to_nchw(ops.upsample_dw_fwd(x, wt, b), c)


def old_bwd():
    d = ops.nchw_grad_to_nhwc(gy, (n, 2 * h, 2 * w, c), c)
    return ops.upsample_dw_bwd(d, x, wt, dw, db)


res = {'fwd_unfused_us': timeit(old_fwd), 'fwd_fused_us': timeit(lambda: ops.upsample_dw_fwd_nchw(x, wt, b)),
       'bwd_unfused_us': timeit(old_bwd), 'bwd_fused_us': timeit(lambda: ops.upsample_dw_bwd_nchw(gy, x, wt, dw, db))}
res['fwd_fused_GBps'] = (x.numel() * 2 + gy.numel() * 4) / res['fwd_fused_us'] / 1e3
res['bwd_fused_GBps'] = (2 * x.numel() * 2 + gy.numel() * 4) / res['bwd_fused_us'] / 1e3
print(json.dumps(res))
