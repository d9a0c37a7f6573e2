"""Is the C=64 halo conv's launch time address dependent?  (ncu and the in-step launch lists show the same kernel at
36-38 us on some runs and 46-48 us on others.)  One pool, input at offset 0, output at input_end + delta for a sweep of
deltas; 30 back-to-back launches per point in one CUDA-event pair (78 MB in + 78 MB out per launch: > L2 with rotation
between two input copies).  Usage: python scripts/conv64_alignment.py [C]"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from emsanet_b200 import ops  # noqa: E402


def main():
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n, h, w = (32, 120, 160) if c == 64 else (32, 60, 80)
    torch.manual_seed(0)
    numel = n * h * w * c
    pool = torch.empty(3 * numel + (64 << 20), dtype=torch.bfloat16, device='cuda')
    base = pool.data_ptr()
    x0 = pool[:numel].view(n, h, w, c)
    x0.copy_(torch.randn(n, h, w, c, device='cuda').clamp_min(0))
    x1 = pool[numel:2 * numel].view(n, h, w, c)
    x1.copy_(x0)
    wt = torch.randn(c, c, 3, 1, device='cuda') / math.sqrt(3 * c)
    bias = torch.randn(c, device='cuda')
    pw = ops.pack_weight(wt)
    res = []
    for delta in (0, 128, 256, 1024, 4096, 16384, 65536, 262144, 1 << 20, (1 << 20) + 4096, 2 << 20, (2 << 20) + 65536,
                  8 << 20, 32 << 20):
        off = 2 * numel + delta // 2
        y = pool[off:off + numel].view(n, h, w, c)
        for _ in range(3):
            ops.conv2d(x0, pw, bias=bias, relu=True, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            ops.conv2d(x0 if i & 1 else x1, pw, bias=bias, relu=True, out=y)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        res.append({'delta': delta, 'out_addr_mod_2MB': (y.data_ptr() - base) % (2 << 20), 'us': round(us, 2),
                    'TBps': round(2 * numel * 2 / us / 1e6, 3)})
        print(res[-1], flush=True)
    print(json.dumps({'c': c, 'base_mod_2MB': base % (2 << 20), 'sweep': res}))


if __name__ == '__main__':
    main()
