"""Graph-replayed micro-benchmarks of the tensor-core kernels at config-2 layer shapes (no host launch overhead in
the timing).  Inputs rotate over ROT buffers so consecutive launches do not find their operands in L2 unless the
tensor is small enough that the previous kernel of a real step would have left it there too (ROT=1: always warm).

    python scripts/microbench.py [conv|wgrad|all] [ROT]
"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
ROT = int(sys.argv[2]) if len(sys.argv) > 2 else 4
REP = 24


def graph_time(fns):
    """fns: list of callables (one per rotation slot); returns us per call"""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            fns[i % len(fns)]()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * REP)


def main():
    torch.manual_seed(0)
    shapes = [(32, 120, 160, 64, 3, 1), (32, 120, 160, 64, 1, 3), (32, 60, 80, 128, 3, 1), (32, 60, 80, 128, 1, 3),
              (32, 30, 40, 256, 3, 1), (32, 30, 40, 256, 1, 3), (32, 15, 20, 512, 3, 1), (32, 15, 20, 512, 1, 3)]
    extra = [(32, 30, 40, 512, 256, 3, 3), (32, 60, 80, 256, 128, 3, 3), (32, 15, 20, 512, 512, 3, 3), (32, 120, 160, 128, 96, 3, 3)]
    dbg = os.environ.get('EB200_CONV_DEBUG', '0')
    for (n, h, w, c, kh, kw) in shapes:
        xs = [torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16) for _ in range(ROT)]
        dys = [torch.randn(n, h, w, c, device='cuda').to(torch.bfloat16) for _ in range(ROT)]
        wt = torch.randn(c, c, kh, kw, device='cuda') / math.sqrt(3 * c)
        bias = torch.randn(c, device='cuda')
        stats = torch.zeros(2 * c, device='cuda')
        dw = torch.zeros_like(wt)
        pw = ops.pack_weight(wt)
        outs = [torch.empty_like(xs[0]) for _ in range(ROT)]
        flops = 2.0 * n * h * w * c * c * 3
        res = []
        if what in ('conv', 'all'):
            modes = {
                'plain': lambda i: ops.conv2d(xs[i], pw, out=outs[i]),
                'bias_relu': lambda i: ops.conv2d(xs[i], pw, bias=bias, relu=True, out=outs[i]),
                'stats': lambda i: ops.conv2d(xs[i], pw, stats=stats, out=outs[i]),
                'dgrad_mask_stats': lambda i: ops.conv2d_dgrad(dys[i], pw, tuple(xs[i].shape), aux=xs[i], aux_mode='mask', stats=stats),
                'dgrad_add': lambda i: ops.conv2d_dgrad(dys[i], pw, tuple(xs[i].shape), aux=xs[i], aux_mode='add'),
            }
            for name, fn in modes.items():
                us = graph_time([(lambda i=i, fn=fn: fn(i)) for i in range(ROT)])
                res.append(f'{name} {us:6.1f}us {flops / us / 1e6:5.0f}TF')
        if what in ('wgrad', 'all'):
            us = graph_time([(lambda i=i: ops.conv2d_wgrad(dys[i], xs[i], dw, kh, kw)) for i in range(ROT)])
            res.append(f'wgrad {us:6.1f}us {flops / us / 1e6:5.0f}TF')
        print(f'dbg={dbg} rot={ROT} C={c:3d} {kh}x{kw} {h}x{w}: ' + ' | '.join(res), flush=True)
    if what in ('conv', 'all', 'wgrad'):
        for (n, h, w, ci, co, kh, kw) in extra:
            xs = [torch.randn(n, h, w, ci, device='cuda').clamp_min(0).to(torch.bfloat16) for _ in range(ROT)]
            dys = [torch.randn(n, h, w, co, device='cuda').to(torch.bfloat16) for _ in range(ROT)]
            wt = torch.randn(co, ci, kh, kw, device='cuda') / math.sqrt(9 * ci)
            stats = torch.zeros(2 * co, device='cuda')
            dw = torch.zeros_like(wt)
            pw = ops.pack_weight(wt)
            flops = 2.0 * n * h * w * ci * co * 9
            res = []
            if what != 'wgrad':
                us = graph_time([(lambda i=i: ops.conv2d(xs[i], pw, stats=stats)) for i in range(ROT)])
                res.append(f'stats {us:6.1f}us {flops / us / 1e6:5.0f}TF')
                us = graph_time([(lambda i=i: ops.conv2d_dgrad(dys[i], pw, tuple(xs[i].shape))) for i in range(ROT)])
                res.append(f'dgrad {us:6.1f}us {flops / us / 1e6:5.0f}TF')
            if what != 'conv':
                us = graph_time([(lambda i=i: ops.conv2d_wgrad(dys[i], xs[i], dw, kh, kw)) for i in range(ROT)])
                res.append(f'wgrad {us:6.1f}us {flops / us / 1e6:5.0f}TF')
            print(f'dbg={dbg} rot={ROT} {ci}->{co} {kh}x{kw} {h}x{w}: ' + ' | '.join(res), flush=True)


if __name__ == '__main__':
    main()
