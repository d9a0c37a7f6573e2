"""Same-box A/B of single halo-conv launches between two source trees (EB200_TREE=<path of a checkout with a built lib>,
default this one): 3x1 / 1x3 conv bias+ReLU and conv + BN statistics at the four config-2 layer classes, 30 back-to-back
launches per point in one CUDA-event pair, inputs alternating between two copies."""
import json
import math
import os
import sys

tree = os.environ.get('EB200_TREE') or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, tree)
import torch  # noqa: E402

from emsanet_b200 import _lib  # noqa: E402
if os.environ.get('EB200_FORCE_LIB'):       # an older tree's Python with another build of the library
    _lib.LIB_PATH = os.environ['EB200_FORCE_LIB']
    _lib.load.__defaults__ = (_lib.LIB_PATH,)
from emsanet_b200 import ops  # noqa: E402


def main():
    torch.manual_seed(0)
    res = {}
    for c, h, w, k in ((64, 120, 160, (3, 1)), (128, 60, 80, (1, 3)), (256, 30, 40, (3, 1)), (512, 15, 20, (1, 3))):
        n = 32
        xs = [torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16) for _ in range(2)]
        wt = torch.randn(c, c, *k, device='cuda') / math.sqrt(3 * c)
        bias = torch.randn(c, device='cuda')
        pw = ops.pack_weight(wt)
        y = torch.empty_like(xs[0])
        stats = torch.zeros(2 * c, device='cuda')
        for name, kw in (('bias_relu', dict(bias=bias, relu=True)), ('stats', dict(stats=stats))):
            for _ in range(3):
                ops.conv2d(xs[0], pw, out=y, **kw)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()          # graph replay: no host launch overhead between the 30 kernels
            with torch.cuda.graph(g):
                for i in range(30):
                    ops.conv2d(xs[i & 1], pw, out=y, **kw)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            res[f'C{c}_{name}'] = round(e0.elapsed_time(e1) / 30 * 1e3, 2)
    print(json.dumps({'tree': tree, 'lib': _lib.LIB_PATH, 'us': res}))


if __name__ == '__main__':
    main()
