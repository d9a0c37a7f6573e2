"""1-CTA vs CTA-pair (cta_group::2) halo conv at growing problem sizes (C=256 / 512)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
from microbench import graph_time
torch.manual_seed(0)
for (n, h, w, c) in [(32, 30, 40, 256), (32, 60, 80, 256), (8, 96, 128, 256), (32, 15, 20, 512), (32, 30, 40, 512), (8, 48, 64, 512)]:
    xs = [torch.randn(n, h, w, c, device='cuda').clamp_min(0).to(torch.bfloat16) for _ in range(4)]
    wt = torch.randn(c, c, 3, 1, device='cuda') / math.sqrt(3 * c)
    pw = ops.pack_weight(wt)
    outs = [torch.empty_like(xs[0]) for _ in range(4)]
    flops = 2.0 * n * h * w * c * c * 3
    res = []
    for mode in ('1cta', '2cta'):
        if mode == '2cta': os.environ['EB200_CONV3_2CTA'] = '1'
        else: os.environ.pop('EB200_CONV3_2CTA', None)
        us = graph_time([(lambda i=i: ops.conv2d(xs[i], pw, out=outs[i])) for i in range(4)])
        res.append(f'{mode} {us:6.1f}us {flops / us / 1e6:5.0f}TF')
    print(f'C={c} {h}x{w} N={n}: ' + ' | '.join(res), flush=True)
