"""Graph-replayed timing of the learned-upsampling kernels at config-2 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import graph_time  # noqa (runs nothing: guarded below)
torch.manual_seed(0)
for (n, h, w, c, creal) in [(32, 240, 320, 40, 40), (32, 120, 160, 40, 40), (32, 240, 320, 8, 5), (32, 120, 160, 8, 5),
                            (32, 60, 80, 128, 128), (32, 30, 40, 256, 256), (32, 15, 20, 512, 512)]:
    x = torch.randn(n, h, w, c, device='cuda').to(torch.bfloat16)
    dy = torch.randn(n, 2 * h, 2 * w, c, device='cuda').to(torch.bfloat16)
    wt = torch.randn(creal, 1, 3, 3, device='cuda')
    b = torch.randn(creal, device='cuda')
    dw, db = torch.zeros_like(wt), torch.zeros_like(b)
    t_f = graph_time([lambda: ops.upsample_dw_fwd(x, wt, b)])
    t_b = graph_time([lambda: ops.upsample_dw_bwd(dy, x, wt, dw, db)])
    from emsanet_b200 import _lib
    dxb = torch.empty_like(x)
    t_bi = graph_time([lambda: _lib.call('eb200_upsample_dw_bwd_input', dy.data_ptr(), wt.data_ptr(), dxb.data_ptr(), n, h, w, c, creal, ops._stream())])
    mb_f = (x.numel() + dy.numel()) * 2 / 1e6
    print(f'C={c:3d} {h}x{w}: fwd {t_f:7.1f} us ({mb_f / t_f:5.2f} TB/s)   bwd(in+w) {t_b:7.1f} us ({(2 * dy.numel() + 2 * x.numel()) * 2 / 1e6 / t_b:5.2f} TB/s)  of which bwd_input {t_bi:7.1f} us', flush=True)
