"""Where the module-API (e2e) step spends its time: CUDA-event segments around H2D, forward, loss, backward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
torch.manual_seed(0)
model = EMSANetB200(default_args(), simple_dataset_config()).cuda().train()
N = 32
rgb_h = torch.randn(N, 3, 480, 640).pin_memory(); depth_h = torch.randn(N, 1, 480, 640).pin_memory()
def flatten(o):
    if o is None: return []
    if isinstance(o, (list, tuple)): return [t for x in o for t in flatten(x)]
    return [o]
def ev(): 
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(6):
    t0 = time.perf_counter(); e0 = ev()
    batch = {'rgb': rgb_h.cuda(non_blocking=True), 'depth': depth_h.cuda(non_blocking=True)}
    e1 = ev(); out = model(batch); e2 = ev()
    loss = sum((o.float() ** 2).mean() for o in flatten(out)); e3 = ev()
    for p in model.parameters(): p.grad = None
    loss.backward(); e4 = ev()
    l = float(loss.item()); e5 = ev(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if it >= 3:
        print(f'h2d {e0.elapsed_time(e1):6.2f}  fwd {e1.elapsed_time(e2):6.2f}  loss {e2.elapsed_time(e3):6.2f}  bwd(total) {e3.elapsed_time(e4):6.2f}  item {e4.elapsed_time(e5):5.2f}  | gpu total {e0.elapsed_time(e5):6.2f} ms, wall {1e3*(t1-t0):6.2f} ms')
