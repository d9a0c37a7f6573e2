"""Config 5: batch-1 640x480 inference latency of the full model through the nn.Module API (CUDA-graph replay),
including the H2D copy of the inputs and a D2H read of the semantic arg-max, as inference_time_whole_model.py does
for the PyTorch path (reference: inference_time_whole_model.py:297-347)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model = EMSANetB200(default_args(), simple_dataset_config()).cuda().eval()
rgb_h = torch.randn(n, 3, 480, 640).pin_memory(); depth_h = torch.randn(n, 1, 480, 640).pin_memory()
def flatten(o):
    if o is None: return []
    if isinstance(o, (list, tuple)): return [t for x in o for t in flatten(x)]
    return [o]
def once():
    with torch.no_grad():
        out = model({'rgb': rgb_h.cuda(non_blocking=True), 'depth': depth_h.cuda(non_blocking=True)})
        sem = flatten(out)[0]
        return sem.argmax(1).to(torch.uint8).cpu()
for _ in range(5): once()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    t0 = time.perf_counter(); once(); ts.append(time.perf_counter() - t0)
ts.sort()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    batch = {'rgb': rgb_h.cuda(), 'depth': depth_h.cuda()}
    e0.record()
    for _ in range(20): model(batch)
    e1.record(); torch.cuda.synchronize()
print(json.dumps({'config': f'full EMSANet RGB-D r34-NBt1D eval, batch {n}, 640x480, bf16', 'latency_ms_median_incl_h2d_d2h': 1e3 * ts[len(ts) // 2],
                  'latency_ms_min': 1e3 * ts[0], 'device_ms_per_forward': e0.elapsed_time(e1) / 20, 'fps': n / ts[len(ts) // 2]}))
