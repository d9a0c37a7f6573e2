"""One profiled training step (config 2) between cudaProfilerStart/Stop, for ncu --profile-from-start off.
Also prints host-side (launch) time vs device time of a step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
from emsanet_b200.patch import _engine_for
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
model = EMSANetB200(default_args(), simple_dataset_config()).cuda().train()
with torch.no_grad():
    for k, p in model.named_parameters():
        if k.endswith('norm2.weight'):
            p.fill_(0.15)
eng = _engine_for(model)
eng.force_repack = True
rgb = torch.randn(n, 3, 480, 640, device='cuda'); depth = torch.randn(n, 1, 480, 640, device='cuda')
def step():
    res = eng.forward(rgb, depth, True)
    gouts = {t: [o * (2.0 / o.numel()) for o in outs] for t, outs in res.items()}
    eng.backward(gouts)
for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'host enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms')
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
