"""Config 5: batch-1 640x480 inference latency of the full model through the nn.Module API (CUDA-graph replay),
including the H2D copy of the inputs and the D2H copy of the results, as inference_time_whole_model.py does for the
PyTorch path (reference: inference_time_whole_model.py:297-347).

    python scripts/infer_latency.py [batch] [--with-postprocessing]

Without the flag: network only, D2H of the semantic arg-max.  With it (the reference's --with-postprocessing,
inference_time_whole_model.py:93-96,321,337-339): model(batch, do_postprocessing=True) on the GPU post-processing
(emsanet_b200/postprocessing.py) and the WHOLE result dictionary moved to the CPU.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config  # noqa: E402

torch.manual_seed(0)
argv = [a for a in sys.argv[1:] if not a.startswith('--')]
with_pp = '--with-postprocessing' in sys.argv
n = int(argv[0]) if argv else 1
model = EMSANetB200(default_args(), simple_dataset_config()).cuda().eval()
rgb_h = torch.randn(n, 3, 480, 640).pin_memory()
depth_h = torch.randn(n, 1, 480, 640).pin_memory()
# what post-processing reads from the batch (MT/data/preprocessing/resize.py:30-78): NYUv2 is 480x640, no resize
meta = [[{'type': 'Resize', 'valid_region_slice_y': slice(0, 480), 'valid_region_slice_x': slice(0, 640)}]] * n
fullres = torch.zeros(n, 3, 480, 640)


def flatten(o):
    if o is None:
        return []
    if isinstance(o, (list, tuple)):
        return [t for x in o for t in flatten(x)]
    return [o]


def to_cpu(o):
    if torch.is_tensor(o):
        return o.cpu()
    if isinstance(o, dict):
        return {k: to_cpu(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(to_cpu(v) for v in o)
    return o


def once(move_all=True):
    with torch.no_grad():
        batch = {'rgb': rgb_h.cuda(non_blocking=True), 'depth': depth_h.cuda(non_blocking=True)}
        if with_pp:
            batch.update({'_applied_preprocessing': meta, 'rgb_fullres': fullres})
            r = model(batch, do_postprocessing=True)
            if move_all:
                return to_cpu(r)
            torch.cuda.synchronize()       # results where the post-processing leaves them (panoptic maps on the CPU)
            return r
        sem = flatten(model(batch))[0]
        return sem.argmax(1).to(torch.uint8).cpu()


for _ in range(5):
    once()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    t0 = time.perf_counter()
    once()
    ts.append(time.perf_counter() - t0)
ts.sort()
extra = {}
if with_pp:
    t2 = []
    for _ in range(30):
        t0 = time.perf_counter()
        once(move_all=False)
        t2.append(time.perf_counter() - t0)
    t2.sort()
    extra = {'latency_ms_median_results_in_place': 1e3 * t2[len(t2) // 2],
             'note': 'the whole-dict variant copies ~300 MB of pageable fp32 score / logit tensors per image to the host '
                     '(inference_time_whole_model.py:337-339); in place = dense fp32 tensors stay on the device, '
                     'panoptic maps and meta dictionaries on the CPU like the reference'}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    batch = {'rgb': rgb_h.cuda(), 'depth': depth_h.cuda()}
    e0.record()
    for _ in range(20):
        model(batch)
    e1.record()
    torch.cuda.synchronize()
print(json.dumps({'config': f'full EMSANet RGB-D r34-NBt1D eval, batch {n}, 640x480, bf16'
                            + (', GPU post-processing, whole result dict to CPU' if with_pp else ''),
                  'latency_ms_median_incl_h2d_d2h': 1e3 * ts[len(ts) // 2], 'latency_ms_min': 1e3 * ts[0],
                  'device_ms_per_forward': e0.elapsed_time(e1) / 20, 'fps': n / ts[len(ts) // 2], **extra}))
