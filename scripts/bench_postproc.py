"""Time the GPU post-processing chain at the bench batch and the CPU restatement beside it.

    python scripts/bench_postproc.py [--batch 32] [--iters 10]

Prints one JSON object: ms per batch for the whole PanopticPostprocessingB200.postprocess call (device-resident
result maps, and with the reference's CPU placement of the panoptic maps), the softmax/arg-max kernel alone against
the HBM roofline (algorithmic bytes = logits read once + scores written once + 17 B/pixel of maps), and the oracle
port on the host cores for a 2-image sample.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emsanet_b200 import postprocessing as pp   # noqa: E402
from oracle import postprocessing_oracle as P   # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--ncu', action='store_true', help='one call inside a cudaProfilerStart/Stop range (ncu --profile-from-start off)')
    a = ap.parse_args()
    h, w, c = 480, 640, 40
    base = P.make_inputs(2, h, w, seed=11, n_blobs=40)
    reps = a.batch // 2
    d = {k: v.repeat(reps, *([1] * (v.ndim - 1))).cuda() for k, v in base.items()}
    n = d['semantic'].shape[0]
    batch = P.make_batch((0, h, 0, w), (h, w), n, device='cuda')
    data = ((d['semantic'], (d['center'], d['offset'], d['orientation'])), (None, None))
    res = {'batch': n, 'resolution': [h, w], 'classes': c, 'EB200_PP_LD': os.environ.get('EB200_PP_LD', 'default')}
    if a.ncu:
        sem = pp.SemanticPostprocessingB200()
        ins = pp.InstancePostprocessingB200(heatmap_threshold=0.1, heatmap_nms_kernel_size=17, top_k_instances=64)
        pan = pp.PanopticPostprocessingB200(sem, ins, P.golden_is_thing(c), P.golden_has_orientation(c),
                                            compute_scores=True, mirror_host_placement=False)
        pan.postprocess(data, batch, is_training=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        pan.postprocess(data, batch, is_training=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    for mirror in (False, True):
        sem = pp.SemanticPostprocessingB200()
        ins = pp.InstancePostprocessingB200(heatmap_threshold=0.1, heatmap_nms_kernel_size=17, top_k_instances=64)
        pan = pp.PanopticPostprocessingB200(sem, ins, P.golden_is_thing(c), P.golden_has_orientation(c),
                                            compute_scores=True, mirror_host_placement=mirror)
        ms = timed(lambda: pan.postprocess(data, batch, is_training=False), a.iters)
        res['panoptic_ms_per_batch_' + ('cpu_placement' if mirror else 'device_maps')] = ms
        res['panoptic_images_per_s_' + ('cpu_placement' if mirror else 'device_maps')] = n / ms * 1e3
        t0 = time.perf_counter()
        for _ in range(a.iters):
            pan.postprocess(data, batch, is_training=False)
        torch.cuda.synchronize()
        res['panoptic_wall_ms_' + ('cpu_placement' if mirror else 'device_maps')] = (time.perf_counter() - t0) / a.iters * 1e3
    # the C-ABI calls one by one (device time, CUDA events)
    flags = torch.tensor([int(t) | (int(o) << 1) for t, o in zip(P.golden_is_thing(c), P.golden_has_orientation(c))],
                         dtype=torch.uint8, device='cuda')
    _, scores, _, sem_idx, fg = pp.softmax_argmax(d['semantic'], cls_flags=flags)
    tab = pp.InstanceTables(n, 'cuda', c)
    stages = {}
    stages['instance_centers_ms'] = timed(lambda: pp.instance_centers(d['center'], tab, 0.1, 17, 64), a.iters)
    stages['instance_assign_ms'] = timed(lambda: pp.instance_assign(d['offset'], fg, tab, float(h), float(w), None,
                                                                    sem_idx, c), a.iters)
    seg = pp.instance_assign(d['offset'], fg, tab, float(h), float(w), None, sem_idx, c)
    stages['panoptic_merge_ms'] = timed(lambda: pp.panoptic_merge(seg, sem_idx, flags, tab, scores, d['orientation'], c),
                                        a.iters)
    stages['instances_per_image'] = tab.counts.float().mean().item()
    res['stages'] = stages
    ms = timed(lambda: pp.softmax_argmax(d['semantic']), a.iters)
    bytes_ = n * h * w * (2 * c * 4 + 4 + 8)
    peak, peak_source = 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        peak_source = 'MEASURED_PEAKS.json hbm_gbs'
    except Exception:
        pass
    res['softmax_argmax'] = {'ms': ms, 'algorithmic_GB': bytes_ / 1e9, 'achieved_GBps': bytes_ / ms / 1e6,
                             'peak_GBps': peak, 'peak_source': peak_source, 'frac': bytes_ / ms / 1e6 / peak}
    t0 = time.perf_counter()
    P.panoptic_postprocess(base['semantic'], base['center'], base['offset'], base['orientation'],
                           P.golden_is_thing(c), P.golden_has_orientation(c), (slice(0, h), slice(0, w)), (h, w))
    dt = time.perf_counter() - t0
    res['cpu_port'] = {'images_per_s': 2 / dt, 'cores': torch.get_num_threads(),
                       'sample': '2 images 480x640, oracle port (numpy/torch CPU)'}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
