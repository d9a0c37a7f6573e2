"""Pin oracle/postprocessing_oracle.py against the UNMODIFIED reference post-processing classes
(/root/reference, importable only in the build container) and write tests/golden/postproc/*.npz.
TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_postproc.py

For every case the reference's PanopticPostprocessing / ScenePostprocessing run on the oracle's seeded
synthetic decoder outputs; every key of the reference result is compared with the oracle's (integer /
index outputs must be identical, floats within 1e-6) and the reference's outputs are stored as the
fixture the tests re-check on any box (the inputs are regenerated from the seed, not stored).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import postprocessing_oracle as P          # noqa: E402
from oracle.make_golden import install_reference_shim  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'postproc')

# name -> (make_inputs kwargs, network-resolution crop (y0,y1,x0,x1), fullres (h,w), nms kernel, top_k)
CASES = {
    'nyuv2_like_noresize': (dict(n=2, h=96, w=128, seed=3), (0, 96, 0, 128), (96, 128), 17, 64),
    'upscale_crop': (dict(n=2, h=80, w=112, seed=4, n_blobs=12), (8, 72, 0, 112), (131, 229), 9, 64),
    'ties_quantised_topk': (dict(n=3, h=64, w=96, seed=5, n_blobs=30, quantise=16), (0, 64, 0, 96), (48, 72), 3, 8),
    'few_classes_ragged': (dict(n=1, h=50, w=70, n_classes=13, seed=6, n_blobs=4), (0, 50, 3, 67), (75, 96), 5, 64),
}

GT_FOREGROUND_CASES = P.GT_FOREGROUND_CASES

is_thing, has_orientation, make_batch = P.golden_is_thing, P.golden_has_orientation, P.make_batch


def to_np(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def main():
    install_reference_shim()
    from nicr_mt_scene_analysis.model.postprocessing import get_postprocessing_class
    os.makedirs(OUT, exist_ok=True)
    worst = 0.0
    for name, (kw, crop, fullres, k, top_k) in CASES.items():
        inp = P.make_inputs(**kw)
        n, c = inp['semantic'].shape[:2]
        sem_pp = get_postprocessing_class('semantic')()
        ins_pp = get_postprocessing_class('instance', heatmap_threshold=0.1, heatmap_nms_kernel_size=k,
                                          heatmap_apply_foreground_mask=False, top_k_instances=top_k,
                                          normalized_offset=True, offset_distance_threshold=None)()
        pan_pp = get_postprocessing_class('panoptic', semantic_postprocessing=sem_pp,
                                          instance_postprocessing=ins_pp, semantic_classes_is_thing=is_thing(c),
                                          semantic_class_has_orientation=has_orientation(c), compute_scores=True)()
        scene_pp = get_postprocessing_class('scene')()
        batch = make_batch(crop, fullres, n)
        gt_fg = None
        if name in GT_FOREGROUND_CASES:      # dataset-evaluation branch of instance.py:365-400
            gt_fg = P.golden_instance_foreground(inp)
            batch['instance_foreground'] = gt_fg.clone()
        data = ((inp['semantic'].clone(), (inp['center'].clone(), inp['offset'].clone(),
                                           inp['orientation'].clone())), (None, None))
        with torch.no_grad():
            ref = pan_pp.postprocess(data, batch, is_training=False)
            ref_scene = scene_pp.postprocess((inp['scene'].clone(), None), batch, is_training=False)
        ora = P.panoptic_postprocess(inp['semantic'], inp['center'], inp['offset'], inp['orientation'],
                                     is_thing(c), has_orientation(c), (slice(crop[0], crop[1]),
                                                                       slice(crop[2], crop[3])), fullres,
                                     threshold=0.1, k=k, top_k=top_k, instance_foreground=gt_fg)
        ora.update(P.scene_postprocess(inp['scene']))
        ref = {**ref, **ref_scene}
        fix = {}
        for key, rv in ref.items():
            if key in ('instance_output', 'instance_side_outputs', 'semantic_side_outputs', 'instance_centers',
                       'instance_offsets', 'instance_orientation', 'semantic_output', 'scene_output'):
                continue                                             # pass-through of the inputs
            ov = ora[key]
            if isinstance(rv, list):                                  # per-image dicts
                assert len(rv) == len(ov), key
                for a, b in zip(rv, ov):
                    assert set(a.keys()) == set(b.keys()), (name, key, sorted(a), sorted(b))
                    for kk in a:
                        if isinstance(a[kk], dict):
                            for f in a[kk]:
                                x, y = a[kk][f], b[kk][f]
                                if isinstance(x, float):
                                    if not (np.isnan(x) and np.isnan(y)):
                                        assert abs(x - y) <= 1e-6 * max(1, abs(x)), (name, key, kk, f, x, y)
                                else:
                                    assert tuple(np.atleast_1d(x)) == tuple(np.atleast_1d(y)), (name, key, kk, f, x, y)
                        elif isinstance(a[kk], float):
                            assert abs(a[kk] - b[kk]) <= 1e-6, (name, key, kk, a[kk], b[kk])
                        else:
                            assert a[kk] == b[kk], (name, key, kk)
                fix[key] = np.frombuffer(json.dumps(
                    [{str(kk): vv for kk, vv in d.items()} for d in rv]).encode(), dtype=np.uint8)
                continue
            r, o = to_np(rv), to_np(ov)
            assert r.shape == o.shape, (name, key, r.shape, o.shape)
            if r.dtype.kind in 'iub':
                assert (r.astype(np.int64) == o.astype(np.int64)).all(), (name, key, int((r != o).sum()))
                small = r.astype(np.uint8) if r.max() < 256 and key != 'panoptic_segmentation_deeplab' else r
                fix[key] = small
            else:
                err = float(np.abs(r.astype(np.float64) - o).max())
                worst = max(worst, err)
                assert err <= 1e-6, (name, key, err)
                if r.ndim == 4:            # [N,C,H,W] score tensors: keep a strided sample only
                    fix[key + '__sample'] = r[:, ::5, ::7, ::9].copy()
                else:
                    fix[key] = r
        n_inst = [len(m) for m in ref['panoptic_segmentation_deeplab_instance_meta']]
        fix['meta'] = np.frombuffer(json.dumps({'inputs': kw, 'crop': crop, 'fullres': fullres, 'k': k,
                                                'top_k': top_k, 'n_instances': n_inst}).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **fix)
        print(f'{name}: {len(fix)} entries, instances per image {n_inst}, '
              f'size {os.path.getsize(os.path.join(OUT, name + ".npz")) / 1024:.0f} KB')
    print(f'oracle == reference on all integer outputs; max float deviation {worst:.3e}')


if __name__ == '__main__':
    main()
