"""The reference's OWN data pipeline and model in front of the emsanet_b200 post-processing mirrors, on the synthetic
mini-NYUv2.  Runs only where the reference checkout exists (the build container):

  synthetic_nyuv2.write_dataset -> emsanet.data.get_datahelper / emsanet.preprocessing.get_preprocessor (Resize to
  384x512, full resolution 480x640) -> a real validation batch (AppliedPreprocessingMeta, *_fullres, ground-truth
  foreground masks) -> the UNMODIFIED reference EMSANet on CPU with do_postprocessing=True
  vs. the mirror classes (C-ABI calls replaced by their CPU restatements) on the same network outputs and batch.

This pins what unit fixtures cannot: that the mirrors read the batch exactly as the reference's data pipeline
produces it (meta format, key names, dtypes of the ground-truth masks, full-resolution shapes).
"""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout only exists in the build container')


def _args(root, **over):
    a = dict(dataset='nyuv2', dataset_path=root, split='train', validation_split='test',
             tasks=('semantic', 'scene', 'instance', 'orientation'), input_modalities=('rgb', 'depth'), raw_depth=False,
             cache_dataset=False, batch_size=2, validation_batch_size=2, n_workers=0, subset_train=1.0,
             subset_deterministic=False, hypersim_subsample=1, hypersim_use_old_depth_stats=False, scannet_subsample=50,
             validation_scannet_subsample=100, aug_scale_min=1.0, aug_scale_max=1.4, debug=False, input_height=384,
             input_width=512, instance_center_sigma=8, instance_no_multiscale_supervision=False,
             instance_offset_encoding='tanh', normal_no_multiscale_supervision=False, scannet_semantic_n_classes=40,
             semantic_no_multiscale_supervision=False, validation_full_resolution=False, validation_input_height=384,
             validation_input_width=512, validation_scannet_benchmark_mode=False, visualize_validation=False,
             enable_panoptic=True, use_original_scene_labels=False, sunrgbd_depth_do_not_force_mm=False,
             sunrgbd_instances_version='panopticndt')
    a.update(over)
    return argparse.Namespace(**a)


@pytest.fixture(scope='module')
def reference_batch_and_model(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import make_golden as mg
    from oracle import emsanet_oracle as O
    mg.install_reference_shim()
    import emsanet.data as D
    import emsanet.preprocessing as PP
    from emsanet.model import EMSANet
    from emsanet_b200 import synthetic_nyuv2 as S
    root = str(tmp_path_factory.mktemp('nyuv2_synth'))
    assert S.write_dataset(root, n_train=4, n_test=3, seed=0) == (4, 3)
    args = _args(root)
    helper = D.get_datahelper(args)
    helper.set_valid_preprocessor(PP.get_preprocessor(args, dataset=helper.datasets_valid[0], phase='test',
                                                      multiscale_downscales=None))
    batch = next(iter(helper.valid_dataloaders[0]))
    cfg = O.OracleConfig(backbone='resnet18')
    torch.manual_seed(0)
    model = EMSANet(mg.make_args(cfg, 384, 512), helper.datasets_valid[0].config).eval()
    return batch, model, helper.datasets_valid[0]


def test_synthetic_dataset_feeds_the_reference_pipeline(reference_batch_and_model):
    batch, _, ds = reference_batch_and_model
    assert len(ds) == 3
    assert tuple(batch['rgb'].shape) == (2, 3, 384, 512) and batch['rgb'].dtype == torch.float32
    assert tuple(batch['depth'].shape) == (2, 1, 384, 512)
    assert tuple(batch['rgb_fullres'].shape) == (2, 3, 480, 640)
    assert batch['semantic'].dtype == torch.uint8 and int(batch['semantic'].max()) <= 40
    assert batch['instance_foreground'].dtype == torch.bool and bool(batch['instance_foreground'].any())
    assert bool(batch['orientation_foreground'].any())
    resize = [p for p in batch['_applied_preprocessing'][0] if p['type'] == 'Resize'][0]
    assert (resize['old_height'], resize['old_width'], resize['new_height'], resize['new_width']) == (480, 640, 384, 512)


def test_mirrors_on_a_real_validation_batch_match_the_reference(reference_batch_and_model, monkeypatch):
    batch, model, _ = reference_batch_and_model
    from oracle import postproc_abi_oracle as A
    from emsanet_b200 import postprocessing as pp
    for fn in ('softmax_argmax', 'nearest_resize', 'instance_centers', 'instance_assign', 'panoptic_merge',
               'orientation_sums'):
        monkeypatch.setattr(pp, fn, getattr(A, fn))
    monkeypatch.setattr(pp, '_dev', lambda t, dtype, what: t.detach().to(dtype).contiguous())
    with torch.no_grad():
        raw = model(batch, do_postprocessing=False)
        ref = model(batch, do_postprocessing=True)
    # the mirrors, built from the reference objects' own settings
    pan_ref = model.decoders['panoptic_helper'].postprocessing
    sem, ins = pp.build_for(pan_ref._semantic_postprocessing), pp.build_for(pan_ref._instance_postprocessing)
    n_cls = 40
    is_thing, has_or = np.zeros(n_cls, bool), np.zeros(n_cls, bool)
    is_thing[pan_ref._thing_class_ids] = True
    has_or[pan_ref._orientation_ids - 1] = True
    pan = pp.PanopticPostprocessingB200(sem, ins, tuple(is_thing), tuple(has_or),
                                        normalized_offset=pan_ref._normalized_offset,
                                        compute_scores=pan_ref._compute_scores)
    scene = pp.ScenePostprocessingB200()
    got = {**pan.postprocess(raw[0], batch, is_training=False), **scene.postprocess(raw[1], batch, is_training=False)}
    skipped = set()
    assert set(ref) == set(got), (set(ref) - set(got), set(got) - set(ref))
    assert sum(len(d) for d in ref['orientations_gt_instance_gt_orientation_foreground']) > 0
    n_checked = 0
    for key, want in ref.items():
        if key in skipped:
            continue
        have = got[key]
        if isinstance(want, torch.Tensor):
            assert tuple(have.shape) == tuple(want.shape) and have.dtype == want.dtype, (key, have.dtype, want.dtype)
            if want.dtype.is_floating_point:
                assert float((have - want).abs().max()) <= 1e-5, key
            else:
                assert torch.equal(have, want), key
            n_checked += 1
        elif isinstance(want, list) and want and isinstance(want[0], dict):
            for g, w in zip(have, want):
                assert set(g) == set(w), key
                for k in w:
                    if isinstance(w[k], dict):
                        for f, x in w[k].items():
                            y = g[k][f]
                            if isinstance(x, float):
                                assert (np.isnan(x) and np.isnan(y)) or abs(x - y) <= 1e-4, (key, k, f, x, y)
                            else:
                                assert tuple(np.atleast_1d(x)) == tuple(np.atleast_1d(y)), (key, k, f)
                    elif isinstance(w[k], float):
                        assert abs(w[k] - g[k]) <= 1e-4, (key, k)
                    else:
                        assert w[k] == g[k], (key, k)
            n_checked += 1
    assert n_checked >= 25
    assert tuple(got['semantic_segmentation_idx_fullres'].shape) == (2, 480, 640)
    assert tuple(got['panoptic_segmentation_deeplab_fullres'].shape) == (2, 480, 640)
    assert tuple(got['instance_segmentation_gt_foreground_fullres'].shape) == (2, 480, 640)
    # the comparison is only meaningful if the random network produced instances at all
    n_inst = [len(m) for m in ref['panoptic_segmentation_deeplab_instance_meta']]
    n_gt = [len(m) for m in ref['instance_segmentation_gt_meta']]
    print('instances per image (panoptic / gt foreground):', n_inst, n_gt)
    assert sum(n_inst) > 0 and sum(n_gt) > 0


@pytest.mark.parametrize('apply_fg,dist_thr', [(True, None), (False, 6), (True, 4)])
def test_instance_options_match_the_reference(apply_fg, dist_thr, monkeypatch):
    """heatmap_apply_foreground_mask and offset_distance_threshold (added after the EMSANet release,
    instance.py:47-49,139-140,236-238) against the unmodified reference class: the oracle, and the mirror on the
    restated C-ABI calls"""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import make_golden as mg
    mg.install_reference_shim()
    from nicr_mt_scene_analysis.model.postprocessing import get_postprocessing_class
    from oracle import postprocessing_oracle as P
    from oracle import postproc_abi_oracle as A
    from emsanet_b200 import postprocessing as pp
    inp = P.make_inputs(2, 80, 112, seed=31, n_blobs=14, quantise=32)
    fg = P.golden_instance_foreground(inp)
    h, w = 80, 112
    off = inp['offset'].clone()
    off[:, 0] *= h
    off[:, 1] *= w
    kw = dict(heatmap_threshold=0.1, heatmap_nms_kernel_size=5, heatmap_apply_foreground_mask=apply_fg,
              top_k_instances=16, normalized_offset=True, offset_distance_threshold=dist_thr)
    ref = get_postprocessing_class('instance', **kw)()
    seg_ref, meta_ref = ref._get_instance_segmentation(inp['center'].clone(), off.clone(), fg.clone())
    seg_o, meta_o = P.instance_segmentation(inp['center'], off, fg, 0.1, 5, 16, apply_foreground_mask=apply_fg,
                                            distance_threshold=dist_thr)
    assert np.array_equal(seg_ref.numpy(), seg_o)
    assert [sorted(m) for m in meta_ref] == [sorted(m) for m in meta_o]
    for fn in ('instance_centers', 'instance_assign'):
        monkeypatch.setattr(pp, fn, getattr(A, fn))
    monkeypatch.setattr(pp, '_dev', lambda t, dtype, what: t.detach().to(dtype).contiguous())
    mirror = pp.build_for(ref)
    seg_m, meta_m = mirror._get_instance_segmentation(inp['center'], inp['offset'], fg)   # the mirror scales itself
    assert torch.equal(seg_m, seg_ref)
    for m_ref, m_m in zip(meta_ref, meta_m):
        assert set(m_ref) == set(m_m)
        for i in m_ref:
            assert m_ref[i]['center_yx'] == m_m[i]['center_yx'] and m_ref[i]['area'] == m_m[i]['area']
            assert m_ref[i]['score'] == m_m[i]['score']
    assert sum(len(m) for m in meta_ref) > 0
    if dist_thr is not None:
        assert bool(((seg_ref == 0) & fg).any())          # the threshold really un-assigns some foreground pixels
