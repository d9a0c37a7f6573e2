"""Inference post-processing.

not-gpu:  * the CPU oracle (oracle/postprocessing_oracle.py) against the golden fixtures made from the UNMODIFIED
            reference classes (oracle/make_golden_postproc.py) — integer outputs identical, floats <= 1e-6;
          * the host side of emsanet_b200/postprocessing.py (key set, table unpacking, meta dictionaries, placement)
            with the five C-ABI calls replaced by their CPU restatements (oracle/postproc_abi_oracle.py).
gpu:      * every eb200_pp_* C-ABI call against its restatement on the same inputs;
          * the mirror classes end to end against the golden fixtures (integer maps identical; tolerances below);
          * a 480x640 batch against the oracle, plus size-independent properties at the bench size.

Tolerances (floats): softmax scores 2e-6 abs, resampled logits 2e-5 abs (|logit| ~ 10, FMA contraction differs
between ATen's AVX2 kernels and nvcc), per-instance mean scores 1e-5, orientation angles 1e-4 rad.
Index maps: identical; up to MAX_FLIPS pixels per map are tolerated and reported where the two best candidates
are closer than float rounding (1 ulp differences of expf between Sleef and CUDA).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import postprocessing_oracle as P
from oracle import postproc_abi_oracle as A

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'postproc')
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz'))
MAX_FLIPS = 3
FLOAT_TOL = {'semantic_output_fullres': 2e-5, 'default': 2e-6,
             'panoptic_segmentation_deeplab_panoptic_score': 1e-5,
             'panoptic_segmentation_deeplab_panoptic_score_fullres': 1e-5}
PASS_THROUGH = ('instance_output', 'instance_side_outputs', 'semantic_side_outputs', 'instance_centers',
                'instance_offsets', 'instance_orientation', 'semantic_output', 'scene_output')


def _load(name):
    fix = dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))
    meta = json.loads(bytes(fix.pop('meta')).decode())
    inp = P.make_inputs(**meta['inputs'])
    return fix, meta, inp


def _np(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def _compare_dict_lists(key, got, want, report):
    assert len(got) == len(want), key
    for b, (g, w) in enumerate(zip(got, want)):
        g = {str(k): v for k, v in g.items()}
        assert set(g) == set(w), (key, b, sorted(g), sorted(w))
        for k in w:
            if isinstance(w[k], dict):
                assert set(g[k]) == set(w[k]), (key, b, k, sorted(g[k]), sorted(w[k]))
                for f, x in w[k].items():
                    y = g[k][f]
                    if isinstance(x, float):
                        if np.isnan(x):
                            assert np.isnan(y), (key, b, k, f)
                        else:
                            tol = 0.0 if f == 'score' else (1e-4 if f == 'orientation' else 1e-5)
                            assert abs(x - y) <= tol, (key, b, k, f, x, y)
                            report[f'{key}.{f}'] = max(report.get(f'{key}.{f}', 0.0), abs(x - y))
                    else:
                        assert list(np.atleast_1d(x)) == list(np.atleast_1d(y)), (key, b, k, f, x, y)
            elif isinstance(w[k], float):
                assert abs(w[k] - g[k]) <= 1e-4, (key, b, k, w[k], g[k])
            else:
                assert w[k] == g[k], (key, b, k)


def _compare(result, fix, exact=False):
    """result of a post-processing run vs a golden fixture; returns a report of the deviations"""
    report = {}
    for fkey, want in fix.items():
        sample = fkey.endswith('__sample')
        key = fkey[:-8] if sample else fkey
        assert key in result, f'missing key {key}'
        if want.dtype == np.uint8 and want.ndim == 1 and isinstance(result[key], list):
            _compare_dict_lists(key, result[key], json.loads(bytes(want).decode()), report)
            continue
        got = _np(result[key])
        if sample:
            got = got[:, ::5, ::7, ::9]
        assert got.shape == want.shape, (key, got.shape, want.shape)
        if want.dtype.kind in 'iub':
            bad = int((got.astype(np.int64) != want.astype(np.int64)).sum())
            report[key] = bad
            assert bad <= (0 if exact else MAX_FLIPS), f'{key}: {bad} of {want.size} entries differ'
        else:
            err = float(np.abs(got.astype(np.float64) - want).max())
            report[key] = err
            tol = 1e-6 if exact else FLOAT_TOL.get(key, FLOAT_TOL['default'])
            if not exact and key.startswith('panoptic') and 'score' in key:
                # a flipped pixel changes a score by O(1): bound the number of deviating pixels instead
                assert int((np.abs(got - want) > 1e-5).sum()) <= MAX_FLIPS, (key, err)
            else:
                assert err <= tol, (key, err)
    for key in result:
        assert key in PASS_THROUGH or key in fix or key + '__sample' in fix, f'unexpected key {key}'
    return report


def _crop_slices(meta):
    c = meta['crop']
    return (slice(c[0], c[1]), slice(c[2], c[3]))


# ------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_fixture(name):
    fix, meta, inp = _load(name)
    c = inp['semantic'].shape[1]
    gt_fg = P.golden_instance_foreground(inp) if name in P.GT_FOREGROUND_CASES else None
    r = P.panoptic_postprocess(inp['semantic'], inp['center'], inp['offset'], inp['orientation'],
                               P.golden_is_thing(c), P.golden_has_orientation(c), _crop_slices(meta),
                               tuple(meta['fullres']), threshold=0.1, k=meta['k'], top_k=meta['top_k'],
                               instance_foreground=gt_fg)
    r.update(P.scene_postprocess(inp['scene']))
    r.pop('semantic_output')
    r.pop('scene_output')
    _compare(r, fix, exact=True)
    assert [len(m) for m in r['panoptic_segmentation_deeplab_instance_meta']] == meta['n_instances']


def _mirror_objects(pp, meta, c, mirror_host_placement=True):
    sem = pp.SemanticPostprocessingB200()
    ins = pp.InstancePostprocessingB200(heatmap_threshold=0.1, heatmap_nms_kernel_size=meta['k'],
                                        heatmap_apply_foreground_mask=False, top_k_instances=meta['top_k'],
                                        normalized_offset=True, offset_distance_threshold=None)
    pan = pp.PanopticPostprocessingB200(sem, ins, P.golden_is_thing(c), P.golden_has_orientation(c),
                                        compute_scores=True, mirror_host_placement=mirror_host_placement)
    return pan, pp.ScenePostprocessingB200()


def _run_mirror(pp, meta, inp, device, mirror_host_placement=True, gt_foreground=False):
    c = inp['semantic'].shape[1]
    pan, scene = _mirror_objects(pp, meta, c, mirror_host_placement)
    d = {k: v.to(device) for k, v in inp.items()}
    batch = P.make_batch(meta['crop'], tuple(meta['fullres']), d['semantic'].shape[0], device=device)
    if gt_foreground:                                   # dataset-evaluation branch, instance.py:365-400
        batch['instance_foreground'] = P.golden_instance_foreground(inp).to(device)
    data = ((d['semantic'], (d['center'], d['offset'], d['orientation'])), (None, None))
    r = pan.postprocess(data, batch, is_training=False)
    r.update(scene.postprocess((d['scene'], None), batch, is_training=False))
    return r


@pytest.fixture
def emulated_abi(monkeypatch):
    """the host side of emsanet_b200.postprocessing on CPU tensors, C-ABI calls replaced by their restatements"""
    from emsanet_b200 import postprocessing as pp
    for fn in ('softmax_argmax', 'nearest_resize', 'instance_centers', 'instance_assign', 'panoptic_merge'):
        monkeypatch.setattr(pp, fn, getattr(A, fn))
    monkeypatch.setattr(pp, '_dev', lambda t, dtype, what: t.detach().to(dtype).contiguous())
    return pp


@pytest.mark.parametrize('name', CASES)
def test_host_side_with_emulated_abi(name, emulated_abi):
    fix, meta, inp = _load(name)
    r = _run_mirror(emulated_abi, meta, inp, 'cpu', gt_foreground=name in P.GT_FOREGROUND_CASES)
    _compare(r, fix, exact=False)
    tr = _mirror_objects(emulated_abi, meta, inp['semantic'].shape[1])[0].postprocess(
        ((inp['semantic'], (inp['center'], inp['offset'], inp['orientation'])), ((None,), (None,))), {},
        is_training=True)
    assert set(tr) == {'semantic_output', 'semantic_side_outputs', 'instance_output', 'instance_side_outputs'}


def test_valid_region_helpers():
    from emsanet_b200 import postprocessing as pp
    batch = P.make_batch((8, 72, 0, 112), (131, 229), 2)
    crop, shape = pp.valid_region_and_fullres_shape(batch, 'semantic')
    assert pp._crop_box(crop, 80, 112) == (8, 0, 64, 112) and shape == (131, 229)
    with pytest.raises(ValueError, match='valid region'):
        pp.valid_region_and_fullres_shape({'rgb_fullres': torch.zeros(1, 3, 4, 4)}, 'semantic')
    with pytest.raises(ValueError, match='fullres shape'):
        pp.valid_region_and_fullres_shape({'_applied_preprocessing': batch['_applied_preprocessing']}, 'semantic')


def test_no_cpu_fallback():
    from emsanet_b200 import _lib, postprocessing as pp
    with pytest.raises(_lib.EB200Error, match='no CPU path'):
        pp.softmax_argmax(torch.zeros(1, 4, 8, 8))


def test_module_mirror_carries_postprocessing():
    from emsanet_b200 import postprocessing as pp
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    m = EMSANetB200(default_args(rgb_encoder_backbone='resnet18', depth_encoder_backbone='resnet18'),
                    simple_dataset_config())
    assert isinstance(m.decoders['panoptic_helper'].postprocessing, pp.PanopticPostprocessingB200)
    assert isinstance(m.decoders['scene_decoder'].postprocessing, pp.ScenePostprocessingB200)
    assert m.decoders['panoptic_helper'].postprocessing._instance_postprocessing._heatmap_nms_kernel_size == 17
    assert not any('postprocessing' in k for k in m.state_dict())


# ------------------------------------------------------------------------------------------------- GPU
def _tables_close(a, b, n):
    for b_ in range(n):
        k = int(a.counts[b_])
        assert k == int(b.counts[b_]), ('counts', b_, k, int(b.counts[b_]))
        assert torch.equal(a.centers[b_, :k].cpu(), b.centers[b_, :k].cpu())
        assert torch.equal(a.scores[b_, :k].cpu(), b.scores[b_, :k].cpu())


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_each_abi_call_matches_its_restatement(name):
    from emsanet_b200 import postprocessing as pp
    fix, meta, inp = _load(name)
    dev = 'cuda'
    n, c, h, w = inp['semantic'].shape
    flags_h = torch.from_numpy((np.asarray(P.golden_is_thing(c), np.uint8) |
                                (np.asarray(P.golden_has_orientation(c), np.uint8) << 1)))
    y0, y1, x0, x1 = meta['crop']
    box, out_hw = (y0, x0, y1 - y0, x1 - x0), tuple(meta['fullres'])
    # softmax / arg-max, native and resampled
    for kw in (dict(cls_flags=True), dict(box=box, out_hw=out_hw, want_logits=True)):
        use_flags = kw.pop('cls_flags', False)
        g = pp.softmax_argmax(inp['semantic'].to(dev), cls_flags=flags_h.to(dev) if use_flags else None, **kw)
        r = A.softmax_argmax(inp['semantic'], cls_flags=flags_h if use_flags else None, **kw)
        if g[0] is not None:
            assert float((g[0].cpu() - r[0]).abs().max()) <= 2e-5
        assert float((g[1].cpu() - r[1]).abs().max()) <= 2e-6
        assert float((g[2].cpu() - r[2]).abs().max()) <= 2e-6
        assert int((g[3].cpu() != r[3]).sum()) <= MAX_FLIPS
        if use_flags:
            assert int((g[4].cpu() != r[4]).sum()) <= MAX_FLIPS
            sem_idx_d, fg_d, sem_idx_h, fg_h = g[3], g[4], r[3], r[4]
    g = pp.softmax_argmax(inp['scene'].to(dev), want_scores=False)
    r = A.softmax_argmax(inp['scene'], want_scores=False)
    assert torch.equal(g[3].cpu(), r[3]) and float((g[2].cpu() - r[2]).abs().max()) <= 2e-6
    # centres (threshold, NMS tie rule, top-k with ties, row-major order)
    td, th = pp.InstanceTables(n, dev, c), pp.InstanceTables(n, 'cpu', c)
    pp.instance_centers(inp['center'].to(dev), td, 0.1, meta['k'], meta['top_k'])
    A.instance_centers(inp['center'], th, 0.1, meta['k'], meta['top_k'])
    _tables_close(td, th, n)
    assert int(td.status.abs().sum()) == 0
    # with the foreground mask applied to the centres (instance.py:139-140)
    td2, th2 = pp.InstanceTables(n, dev), pp.InstanceTables(n, 'cpu')
    pp.instance_centers(inp['center'].to(dev), td2, 0.1, meta['k'], meta['top_k'], fg=fg_h.to(dev))
    A.instance_centers(inp['center'], th2, 0.1, meta['k'], meta['top_k'], fg=fg_h)
    _tables_close(td2, th2, n)
    # assignment (use the restatement's foreground / classes on both sides so that this step is compared alone)
    for thr in (None, 6.0):
        seg_d = pp.instance_assign(inp['offset'].to(dev), fg_h.to(dev), td, float(h), float(w), thr,
                                   sem_idx_h.to(dev), c)
        seg_h = A.instance_assign(inp['offset'], fg_h, th, float(h), float(w), thr, sem_idx_h, c)
        assert int((seg_d.cpu() != seg_h).sum()) <= MAX_FLIPS
        assert int((td.areas.cpu() - th.areas).abs().sum()) <= 2 * MAX_FLIPS
        assert int((td.votes.cpu() - th.votes).abs().sum()) <= 2 * MAX_FLIPS
    seg_d = pp.instance_assign(inp['offset'].to(dev), fg_h.to(dev), td, float(h), float(w), None, sem_idx_h.to(dev), c)
    seg_h = A.instance_assign(inp['offset'], fg_h, th, float(h), float(w), None, sem_idx_h, c)
    # panoptic merge on identical inputs: integer outputs identical
    scores_h = A.softmax_argmax(inp['semantic'])[1]
    td.votes.copy_(th.votes)
    out_d = pp.panoptic_merge(seg_h.to(dev), sem_idx_h.to(dev), flags_h.to(dev), td, scores_h.to(dev),
                              inp['orientation'].to(dev), c)
    out_h = A.panoptic_merge(seg_h, sem_idx_h, flags_h, th, scores_h, inp['orientation'], c)
    assert torch.equal(td.inst_pan.cpu(), th.inst_pan)
    assert torch.equal(out_d[0].cpu(), out_h[0]) and torch.equal(out_d[1].cpu(), out_h[1])
    assert torch.equal(out_d[2].cpu(), out_h[2])                       # gathered semantic score: a copy
    assert torch.equal(out_d[3].cpu(), out_h[3])                       # instance score: a copy
    assert float((out_d[4].cpu() - out_h[4]).abs().max()) <= 1e-5
    acc_d, acc_h = td.inst_acc.cpu().numpy(), th.inst_acc.numpy()
    assert np.array_equal(acc_d[..., 1], acc_h[..., 1]) and np.array_equal(acc_d[..., 4], acc_h[..., 4])
    assert np.abs(acc_d - acc_h).max() <= 1e-3 * max(1.0, np.abs(acc_h).max() * 1e-3)
    # nearest resize of 1 / 4 / 8 byte maps
    for t in (seg_h, out_h[2], out_h[0]):
        assert torch.equal(pp.nearest_resize(t.to(dev), box, out_hw).cpu(), A.nearest_resize(t, box, out_hw))


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('mirror_host_placement', [True, False])
def test_mirror_classes_match_reference_fixture(name, mirror_host_placement):
    from emsanet_b200 import postprocessing as pp
    fix, meta, inp = _load(name)
    r = _run_mirror(pp, meta, inp, 'cuda', mirror_host_placement, gt_foreground=name in P.GT_FOREGROUND_CASES)
    report = _compare(r, fix, exact=False)
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f'parity_postproc_{name}.json'), 'w') as f:
        json.dump(report, f, indent=1)
    on_cpu = not r['panoptic_segmentation_deeplab'].is_cuda
    assert on_cpu == mirror_host_placement                              # panoptic.py:140-147 keeps these on the CPU
    assert r['panoptic_segmentation_deeplab_instance_idx'].is_cuda and r['semantic_segmentation_idx'].is_cuda
    assert r['panoptic_segmentation_deeplab'].dtype == torch.int64
    assert r['panoptic_segmentation_deeplab_instance_idx'].dtype == torch.uint8
    assert r['panoptic_foreground_mask'].dtype == torch.bool


@pytest.mark.gpu
def test_full_resolution_batch_against_oracle_and_properties():
    """480x640 (config-2 resolution): 2 images against the oracle, then the bench batch through invariants"""
    from emsanet_b200 import postprocessing as pp
    meta = {'k': 17, 'top_k': 64, 'crop': (0, 480, 0, 640), 'fullres': (480, 640)}
    inp = P.make_inputs(2, 480, 640, seed=11, n_blobs=40)
    c = 40
    want = P.panoptic_postprocess(inp['semantic'], inp['center'], inp['offset'], inp['orientation'],
                                  P.golden_is_thing(c), P.golden_has_orientation(c), (slice(0, 480), slice(0, 640)),
                                  (480, 640), k=17, top_k=64)
    got = _run_mirror(pp, meta, inp, 'cuda', mirror_host_placement=False)
    for key in ('semantic_segmentation_idx', 'panoptic_segmentation_deeplab_instance_idx',
                'panoptic_segmentation_deeplab', 'panoptic_segmentation_deeplab_semantic_idx'):
        bad = int((_np(got[key]).astype(np.int64) != _np(want[key]).astype(np.int64)).sum())
        assert bad <= 4 * MAX_FLIPS, (key, bad)
    assert float((got['semantic_softmax_scores'].cpu() - want['semantic_softmax_scores']).abs().max()) <= 2e-6
    assert [len(m) for m in got['panoptic_segmentation_deeplab_instance_meta']] == \
           [len(m) for m in want['panoptic_segmentation_deeplab_instance_meta']]
    assert got['panoptic_segmentation_deeplab_ids'] == want['panoptic_segmentation_deeplab_ids']
    # properties at the bench batch (size-independent)
    big = {k: v.repeat(8, *([1] * (v.ndim - 1))) for k, v in inp.items()}
    r = _run_mirror(pp, meta, big, 'cuda', mirror_host_placement=False)
    seg, pan, fg = (r['panoptic_segmentation_deeplab_instance_idx'], r['panoptic_segmentation_deeplab'],
                    r['panoptic_foreground_mask'])
    assert bool((seg[~fg] == 0).all())
    assert torch.equal(pan >> 16, r['panoptic_segmentation_deeplab_semantic_idx'])
    assert torch.equal(r['semantic_softmax_scores'].argmax(dim=1), r['semantic_segmentation_idx'])
    assert float((r['semantic_softmax_scores'].sum(dim=1) - 1).abs().max()) <= 1e-5
    for b in range(16):                                                  # replicas of the two images agree exactly
        assert torch.equal(seg[b], seg[b % 2]) and torch.equal(pan[b], pan[b % 2])
    areas = [sum(v['area'] for v in m.values()) for m in r['panoptic_segmentation_deeplab_instance_meta']]
    assert areas == [int(((seg[b] > 0)).sum()) for b in range(16)]
    r2 = _run_mirror(pp, meta, big, 'cuda', mirror_host_placement=False)   # deterministic integer outputs
    assert torch.equal(r2['panoptic_segmentation_deeplab'], pan)


@pytest.mark.gpu
@pytest.mark.parametrize('nms_k', [5, 17])      # 17 (the reference default) finds no centre on this tiny random net:
def test_model_forward_with_postprocessing_matches_oracle_on_its_outputs(nms_k):   # the "no instances" path
    """EMSANetB200.forward(batch, do_postprocessing=True) (emsanet/model.py:214-233): network on the engine, then the
    GPU post-processing; compared with the CPU oracle applied to the same network outputs"""
    from oracle import emsanet_oracle as O
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    cfg = O.OracleConfig(backbone='resnet18')
    sd = O.make_state_dict(cfg, seed=0)
    rgb, depth = O.make_inputs(2, 64, 96, seed=1)
    with torch.no_grad():      # calibrated running statistics (random ones blow eval-mode activations up by 1e5)
        _, stats = O.forward(sd, O.OracleConfig(backbone='resnet18', bn_momentum=1.0), rgb, depth, True)
    sd.update({k: v for k, v in stats.items() if 'num_batches' not in k})
    args = default_args(input_height=64, input_width=96, rgb_encoder_backbone='resnet18',
                        depth_encoder_backbone='resnet18', instance_center_heatmap_nms_kernel_size=nms_k)
    model = EMSANetB200(args, simple_dataset_config())
    model.load_state_dict(sd, strict=True)
    model.cuda().eval()
    batch = P.make_batch((0, 64, 0, 96), (64, 96), 2, device='cuda')
    batch.update(rgb=rgb.cuda(), depth=depth.cuda())
    with torch.no_grad():
        raw = model(batch)
        (sem, inst), _ = raw[0]
        sem, inst = sem.detach().cpu().clone(), [t.detach().cpu().clone() for t in inst]
        scene = raw[1][0].detach().cpu().clone()
        r = model(batch, do_postprocessing=True)
    assert isinstance(r, dict)
    want = P.panoptic_postprocess(sem, inst[0], inst[1], inst[2], (True,) * 40, (True,) * 40,
                                  (slice(0, 64), slice(0, 96)), (64, 96), threshold=0.1, k=nms_k, top_k=64)
    want.update(P.scene_postprocess(scene))
    for key in ('semantic_segmentation_idx', 'panoptic_segmentation_deeplab',
                'panoptic_segmentation_deeplab_instance_idx', 'scene_class_idx',
                'panoptic_segmentation_deeplab_semantic_idx_fullres'):
        bad = int((_np(r[key]).astype(np.int64) != _np(want[key]).astype(np.int64)).sum())
        assert bad <= MAX_FLIPS, (key, bad)
    assert r['panoptic_segmentation_deeplab_ids'] == want['panoptic_segmentation_deeplab_ids']
    assert [len(m) for m in r['panoptic_segmentation_deeplab_instance_meta']] == \
           [len(m) for m in want['panoptic_segmentation_deeplab_instance_meta']]
    assert set(want) <= set(r)
    for key in ('semantic_side_outputs', 'instance_side_outputs', 'instance_output', 'instance_centers'):
        assert key in r


REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout only exists in the build container')
def test_install_on_the_real_reference_model():
    """patch(model, postprocessing=True) on an UNMODIFIED reference EMSANet: every decoder's post-processing object is
    replaced by its mirror, carrying the reference object's settings; state_dict untouched"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), 'oracle'))
    import make_golden as mg
    from oracle import emsanet_oracle as O
    mg.install_reference_shim()
    from emsanet.model import EMSANet
    from emsanet_b200 import patch as patch_mod, postprocessing as pp
    cfg = O.OracleConfig(backbone='resnet18')
    model = EMSANet(mg.make_args(cfg, 96, 128), mg.make_dataset_config(cfg))
    keys = list(model.state_dict().keys())
    ref_pan = model.decoders['panoptic_helper'].postprocessing
    patch_mod.patch(model, postprocessing=True)
    new = model.decoders['panoptic_helper'].postprocessing
    assert isinstance(new, pp.PanopticPostprocessingB200)
    assert isinstance(model.decoders['scene_decoder'].postprocessing, pp.ScenePostprocessingB200)
    ins, ref_ins = new._instance_postprocessing, ref_pan._instance_postprocessing
    assert (ins._heatmap_threshold, ins._heatmap_nms_kernel_size, ins._top_k_instances, ins._normalized_offset) == \
           (ref_ins._heatmap_threshold, ref_ins._heatmap_nms_kernel_size, ref_ins._top_k_instances,
            ref_ins._normalized_offset) == (0.1, 17, 64, True)
    assert new._compute_scores == ref_pan._compute_scores is True
    assert list(new._thing_class_ids) == list(ref_pan._thing_class_ids)
    assert list(new._orientation_ids) == list(ref_pan._orientation_ids)
    assert new.max_instances_per_category == ref_pan.max_instances_per_category
    assert list(model.state_dict().keys()) == keys
    patch_mod.unpatch(model)


def test_crop_without_resize_returns_the_cropped_view(emulated_abi):
    """valid region smaller than the network resolution but already at the full resolution: no resampling, the
    full-resolution logits are the reference's cropped VIEW of the output (dense_base.py:24-31)"""
    inp = P.make_inputs(1, 40, 56, n_classes=7, seed=9)
    sem = emulated_abi.SemanticPostprocessingB200()
    batch = P.make_batch((4, 36, 8, 56), (32, 48), 1)
    r = sem.postprocess((inp['semantic'], (None,)), batch, is_training=False)
    want = P.semantic_postprocess(inp['semantic'], (slice(4, 36), slice(8, 56)), (32, 48))
    assert r['semantic_output_fullres'].data_ptr() == inp['semantic'][..., 4:36, 8:56].data_ptr()
    for key in ('semantic_segmentation_idx_fullres', 'semantic_segmentation_idx'):
        assert torch.equal(r[key], want[key])
    assert float((r['semantic_softmax_scores_fullres'] - want['semantic_softmax_scores_fullres']).abs().max()) <= 1e-6
    assert tuple(r['semantic_segmentation_idx_fullres'].shape) == (1, 32, 48)


@pytest.mark.gpu
@pytest.mark.parametrize('seg_dtype', [torch.uint8, torch.int32, torch.int64])
def test_orientation_sums_kernel_matches_its_restatement(seg_dtype):
    from emsanet_b200 import postprocessing as pp
    inp = P.make_inputs(2, 96, 128, seed=13, n_blobs=12)
    g = torch.Generator().manual_seed(5)
    seg = (torch.rand(2, 6, 8, generator=g) * 9).to(torch.int64)
    seg = seg.repeat_interleave(16, 1).repeat_interleave(16, 2).to(seg_dtype)          # blocky instance map, ids 0..8
    fg = torch.rand(2, 96, 128, generator=g) > 0.3
    max_id = 255 if seg_dtype == torch.uint8 else int(seg.max())
    got = pp.orientation_sums(inp['orientation'].cuda(), seg.cuda(), fg.cuda(), max_id).cpu()
    want = A.orientation_sums(inp['orientation'], seg, fg, max_id)
    assert torch.equal(got[..., 2], want[..., 2])
    assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))
    got = pp.orientation_sums(inp['orientation'].cuda(), seg.cuda(), None, max_id).cpu()
    assert torch.equal(got[..., 2], A.orientation_sums(inp['orientation'], seg, None, max_id)[..., 2])
