"""GPU-side input normalisation (MT/data/preprocessing/normalize.py:14-124).
not gpu: the oracle against the reference's own `normalize` / NormalizeRGB / NormalizeDepth where a reference install
         exists (bit-identical), parameter extraction from the reference objects, refusal of CPU tensors.
gpu:     the kernels against the oracle, bit for bit, incl. raw depth with invalid pixels and ragged sizes."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import preprocessing_oracle as PO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(n, h, w, seed):
    rng = np.random.default_rng(seed)
    rgb = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    depth = rng.integers(0, 40000, (n, h, w), dtype=np.uint16)
    depth[rng.random((n, h, w)) < 0.1] = 0
    return rgb, depth


def _reference():
    from emsanet_b200 import run
    try:
        root = run.find_reference()
    except FileNotFoundError:
        return None
    run.setup_paths(root)
    from nicr_mt_scene_analysis.data.preprocessing import normalize as ref
    return ref


def test_oracle_equals_reference_functions():
    ref = _reference()
    if ref is None:
        pytest.skip('no reference install')
    rgb, depth = _inputs(2, 24, 32, 0)
    nrgb = ref.NormalizeRGB()
    ndep = ref.NormalizeDepth(depth_mean=2841.9, depth_std=1417.3, raw_depth=True)
    for i in range(2):
        s = {'rgb': rgb[i].copy(), 'depth': depth[i].copy()}
        s, _ = nrgb._preprocess(s)
        s, _ = ndep._preprocess(s)
        want_rgb = np.ascontiguousarray(s['rgb'].transpose(2, 0, 1))
        got_rgb = PO.normalize_rgb(rgb[i:i + 1], nrgb._rgb_mean, nrgb._rgb_std)[0]
        assert got_rgb.dtype == np.float32 and np.array_equal(got_rgb, want_rgb)
        got_d = PO.normalize_depth(depth[i:i + 1], ndep._depth_mean, ndep._depth_std, True, 0.0)[0, 0]
        assert np.array_equal(got_d, s['depth'])
    from emsanet_b200.preprocessing import NormalizeB200
    nb = NormalizeB200.from_reference([nrgb, ndep])
    assert list(nb.rgb_mean) == nrgb._rgb_mean.tolist() and nb.depth_std == float(ndep._depth_std) and nb.raw_depth


def test_no_cpu_path_and_argument_checks():
    from emsanet_b200 import _lib
    from emsanet_b200.preprocessing import NormalizeB200
    nb = NormalizeB200(depth_mean=1.0, depth_std=2.0)
    with pytest.raises(_lib.EB200Error, match='no CPU path'):
        nb(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))
    with pytest.raises(ValueError):
        NormalizeB200(depth_std=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 48, 64), (1, 37, 51), (3, 480, 640)])
@pytest.mark.parametrize('raw', [False, True])
def test_kernels_are_bit_identical_to_the_oracle(shape, raw):
    from emsanet_b200.preprocessing import IMAGENET_MEAN, IMAGENET_STD, NormalizeB200
    n, h, w = shape
    rgb, depth = _inputs(n, h, w, 7)
    nb = NormalizeB200(depth_mean=2841.94941272766, depth_std=1417.2594281672277, raw_depth=raw)
    out = nb(torch.from_numpy(rgb).cuda(), torch.from_numpy(depth.astype(np.int32)).cuda())
    want_rgb = PO.normalize_rgb(rgb, np.float32(IMAGENET_MEAN), np.float32(IMAGENET_STD))
    want_d = PO.normalize_depth(depth, 2841.94941272766, 1417.2594281672277, raw, 0.0)
    assert np.array_equal(out['rgb'].cpu().numpy(), want_rgb)
    assert np.array_equal(out['depth'].cpu().numpy(), want_d)
    out16 = nb.depth(torch.from_numpy(depth.view(np.int16)).cuda())          # uint16 payload in a 2-byte tensor
    assert np.array_equal(out16.cpu().numpy(), want_d)
    # the timing script's recipe (inference_time_whole_model.py:536-537): /255 and /20000
    t = NormalizeB200(rgb_mean=(0, 0, 0), rgb_std=(255, 255, 255), depth_mean=0.0, depth_std=20000.0)
    o = t(torch.from_numpy(rgb).cuda(), torch.from_numpy(depth.view(np.int16)).cuda())
    assert np.array_equal(o['rgb'].cpu().numpy(), (rgb / 255).astype('float32').transpose(0, 3, 1, 2))
    assert np.array_equal(o['depth'].cpu().numpy()[:, 0], depth.astype('float32') / 20000)
