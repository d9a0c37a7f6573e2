"""The algorithmic shortcuts inside csrc/postproc.cu, modelled in numpy float32 (IEEE add / mul / sqrt, the same
arithmetic the kernels are pinned to with __fadd_rn / __fmul_rn / sqrtf) and checked against the plain formulation
of the reference on random AND adversarial inputs — the near-tie cases random GPU test data does not reach.

* pp_select_kernel: the top_k-th largest value by a 31-step search over float bit patterns;
* pp_assign_kernel<LAZY_SQRT>: first arg-min of the ROUNDED distances with one sqrt of the minimum squared distance
  plus a sqrt only for centres within float rounding of it (torch.norm + torch.min, instance.py:220-226);
* pp_nms_kernel: the candidate-list capacity bound (two survivors are never inside each other's window) and the
  empty-tile early exit;
* pp_merge_table_kernel: first maximum of the votes row == torch.mode (smallest most frequent label).
"""
import numpy as np
import pytest
import torch

from oracle import postprocessing_oracle as P

f32 = np.float32


def kth_largest_bitwise(v: np.ndarray, k: int) -> np.float32:
    """pp_select_kernel: largest T (as a bit pattern) with count(v >= T) >= k; v > 0"""
    bits = v.view(np.uint32)
    prefix = np.uint32(0)
    for bit in range(30, -1, -1):
        trial = prefix | np.uint32(1 << bit)
        if int((bits >= trial).sum()) >= k:
            prefix = trial
    return np.array([prefix], np.uint32).view(np.float32)[0]


@pytest.mark.parametrize('seed', range(5))
def test_bitwise_search_finds_the_kth_largest_value(seed):
    rng = np.random.default_rng(seed)
    for n, k in ((1, 1), (7, 7), (64, 8), (500, 64), (5000, 254)):
        v = (rng.random(n).astype(f32) * f32(0.9) + f32(0.1000001))
        if seed % 2:                                   # plateaus: many exactly equal values around the k-th
            v = np.round(v * 16).astype(f32) / f32(16) + f32(0.125)
        want = np.sort(v)[::-1][k - 1]
        assert kth_largest_bitwise(v, k) == want
        assert int((v >= want).sum()) >= k            # '>=' keeps ties: possibly more than k centres (instance.py:152)


def first_argmin_reference(cy, cx, ly, lx):
    """torch.norm(centers - loc, dim=-1) then torch.min(dim=0): instance.py:220-226 (first minimum on CPU)"""
    dy, dx = cy - ly, cx - lx
    d = np.sqrt(dy * dy + dx * dx, dtype=f32)
    return int(np.argmin(d)), d.min()


def first_argmin_lazy(cy, cx, ly, lx):
    """pp_assign_kernel<true>"""
    dy, dx = cy - ly, cx - lx
    d2 = dy * dy + dx * dx
    m2 = d2.min()
    best = np.sqrt(m2, dtype=f32)
    lim = m2 * f32(1.000001)
    for j in range(len(d2)):
        if d2[j] <= lim and np.sqrt(d2[j], dtype=f32) == best:
            return j, best
    raise AssertionError('no candidate found')


def test_lazy_sqrt_argmin_equals_reference_on_random_points():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        k = int(rng.integers(1, 65))
        cy, cx = rng.integers(0, 480, k).astype(f32), rng.integers(0, 640, k).astype(f32)
        ly, lx = f32(rng.random() * 480), f32(rng.random() * 640)
        assert first_argmin_lazy(cy, cx, ly, lx) == first_argmin_reference(cy, cx, ly, lx)


def test_lazy_sqrt_argmin_equals_reference_on_adversarial_near_ties():
    """squared distances that differ by one or a few ulps: their square roots round to the SAME float, so the reference
    takes the FIRST of them even when a later one has the (slightly) smaller squared distance"""
    rng = np.random.default_rng(1)
    merged = 0
    for _ in range(3000):
        ly, lx = f32(rng.random() * 400 + 20), f32(rng.random() * 600 + 20)
        # symmetric / mirrored centres give (nearly) equal distances; jitter the pixel location by ulps
        r, c = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        base_y, base_x = f32(round(float(ly))), f32(round(float(lx)))
        cy = np.array([base_y + r, base_y - r, base_y + c, base_y - c, base_y + r], f32)
        cx = np.array([base_x + c, base_x - c, base_x + r, base_x - r, base_x - c], f32)
        ly2 = np.nextafter(base_y, f32(np.inf) if rng.random() < 0.5 else f32(-np.inf), dtype=f32) if rng.random() < 0.7 else base_y
        lx2 = np.nextafter(base_x, f32(np.inf) if rng.random() < 0.5 else f32(-np.inf), dtype=f32) if rng.random() < 0.7 else base_x
        perm = rng.permutation(5)
        got = first_argmin_lazy(cy[perm], cx[perm], ly2, lx2)
        want = first_argmin_reference(cy[perm], cx[perm], ly2, lx2)
        assert got == want
        dy, dx = cy - ly2, cx - lx2
        d2 = dy * dy + dx * dx
        d = np.sqrt(d2, dtype=f32)
        merged += int(len(np.unique(d2[d == d.min()])) > 1)
    assert merged > 50, 'the adversarial generator should produce distinct squared distances with equal rounded roots'


def test_sqrt_merges_only_values_within_the_prefilter_window():
    """the claim behind `d2 <= m2 * 1.000001f`: if sqrtf(a) == sqrtf(b) with a <= b then b <= a * (1 + 2.4e-7)"""
    rng = np.random.default_rng(2)
    a = (rng.random(200000).astype(f32) * f32(1e6) + f32(1e-3))
    for step in (1, 2, 3, 4):
        b = a.copy()
        for _ in range(step):
            b = np.nextafter(b, f32(np.inf), dtype=f32)
        same = np.sqrt(a, dtype=f32) == np.sqrt(b, dtype=f32)
        assert (b[same] <= a[same] * f32(1.000001)).all()
        if step == 4:
            assert not same.any()      # four ulps apart never merge (2 ulps of the root at most)


@pytest.mark.parametrize('k,quantise', [(3, None), (3, 4), (5, 8), (17, None), (9, 2)])
def test_nms_survivors_respect_the_capacity_bound_and_empty_tiles(k, quantise):
    rng = np.random.default_rng(k)
    h, w = 70, 101
    heat = rng.random((h, w)).astype(f32)
    if quantise:
        heat = np.round(heat * quantise).astype(f32) / f32(quantise)     # plateaus: the tie rule is what bounds the count
    nms = P.nms_heatmap(heat, 0.1, k)
    pad = (k - 1) // 2
    cap = ((h + pad) // (pad + 1)) * ((w + pad) // (pad + 1))           # eb200_pp_centers_ws_bytes
    ys, xs = np.nonzero(nms != -1)
    assert len(ys) <= cap
    # pairwise Chebyshev distance > pad: no survivor lies inside another survivor's window
    for i in range(len(ys)):
        d = np.maximum(np.abs(ys - ys[i]), np.abs(xs - xs[i]))
        d[i] = 10 ** 6
        assert d.min() > pad
    # early exit: a 32x32 tile without a pixel above the threshold has no survivor (survivors are above it)
    assert (heat[ys, xs] > 0.1).all()


def test_votes_first_maximum_is_torch_mode():
    rng = np.random.default_rng(3)
    for _ in range(300):
        labels = rng.integers(1, 8, int(rng.integers(1, 60)))
        votes = np.bincount(labels, minlength=9)
        best, cls = 0, 0
        for c in range(len(votes)):                    # pp_merge_table_kernel: strict '>' keeps the first maximum
            if votes[c] > best:
                best, cls = votes[c], c
        assert cls == int(torch.mode(torch.from_numpy(labels)).values)
