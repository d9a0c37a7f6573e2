"""CPU oracle of the inference post-processing that sits directly behind the EMSANet hot path
(SURVEY.md §8(f) row 1).  TEST INFRASTRUCTURE ONLY: imported by tests/, never by emsanet_b200/.

A functional numpy restatement (torch only for softmax / bilinear interpolate / atan2, i.e. the
floating-point ATen ops the reference itself dispatches to) of
    MT/model/postprocessing/semantic.py:35-82     SemanticPostprocessing._postprocess_inference
    MT/model/postprocessing/instance.py:74-160    InstancePostprocessing._get_instance_centers
    MT/model/postprocessing/instance.py:162-273   InstancePostprocessing._get_instance_segmentation
    MT/model/postprocessing/instance.py:275-323   InstancePostprocessing._get_instance_orientation
    MT/model/postprocessing/panoptic.py:77-316    PanopticPostprocessing._postprocess_inference
    MT/utils/panoptic_merge.py:168-225            deeplab_merge_semantic_and_instance
    MT/model/postprocessing/scene.py:32-53        ScenePostprocessing._postprocess_inference
(MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/).

Pinned: oracle/make_golden_postproc.py runs the UNMODIFIED reference classes on the same synthetic
inputs in the build container; every integer / index output is identical and every float output is
within 1e-6 (fixtures: tests/golden/postproc/, re-checked by tests/test_postproc.py).

Written the vectorised way a GPU kernel computes it (per-pixel rules, per-instance tables), not as
the reference's Python loops over batch and instances, so the kernels can be compared rule by rule.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MAX_INSTANCES_PER_CATEGORY = 1 << 16      # panoptic.py:52


# ---------------------------------------------------------------------------------------------- inputs
def golden_is_thing(n_classes: int) -> Tuple[bool, ...]:
    """class flags of the golden cases (every third class is stuff; every class = 1 mod 3 has an orientation)"""
    return tuple((c % 3) != 0 for c in range(n_classes))


def golden_has_orientation(n_classes: int) -> Tuple[bool, ...]:
    return tuple((c % 3) == 1 for c in range(n_classes))


GT_FOREGROUND_CASES = ('upscale_crop', 'ties_quantised_topk')   # golden cases that carry batch['instance_foreground']


def golden_instance_foreground(inp: Dict[str, torch.Tensor]) -> torch.Tensor:
    """a ground-truth-like boolean foreground mask [N,H,W] derived from the synthetic inputs"""
    return (inp['center'][:, 0] > 0.12) | (inp['semantic'][:, 1] > inp['semantic'][:, 0])


def make_batch(crop: Tuple[int, int, int, int], fullres: Tuple[int, int], n: int, device=None) -> Dict:
    """the two things post-processing reads from the batch (MT/data/preprocessing/resize.py:30-78):
    the Resize entry of the applied-preprocessing meta and the shape of a *_fullres tensor"""
    meta = [[{'type': 'Resize', 'valid_region_slice_y': slice(crop[0], crop[1]),
              'valid_region_slice_x': slice(crop[2], crop[3])}] for _ in range(n)]
    return {'_applied_preprocessing': meta, 'rgb_fullres': torch.zeros(n, 3, *fullres, device=device)}


def make_inputs(n: int, h: int, w: int, n_classes: int = 40, seed: int = 3, n_blobs: int = 9,
                quantise: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Synthetic decoder outputs with the statistics post-processing cares about: blobby class maps,
    a centre heat map in [0,1] made of Gaussian bumps over low noise, normalised offsets that point to the
    bumps, unit-length orientation vectors.  `quantise=q` rounds the heat map to multiples of 1/q, which
    produces plateaus, i.e. exact ties inside the NMS window (instance.py:86-89)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.randn(n, n_classes, max(h // 16, 2), max(w // 16, 2), generator=g) * 3.0
    logits = F.interpolate(low, size=(h, w), mode='bilinear', align_corners=False)
    logits = logits + 0.3 * torch.randn(n, n_classes, h, w, generator=g)
    ys = torch.arange(h, dtype=torch.float32).view(1, 1, h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w)
    cy = torch.rand(n, n_blobs, 1, 1, generator=g) * (h - 1)
    cx = torch.rand(n, n_blobs, 1, 1, generator=g) * (w - 1)
    amp = 0.3 + 0.7 * torch.rand(n, n_blobs, 1, 1, generator=g)
    sig = 3.0 + 5.0 * torch.rand(n, n_blobs, 1, 1, generator=g)
    d2 = (ys - cy) ** 2 + (xs - cx) ** 2
    bumps = amp * torch.exp(-d2 / (2 * sig * sig))
    heat, near = bumps.max(dim=1, keepdim=True)
    heat = (heat + 0.08 * torch.rand(n, 1, h, w, generator=g)).clamp(0, 1)
    if quantise:
        heat = torch.round(heat * quantise) / quantise
    ncy = torch.gather(cy.expand(n, n_blobs, h, w), 1, near)
    ncx = torch.gather(cx.expand(n, n_blobs, h, w), 1, near)
    off = torch.cat([(ncy - ys) / h, (ncx - xs) / w], dim=1)
    off = (off + 0.01 * torch.randn(n, 2, h, w, generator=g)).clamp(-1, 1)
    ang = F.interpolate(torch.rand(n, 1, max(h // 8, 2), max(w // 8, 2), generator=g) * 6.28318,
                        size=(h, w), mode='bilinear', align_corners=False)
    ang = ang + 0.05 * torch.randn(n, 1, h, w, generator=g)
    orient = torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)
    scene = torch.randn(n, 10, generator=g) * 2
    return {'semantic': logits.contiguous(), 'center': heat.contiguous(), 'offset': off.contiguous(),
            'orientation': orient.contiguous(), 'scene': scene}


# -------------------------------------------------------------------------------------------- semantic
def softmax_max(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """semantic.py:55-56 / scene.py:41-42: softmax over dim 1, then (max, first arg-max) of the SOFTMAX values"""
    pred = F.softmax(logits, dim=1)
    score, idx = torch.max(pred, dim=1)
    return pred, score, idx


def crop_resize(t: torch.Tensor, crop: Tuple[slice, slice], shape: Tuple[int, int], mode: str) -> torch.Tensor:
    """dense_base.py:15-58 (_crop_to_valid_region_and_resize_prediction)"""
    t = t[..., crop[0], crop[1]]
    if tuple(shape) == tuple(t.shape[-2:]):
        return t
    nd, dt = t.ndim, t.dtype
    if nd == 3:
        t = t.unsqueeze(1)
    if not t.is_floating_point():
        t = t.to(torch.float32)
    kw = {'align_corners': False} if mode != 'nearest' else {}
    t = F.interpolate(t, size=tuple(shape), mode=mode, **kw).to(dt)
    return t.squeeze(1) if nd == 3 else t


def semantic_postprocess(logits: torch.Tensor, crop: Tuple[slice, slice], fullres: Tuple[int, int]) -> Dict:
    """semantic.py:35-82"""
    pred, score, idx = softmax_max(logits)
    out_f = crop_resize(logits, crop, fullres, 'bilinear')
    pred_f, score_f, idx_f = softmax_max(out_f)
    return {'semantic_output': logits, 'semantic_softmax_scores': pred, 'semantic_segmentation_score': score,
            'semantic_segmentation_idx': idx, 'semantic_output_fullres': out_f,
            'semantic_softmax_scores_fullres': pred_f, 'semantic_segmentation_score_fullres': score_f,
            'semantic_segmentation_idx_fullres': idx_f}


# -------------------------------------------------------------------------------------------- instance
def nms_heatmap(heat: np.ndarray, threshold: float, k: int) -> np.ndarray:
    """instance.py:79-128 for one image `heat` [H,W] -> [H,W] with the surviving centre values, -1 elsewhere.

    Per-pixel rule: p survives iff it lies >= (k-1)/2 pixels away from every border, heat[p] > threshold
    (F.threshold keeps x > threshold), heat[p] equals the maximum of its k x k window AND no pixel that comes
    EARLIER in row-major order inside that window has the same value (max_pool2d's returned index is the first
    maximum; `pooling_indices != pixel_index_map` removes every later tie)."""
    h, w = heat.shape
    pad = (k - 1) // 2
    t = np.where(heat > threshold, heat, np.float32(-1)).astype(np.float32)
    out = np.full((h, w), -1, np.float32)
    if h < k or w < k:
        return out
    win = np.lib.stride_tricks.sliding_window_view(t, (k, k)).reshape(h - k + 1, w - k + 1, k * k)
    first = win.argmax(axis=-1)                       # first maximum in row-major order
    centre = t[pad:h - pad, pad:w - pad]
    keep = (first == pad * k + pad) & (centre != -1)  # an all -1 window has first == 0 != centre for k >= 3
    if k == 1:
        keep = centre != -1
    out[pad:h - pad, pad:w - pad] = np.where(keep, centre, np.float32(-1))
    return out


def instance_centers(heat: torch.Tensor, threshold: float = 0.1, k: int = 17, top_k: int = 64,
                     foreground: Optional[torch.Tensor] = None, apply_foreground_mask: bool = False
                     ) -> Tuple[np.ndarray, List[np.ndarray]]:
    """instance.py:74-160.  heat [N,1,H,W] -> (bool [N,H,W], list of int32 [M,2] (y,x) in row-major order)."""
    n, _, h, w = heat.shape
    masks, lists = [], []
    for b in range(n):
        nms = nms_heatmap(heat[b, 0].numpy(), threshold, k)
        flat = np.sort(nms.reshape(-1))[::-1]
        kth = flat[top_k - 1]                         # torch.topk(...)[0][:, -1]  (instance.py:131-148)
        if apply_foreground_mask:                     # applied AFTER the top-k scores were taken (:139-140)
            nms = np.where(foreground[b].numpy().astype(bool), nms, np.float32(-1))
        lowest = max(float(kth), 0.0)
        m = nms >= np.float32(lowest)
        masks.append(m)
        lists.append(np.argwhere(m).astype(np.int32))  # nonzero(): row-major order (:158-159)
    return np.stack(masks), lists


def instance_segmentation(heat: torch.Tensor, offset_px: torch.Tensor, foreground: torch.Tensor,
                          threshold: float = 0.1, k: int = 17, top_k: int = 64,
                          apply_foreground_mask: bool = False, distance_threshold: Optional[float] = None
                          ) -> Tuple[np.ndarray, List[Dict[int, Dict]]]:
    """instance.py:162-273.  offset_px [N,2,H,W] is the offset ALREADY in pixels (:357-363).
    Per foreground pixel: id = 1 + first arg-min over the centres of || centre - (pixel + offset) ||_2 (fp32)."""
    n, _, h, w = offset_px.shape
    _, centres = instance_centers(heat, threshold, k, top_k, foreground, apply_foreground_mask)
    seg = np.zeros((n, h, w), np.uint8)
    metas: List[Dict[int, Dict]] = [{} for _ in range(n)]
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing='ij')
    for b, cen in enumerate(centres):
        if cen.shape[0] == 0:
            continue
        fg = foreground[b].numpy().astype(bool)
        off = offset_px[b].numpy()
        ly = (yy.astype(np.int64) + off[0])[fg].astype(np.float32)          # mesh_grid (int64) + float32 -> float32
        lx = (xx.astype(np.int64) + off[1])[fg].astype(np.float32)
        dy = cen[:, 0:1].astype(np.float32) - ly[None, :]
        dx = cen[:, 1:2].astype(np.float32) - lx[None, :]
        dist = torch.norm(torch.stack([torch.from_numpy(dy), torch.from_numpy(dx)], -1), dim=-1).numpy()
        ids = (dist.argmin(axis=0) + 1).astype(np.uint8)
        if distance_threshold is not None:
            ids[dist.min(axis=0) > distance_threshold] = 0
        seg[b][fg] = ids
        areas = np.bincount(ids, minlength=cen.shape[0] + 1)
        for i, (y, x) in enumerate(cen.tolist(), start=1):
            metas[b][i] = {'center_yx': (y, x), 'area': int(areas[i]), 'score': float(heat[b, 0, y, x])}
    return seg, metas


def instance_orientation(orientation: torch.Tensor, seg: np.ndarray, foreground: Optional[np.ndarray]
                         ) -> List[Dict[int, float]]:
    """instance.py:275-323: per instance id, atan2 of the summed (cos, sin) vectors over its foreground pixels"""
    res = []
    for b in range(seg.shape[0]):
        o = orientation[b].numpy()
        m = np.ones(seg[b].shape, bool) if foreground is None else foreground[b].astype(bool)
        d = {}
        for i in np.unique(seg[b][m]):
            if i == 0:
                continue
            sel = m & (seg[b] == i)
            s = torch.from_numpy(o[:, sel]).sum(dim=1)
            d[int(i)] = float(torch.atan2(s[1], s[0]))
        res.append(d)
    return res


# -------------------------------------------------------------------------------------------- panoptic
def panoptic_merge(sem_with_void: np.ndarray, seg: np.ndarray, foreground: np.ndarray,
                   thing_ids: Sequence[int], void_label: int = 0) -> Tuple[np.ndarray, List[Dict[int, int]]]:
    """panoptic_merge.py:168-225 over a batch.  Per instance (ascending id): class = smallest most frequent
    semantic label of its pixels, new id = running count of that class; per pixel:
    instance pixel -> class * 2^16 + new id, stuff-class pixel without instance -> class * 2^16, else void."""
    n = seg.shape[0]
    pan = np.full(seg.shape, void_label, np.int64)
    dicts: List[Dict[int, int]] = []
    things = set(int(t) for t in thing_ids)
    for b in range(n):
        is_thing = (seg[b] > 0) & (foreground[b] > 0)
        counter: Dict[int, int] = {}
        id_dict: Dict[int, int] = {}
        for i in np.unique(seg[b]):
            if i == 0:
                continue
            m = (seg[b] == i) & is_thing
            if not m.any():
                continue
            votes = np.bincount(sem_with_void[b][m].astype(np.int64))
            cls = int(votes.argmax())                 # torch.mode: the smallest of the most frequent values
            if cls == 0:
                continue
            counter[cls] = counter.get(cls, 0) + 1
            pid = cls * MAX_INSTANCES_PER_CATEGORY + counter[cls]
            id_dict[pid] = int(i)
            pan[b][m] = pid
        for cls in np.unique(sem_with_void[b]):
            if cls == 0 or int(cls) in things:
                continue
            pan[b][(sem_with_void[b] == cls) & (seg[b] == 0)] = int(cls) * MAX_INSTANCES_PER_CATEGORY
        dicts.append(id_dict)
    return pan, dicts


def panoptic_postprocess(sem_logits: torch.Tensor, center: torch.Tensor, offset: torch.Tensor,
                         orientation: Optional[torch.Tensor], classes_is_thing: Sequence[bool],
                         classes_has_orientation: Sequence[bool], crop: Tuple[slice, slice],
                         fullres: Tuple[int, int], threshold: float = 0.1, k: int = 17, top_k: int = 64,
                         normalized_offset: bool = True, compute_scores: bool = True,
                         instance_foreground: Optional[torch.Tensor] = None) -> Dict:
    """panoptic.py:77-316; `instance_foreground` = batch['instance_foreground'], the ground-truth mask that adds the
    dataset-evaluation outputs of instance.py:365-400."""
    r = semantic_postprocess(sem_logits, crop, fullres)
    n, _, h, w = offset.shape
    off = offset.clone()
    if normalized_offset:                                       # panoptic.py:106-112
        off[:, 0] = off[:, 0] * h
        off[:, 1] = off[:, 1] * w
    if instance_foreground is not None:                         # instance.py:365-400
        seg_gt, meta_gt = instance_segmentation(center, off, instance_foreground, threshold, k, top_k)
        r['instance_segmentation_gt_foreground'] = seg_gt
        r['instance_segmentation_gt_meta'] = meta_gt
        r['instance_segmentation_gt_foreground_fullres'] = crop_resize(torch.from_numpy(seg_gt), crop, fullres,
                                                                       'nearest').numpy()
    thing_cls = np.where(np.asarray(classes_is_thing))[0]
    sem_idx = r['semantic_segmentation_idx'].numpy()
    fg = np.isin(sem_idx, thing_cls)                            # :123-128
    seg, meta = instance_segmentation(center, off, torch.from_numpy(fg), threshold, k, top_k)
    pan, ids = panoptic_merge(sem_idx + 1, seg, fg, thing_cls + 1)          # :140-147
    pan_sem = pan // MAX_INSTANCES_PER_CATEGORY
    r.update({'panoptic_foreground_mask': fg, 'panoptic_segmentation_deeplab': pan,
              'panoptic_segmentation_deeplab_ids': ids, 'panoptic_segmentation_deeplab_semantic_idx': pan_sem,
              'panoptic_segmentation_deeplab_instance_idx': seg,
              'panoptic_segmentation_deeplab_instance_meta': meta})
    if compute_scores:                                          # :167-236
        scores = r['semantic_softmax_scores'].numpy()
        void = pan_sem == 0
        cls0 = np.where(void, 0, pan_sem - 1)
        sem_score = np.take_along_axis(scores, cls0[:, None], axis=1)[:, 0]
        sem_score = np.where(void, np.float32(0), sem_score).astype(np.float32)
        ins_score = np.zeros(pan.shape, np.float32)
        pan_score = sem_score.copy()
        for b in range(n):
            for pid, iid in ids[b].items():
                m = pan[b] == pid
                s_i = np.float32(meta[b][iid]['score'])
                ins_score[b][m] = s_i
                s_s = torch.mean(torch.from_numpy(sem_score[b][m]))
                meta[b][iid]['semantic_score'] = float(s_s)
                meta[b][iid]['semantic_idx'] = int(pan_sem[b][m][0])
                p = s_s * float(s_i)
                pan_score[b][m] = float(p)
                meta[b][iid]['panoptic_score'] = float(p)
                meta[b][iid]['panoptic_id'] = pid
        r.update({'panoptic_segmentation_deeplab_semantic_score': sem_score,
                  'panoptic_segmentation_deeplab_instance_score': ins_score,
                  'panoptic_segmentation_deeplab_panoptic_score': pan_score})
    for key in ('panoptic_segmentation_deeplab', 'panoptic_segmentation_deeplab_instance_idx',
                'panoptic_segmentation_deeplab_semantic_idx') + (
            ('panoptic_segmentation_deeplab_semantic_score', 'panoptic_segmentation_deeplab_instance_score',
             'panoptic_segmentation_deeplab_panoptic_score') if compute_scores else ()):
        r[key + '_fullres'] = crop_resize(torch.from_numpy(np.ascontiguousarray(r[key])), crop, fullres,
                                          'nearest').numpy()                # :239-286
    if orientation is not None:                                 # :289-314
        fg_o = np.isin(pan_sem, np.where(np.asarray(classes_has_orientation))[0] + 1)
        ors = instance_orientation(orientation, seg, fg_o)
        r['orientations_panoptic_segmentation_deeplab_instance'] = ors
        for b in range(n):
            for iid in meta[b]:
                meta[b][iid]['orientation'] = ors[b].get(iid, float('nan'))
    return r


def scene_postprocess(logits: torch.Tensor) -> Dict:
    """scene.py:32-53"""
    _, score, idx = softmax_max(logits)
    return {'scene_class_score': score, 'scene_class_idx': idx, 'scene_output': logits}
