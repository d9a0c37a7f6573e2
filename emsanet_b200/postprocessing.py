"""Inference post-processing on the GPU, behind the reference's own interface.

Mirrors of the reference's post-processing classes — same constructor arguments, same
`postprocess(data, batch, is_training=True)` contract, same result keys / shapes / dtypes:

    SemanticPostprocessingB200   MT/model/postprocessing/semantic.py:17-82
    InstancePostprocessingB200   MT/model/postprocessing/instance.py:23-468
    PanopticPostprocessingB200   MT/model/postprocessing/panoptic.py:23-316 (+ MT/utils/panoptic_merge.py:168-225)
    ScenePostprocessingB200      MT/model/postprocessing/scene.py:15-53

(MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/).  Training mode passes the decoder outputs
through like the reference.  Inference mode runs on the kernels of csrc/postproc.cu through the C ABI
(eb200_pp_*): no Python loop over images or instances, no .item() calls, no CPU round trip for the panoptic
merge; everything per-instance comes back in ONE device-to-host copy of small tables from which the meta
dictionaries are built.  There is no CPU / torch fallback: inputs that are not CUDA tensors raise.

`install(model)` swaps the post-processing objects of a reference EMSANet (or an EMSANetB200) for these.

Not covered (raise NotImplementedError instead of returning something else): the debug variants
(instance.py:402-421,453-466).  The orientation variants on ground-truth instance maps (instance.py:431-451, part
of every validation batch of the orientation task) run on eb200_pp_instance_orientation (verified on a B200 in
round 2: profiles/r2_unverified_kernels_first_run.log).
"""
import ctypes as C
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MAX_INST = 256          # EB200_PP_MAX_INSTANCES
ACC = 5                 # EB200_PP_ACC_FIELDS
FULLRES_SUFFIX = '_fullres'                      # MT/data/preprocessing/resize.py:19
MAX_INSTANCES_PER_CATEGORY = 1 << 16             # panoptic.py:52


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_host(tensors: List[torch.Tensor]) -> List[torch.Tensor]:
    """device -> pinned host memory (torch's caching host allocator), all copies asynchronous, ONE synchronisation"""
    out = []
    for t in tensors:
        h = torch.empty(t.shape, dtype=t.dtype, device='cpu', pin_memory=t.is_cuda)
        h.copy_(t, non_blocking=True)
        out.append(h)
    if any(t.is_cuda for t in tensors):
        torch.cuda.current_stream().synchronize()
    return out


def _dev(t: torch.Tensor, dtype, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.EB200Error(f'{what} must be a CUDA tensor: emsanet_b200 post-processing has no CPU path')
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.detach().contiguous()


# ------------------------------------------------------------------------------------------------ batch meta
def valid_region_and_fullres_shape(batch: Dict[str, Any], key: str) -> Tuple[Tuple[int, int, int, int], Tuple[int, int]]:
    """MT/data/preprocessing/resize.py:30-78 without importing the reference: ((y0, y1, x0, x1) | None, (h, w))"""
    meta = batch.get('_applied_preprocessing', None)
    crop = None
    if meta is not None and len(meta) > 0:
        for pre in meta[0]:                       # all samples share the original resolution (resize.py:55-57)
            if pre['type'] == 'Resize':
                crop = (pre['valid_region_slice_y'], pre['valid_region_slice_x'])
                break
    if crop is None:
        raise ValueError('Unable to get get valid region slices.')
    for k in (key, 'rgb', 'depth'):
        img = batch.get(k + FULLRES_SUFFIX, None)
        if img is not None:
            return crop, tuple(img.shape[-2:])
    raise ValueError(f'Unable to get fullres shape for `{key}`.')


def _crop_box(crop, h: int, w: int) -> Tuple[int, int, int, int]:
    ys, xs = (range(h)[crop[0]], range(w)[crop[1]])
    if ys.step != 1 or xs.step != 1 or len(ys) == 0 or len(xs) == 0:
        raise ValueError(f'unsupported valid-region slices {crop}')
    return ys.start, xs.start, len(ys), len(xs)


# ------------------------------------------------------------------------------------------------ raw ops
def softmax_argmax(logits: torch.Tensor, box=None, out_hw=None, want_scores=True, want_logits=False,
                   cls_flags: Optional[torch.Tensor] = None):
    """eb200_pp_softmax_argmax.  logits fp32 [N,C,H,W] (or [N,C]).  Returns (out_logits|None, scores|None, score, idx,
    flags|None) on the crop `box` = (y0, x0, Hc, Wc) resampled to `out_hw`."""
    two_d = logits.ndim == 2
    x = _dev(logits, torch.float32, 'logits')
    if two_d:
        x = x[:, :, None, None]
    n, c, h, w = x.shape
    y0, x0, hc, wc = box or (0, 0, h, w)
    ho, wo = out_hw or (hc, wc)
    dev = x.device
    out_logits = torch.empty((n, c, ho, wo), dtype=torch.float32, device=dev) if want_logits else None
    scores = torch.empty((n, c, ho, wo), dtype=torch.float32, device=dev) if want_scores else None
    score = torch.empty((n, ho, wo), dtype=torch.float32, device=dev)
    idx = torch.empty((n, ho, wo), dtype=torch.int64, device=dev)
    flags = torch.empty((n, ho, wo), dtype=torch.uint8, device=dev) if cls_flags is not None else None
    _lib.call('eb200_pp_softmax_argmax', _p(x), n, c, h, w, y0, x0, hc, wc, ho, wo, _p(out_logits), _p(scores),
              _p(score), _p(idx), _p(cls_flags), _p(flags), _stream())
    if two_d:
        scores = None if scores is None else scores[:, :, 0, 0]
        score, idx = score[:, 0, 0], idx[:, 0, 0]
    return out_logits, scores, score, idx, flags


def nearest_resize(t: torch.Tensor, box, out_hw) -> torch.Tensor:
    """eb200_pp_nearest_resize on [N,H,W] maps of 1/4/8-byte elements"""
    if not t.is_cuda:
        raise _lib.EB200Error('nearest_resize: CUDA tensor expected')
    t = t.contiguous()
    n, h, w = t.shape
    y0, x0, hc, wc = box
    ho, wo = out_hw
    out = torch.empty((n, ho, wo), dtype=t.dtype, device=t.device)
    _lib.call('eb200_pp_nearest_resize', _p(t), _p(out), t.element_size(), n, h, w, y0, x0, hc, wc, ho, wo, _stream())
    return out


def _crop_resize_nearest(t: torch.Tensor, crop, shape) -> torch.Tensor:
    """dense_base.py:15-58 with mode='nearest'"""
    h, w = t.shape[-2:]
    box = _crop_box(crop, h, w)
    if box == (0, 0, h, w) and tuple(shape) == (h, w):
        return t                                          # the reference returns the (full) view itself
    return nearest_resize(t, box, tuple(shape))


class InstanceTables:
    """device tables of one instance post-processing run"""

    def __init__(self, n: int, device, n_classes: int = 0):
        i32 = dict(dtype=torch.int32, device=device)
        self.centers = torch.zeros((n, MAX_INST, 2), **i32)
        self.scores = torch.zeros((n, MAX_INST), dtype=torch.float32, device=device)
        self.counts = torch.zeros((n,), **i32)
        self.status = torch.zeros((n,), **i32)
        self.areas = torch.zeros((n, MAX_INST), **i32)
        self.votes = torch.zeros((n, MAX_INST, n_classes + 1), **i32) if n_classes else None
        self.inst_pan = None
        self.inst_acc = None


def instance_centers(heat: torch.Tensor, tables: InstanceTables, threshold: float, nms_k: int, top_k: int,
                     fg: Optional[torch.Tensor] = None) -> None:
    n, _, h, w = heat.shape
    lib = _lib.load()
    nbytes = int(lib.eb200_pp_centers_ws_bytes(n, h, w, nms_k))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=heat.device)
    _lib.call('eb200_pp_instance_centers', _p(heat), n, h, w, float(threshold), int(nms_k), int(top_k), _p(fg), _p(ws),
              nbytes, _p(tables.centers), _p(tables.scores), _p(tables.counts), _p(tables.status), _stream())


def instance_assign(offset: torch.Tensor, fg: torch.Tensor, tables: InstanceTables, scale_y: float, scale_x: float,
                    dist_thr: Optional[float], sem_idx: Optional[torch.Tensor] = None, n_classes: int = 0
                    ) -> torch.Tensor:
    n, _, h, w = offset.shape
    seg = torch.empty((n, h, w), dtype=torch.uint8, device=offset.device)
    _lib.call('eb200_pp_instance_assign', _p(offset), _p(fg), _p(tables.centers), _p(tables.counts), n, h, w,
              float(scale_y), float(scale_x), -1.0 if dist_thr is None else float(dist_thr), _p(sem_idx),
              int(n_classes), _p(seg), _p(tables.areas), _p(tables.votes if sem_idx is not None else None), _stream())
    return seg


def panoptic_merge(seg: torch.Tensor, sem_idx: torch.Tensor, cls_flags: torch.Tensor, tables: InstanceTables,
                   scores: Optional[torch.Tensor], orient: Optional[torch.Tensor], n_classes: int):
    """eb200_pp_panoptic_merge: fills tables.inst_pan / tables.inst_acc, returns (pan, pan_sem, sem_score, ins_score,
    pan_score) — the three score maps are None without `scores`."""
    n, h, w = seg.shape
    dev = seg.device
    tables.inst_pan = torch.empty((n, MAX_INST), dtype=torch.int32, device=dev)
    tables.inst_acc = torch.empty((n, MAX_INST, ACC), dtype=torch.float64, device=dev)
    pan = torch.empty((n, h, w), dtype=torch.int64, device=dev)
    pan_sem = torch.empty((n, h, w), dtype=torch.int64, device=dev)
    sem_score = ins_score = pan_score = None
    if scores is not None:
        sem_score, ins_score, pan_score = (torch.empty((n, h, w), dtype=torch.float32, device=dev) for _ in range(3))
    _lib.call('eb200_pp_panoptic_merge', _p(seg), _p(sem_idx), _p(cls_flags), _p(tables.counts), _p(tables.votes),
              _p(scores), _p(orient), _p(tables.scores), n, int(n_classes), h, w, _p(tables.inst_pan), _p(pan),
              _p(pan_sem), _p(sem_score), _p(ins_score), _p(pan_score), _p(tables.inst_acc), _stream())
    return pan, pan_sem, sem_score, ins_score, pan_score


def orientation_sums(orientation: torch.Tensor, seg: torch.Tensor, fg: Optional[torch.Tensor], max_id: int
                     ) -> torch.Tensor:
    """eb200_pp_instance_orientation -> acc fp64 [N, max_id+1, 3] (sum cos, sum sin, pixels) on the device"""
    n, _, h, w = orientation.shape
    acc = torch.empty((n, max_id + 1, 3), dtype=torch.float64, device=orientation.device)
    _lib.call('eb200_pp_instance_orientation', _p(orientation), _p(seg), seg.element_size(), _p(fg), n, h, w,
              int(max_id), _p(acc), _stream())
    return acc


# ------------------------------------------------------------------------------------------------ classes
class _Base:
    def postprocess(self, data, batch, is_training: bool = True):          # postprocessing/base.py:14-24
        if is_training:
            return self._postprocess_training(data, batch)
        return self._postprocess_inference(data, batch)


class SemanticPostprocessingB200(_Base):
    def __init__(self, **kwargs) -> None:
        pass

    def _postprocess_training(self, data, batch):
        output, side_outputs = data
        return {'semantic_output': output, 'semantic_side_outputs': side_outputs}

    def _postprocess_inference(self, data, batch, cls_flags: Optional[torch.Tensor] = None, _keep: Optional[dict] = None):
        output, side_outputs = data
        r = {'semantic_output': output, 'semantic_side_outputs': side_outputs}
        _, pred, score, idx, flags = softmax_argmax(output, cls_flags=cls_flags)
        if _keep is not None:
            _keep['foreground'] = flags
        r.update({'semantic_softmax_scores': pred, 'semantic_segmentation_score': score,
                  'semantic_segmentation_idx': idx})
        crop, shape = valid_region_and_fullres_shape(batch, 'semantic')
        h, w = output.shape[-2:]
        box = _crop_box(crop, h, w)
        if box == (0, 0, h, w) and tuple(shape) == (h, w):
            out_f, pred_f, score_f, idx_f = output, pred, score, idx      # no-op resize: identical values
        else:
            resample = (box[2], box[3]) != tuple(shape)
            out_f, pred_f, score_f, idx_f, _ = softmax_argmax(output, box=box, out_hw=tuple(shape),
                                                              want_logits=resample)
            if not resample:
                out_f = output[..., box[0]:box[0] + box[2], box[1]:box[1] + box[3]]   # the reference's cropped view
        r.update({'semantic_output' + FULLRES_SUFFIX: out_f, 'semantic_softmax_scores' + FULLRES_SUFFIX: pred_f,
                  'semantic_segmentation_score' + FULLRES_SUFFIX: score_f,
                  'semantic_segmentation_idx' + FULLRES_SUFFIX: idx_f})
        return r


class ScenePostprocessingB200(_Base):
    def __init__(self, **kwargs) -> None:
        pass

    def _postprocess_training(self, data, batch):
        output, _ = data
        return {'scene_output': output}

    def _postprocess_inference(self, data, batch):
        output, _ = data
        _, _, score, idx, _ = softmax_argmax(output, want_scores=False)
        return {'scene_class_score': score, 'scene_class_idx': idx, 'scene_output': output}


class InstancePostprocessingB200(_Base):
    def __init__(self, heatmap_threshold: float = 0.1, heatmap_nms_kernel_size: int = 3,
                 heatmap_apply_foreground_mask: bool = False, top_k_instances: int = 64,
                 normalized_offset: bool = True, offset_distance_threshold: Optional[float] = None, **kwargs) -> None:
        assert heatmap_nms_kernel_size % 2 == 1
        assert top_k_instances <= 254
        self._heatmap_nms_kernel_size = heatmap_nms_kernel_size
        self._heatmap_threshold = heatmap_threshold
        self._top_k_instances = top_k_instances
        self._normalized_offset = normalized_offset
        self._heatmap_apply_foreground_mask = heatmap_apply_foreground_mask
        self._offset_distance_threshold = offset_distance_threshold
        self.debug = kwargs.get('debug', False)

    # -- device part: centres + assignment; returns (seg uint8 [N,H,W] on device, tables)
    def segment(self, center_heatmap, center_offset, foreground_mask, sem_idx=None, n_classes: int = 0):
        heat = _dev(center_heatmap, torch.float32, 'center heat map')
        off = _dev(center_offset, torch.float32, 'center offsets')
        if foreground_mask.dtype not in (torch.bool, torch.uint8):
            foreground_mask = foreground_mask != 0
        fg = _dev(foreground_mask, foreground_mask.dtype, 'foreground mask')       # 1 byte per pixel either way
        n, _, h, w = off.shape
        tables = InstanceTables(n, heat.device, n_classes if sem_idx is not None else 0)
        instance_centers(heat, tables, self._heatmap_threshold, self._heatmap_nms_kernel_size, self._top_k_instances,
                         fg if self._heatmap_apply_foreground_mask else None)
        sy, sx = (float(h), float(w)) if self._normalized_offset else (1.0, 1.0)      # instance.py:357-363
        seg = instance_assign(off, fg, tables, sy, sx, self._offset_distance_threshold, sem_idx, n_classes)
        return seg, tables

    @staticmethod
    def meta_from_tables(centers, scores, counts, areas) -> List[Dict[int, Dict]]:
        """instance.py:255-271 from host copies of the tables"""
        cnt = [int(k) for k in counts]
        kmax = max(cnt, default=0)                  # only the populated rows of the 256-row tables are converted
        cyx, sc, ar = centers[:, :kmax].tolist(), scores[:, :kmax].tolist(), areas[:, 1:kmax + 1].tolist()
        metas = [{i + 1: {'center_yx': (cyx[b][i][0], cyx[b][i][1]), 'area': ar[b][i], 'score': sc[b][i]}
                  for i in range(cnt[b])} for b in range(len(cnt))]
        return metas

    def _get_instance_segmentation(self, center_heatmap, center_offset, foreground_mask):
        seg, t = self.segment(center_heatmap, center_offset, foreground_mask)
        host = torch.cat([t.centers.reshape(len(t.counts), -1).float(), t.scores, t.areas.float(),
                          t.counts[:, None].float(), t.status[:, None].float()], dim=1).cpu().numpy()   # one D2H
        m = MAX_INST
        _check_status(host[:, 4 * m + 1])
        metas = self.meta_from_tables(host[:, :2 * m].reshape(-1, m, 2).astype(np.int64), host[:, 2 * m:3 * m],
                                      host[:, 4 * m].astype(np.int64), host[:, 3 * m:4 * m].astype(np.int64))
        return seg, metas

    def _get_instance_orientation(self, orientation, instance_segmentation, foreground_mask) -> List[Dict[int, float]]:
        """instance.py:275-323 on an arbitrary (e.g. ground-truth) instance map: per instance id present inside the
        foreground mask, atan2 of the summed (cos, sin) vectors"""
        orient = _dev(orientation, torch.float32, 'orientation')
        seg = _dev(instance_segmentation, instance_segmentation.dtype, 'instance segmentation')
        if seg.dtype not in (torch.uint8, torch.int16, torch.int32, torch.int64):
            seg = seg.to(torch.int32)
        fg = None
        if foreground_mask is not None:
            fg = foreground_mask if foreground_mask.dtype in (torch.bool, torch.uint8) else foreground_mask != 0
            fg = _dev(fg, fg.dtype, 'orientation foreground mask')
        max_id = 255 if seg.dtype == torch.uint8 else max(int(seg.max().item()), 0)
        acc = orientation_sums(orient, seg, fg, max_id)
        hit = torch.nonzero(acc[..., 2] > 0)                                   # (image, id) pairs, ids ascending
        vals = acc[hit[:, 0], hit[:, 1]]
        host = torch.cat([hit.double(), vals], dim=1).cpu().numpy()           # one D2H
        ang = np.arctan2(host[:, 3].astype(np.float32), host[:, 2].astype(np.float32)).tolist()
        res: List[Dict[int, float]] = [{} for _ in range(orient.shape[0])]
        for (b, i), a in zip(host[:, :2].astype(np.int64).tolist(), ang):
            res[b][i] = a
        return res

    def _postprocess_training(self, data, batch):
        output, side_outputs = data
        return {'instance_output': output, 'instance_side_outputs': side_outputs}

    def _postprocess_inference(self, data, batch):
        output, side_outputs = data
        with_orientation = len(output) == 3
        center_heatmap, center_offset = output[0], output[1]
        r = {'instance_output': output, 'instance_side_outputs': side_outputs, 'instance_centers': center_heatmap,
             'instance_offsets': center_offset}
        if with_orientation:
            r['instance_orientation'] = output[2]
        if self.debug:
            raise NotImplementedError('emsanet_b200 post-processing: debug variants (instance.py:402-421) not covered')
        if 'instance_foreground' in batch:                                  # instance.py:367-400
            seg, meta = self._get_instance_segmentation(center_heatmap, center_offset, batch['instance_foreground'])
            r['instance_segmentation_gt_foreground'] = seg
            r['instance_segmentation_gt_meta'] = meta
            crop, shape = valid_region_and_fullres_shape(batch, 'instance')
            r['instance_segmentation_gt_foreground' + FULLRES_SUFFIX] = _crop_resize_nearest(seg, crop, shape)
        if with_orientation and 'orientation_foreground' in batch and ('instance' in batch or
                                                                       'instance_foreground' in batch):
            if 'instance' in batch:                                         # o-1, instance.py:434-440
                r['orientations_gt_instance_gt_orientation_foreground'] = self._get_instance_orientation(
                    output[2], batch['instance'], batch['orientation_foreground'])
            if 'instance_foreground' in batch:                              # o-2, instance.py:444-451
                r['orientations_instance_segmentation_gt_orientation_foreground'] = self._get_instance_orientation(
                    output[2], r['instance_segmentation_gt_foreground'], batch['orientation_foreground'])
        return r


def _check_status(status) -> None:
    st = np.asarray(status).astype(np.int64)
    if (st & 2).any():
        raise _lib.EB200Error('instance post-processing: more than 255 instance centres in one image (uint8 ids; the '
                              'reference would silently wrap around)')
    if (st & 1).any():       # cannot happen: the workspace is sized by the survivor bound (eb200_pp_centers_ws_bytes)
        raise _lib.EB200Error('instance post-processing: candidate list overflow (internal error)')


class PanopticPostprocessingB200(_Base):
    """`mirror_host_placement=True` returns the panoptic maps as CPU tensors like the reference does (its merge runs
    on the CPU, panoptic.py:140-147); False keeps every map on the device."""

    def __init__(self, semantic_postprocessing: SemanticPostprocessingB200,
                 instance_postprocessing: InstancePostprocessingB200, semantic_classes_is_thing: Sequence[bool],
                 semantic_class_has_orientation: Sequence[bool], normalized_offset: bool = True,
                 compute_scores: bool = False, mirror_host_placement: bool = True, **kwargs) -> None:
        self._semantic_postprocessing = semantic_postprocessing
        self._instance_postprocessing = instance_postprocessing
        self._thing_class_ids = np.where(semantic_classes_is_thing)[0]
        self._thing_ids_panoptic = self._thing_class_ids + 1
        self._orientation_ids = np.where(semantic_class_has_orientation)[0] + 1
        flags = np.zeros(len(semantic_classes_is_thing), np.uint8)
        flags[np.asarray(semantic_classes_is_thing, bool)] |= 1
        flags[np.asarray(semantic_class_has_orientation, bool)[:len(flags)]] |= 2
        self._cls_flags_host = torch.from_numpy(flags)
        self._cls_flags: Dict[Any, torch.Tensor] = {}
        self._normalized_offset = normalized_offset
        self._compute_scores = compute_scores
        self._mirror_host_placement = mirror_host_placement
        self._max_instances_per_category = MAX_INSTANCES_PER_CATEGORY

    @property
    def max_instances_per_category(self):
        return self._max_instances_per_category

    def _flags(self, device) -> torch.Tensor:
        if device not in self._cls_flags:
            self._cls_flags[device] = self._cls_flags_host.to(device)
        return self._cls_flags[device]

    def _postprocess_training(self, data, batch):
        (s_output, i_output), (s_side_outputs, i_side_outputs) = data
        return {**self._semantic_postprocessing._postprocess_training((s_output, s_side_outputs), batch),
                **self._instance_postprocessing._postprocess_training((i_output, i_side_outputs), batch)}

    def _postprocess_inference(self, data, batch):
        (s_output, i_output), (s_side_outputs, i_side_outputs) = data
        post = self._instance_postprocessing
        dev = s_output.device
        flags = self._flags(dev)
        if flags.numel() != s_output.shape[1]:
            raise ValueError(f'{flags.numel()} is-thing flags for {s_output.shape[1]} semantic classes')
        keep: Dict[str, torch.Tensor] = {}
        r = {**self._semantic_postprocessing._postprocess_inference((s_output, s_side_outputs), batch, cls_flags=flags,
                                                                    _keep=keep),
             **post._postprocess_inference((i_output, i_side_outputs), batch)}
        with_orientation = len(i_output) == 3
        sem_idx = r['semantic_segmentation_idx']
        fg = keep['foreground']                                             # isin(idx, thing classes), panoptic.py:123-128
        r['panoptic_foreground_mask'] = fg.view(torch.bool)
        n_classes = s_output.shape[1]
        if post._normalized_offset != self._normalized_offset:
            raise ValueError('normalized_offset differs between the instance and the panoptic post-processing')
        seg, t = post.segment(i_output[0], i_output[1], fg, sem_idx=sem_idx, n_classes=n_classes)
        n, h, w = seg.shape
        scores = _dev(r['semantic_softmax_scores'], torch.float32, 'scores') if self._compute_scores else None
        orient = _dev(i_output[2], torch.float32, 'orientation') if with_orientation else None
        pan, pan_sem, sem_score, ins_score, pan_score = panoptic_merge(seg, sem_idx, flags, t, scores, orient, n_classes)

        # ---- ONE device-to-host copy of every per-instance table
        m = MAX_INST
        host = torch.cat([t.centers.reshape(n, -1).double(), t.scores.double(), t.areas.double(), t.inst_pan.double(),
                          t.inst_acc.reshape(n, -1), t.counts[:, None].double(), t.status[:, None].double()],
                         dim=1).cpu().numpy()
        centers = host[:, :2 * m].reshape(n, m, 2).astype(np.int64)
        cscore = host[:, 2 * m:3 * m].astype(np.float32)
        areas = host[:, 3 * m:4 * m].astype(np.int64)
        inst_pan = host[:, 4 * m:5 * m].astype(np.int64)
        acc = host[:, 5 * m:5 * m + m * ACC].reshape(n, m, ACC)
        counts = host[:, 5 * m + m * ACC].astype(np.int64)
        _check_status(host[:, 5 * m + m * ACC + 1])
        meta = InstancePostprocessingB200.meta_from_tables(centers, cscore, counts, areas)
        # per-instance scalars for the whole batch at once (fp32 like the reference's tensors), then plain Python
        # objects for the dictionaries: panoptic.py:204-233, instance.py:300-321
        kk = int(counts.max()) + 1 if n else 1                              # populated table rows only
        acc, inst_pan = acc[:, :kk], inst_pan[:, :kk]
        cs_id = np.zeros((n, kk), np.float32)
        cs_id[:, 1:] = cscore[:, :kk - 1]                                   # centre score by instance id
        s_sem = (acc[..., 0] / np.maximum(acc[..., 1], 1.0)).astype(np.float32)
        p_sc = (s_sem * cs_id).astype(np.float32)
        angle = np.arctan2(acc[..., 3].astype(np.float32), acc[..., 2].astype(np.float32))
        pid_l, sem_l, psc_l, ang_l, ocnt_l = (inst_pan.tolist(), s_sem.tolist(), p_sc.tolist(), angle.tolist(),
                                              acc[..., 4].tolist())
        ids: List[Dict[int, int]] = []
        orientations: List[Dict[int, float]] = []
        for b in range(n):
            d, o = {}, {}
            for i in range(1, int(counts[b]) + 1):
                pid = pid_l[b][i]
                if pid:
                    d[pid] = i
                    if self._compute_scores:
                        meta[b][i].update({'semantic_score': sem_l[b][i], 'semantic_idx': pid >> 16,
                                           'panoptic_score': psc_l[b][i], 'panoptic_id': pid})
                if with_orientation and ocnt_l[b][i] > 0:
                    o[i] = ang_l[b][i]
            ids.append(d)
            orientations.append(o)

        # ---- full-resolution copies (panoptic.py:239-286) are taken on the device, then everything the reference
        # keeps on the CPU is moved there in one go
        maps = {'panoptic_segmentation_deeplab': pan, 'panoptic_segmentation_deeplab_semantic_idx': pan_sem,
                'panoptic_segmentation_deeplab_instance_idx': seg}
        if self._compute_scores:
            maps.update({'panoptic_segmentation_deeplab_semantic_score': sem_score,
                         'panoptic_segmentation_deeplab_instance_score': ins_score,
                         'panoptic_segmentation_deeplab_panoptic_score': pan_score})
        crop, shape = valid_region_and_fullres_shape(batch, 'instance')
        box = _crop_box(crop, h, w)
        same = box == (0, 0, h, w) and tuple(shape) == (h, w)
        keep_on_device = 'panoptic_segmentation_deeplab_instance_idx'     # the raw instance map stays on the device
        fulls = {key: (src if same else nearest_resize(src, box, tuple(shape))) for key, src in maps.items()}
        if self._mirror_host_placement:
            keys = [k for k in maps if k != keep_on_device]
            moved = _to_host([maps[k] for k in keys] + ([] if same else [fulls[k] for k in keys]))
            for j, k in enumerate(keys):
                maps[k] = moved[j]
                fulls[k] = moved[j] if same else moved[len(keys) + j]
        for key in maps:
            r[key] = maps[key]
            r[key + FULLRES_SUFFIX] = fulls[key]
        r['panoptic_segmentation_deeplab_ids'] = ids
        r['panoptic_segmentation_deeplab_instance_meta'] = meta
        if with_orientation:                                                # panoptic.py:289-314
            r['orientations_panoptic_segmentation_deeplab_instance'] = orientations
            for b in range(n):
                for i in meta[b]:
                    meta[b][i]['orientation'] = orientations[b].get(i, float('nan'))
        return r


# ------------------------------------------------------------------------------------------------ installation
def _kind(obj) -> str:
    """class name of a post-processing object, looking through the reference's partial_class wrapper
    (MT/utils: get_postprocessing_class returns an anonymous subclass named 'PartialClass')"""
    for cls in type(obj).__mro__:
        if cls.__name__.endswith('Postprocessing') or cls.__name__.endswith('PostprocessingB200'):
            return cls.__name__
    return type(obj).__name__


def build_for(obj):
    """B200 mirror of one reference post-processing object (read through its private attributes)."""
    name = _kind(obj)
    if name.startswith('SemanticPostprocessing'):
        return SemanticPostprocessingB200()
    if name.startswith('ScenePostprocessing'):
        return ScenePostprocessingB200()
    if name.startswith('InstancePostprocessing'):
        return InstancePostprocessingB200(
            heatmap_threshold=obj._heatmap_threshold, heatmap_nms_kernel_size=obj._heatmap_nms_kernel_size,
            heatmap_apply_foreground_mask=obj._heatmap_apply_foreground_mask, top_k_instances=obj._top_k_instances,
            normalized_offset=obj._normalized_offset, offset_distance_threshold=obj._offset_distance_threshold,
            debug=getattr(obj, 'debug', False))
    raise NotImplementedError(f'emsanet_b200 post-processing: no mirror for {name}')


def install(model, semantic_n_classes: Optional[int] = None, mirror_host_placement: bool = True):
    """Swap `decoder.postprocessing` of every decoder of `model` (a reference EMSANet or an EMSANetB200 carrying
    reference post-processing objects) for the B200 mirrors.  Returns the model."""
    for name, dec in model.decoders.items():
        old = dec.postprocessing
        if old is None or _kind(old).endswith('B200'):
            continue
        if _kind(old).startswith('PanopticPostprocessing'):
            sem = build_for(old._semantic_postprocessing)
            ins = build_for(old._instance_postprocessing)
            n_cls = semantic_n_classes or len(model.dataset_config.semantic_label_list_without_void)
            is_thing = np.zeros(n_cls, bool)
            is_thing[old._thing_class_ids] = True
            has_or = np.zeros(n_cls, bool)
            has_or[old._orientation_ids - 1] = True
            new = PanopticPostprocessingB200(sem, ins, tuple(is_thing), tuple(has_or),
                                             normalized_offset=old._normalized_offset,
                                             compute_scores=old._compute_scores,
                                             mirror_host_placement=mirror_host_placement)
            for sub, pp in (('semantic_decoder', sem), ('instance_decoder', ins)):
                d = getattr(dec, sub, None)
                if d is not None and hasattr(d, '_postprocessing'):
                    d._postprocessing = pp
        else:
            new = build_for(old)
        if hasattr(dec, '_postprocessing'):
            dec._postprocessing = new            # reference decoders: read-only property over _postprocessing
        else:
            dec.postprocessing = new
    return model


def make_postprocessing(args, semantic_classes_is_thing: Sequence[bool], semantic_class_has_orientation: Sequence[bool],
                        mirror_host_placement: bool = True) -> Dict[str, Any]:
    """The post-processing objects emsanet/decoder.py:60-167 attaches, from the same `args` fields
    (instance_center_heatmap_*, instance_offset_*; emsanet/model.py:112-135)."""
    out: Dict[str, Any] = {}
    tasks = tuple(args.tasks)
    sem = SemanticPostprocessingB200() if 'semantic' in tasks else None
    ins = None
    if 'instance' in tasks:
        enc = getattr(args, 'instance_offset_encoding', 'tanh')
        if enc not in ('tanh', 'relative', 'deeplab'):
            raise NotImplementedError(enc)
        ins = InstancePostprocessingB200(
            heatmap_threshold=getattr(args, 'instance_center_heatmap_threshold', 0.1),
            heatmap_nms_kernel_size=getattr(args, 'instance_center_heatmap_nms_kernel_size', 17),
            heatmap_apply_foreground_mask=getattr(args, 'instance_center_heatmap_apply_foreground_mask', False),
            top_k_instances=getattr(args, 'instance_center_heatmap_top_k', 64),
            normalized_offset=enc != 'deeplab',
            offset_distance_threshold=getattr(args, 'instance_offset_distance_threshold', None),
            debug=getattr(args, 'debug', False))
    if getattr(args, 'enable_panoptic', False):
        out['panoptic_helper'] = PanopticPostprocessingB200(
            sem, ins, semantic_classes_is_thing, semantic_class_has_orientation,
            normalized_offset=ins._normalized_offset, compute_scores=True, mirror_host_placement=mirror_host_placement)
        out['panoptic_helper.semantic_decoder'] = sem
        out['panoptic_helper.instance_decoder'] = ins
    else:
        if sem is not None:
            out['semantic_decoder'] = sem
        if ins is not None:
            out['instance_decoder'] = ins
    if 'scene' in tasks:
        out['scene_decoder'] = ScenePostprocessingB200()
    return out


__all__ = ['SemanticPostprocessingB200', 'InstancePostprocessingB200', 'PanopticPostprocessingB200',
           'ScenePostprocessingB200', 'install', 'make_postprocessing', 'softmax_argmax', 'nearest_resize',
           'valid_region_and_fullres_shape']
