"""CPU restatement of the five eb200_pp_* C-ABI calls, one function per call, with the signatures of the
Python wrappers in emsanet_b200/postprocessing.py.  TEST INFRASTRUCTURE ONLY.

Two uses: (1) `-m "not gpu"` tests monkeypatch the wrappers with these and run the host-side classes
(table unpacking, meta dictionaries, key set, placement) against the golden fixtures made from the
reference; (2) `-m gpu` tests compare every C-ABI call with its restatement on the same inputs.
Built on oracle/postprocessing_oracle.py (which is pinned against the reference).
"""
from typing import Optional

import numpy as np
import torch

from . import postprocessing_oracle as P

MAX_INST, ACC = 256, 5


def softmax_argmax(logits, box=None, out_hw=None, want_scores=True, want_logits=False, cls_flags=None):
    two_d = logits.ndim == 2
    x = logits.detach().float().cpu()
    if two_d:
        x = x[:, :, None, None]
    n, c, h, w = x.shape
    y0, x0, hc, wc = box or (0, 0, h, w)
    ho, wo = out_hw or (hc, wc)
    v = P.crop_resize(x, (slice(y0, y0 + hc), slice(x0, x0 + wc)), (ho, wo), 'bilinear').contiguous()
    pred, score, idx = P.softmax_max(v)
    flags = None
    if cls_flags is not None:
        flags = (cls_flags.cpu()[idx] & 1).to(torch.uint8)
    if two_d:
        pred, score, idx = pred[:, :, 0, 0], score[:, 0, 0], idx[:, 0, 0]
    return (v if want_logits else None), (pred if want_scores else None), score, idx, flags


def nearest_resize(t, box, out_hw):
    y0, x0, hc, wc = box
    return P.crop_resize(t.cpu(), (slice(y0, y0 + hc), slice(x0, x0 + wc)), tuple(out_hw), 'nearest').contiguous()


def instance_centers(heat, tables, threshold, nms_k, top_k, fg: Optional[torch.Tensor] = None) -> None:
    heat = heat.detach().float().cpu()
    _, lists = P.instance_centers(heat, threshold, nms_k, top_k, None if fg is None else fg.cpu(),
                                  apply_foreground_mask=fg is not None)
    tables.centers.zero_()
    tables.scores.zero_()
    for b, cen in enumerate(lists):
        k = min(len(cen), MAX_INST - 1)
        tables.status[b] = 2 if len(cen) > MAX_INST - 1 else 0
        tables.counts[b] = k
        if k:
            tables.centers[b, :k] = torch.from_numpy(cen[:k].astype(np.int32))
            tables.scores[b, :k] = heat[b, 0][cen[:k, 0], cen[:k, 1]]


def instance_assign(offset, fg, tables, scale_y, scale_x, dist_thr, sem_idx=None, n_classes=0):
    off = offset.detach().float().cpu().numpy()
    fgm = fg.cpu().numpy().astype(bool)
    n, _, h, w = off.shape
    seg = np.zeros((n, h, w), np.uint8)
    tables.areas.zero_()
    if sem_idx is not None:
        tables.votes.zero_()
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing='ij')
    for b in range(n):
        k = int(tables.counts[b])
        if k == 0:
            continue
        cen = tables.centers[b, :k].cpu().numpy().astype(np.float32)
        ly = (yy + off[b, 0] * np.float32(scale_y))[fgm[b]]
        lx = (xx + off[b, 1] * np.float32(scale_x))[fgm[b]]
        dy, dx = cen[:, 0:1] - ly[None], cen[:, 1:2] - lx[None]
        dist = np.sqrt(dy * dy + dx * dx)
        ids = (dist.argmin(axis=0) + 1).astype(np.uint8)
        if dist_thr is not None:
            ids[dist.min(axis=0) > np.float32(dist_thr)] = 0
        seg[b][fgm[b]] = ids
        tables.areas[b] = torch.from_numpy(np.bincount(ids, minlength=MAX_INST)[:MAX_INST].astype(np.int32))
        if sem_idx is not None:
            cls = sem_idx[b].cpu().numpy()[fgm[b]] + 1
            sel = ids > 0
            votes = np.zeros((MAX_INST, n_classes + 1), np.int32)
            np.add.at(votes, (ids[sel].astype(np.int64), cls[sel]), 1)
            tables.votes[b] = torch.from_numpy(votes)
    return torch.from_numpy(seg)


def panoptic_merge(seg, sem_idx, cls_flags, tables, scores, orient, n_classes):
    segn = seg.cpu().numpy()
    sem = sem_idx.cpu().numpy()
    flags = cls_flags.cpu().numpy()
    n, h, w = segn.shape
    votes = tables.votes.cpu().numpy()
    inst_pan = np.zeros((n, MAX_INST), np.int32)
    acc = np.zeros((n, MAX_INST, ACC), np.float64)
    pan = np.zeros((n, h, w), np.int64)
    for b in range(n):
        tracker = {}
        for i in range(1, int(tables.counts[b]) + 1):
            v = votes[b, i]
            if v.sum() == 0:
                continue
            cls = int(v.argmax())
            if cls == 0:
                continue
            tracker[cls] = tracker.get(cls, 0) + 1
            inst_pan[b, i] = cls * 65536 + tracker[cls]
        stuff = np.where((flags[sem[b]] & 1) != 0, 0, (sem[b] + 1) << 16)
        pan[b] = np.where(segn[b] > 0, inst_pan[b][segn[b]], stuff)
    pan_sem = pan >> 16
    sem_score = ins_score = pan_score = None
    if scores is not None:
        sc = scores.detach().cpu().numpy()
        void = pan_sem == 0
        sem_score = np.where(void, np.float32(0), np.take_along_axis(
            sc, np.where(void, 0, pan_sem - 1)[:, None], axis=1)[:, 0]).astype(np.float32)
    o = None if orient is None else orient.detach().cpu().numpy()
    for b in range(n):
        for i in range(1, int(tables.counts[b]) + 1):
            m = (segn[b] == i) & (pan[b] > 0)
            acc[b, i, 1] = m.sum()
            if sem_score is not None:
                acc[b, i, 0] = sem_score[b][m].astype(np.float64).sum()
            if o is not None:
                mo = m & ((flags[np.maximum(pan_sem[b] - 1, 0)] & 2) != 0) & (pan_sem[b] > 0)
                acc[b, i, 2], acc[b, i, 3], acc[b, i, 4] = o[b, 0][mo].astype(np.float64).sum(), \
                    o[b, 1][mo].astype(np.float64).sum(), mo.sum()
    if scores is not None:
        ins_score = np.zeros((n, h, w), np.float32)
        pan_score = sem_score.copy()
        cs = tables.scores.cpu().numpy()
        for b in range(n):
            for i in range(1, int(tables.counts[b]) + 1):
                if inst_pan[b, i]:
                    m = segn[b] == i
                    ins_score[b][m] = cs[b, i - 1]
                    pan_score[b][m] = np.float32(acc[b, i, 0] / acc[b, i, 1]) * cs[b, i - 1]
    tables.inst_pan = torch.from_numpy(inst_pan)
    tables.inst_acc = torch.from_numpy(acc)
    t = torch.from_numpy
    return (t(pan), t(pan_sem), None if sem_score is None else t(sem_score),
            None if ins_score is None else t(ins_score), None if pan_score is None else t(pan_score))


def orientation_sums(orientation, seg, fg, max_id):
    o = orientation.detach().double().cpu().numpy()
    sg = seg.cpu().numpy().astype(np.int64)
    m = np.ones(sg.shape, bool) if fg is None else fg.cpu().numpy().astype(bool)
    n = sg.shape[0]
    acc = np.zeros((n, max_id + 1, 3), np.float64)
    for b in range(n):
        sel = m[b] & (sg[b] > 0) & (sg[b] <= max_id)
        ids = sg[b][sel]
        np.add.at(acc[b, :, 0], ids, o[b, 0][sel])
        np.add.at(acc[b, :, 1], ids, o[b, 1][sel])
        np.add.at(acc[b, :, 2], ids, 1.0)
    return torch.from_numpy(acc)
