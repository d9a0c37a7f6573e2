"""A synthetic mini-NYUv2 on disk, in the exact layout the reference's dataset class reads
(SURVEY.md §8(f) row 4), so that the reference's own data pipeline — `emsanet/data.py:get_dataset`,
`emsanet/preprocessing.py`, `main.py`, `inference_dataset.py` — runs unchanged in front of the
emsanet_b200 path where the real dataset cannot be downloaded.

Layout (DS/datasets/nyuv2/nyuv2.py:14-56, dataset.py:60-178; DS = lib/nicr-scene-analysis-datasets/src/
nicr_scene_analysis_datasets):

    <root>/train.txt, <root>/test.txt                      one sample name per line
    <root>/<split>/rgb/<name>.png                          uint8  HxWx3
    <root>/<split>/depth/<name>.png, depth_raw/<name>.png  uint16 HxW, millimetres
    <root>/<split>/semantic_40/<name>.png                  uint8  HxW, 0 = void, 1..40
    <root>/<split>/instance/<name>.png                     uint16 HxW, 0 = no instance
    <root>/<split>/orientations/<name>.json                {"<instance id>": angle in rad}
    <root>/<split>/scene_class/<name>.txt                  a scene label of NYUv2Meta.SCENE_LABEL_LIST
    <root>/<split>/normal/<name>.png                       uint8  HxWx3 (optional)

Content: a floor / wall / ceiling layout (stuff classes) with a handful of box-shaped objects (thing classes,
one instance id each, some with an orientation), depth consistent with the layout, colours tied to the class with
pixel noise.  Deterministic for a given seed.  This is data generation for the reference's loader, not a model of
the real dataset's statistics.

    python -m emsanet_b200.synthetic_nyuv2 /tmp/nyuv2_synth --train 16 --test 8
"""
import argparse
import json
import os
from typing import Dict, Tuple

import numpy as np

SPLITS = ('train', 'test')
SCENES = ('bedroom', 'kitchen', 'living_room', 'office', 'bathroom')         # members of NYUv2Meta.SCENE_LABEL_LIST
STUFF = {'wall': 1, 'floor': 2, 'ceiling': 22}                                # NYUv2-40 ids of the layout classes
THINGS = (3, 4, 5, 6, 7, 10, 14, 24, 25, 33)                                   # cabinet, bed, chair, sofa, table, ...
DEPTH_RANGE_MM = (713, 9995)                                                   # NYUv2Meta.TRAIN_SPLIT_DEPTH_STATS


def _class_colour(c: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + c)
    return rng.integers(40, 216, 3)


def make_sample(rng: np.random.Generator, height: int, width: int) -> Dict[str, object]:
    """one scene: arrays in the dtypes the loader expects + orientation dict + scene label"""
    ys, xs = np.mgrid[0:height, 0:width]
    horizon = int(height * rng.uniform(0.35, 0.55))
    ceiling = int(height * rng.uniform(0.05, 0.15))
    semantic = np.full((height, width), STUFF['wall'], np.uint8)
    semantic[ys >= horizon] = STUFF['floor']
    semantic[ys < ceiling] = STUFF['ceiling']
    far = rng.uniform(3500, 6000)
    depth = np.full((height, width), far, np.float32)
    floor = ys >= horizon
    depth[floor] = far - (far - 900) * ((ys[floor] - horizon) / max(1, height - horizon))
    instance = np.zeros((height, width), np.uint16)
    orientations: Dict[str, float] = {}
    n_obj = int(rng.integers(2, 7))
    for k in range(1, n_obj + 1):
        h = int(rng.integers(height // 8, height // 3))
        w = int(rng.integers(width // 10, width // 3))
        y0 = int(rng.integers(ceiling, max(ceiling + 1, height - h)))
        x0 = int(rng.integers(0, max(1, width - w)))
        cls = int(THINGS[int(rng.integers(0, len(THINGS)))])
        box = (slice(y0, y0 + h), slice(x0, x0 + w))
        d_obj = float(np.clip(depth[min(height - 1, y0 + h - 1), x0 + w // 2] - rng.uniform(100, 600), *DEPTH_RANGE_MM))
        closer = depth[box] > d_obj                                            # objects occlude what lies behind them
        semantic[box][closer] = cls
        instance[box][closer] = k
        depth[box][closer] = d_obj
        if rng.random() < 0.7:
            orientations[str(k)] = float(rng.uniform(0, 2 * np.pi))
    void = rng.random((height, width)) < 0.02                                  # unlabeled pixels (holes)
    semantic[void] = 0
    instance[void] = 0
    for k in list(orientations):                                               # drop orientations of fully occluded objects
        if not (instance == int(k)).any():
            del orientations[k]
    lut = np.stack([_class_colour(c) for c in range(41)]).astype(np.float32)
    rgb = lut[semantic] * (0.6 + 0.4 * (1 - depth[..., None] / 6500.0)) + rng.normal(0, 6, (height, width, 3))
    depth_raw = depth.copy()
    depth_raw[rng.random((height, width)) < 0.05] = 0                          # sensor holes in the raw depth
    normal = np.zeros((height, width, 3), np.float32)
    normal[..., 2] = -1.0
    normal[floor] = (0.0, -1.0, 0.0)
    return {'rgb': np.clip(rgb, 0, 255).astype(np.uint8),
            'depth': np.clip(depth, *DEPTH_RANGE_MM).astype(np.uint16),
            'depth_raw': np.clip(depth_raw, 0, DEPTH_RANGE_MM[1]).astype(np.uint16),
            'semantic': semantic, 'instance': instance, 'orientations': orientations,
            'scene': SCENES[int(rng.integers(0, len(SCENES)))],
            'normal': np.clip((normal + 1) * 127, 0, 254).astype(np.uint8)}


def write_dataset(root: str, n_train: int = 16, n_test: int = 8, height: int = 480, width: int = 640, seed: int = 0,
                  with_normal: bool = False) -> Tuple[int, int]:
    """Write the dataset under `root` (created if needed).  Returns (n_train, n_test)."""
    import cv2
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    for split, count in zip(SPLITS, (n_train, n_test)):
        names = [f'{i:04d}' for i in range(count)]
        dirs = ['rgb', 'depth', 'depth_raw', 'semantic_40', 'instance', 'orientations', 'scene_class']
        if with_normal:
            dirs.append('normal')
        for d in dirs:
            os.makedirs(os.path.join(root, split, d), exist_ok=True)
        for name in names:
            s = make_sample(rng, height, width)
            base = os.path.join(root, split)
            ok = cv2.imwrite(os.path.join(base, 'rgb', name + '.png'), cv2.cvtColor(s['rgb'], cv2.COLOR_RGB2BGR))
            ok &= cv2.imwrite(os.path.join(base, 'depth', name + '.png'), s['depth'])
            ok &= cv2.imwrite(os.path.join(base, 'depth_raw', name + '.png'), s['depth_raw'])
            ok &= cv2.imwrite(os.path.join(base, 'semantic_40', name + '.png'), s['semantic'])
            ok &= cv2.imwrite(os.path.join(base, 'instance', name + '.png'), s['instance'])
            if with_normal:
                ok &= cv2.imwrite(os.path.join(base, 'normal', name + '.png'), cv2.cvtColor(s['normal'], cv2.COLOR_RGB2BGR))
            if not ok:
                raise IOError(f'could not write sample {name} under {base}')
            with open(os.path.join(base, 'orientations', name + '.json'), 'w') as f:
                json.dump(s['orientations'], f)
            with open(os.path.join(base, 'scene_class', name + '.txt'), 'w') as f:
                f.write(s['scene'])
        with open(os.path.join(root, f'{split}.txt'), 'w') as f:
            f.write('\n'.join(names) + '\n')
    return n_train, n_test


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('root')
    ap.add_argument('--train', type=int, default=16)
    ap.add_argument('--test', type=int, default=8)
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--with-normal', action='store_true')
    a = ap.parse_args()
    n = write_dataset(a.root, a.train, a.test, a.height, a.width, a.seed, a.with_normal)
    print(f'wrote {n[0]} train / {n[1]} test samples ({a.width}x{a.height}) to {a.root}')


if __name__ == '__main__':
    main()
