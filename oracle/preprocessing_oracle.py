"""CPU oracle of the input normalisation.  TEST INFRASTRUCTURE ONLY.

Restates  MT/data/preprocessing/normalize.py:14-31 (`normalize`: astype(float32) copy, `value -= mean`, `value /= std`
with mean / std broadcast over the spatial axes), :34-71 (NormalizeRGB: ImageNet mean / std x 255 as float32) and
:74-124 (NormalizeDepth: raw depth keeps the invalid value), followed by the HWC -> CHW change of ToTorchTensors
(MT/data/preprocessing/torch.py).  numpy float32, the same operations in the same order -> bit-identical to the reference;
pinned live against the reference's own functions wherever a reference install exists
(tests/test_preprocessing.py::test_oracle_equals_reference_functions)."""
import numpy as np


def normalize_rgb(rgb_u8_nhwc: np.ndarray, mean, std) -> np.ndarray:
    """uint8 [N,H,W,3] -> float32 [N,3,H,W]"""
    assert rgb_u8_nhwc.dtype == np.uint8
    mean, std = np.asarray(mean, np.float32), np.asarray(std, np.float32)
    v = rgb_u8_nhwc.astype(np.float32)
    v -= mean[np.newaxis, np.newaxis, np.newaxis, :]
    v /= std[np.newaxis, np.newaxis, np.newaxis, :]
    return np.ascontiguousarray(v.transpose(0, 3, 1, 2))


def normalize_depth(depth: np.ndarray, mean: float, std: float, raw_depth: bool = False, invalid: float = 0.0) -> np.ndarray:
    """uint16 / int32 [N,H,W] -> float32 [N,1,H,W]"""
    mean, std = np.float32(mean), np.float32(std)
    mask = depth == invalid if raw_depth else None
    v = depth.astype(np.float32)
    v -= mean
    v /= std
    if raw_depth:
        v[mask] = invalid
    return v[:, np.newaxis]
