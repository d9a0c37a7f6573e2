"""GPU-side input normalisation behind the reference's preprocessing parameters.

`NormalizeB200` does on the device, for a whole batch of RAW images, what the reference's CPU data pipeline does per
sample: `NormalizeRGB` / `NormalizeDepth` (MT/data/preprocessing/normalize.py:34-124, MT/ = lib/nicr-multitask-scene-
analysis/src/nicr_mt_scene_analysis/) followed by the HWC -> CHW change of `ToTorchTensors`.  Same parameters (ImageNet
mean / std x 255 for RGB, the dataset's depth mean / std, `raw_depth` + `invalid_depth_value`), bit-identical results
(IEEE float32 subtract and divide).  A loader can then ship uint8 / uint16 images (1.5 MB instead of 4.9 MB per 640x480
RGB-D image across PCIe) and skip the two normalisation steps on its CPU workers — at B200 throughput (> 800 img/s per
GPU) the 8 cv2 + numpy workers of `emsanet/data.py:369-378` are the bottleneck long before the network is.

    norm = NormalizeB200.from_reference(preprocessor)          # parameters of an emsanet.preprocessing pipeline
    batch = norm(rgb_u8.cuda(non_blocking=True), depth_u16.cuda(non_blocking=True))     # {'rgb': ..., 'depth': ...}
    model(batch)

`inference_time_whole_model.py:519-545` scales its random inputs as `/255` and `/20000`: NormalizeB200(rgb_mean=0,
rgb_std=255, depth_mean=0, depth_std=20000) is that recipe.  No CPU path: inputs must be CUDA tensors.
"""
import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib

IMAGENET_MEAN = (0.485 * 255, 0.456 * 255, 0.406 * 255)      # normalize.py:43-46
IMAGENET_STD = (0.229 * 255, 0.224 * 255, 0.225 * 255)


class NormalizeB200:
    def __init__(self, rgb_mean: Sequence[float] = IMAGENET_MEAN, rgb_std: Sequence[float] = IMAGENET_STD,
                 depth_mean: float = 0.0, depth_std: float = 1.0, raw_depth: bool = False,
                 invalid_depth_value: float = 0.0) -> None:
        import numpy as np
        # the reference keeps its parameters as float32 numpy arrays (normalize.py:43-46,90-91): round the same way
        self.rgb_mean = (C.c_float * 3)(*[float(np.float32(v)) for v in rgb_mean])
        self.rgb_std = (C.c_float * 3)(*[float(np.float32(v)) for v in rgb_std])
        self.depth_mean, self.depth_std = float(np.float32(depth_mean)), float(np.float32(depth_std))
        if self.depth_std == 0.0 or any(v == 0.0 for v in self.rgb_std):
            raise ValueError('std must not be zero')
        self.raw_depth, self.invalid_depth_value = bool(raw_depth), float(invalid_depth_value)

    @classmethod
    def from_reference(cls, preprocessor) -> 'NormalizeB200':
        """pick the parameters out of a reference preprocessing pipeline (a torchvision Compose / list of MT
        preprocessors as `emsanet.preprocessing.get_preprocessor` builds it) or of single Normalize* objects"""
        steps = getattr(preprocessor, 'transforms', None) or (preprocessor if isinstance(preprocessor, (list, tuple))
                                                              else [preprocessor])
        kw = {}
        for s in steps:
            name = type(s).__name__
            if name == 'NormalizeRGB':
                kw['rgb_mean'], kw['rgb_std'] = tuple(s._rgb_mean.tolist()), tuple(s._rgb_std.tolist())
            elif name == 'NormalizeDepth':
                kw.update(depth_mean=float(s._depth_mean), depth_std=float(s._depth_std), raw_depth=bool(s._raw_depth),
                          invalid_depth_value=float(s._invalid_depth_value))
        if not kw:
            raise ValueError('no NormalizeRGB / NormalizeDepth step found')
        return cls(**kw)

    def rgb(self, rgb_u8_nhwc: torch.Tensor) -> torch.Tensor:
        x = self._check(rgb_u8_nhwc, 'rgb')
        if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError(f'rgb must be uint8 [N,H,W,3] (NormalizeRGB asserts uint8, normalize.py:64), got '
                             f'{x.dtype} {tuple(x.shape)}')
        n, h, w, _ = x.shape
        out = torch.empty(n, 3, h, w, dtype=torch.float32, device=x.device)
        _lib.call('eb200_normalize_rgb', x.data_ptr(), out.data_ptr(), n, h, w, self.rgb_mean, self.rgb_std,
                  torch.cuda.current_stream().cuda_stream)
        return out

    def depth(self, depth: torch.Tensor) -> torch.Tensor:
        d = self._check(depth, 'depth')
        if d.dim() == 4 and d.shape[1] == 1:
            d = d[:, 0]
        if d.dim() != 3 or d.dtype not in (torch.uint16, torch.int16, torch.int32):
            raise ValueError(f'depth must be uint16 / int32 [N,H,W], got {d.dtype} {tuple(d.shape)}')
        d = d.contiguous()
        n, h, w = d.shape
        out = torch.empty(n, 1, h, w, dtype=torch.float32, device=d.device)
        _lib.call('eb200_normalize_depth', d.data_ptr(), d.element_size(), out.data_ptr(), n, h, w, self.depth_mean,
                  self.depth_std, int(self.raw_depth), self.invalid_depth_value, torch.cuda.current_stream().cuda_stream)
        return out

    def __call__(self, rgb_u8_nhwc: Optional[torch.Tensor] = None, depth: Optional[torch.Tensor] = None
                 ) -> Dict[str, torch.Tensor]:
        batch = {}
        if rgb_u8_nhwc is not None:
            batch['rgb'] = self.rgb(rgb_u8_nhwc)
        if depth is not None:
            batch['depth'] = self.depth(depth)
        return batch

    @staticmethod
    def _check(t: torch.Tensor, what: str) -> torch.Tensor:
        if not t.is_cuda:
            raise _lib.EB200Error(f'{what} must be a CUDA tensor: emsanet_b200 preprocessing has no CPU path')
        return t.contiguous()
