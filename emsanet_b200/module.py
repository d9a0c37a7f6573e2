"""`EMSANetB200`: host-side mirror of the reference's `EMSANet` nn.Module (emsanet/model.py:26-233).

Same constructor arguments (`args` namespace with the fields of SURVEY.md App. F, `dataset_config` with the two
label lists), same `forward(batch, do_postprocessing=False)` contract, same state_dict keys / shapes / dtypes
 — so reference checkpoints load with `load_state_dict(strict=True)` and checkpoints written
here load into the reference.  Unlike `patch()` (which re-uses a reference module instance) this class does not
need the reference package to be importable; it is what `bench.py`, `smoke()` and the GPU tests run where
/root/reference does not exist.  The forward/backward is always the CUDA engine; there is no PyTorch fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from .engine import RESNET_LAYERS, EngineConfig
from . import patch as _patch


def param_inventory(cfg: EngineConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(state_dict key, shape, kind) in the reference's registration order.
    kind: conv | bias | bn_weight | bn_bias | running_mean | running_var | num_batches | upsample | linear"""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(p, c):
        out.extend([(p + 'weight', (c,), 'bn_weight'), (p + 'bias', (c,), 'bn_bias'),
                    (p + 'running_mean', (c,), 'running_mean'), (p + 'running_var', (c,), 'running_var'),
                    (p + 'num_batches_tracked', (), 'num_batches')])

    def nbt1d(p, cin, c, ds):   # MT/model/block.py:174-199
        out.extend([(p + 'conv1_1.weight', (c, cin, 3, 1), 'conv'), (p + 'conv1_1.bias', (c,), 'bias'),
                    (p + 'conv1_2.weight', (c, c, 1, 3), 'conv')])
        bn(p + 'norm1.', c)
        out.extend([(p + 'conv2_1.weight', (c, c, 3, 1), 'conv'), (p + 'conv2_1.bias', (c,), 'bias'),
                    (p + 'conv2_2.weight', (c, c, 1, 3), 'conv')])
        bn(p + 'norm2.', c)
        if ds:                  # MT/model/backbone/resnet.py:139-143
            out.append((p + 'downsample.0.weight', (c, cin, 1, 1), 'conv'))
            bn(p + 'downsample.1.', c)

    def backbone(p, cin):
        out.append((p + 'conv1.weight', (64, cin, 7, 7), 'conv'))
        bn(p + 'norm1.', 64)
        inpl = 64
        for li, (n, c) in enumerate(zip(cfg.layers, (64, 128, 256, 512)), start=1):
            for b in range(n):
                nbt1d(f'{p}layer{li}.{b}.', inpl if b == 0 else c, c, b == 0 and li > 1)
            inpl = c

    def dense_decoder(p):
        chans = cfg.decoder_n_channels
        for i, (ci, co) in enumerate(zip((512,) + chans[:-1], chans)):
            mp = f'{p}decoder_modules.{i}.'
            out.append((mp + 'conv.conv.weight', (co, ci, 3, 3), 'conv'))
            bn(mp + 'conv.norm.', co)
            for b in range(cfg.decoder_n_blocks):
                nbt1d(f'{mp}blocks.{b}.', co, co, False)
            out.extend([(mp + 'upsample.conv.weight', (co, 1, 3, 3), 'upsample'),
                        (mp + 'upsample.conv.bias', (co,), 'bias0')])
        for i, (cs, cd) in enumerate(zip((256, 128, 64), chans)):
            fp = f'{p}fusions.{i}.layer.'
            out.append((fp + 'conv.weight', (cd, cs, 1, 1), 'conv'))
            bn(fp + 'norm.', cd)

    if 'rgb' in cfg.modalities:
        backbone(cfg.backbone_prefix('rgb'), 3)
    if 'depth' in cfg.modalities:
        backbone(cfg.backbone_prefix('depth'), 1)
    if len(cfg.modalities) == 2:
        for i, c in enumerate((64, 64, 128, 256, 512)):
            for m in ('rgb', 'depth'):   # MT/model/utils.py:84-89
                p = f'encoder.fusions.{i}.weighting_{m}.layers.'
                out.extend([(p + '0.weight', (c // 16, c, 1, 1), 'conv'), (p + '0.bias', (c // 16,), 'bias'),
                            (p + '2.weight', (c, c // 16, 1, 1), 'conv'), (p + '2.bias', (c,), 'bias')])
    cred = 512 // len(cfg.ppm_bins)
    for i in range(len(cfg.ppm_bins)):
        out.append((f'context_module.features.{i}.1.conv.weight', (cred, 512, 1, 1), 'conv'))
        bn(f'context_module.features.{i}.1.norm.', cred)
    out.append(('context_module.final_conv.conv.weight', (512, 512 + cred * len(cfg.ppm_bins), 1, 1), 'conv'))
    bn('context_module.final_conv.norm.', 512)
    pre = cfg.decoder_prefixes
    c_last = cfg.decoder_n_channels[-1]
    if 'semantic' in pre:
        p = pre['semantic']
        dense_decoder(p)
        nc = cfg.semantic_n_classes
        out.extend([(p + '_task_head.conv.weight', (nc, c_last, 3, 3), 'conv'), (p + '_task_head.conv.bias', (nc,), 'bias')])
        for u in range(2):
            out.extend([(p + f'_task_head.upsample_{u}.conv.weight', (nc, 1, 3, 3), 'upsample'),
                        (p + f'_task_head.upsample_{u}.conv.bias', (nc,), 'bias0')])
        for i, c in enumerate(cfg.decoder_n_channels):
            out.extend([(p + f'_side_output_heads.{i}.conv.weight', (nc, c, 1, 1), 'conv'),
                        (p + f'_side_output_heads.{i}.conv.bias', (nc,), 'bias')])
    if 'instance' in pre:
        p = pre['instance']
        dense_decoder(p)
        nt = 3 if cfg.with_orientation else 2
        touts = (1, 2, 2)[:nt]

        def head(hp, cin, k, n_up):
            out.append((hp + 'shared_conv.conv.weight', (32 * nt, cin, 3, 3), 'conv'))
            bn(hp + 'shared_conv.norm.', 32 * nt)
            for t, co in enumerate(touts):
                out.extend([(hp + f'task_convs.{t}.weight', (co, 32, k, k), 'conv'),
                            (hp + f'task_convs.{t}.bias', (co,), 'bias')])
            for u in range(n_up):
                out.extend([(hp + f'upsampling.{u}.conv.weight', (sum(touts), 1, 3, 3), 'upsample'),
                            (hp + f'upsampling.{u}.conv.bias', (sum(touts),), 'bias0')])
        head(p + '_task_head.', c_last, 3, 2)
        for i, c in enumerate(cfg.decoder_n_channels):
            head(p + f'_side_output_heads.{i}.', c, 1, 0)
    if 'scene' in pre:
        p = pre['scene']
        out.extend([(p + '_task_head.weight', (cfg.scene_n_classes, cred), 'linear'),
                    (p + '_task_head.bias', (cfg.scene_n_classes,), 'bias')])
    return out


class _Node(nn.Module):
    """anonymous container: the reference's module tree is rebuilt from the dotted state_dict keys"""


class _Decoder(_Node):
    postprocessing = None
    side_output_downscales: Tuple[int, ...] = ()


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, is_buffer: bool) -> None:
    parts = dotted.split('.')
    mod = root
    for depth, name in enumerate(parts[:-1]):
        if name not in mod._modules:
            if depth == 0 and name == 'decoders':
                child = nn.ModuleDict()          # iterated by EMSANet.forward (emsanet/model.py:220-227)
            elif parts[0] == 'decoders' and depth in (1, 2) and (name.endswith('_decoder') or name == 'panoptic_helper'):
                child = _Decoder()
            else:
                child = _Node()
            mod.add_module(name, child)
        mod = mod._modules[name]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor))


def default_args(**overrides):
    """argparse.Namespace with the reference defaults that shape the hot path."""
    import argparse
    a = dict(input_modalities=('rgb', 'depth'), input_height=480, input_width=640,
             tasks=('semantic', 'scene', 'instance', 'orientation'), enable_panoptic=True, activation='relu',
             encoder_normalization='batchnorm', decoder_normalization='batchnorm', no_pretrained_backbone=True,
             dropout_p=0.1, encoder_fusion='se-add-uni-rgb', encoder_decoder_skip_downsamplings=(4, 8, 16),
             context_module='ppm', upsampling_context_module='bilinear', upsampling_prediction='learned-3x3-zeropad',
             instance_offset_encoding='tanh', instance_center_encoding='sigmoid', he_init=('encoder-fusion',),
             no_zero_init_decoder_residuals=False, debug=False)
    for m in ('rgb', 'depth', 'rgbd'):
        a[f'{m}_encoder_backbone'] = 'resnet34'
        a[f'{m}_encoder_backbone_resnet_block'] = 'nonbottleneck1d'
    for d in ('semantic', 'instance', 'normal'):
        a.update({f'{d}_decoder': 'emsanet', f'{d}_decoder_n_channels': (512, 256, 128),
                  f'{d}_decoder_downsamplings': (16, 8, 4), f'{d}_decoder_block': 'nonbottleneck1d',
                  f'{d}_decoder_block_dropout_p': 0.2, f'{d}_decoder_n_blocks': 3,
                  f'{d}_decoder_upsampling': 'learned-3x3-zeropad', f'{d}_encoder_decoder_fusion': 'add-rgb'})
    a.update(overrides)
    if len(a['input_modalities']) == 1:
        a['encoder_fusion'] = 'none'            # emsanet/args.py:1318-1321
    return argparse.Namespace(**a)


def simple_dataset_config(semantic_n_classes: int = 40, scene_n_classes: int = 10):
    """the two attributes EMSANet.__init__ reads from a DatasetConfig (emsanet/model.py:39-43)"""
    import types
    return types.SimpleNamespace(semantic_label_list_without_void=list(range(semantic_n_classes)),
                                 scene_label_list_without_void=list(range(scene_n_classes)))


class EMSANetB200(nn.Module):
    def __init__(self, args, dataset_config) -> None:
        super().__init__()
        self.args = args
        self.dataset_config = dataset_config
        cfg = _patch.config_from_model(self)
        g = torch.Generator().manual_seed(int(torch.initial_seed()) % (2 ** 31))
        zero_res = not getattr(args, 'no_zero_init_decoder_residuals', False)
        for key, shape, kind in param_inventory(cfg):
            if kind == 'conv' or kind == 'linear':       # nn.Conv2d / nn.Linear default (kaiming_uniform a=sqrt(5))
                fan_in = math.prod(shape[1:])
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
                if 'encoder.fusions' in key:   # he_init=('encoder-fusion',), initialization.py:29-66
                    t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
            elif kind == 'bias':
                t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            elif kind == 'bias0':
                t = torch.zeros(shape)
            elif kind == 'upsample':                      # MT/model/upsampling.py:63-69
                t = (torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16.).expand(shape).clone()
            elif kind == 'bn_weight':                     # zero_residual_initialization, initialization.py:69-81
                t = torch.zeros(shape) if (zero_res and key.startswith('decoders.') and '.blocks.' in key
                                           and key.endswith('norm2.weight')) else torch.ones(shape)
            elif kind in ('bn_bias', 'running_mean'):
                t = torch.zeros(shape)
            elif kind == 'running_var':
                t = torch.ones(shape)
            elif kind == 'num_batches':
                t = torch.zeros((), dtype=torch.long)
            else:
                raise AssertionError(kind)
            _attach(self, key, t, kind in ('running_mean', 'running_var', 'num_batches'))
        # what the reference's scripts read from the decoders (main.py:391-393)
        for name, dec in self.decoders.items():
            scales = tuple(16 // 2 ** i for i in range(len(cfg.decoder_n_channels))) if name != 'scene_decoder' else ()
            dec.side_output_downscales = scales
        self._eb200_cfg = cfg
        # inference post-processing on the GPU (emsanet/decoder.py:60-167 attaches the reference's objects here)
        from . import postprocessing as _pp
        labels = dataset_config.semantic_label_list_without_void
        n_cls = len(labels)
        is_thing = tuple(getattr(labels, 'classes_is_thing', (True,) * n_cls))
        has_orientation = tuple(getattr(labels, 'classes_use_orientations', (True,) * n_cls))
        for path, pp in _pp.make_postprocessing(args, is_thing, has_orientation).items():
            self.decoders.get_submodule(path).postprocessing = pp

    def forward(self, batch: Dict[str, torch.Tensor], do_postprocessing: bool = False):
        mods = self.args.input_modalities
        res = _patch.run_model(self, batch['rgb'] if 'rgb' in mods else None,
                               batch['depth'] if 'depth' in mods else None)
        return _patch.assemble_outputs(self, res, batch, do_postprocessing)
