"""CPU oracle for the EMSANet forward/backward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional, self-contained restatement (plain torch fp32 ops on the
CPU, NCHW) of the reference's `EMSANet.forward` for the default model family
(ResNet-NBt1D dual/single encoder, `se-add-uni-rgb` fusion, PPM, EMSANet dense
decoders with `learned-3x3-zeropad` upsampling and `add-rgb` skip fusion).  It is
the *checker* for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  Nothing in
`emsanet_b200/` imports it, and the product path never falls back to it.

Parity pin: the reference ships no golden vectors for this path, so
the oracle is pinned against the reference modules themselves, imported in the
build container from /root/reference by `oracle/make_golden.py`; that script
commits small fixtures under `tests/golden/` which `tests/test_oracle_golden.py`
re-checks everywhere (the GPU box has no /root/reference).

All conv / batch-norm / interpolate arithmetic lives in the third-party
dependency torch (ATen), version 2.11.0 here (reference CI pins 2.10.0); the
oracle calls the same ATen ops the reference's nn.Modules dispatch to.

Reference citations use the aliases of SURVEY.md:
  MT/ = lib/nicr-multitask-scene-analysis/src/nicr_mt_scene_analysis/
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

RESNET_LAYERS = {  # MT/model/backbone/resnet.py:241-252
    'resnet18': (2, 2, 2, 2),
    'resnet34': (3, 4, 6, 3),
    'resnet101': (3, 4, 23, 3),
}


@dataclasses.dataclass(frozen=True)
class OracleConfig:
    """The subset of `args` that shapes the hot path."""
    backbone: str = 'resnet34'
    modalities: Tuple[str, ...] = ('rgb', 'depth')
    tasks: Tuple[str, ...] = ('semantic', 'scene', 'instance', 'orientation')
    enable_panoptic: bool = True
    semantic_n_classes: int = 40
    scene_n_classes: int = 10
    decoder_n_channels: Tuple[int, ...] = (512, 256, 128)
    decoder_n_blocks: int = 3
    ppm_bins: Tuple[int, ...] = (1, 5)
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    # False: the reference's fp32 arithmetic (bit-exact with the reference, see oracle/make_golden.py).
    # True : same algorithm, but every tensor the B200 path keeps in HBM as bf16 (activations, tensor-core conv
    #        weights) is rounded to bf16 at that point (straight-through for autograd).  Used by the GPU parity
    #        tests so that ReLU decisions are taken on the same values on both sides.
    emulate_bf16_storage: bool = False

    @property
    def layers(self) -> Tuple[int, ...]:
        return RESNET_LAYERS[self.backbone]

    def backbone_prefix(self, modality: str) -> str:
        # FusedRGBDEncoder names its backbones backbone_rgb/backbone_depth (MT/model/encoder.py:160-161);
        # the single-modality Encoder holds one `backbone` (MT/model/encoder.py:62-80)
        return f'encoder.backbone_{modality}.' if len(self.modalities) == 2 else 'encoder.backbone.'

    @property
    def with_orientation(self) -> bool:
        return 'orientation' in self.tasks

    @property
    def decoder_prefixes(self) -> Dict[str, str]:
        pre = 'decoders.panoptic_helper.' if self.enable_panoptic else 'decoders.'
        out = {}
        if 'semantic' in self.tasks:
            out['semantic'] = pre + 'semantic_decoder.'
        if 'instance' in self.tasks:
            out['instance'] = pre + 'instance_decoder.'
        if 'scene' in self.tasks:
            out['scene'] = 'decoders.scene_decoder.'
        return out


# --------------------------------------------------------------------------------------
# parameter inventory: name -> shape, in reference state_dict order
# --------------------------------------------------------------------------------------
def _bn(prefix: str, c: int, out: List):
    out += [(prefix + 'weight', (c,)), (prefix + 'bias', (c,)),
            (prefix + 'running_mean', (c,)), (prefix + 'running_var', (c,)),
            (prefix + 'num_batches_tracked', ())]


def _nbt1d(prefix: str, cin: int, c: int, downsample: bool, out: List):
    # MT/model/block.py:174-199
    out += [(prefix + 'conv1_1.weight', (c, cin, 3, 1)), (prefix + 'conv1_1.bias', (c,)),
            (prefix + 'conv1_2.weight', (c, c, 1, 3))]
    _bn(prefix + 'norm1.', c, out)
    out += [(prefix + 'conv2_1.weight', (c, c, 3, 1)), (prefix + 'conv2_1.bias', (c,)),
            (prefix + 'conv2_2.weight', (c, c, 1, 3))]
    _bn(prefix + 'norm2.', c, out)
    if downsample:  # MT/model/backbone/resnet.py:139-143
        out += [(prefix + 'downsample.0.weight', (c, cin, 1, 1))]
        _bn(prefix + 'downsample.1.', c, out)


def _backbone(prefix: str, cin: int, layers, out: List):
    out += [(prefix + 'conv1.weight', (64, cin, 7, 7))]
    _bn(prefix + 'norm1.', 64, out)
    inpl = 64
    for li, (n, c) in enumerate(zip(layers, (64, 128, 256, 512)), start=1):
        for b in range(n):
            _nbt1d(f'{prefix}layer{li}.{b}.', inpl if b == 0 else c, c,
                   b == 0 and (li > 1), out)
        inpl = c


def _dense_decoder(prefix: str, cfg: OracleConfig, out: List):
    chans = cfg.decoder_n_channels
    n_in = (512,) + chans[:-1]
    for i, (ci, co) in enumerate(zip(n_in, chans)):
        p = f'{prefix}decoder_modules.{i}.'
        out += [(p + 'conv.conv.weight', (co, ci, 3, 3))]
        _bn(p + 'conv.norm.', co, out)
        for b in range(cfg.decoder_n_blocks):
            _nbt1d(f'{p}blocks.{b}.', co, co, False, out)
        out += [(p + 'upsample.conv.weight', (co, 1, 3, 3)), (p + 'upsample.conv.bias', (co,))]
    skip_c = (256, 128, 64)
    for i, (cs, cd) in enumerate(zip(skip_c, chans)):
        p = f'{prefix}fusions.{i}.layer.'
        out += [(p + 'conv.weight', (cd, cs, 1, 1))]
        _bn(p + 'norm.', cd, out)


def param_shapes(cfg: OracleConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    out: List = []
    if 'rgb' in cfg.modalities:
        _backbone(cfg.backbone_prefix('rgb'), 3, cfg.layers, out)
    if 'depth' in cfg.modalities:
        _backbone(cfg.backbone_prefix('depth'), 1, cfg.layers, out)
    if len(cfg.modalities) == 2:
        for i, c in enumerate((64, 64, 128, 256, 512)):
            for m in ('rgb', 'depth'):  # MT/model/utils.py:84-89
                p = f'encoder.fusions.{i}.weighting_{m}.layers.'
                out += [(p + '0.weight', (c // 16, c, 1, 1)), (p + '0.bias', (c // 16,)),
                        (p + '2.weight', (c, c // 16, 1, 1)), (p + '2.bias', (c,))]
    cred = 512 // len(cfg.ppm_bins)
    for i in range(len(cfg.ppm_bins)):  # MT/model/context_module/ppm.py:38-54
        out += [(f'context_module.features.{i}.1.conv.weight', (cred, 512, 1, 1))]
        _bn(f'context_module.features.{i}.1.norm.', cred, out)
    out += [('context_module.final_conv.conv.weight', (512, 512 + cred * len(cfg.ppm_bins), 1, 1))]
    _bn('context_module.final_conv.norm.', 512, out)
    pre = cfg.decoder_prefixes
    c_last = cfg.decoder_n_channels[-1]
    if 'semantic' in pre:
        p = pre['semantic']
        _dense_decoder(p, cfg, out)
        nc = cfg.semantic_n_classes
        out += [(p + '_task_head.conv.weight', (nc, c_last, 3, 3)), (p + '_task_head.conv.bias', (nc,))]
        for u in range(2):
            out += [(p + f'_task_head.upsample_{u}.conv.weight', (nc, 1, 3, 3)),
                    (p + f'_task_head.upsample_{u}.conv.bias', (nc,))]
        for i, c in enumerate(cfg.decoder_n_channels):
            out += [(p + f'_side_output_heads.{i}.conv.weight', (nc, c, 1, 1)),
                    (p + f'_side_output_heads.{i}.conv.bias', (nc,))]
    if 'instance' in pre:
        p = pre['instance']
        _dense_decoder(p, cfg, out)
        nt = 3 if cfg.with_orientation else 2
        touts = (1, 2, 2)[:nt]

        def head(hp, cin, k, n_up):
            out.append((hp + 'shared_conv.conv.weight', (32 * nt, cin, 3, 3)))
            _bn(hp + 'shared_conv.norm.', 32 * nt, out)
            for t, co in enumerate(touts):
                out.extend([(hp + f'task_convs.{t}.weight', (co, 32, k, k)),
                            (hp + f'task_convs.{t}.bias', (co,))])
            for u in range(n_up):
                out.extend([(hp + f'upsampling.{u}.conv.weight', (sum(touts), 1, 3, 3)),
                            (hp + f'upsampling.{u}.conv.bias', (sum(touts),))])
        head(p + '_task_head.', c_last, 3, 2)
        for i, c in enumerate(cfg.decoder_n_channels):
            head(p + f'_side_output_heads.{i}.', c, 1, 0)
    if 'scene' in pre:
        p = pre['scene']
        out += [(p + '_task_head.weight', (cfg.scene_n_classes, cred)),
                (p + '_task_head.bias', (cfg.scene_n_classes,))]
    return out


def make_state_dict(cfg: OracleConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded random parameters with the reference's keys/shapes/dtypes.

    Conv/linear weights ~ N(0, 2/fan_in) (he-normal, MT/model/initialization.py:29-66),
    conv biases ~ N(0, 0.05); BN tensors are randomised INCLUDING the decoder
    `norm2.weight` that the reference zero-initialises (pitfall P2 of SURVEY.md), so
    every conv on the path contributes to the outputs.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg):
        leaf = name.rsplit('.', 1)[1]
        if leaf == 'num_batches_tracked':
            sd[name] = torch.zeros((), dtype=torch.long)
        elif leaf == 'running_mean':
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == 'running_var':
            sd[name] = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) == 1 and ('norm' in name or 'downsample.1' in name):
            if leaf == 'weight':
                sd[name] = 0.5 + torch.rand(shape, generator=g)
            else:
                sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # conv / linear bias
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif 'upsampl' in name and shape[1:] == (1, 3, 3):
            base = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16.  # MT/model/upsampling.py:63-69
            sd[name] = base * (1.0 + 0.2 * torch.randn(shape, generator=g))
        else:
            fan_in = math.prod(shape[1:])
            sd[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
    return sd


def make_inputs(n: int, h: int, w: int, seed: int = 1) -> Tuple[Tensor, Tensor]:
    """randn rgb N×3×H×W and depth N×1×H×W from one seeded generator."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randn(n, 3, h, w, generator=g)
    depth = torch.randn(n, 1, h, w, generator=g)
    return rgb, depth


def dropout_sites(cfg: OracleConfig) -> List[Tuple[str, int, float]]:
    """(block prefix, channels, default p) for every Dropout2d on the path, in execution order
    of the reference forward (MT/model/block.py:213-214; p: emsanet/args.py:619-624, 332-337)."""
    sites = []
    # execution order: stage-interleaved rgb/depth (MT/model/encoder.py:233-243)
    for li, (n, c) in enumerate(zip(cfg.layers, (64, 128, 256, 512)), start=1):
        for m in ('rgb', 'depth'):
            if m in cfg.modalities:
                for b in range(n):
                    sites.append((f'{cfg.backbone_prefix(m)}layer{li}.{b}.', c, 0.1))
    for task in ('semantic', 'instance'):
        if task in cfg.decoder_prefixes:
            for i, c in enumerate(cfg.decoder_n_channels):
                for b in range(cfg.decoder_n_blocks):
                    sites.append((f'{cfg.decoder_prefixes[task]}decoder_modules.{i}.blocks.{b}.', c, 0.2))
    return sites


# --------------------------------------------------------------------------------------
# functional forward
# --------------------------------------------------------------------------------------
# Test-only switch (oracle/teacher_forced.py): with it on, the BACKWARD pass of a bf16-storage run also rounds the
# gradient that flows through every storage point to bf16 — what a path that stores activation gradients as bf16 does.
ROUND_GRADS = [False]


class _GradStorage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype) if ROUND_GRADS[0] else g


class _Ctx:
    def __init__(self, sd, cfg, training, dropout_masks, track_running_stats, taps, teacher=None, computed=None):
        self.sd, self.cfg, self.training = sd, cfg, training
        self.teacher, self.computed = teacher, computed
        self.masks = dropout_masks
        self.track = track_running_stats
        self.new_stats: Dict[str, Tensor] = {}
        self.taps = taps  # optional dict name -> tensor for per-layer checks

    def tap(self, name, x):
        if self.taps is not None:
            self.taps[name] = x
        return x

    def q(self, t, name=None):
        """bf16 storage point of the B200 path (identity in the fp32 reference arithmetic).

        Teacher forcing (tests only): with `teacher[name]` given, the value that flows on is the CUDA path's own stored
        tensor (straight-through for autograd) and what this oracle computed for it is kept in `computed[name]` — every
        layer is then checked on identical inputs and the backward pass runs on identical activations / ReLU masks, so
        that gradients of two correct implementations agree to rounding instead of drifting apart chaotically."""
        if self.cfg.emulate_bf16_storage:
            t = t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())
            if t.requires_grad:
                t = _GradStorage.apply(t)
        if name is not None and self.computed is not None:
            self.computed[name] = t.detach()
        if name is not None and self.teacher is not None and name in self.teacher:
            forced = self.teacher[name].to(t.dtype)[:, :t.shape[1]]     # the CUDA path pads channels to a multiple of 8
            assert forced.shape == t.shape, (name, tuple(forced.shape), tuple(t.shape))
            t = t + (forced - t.detach())
        return t

    def w(self, key):
        """tensor-core conv weight (consumed as bf16 by the B200 path)"""
        return self.q(self.sd[key])


def _batch_norm(ctx: _Ctx, x: Tensor, p: str) -> Tensor:
    """nn.BatchNorm2d (MT/model/normalization.py:30-31): train = batch stats (biased var) and a
    running update with the unbiased var, momentum 0.1; eval = running stats."""
    sd = ctx.sd
    w, b = sd[p + 'weight'], sd[p + 'bias']
    if not ctx.training:
        return F.batch_norm(x, sd[p + 'running_mean'], sd[p + 'running_var'], w, b, False, 0.0, ctx.cfg.bn_eps)
    rm = sd[p + 'running_mean'].detach().clone()
    rv = sd[p + 'running_var'].detach().clone()
    y = F.batch_norm(x, rm if ctx.track else None, rv if ctx.track else None, w, b, True,
                     ctx.cfg.bn_momentum, ctx.cfg.bn_eps)
    if ctx.track:
        ctx.new_stats[p + 'running_mean'] = rm
        ctx.new_stats[p + 'running_var'] = rv
        ctx.new_stats[p + 'num_batches_tracked'] = sd[p + 'num_batches_tracked'] + 1
    return y


def _nbt1d_fwd(ctx: _Ctx, x: Tensor, p: str, stride: int) -> Tensor:
    """NonBottleneck1D.forward, MT/model/block.py:201-221."""
    sd, q = ctx.sd, ctx.q
    y = q(F.relu(F.conv2d(x, ctx.w(p + 'conv1_1.weight'), sd[p + 'conv1_1.bias'], (stride, 1), (1, 0))), p + 'a11')
    y = q(F.conv2d(y, ctx.w(p + 'conv1_2.weight'), None, (1, stride), (0, 1)), p + 'c12')
    y = q(F.relu(_batch_norm(ctx, y, p + 'norm1.')), p + 'a12')
    y = q(F.relu(F.conv2d(y, ctx.w(p + 'conv2_1.weight'), sd[p + 'conv2_1.bias'], 1, (1, 0))), p + 'a21')
    y = q(F.conv2d(y, ctx.w(p + 'conv2_2.weight'), None, 1, (0, 1)), p + 'c22')
    y = _batch_norm(ctx, y, p + 'norm2.')
    if ctx.training and ctx.masks is not None and p in ctx.masks:
        # Dropout2d: per-(n,c) keep mask already scaled by 1/(1-p) (block.py:213-214)
        y = y * ctx.masks[p][:, :, None, None]
    if (p + 'downsample.0.weight') in sd:  # resnet.py:139-143
        idt = q(F.conv2d(x, ctx.w(p + 'downsample.0.weight'), None, stride), p + 'cds')
        idt = q(_batch_norm(ctx, idt, p + 'downsample.1.'), p + 'idt')
    else:
        idt = x
    return ctx.tap(p + 'out', q(F.relu(y + idt), p + 'out'))


def _backbone_stage(ctx: _Ctx, x: Tensor, p: str, stage: int) -> Tensor:
    """ResNetBackbone stages, MT/model/backbone/resnet.py:79-85."""
    sd = ctx.sd
    if stage == 0:
        y = ctx.q(F.conv2d(ctx.q(x), ctx.w(p + 'conv1.weight'), None, 2, 3), p + 'conv1.c')
        return ctx.q(F.relu(_batch_norm(ctx, y, p + 'norm1.')), p + 'stem.out')
    if stage == 1:
        x = F.max_pool2d(x, 3, 2, 1)
    n_blocks = ctx.cfg.layers[stage - 1]
    for b in range(n_blocks):
        x = _nbt1d_fwd(ctx, x, f'{p}layer{stage}.{b}.', 2 if (b == 0 and stage > 1) else 1)
    return x


def _se(ctx: _Ctx, x: Tensor, p: str) -> Tensor:
    """SqueezeAndExcitation.forward, MT/model/utils.py:91-95."""
    sd = ctx.sd
    w = F.adaptive_avg_pool2d(x, 1)
    w = F.relu(F.conv2d(w, sd[p + 'layers.0.weight'], sd[p + 'layers.0.bias']))
    w = torch.sigmoid(F.conv2d(w, sd[p + 'layers.2.weight'], sd[p + 'layers.2.bias']))
    return x * w


def _encoder(ctx: _Ctx, rgb: Optional[Tensor], depth: Optional[Tensor]):
    """FusedRGBDEncoder.forward (MT/model/encoder.py:220-261) / Encoder.forward (:62-143)."""
    cfg = ctx.cfg
    skips: Dict[str, Tensor] = {}
    x = {}
    if 'rgb' in cfg.modalities:
        x['rgb'] = rgb
    if 'depth' in cfg.modalities:
        x['depth'] = depth
    for stage in range(5):
        for m in list(x.keys()):
            x[m] = _backbone_stage(ctx, x[m], cfg.backbone_prefix(m), stage)
        if len(x) == 2:  # se-add-uni-rgb, MT/model/encoder_fusion.py:63-90
            p = f'encoder.fusions.{stage}.'
            fused = _se(ctx, x['rgb'], p + 'weighting_rgb.') + _se(ctx, x['depth'], p + 'weighting_depth.')
            x = {'rgb': ctx.q(fused, p + 'out'), 'depth': x['depth']}
        key = 'rgb' if 'rgb' in x else 'depth'
        ctx.tap(f'encoder.stage{stage}.{key}', x[key])
        if stage in (1, 2, 3):
            skips[str(4 * 2 ** (stage - 1))] = x[key]
    key = 'rgb' if 'rgb' in x else 'depth'
    return x[key], skips


def _conv_bn_relu(ctx: _Ctx, x: Tensor, p: str, k: int, post: Optional[Tensor] = None) -> Tensor:
    """ConvNormAct, MT/model/utils.py:44-69 (optionally followed by `+ post`, the skip-fusion add)."""
    y = ctx.q(F.conv2d(x, ctx.w(p + 'conv.weight'), None, 1, k // 2), p + 'c')
    y = F.relu(_batch_norm(ctx, y, p + 'norm.'))
    if post is not None:
        y = y + post
    return ctx.q(y, p + 'out')


def _ppm(ctx: _Ctx, x: Tensor):
    """PyramidPoolingModule.forward, MT/model/context_module/ppm.py:57-78."""
    h, w = x.shape[2:]
    out = [x]
    feats = []
    for i, b in enumerate(ctx.cfg.ppm_bins):
        y = ctx.q(F.adaptive_avg_pool2d(x, b), f'context_module.features.{i}.pooled')
        y = _conv_bn_relu(ctx, y, f'context_module.features.{i}.1.', 1)
        feats.append(y)
        out.append(ctx.q(F.interpolate(y, (int(h), int(w)), mode='bilinear', align_corners=False),
                         f'context_module.features.{i}.up'))
    y = _conv_bn_relu(ctx, torch.cat(out, 1), 'context_module.final_conv.', 1)
    return ctx.tap('context_module.out', y), tuple(feats)


def _upsample(ctx: _Ctx, x: Tensor, p: str) -> Tensor:
    """Upsampling.forward 'learned-3x3-zeropad', MT/model/upsampling.py:85-96."""
    x = F.interpolate(x, scale_factor=2., mode='nearest')
    return ctx.q(F.conv2d(x, ctx.sd[p + 'conv.weight'], ctx.sd[p + 'conv.bias'], 1, 1, 1, x.shape[1]), p + 'out')


def _decoder_modules(ctx: _Ctx, x: Tensor, skips, p: str):
    """DenseDecoderBase._forward_decoder_modules, MT/model/decoder/dense_base.py:229-259."""
    sides = []
    for i in range(len(ctx.cfg.decoder_n_channels)):
        mp = f'{p}decoder_modules.{i}.'
        x = _conv_bn_relu(ctx, x, mp + 'conv.', 3)
        for b in range(ctx.cfg.decoder_n_blocks):
            x = _nbt1d_fwd(ctx, x, f'{mp}blocks.{b}.', 1)
        sides.append(x if ctx.training else None)  # dense_base.py:92
        x = _upsample(ctx, x, mp + 'upsample.')
        skip = skips[str(16 // 2 ** i)]
        # EncoderDecoderFusion 'add-rgb', MT/model/encoder_decoder_fusion.py:85-87
        x = _conv_bn_relu(ctx, skip, f'{p}fusions.{i}.layer.', 1, post=x)
        ctx.tap(f'{mp}fused', x)
    return x, sides


def _instance_head(ctx: _Ctx, x: Tensor, p: str, k: int, n_up: int):
    """InstanceHead.forward, MT/model/decoder/instance.py:95-121."""
    sd = ctx.sd
    x = _conv_bn_relu(ctx, x, p + 'shared_conv.', 3)
    nt = 3 if ctx.cfg.with_orientation else 2
    outs = [F.conv2d(x[:, 32 * t:32 * (t + 1)], ctx.w(p + f'task_convs.{t}.weight'),
                     sd[p + f'task_convs.{t}.bias'], 1, (k - 1) // 2) for t in range(nt)]
    cat = ctx.q(torch.cat(outs, 1), p + 'task_convs.out')
    for u in range(n_up):
        cat = _upsample(ctx, cat, p + f'upsampling.{u}.')
    outs = list(torch.split(cat, [o.shape[1] for o in outs], 1))
    outs[0] = torch.sigmoid(outs[0])
    outs[1] = torch.tanh(outs[1])
    if nt == 3:  # OrientationOutputNormalization, MT/utils/_orientation.py:50-57, _torch.py:88-91
        o = outs[2]
        outs[2] = o / (torch.sqrt(torch.sum(o * o, dim=1, keepdim=True)) + 1e-7)
    return tuple(outs)


def forward(sd: Dict[str, Tensor], cfg: OracleConfig, rgb: Optional[Tensor], depth: Optional[Tensor],
            training: bool, dropout_masks: Optional[Dict[str, Tensor]] = None,
            track_running_stats: bool = True, taps: Optional[Dict[str, Tensor]] = None,
            teacher: Optional[Dict[str, Tensor]] = None, computed: Optional[Dict[str, Tensor]] = None):
    """EMSANet.forward(batch, do_postprocessing=False), emsanet/model.py:192-233.

    Returns (outputs, new_running_stats).  `outputs` has the reference's nesting
    (SURVEY.md App. A); with enable_panoptic: [((sem, inst), (sem_sides, inst_sides)), (scene, None)];
    otherwise one (output, side_outputs) per decoder in ModuleDict order
    (emsanet/decoder.py:61-201: semantic, instance, scene).
    """
    ctx = _Ctx(sd, cfg, training, dropout_masks, track_running_stats, taps, teacher, computed)
    enc_out, skips = _encoder(ctx, rgb, depth)
    con_out, con_feats = _ppm(ctx, enc_out)
    pre = cfg.decoder_prefixes
    res = {}
    if 'semantic' in pre:  # SemanticDecoder, MT/model/decoder/semantic.py:26-83
        p = pre['semantic']
        x, sides = _decoder_modules(ctx, con_out, skips, p)
        y = ctx.q(F.conv2d(x, ctx.w(p + '_task_head.conv.weight'), sd[p + '_task_head.conv.bias'], 1, 1),
                  p + '_task_head.conv.out')
        for u in range(2):
            y = _upsample(ctx, y, p + f'_task_head.upsample_{u}.')
        s_out = tuple(
            ctx.q(F.conv2d(s, ctx.w(p + f'_side_output_heads.{i}.conv.weight'),
                           sd[p + f'_side_output_heads.{i}.conv.bias']), p + f'_side_output_heads.{i}.conv.out')
            if s is not None else None for i, s in enumerate(sides))
        res['semantic'] = (y, s_out)
    if 'instance' in pre:
        p = pre['instance']
        x, sides = _decoder_modules(ctx, con_out, skips, p)
        y = _instance_head(ctx, x, p + '_task_head.', 3, 2)
        s_out = tuple(_instance_head(ctx, s, p + f'_side_output_heads.{i}.', 1, 0) if s is not None else None
                      for i, s in enumerate(sides))
        res['instance'] = (y, s_out)
    if 'scene' in pre:  # MT/model/decoder/scene.py:32-65
        p = pre['scene']
        x = torch.flatten(con_feats[0], 1)
        res['scene'] = (F.linear(x, sd[p + '_task_head.weight'], sd[p + '_task_head.bias']), None)
    outputs = []
    if cfg.enable_panoptic and 'semantic' in res and 'instance' in res:
        (s, ss), (i, is_) = res['semantic'], res['instance']
        outputs.append(((s, i), (ss, is_)))
    else:
        for t in ('semantic', 'instance'):
            if t in res:
                outputs.append(res[t])
    if 'scene' in res:
        outputs.append(res['scene'])
    return outputs, ctx.new_stats


def flatten_outputs(outputs) -> List[Tensor]:
    """Depth-first list of every tensor in the nested output structure (None skipped)."""
    flat: List[Tensor] = []

    def rec(o):
        if o is None:
            return
        if isinstance(o, (list, tuple)):
            for x in o:
                rec(x)
        else:
            flat.append(o)
    rec(outputs)
    return flat


def bench_loss(outputs) -> Tensor:
    """sum over all output tensors of mean(o^2) — keeps L3 out of the timing."""
    return sum((o.float() ** 2).mean() for o in flatten_outputs(outputs))


def forward_backward(sd: Dict[str, Tensor], cfg: OracleConfig, rgb, depth,
                     dropout_masks=None, grad_outputs: Optional[List[Tensor]] = None, teacher=None, computed=None):
    """Train-mode forward + autograd backward.  Returns (outputs, grads dict, new stats).
    With grad_outputs=None the loss is `bench_loss`; otherwise the given cotangents are used."""
    leaves = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v)
              for k, v in sd.items()}
    outputs, stats = forward(leaves, cfg, rgb, depth, True, dropout_masks, teacher=teacher, computed=computed)
    flat = flatten_outputs(outputs)
    if grad_outputs is None:
        bench_loss(outputs).backward()
    else:
        torch.autograd.backward(flat, grad_outputs)
    grads = {k: v.grad for k, v in leaves.items() if v.is_floating_point() and v.requires_grad}
    return outputs, grads, stats
