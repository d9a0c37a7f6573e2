"""Drop-in boundary: run a reference `EMSANet` nn.Module (emsanet/model.py:26) on the B200 engine.

`patch(model)` swaps the *instance's* `forward(batch, do_postprocessing=False)` (emsanet/model.py:192-233)
for one that executes the hand-written CUDA path.  Everything else stays the reference's: the module tree,
its nn.Parameter / buffer objects (so `state_dict()`, `load_state_dict(strict=True)`, optimizers created
before `.to(device)`, `track_running_stats` toggling at main.py:486-498 keep working), the decoders'
post-processing objects, and the returned container structure.  Outputs are fp32 NCHW and
autograd-connected: `loss.backward()` runs our backward kernels and deposits fp32 `.grad`s on the parameters.

Unsupported model variants (swish, basicblock/bottleneck, SE-backbones, other fusions /
context modules / upsamplings / decoders, the `normal` task, rgbd single-encoder input) raise
NotImplementedError instead of being silently mis-handled; during ONNX export the stock forward is used.
"""
from __future__ import annotations

import types
from collections import ChainMap
from typing import Dict, List, Optional

import torch

from .engine import RESNET_LAYERS, Engine, EngineConfig


def _require(cond: bool, what: str) -> None:
    if not cond:
        raise NotImplementedError(f'emsanet_b200 does not implement this EMSANet variant: {what}')


def config_from_model(model) -> EngineConfig:
    """Derive the engine configuration from `model.args` and the module tree."""
    a = model.args
    mods = tuple(a.input_modalities)
    _require(set(mods) <= {'rgb', 'depth'} and len(mods) >= 1, f'input_modalities={mods}')
    _require(a.activation == 'relu', f'activation={a.activation}')
    _require(a.encoder_normalization in ('batchnorm', 'bn') and a.decoder_normalization in ('batchnorm', 'bn'),
             'normalization != batchnorm')
    backbones = {getattr(a, f'{m}_encoder_backbone') for m in mods}
    blocks = {getattr(a, f'{m}_encoder_backbone_resnet_block') for m in mods}
    _require(len(backbones) == 1 and next(iter(backbones)) in RESNET_LAYERS, f'backbone={backbones}')
    _require(blocks == {'nonbottleneck1d'}, f'encoder block={blocks}')
    if len(mods) == 2:
        _require(a.encoder_fusion == 'se-add-uni-rgb', f'encoder_fusion={a.encoder_fusion}')
    _require(a.context_module == 'ppm' and a.upsampling_context_module == 'bilinear',
             f'context_module={a.context_module}/{a.upsampling_context_module}')
    _require(tuple(a.encoder_decoder_skip_downsamplings) == (4, 8, 16), 'skip downsamplings != (4, 8, 16)')
    _require(a.upsampling_prediction == 'learned-3x3-zeropad', f'upsampling_prediction={a.upsampling_prediction}')
    tasks = tuple(a.tasks)
    _require('normal' not in tasks, 'normal task')
    n_ch, n_blocks, p_dec = None, None, None
    for t in ('semantic', 'instance'):
        if t in tasks:
            _require(getattr(a, f'{t}_decoder') == 'emsanet', f'{t}_decoder')
            _require(getattr(a, f'{t}_decoder_block') == 'nonbottleneck1d', f'{t}_decoder_block')
            _require(getattr(a, f'{t}_decoder_upsampling') == 'learned-3x3-zeropad', f'{t}_decoder_upsampling')
            _require(getattr(a, f'{t}_encoder_decoder_fusion') == 'add-rgb' or
                     (len(mods) == 1 and getattr(a, f'{t}_encoder_decoder_fusion').startswith('add')),
                     f'{t}_encoder_decoder_fusion')
            _require(tuple(getattr(a, f'{t}_decoder_downsamplings')) == (16, 8, 4), f'{t}_decoder_downsamplings')
            ch = tuple(getattr(a, f'{t}_decoder_n_channels'))
            nb = int(getattr(a, f'{t}_decoder_n_blocks'))
            pd = float(getattr(a, f'{t}_decoder_block_dropout_p'))
            _require(n_ch in (None, ch) and n_blocks in (None, nb) and p_dec in (None, pd),
                     'semantic and instance decoders differ in shape')
            n_ch, n_blocks, p_dec = ch, nb, pd
    if 'instance' in tasks:
        _require(a.instance_offset_encoding == 'tanh' and a.instance_center_encoding == 'sigmoid',
                 'instance encodings other than sigmoid/tanh')
    if 'orientation' in tasks:
        _require('instance' in tasks, 'orientation without instance')
    ds = model.dataset_config
    return EngineConfig(
        backbone=next(iter(backbones)), modalities=mods, tasks=tasks,
        enable_panoptic=bool(a.enable_panoptic),
        semantic_n_classes=len(ds.semantic_label_list_without_void),
        scene_n_classes=len(ds.scene_label_list_without_void),
        decoder_n_channels=n_ch or (512, 256, 128), decoder_n_blocks=n_blocks or 3,
        dropout_p_encoder=float(a.dropout_p), dropout_p_decoder=p_dec if p_dec is not None else 0.2)


class _EMSANetFunction(torch.autograd.Function):
    """One autograd node for the whole network: forward and backward are our kernel programs."""

    @staticmethod
    def forward(ctx, engine: Engine, layout: List, rgb, depth, training: bool, track: bool, *params):
        runner = _runner_for(engine)
        ctx.graphed = runner.usable(rgb, depth, training, track)
        if ctx.graphed:
            res = runner.forward(rgb, depth, training, track)
            ctx.generation = runner.generation
        else:
            res = engine.forward(rgb, depth, training, track)
        flat = []
        layout.clear()
        for task, outs in res.items():
            for i, o in enumerate(outs):
                layout.append((task, i))
                flat.append(o)    # fresh tensors either way: the output boundary runs outside the recorded graph
        ctx.engine, ctx.layout = engine, list(layout)
        ctx.set_materialize_grads(False)
        return tuple(flat)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.engine
        by_task: Dict[str, List[Optional[torch.Tensor]]] = {}
        for (task, i), g in zip(ctx.layout, gouts):
            by_task.setdefault(task, []).append(g)
        if ctx.graphed:
            runner = _runner_for(eng)
            if ctx.generation != runner.generation:
                raise RuntimeError('emsanet_b200: backward() of a forward whose activations were overwritten by a later '
                                   'forward of the same model (graph-replayed buffers are reused between steps)')
            runner.backward(by_task)
        else:
            eng.backward(by_task)
        _finish_reduction(eng)
        return (None, None, None, None, None, None, *fresh_grad_views(eng))


def fresh_grad_views(eng: Engine) -> List[torch.Tensor]:
    """New view objects of the flat gradient buffer, one per parameter (engine.grad_keys order).  Nothing else holds
    them, so autograd's AccumulateGrad adopts them as `.grad` instead of cloning 675 tensors (it clones a gradient
    tensor somebody else still references — e.g. the engine's own `G` views)."""
    flat = eng.flat_grad
    return [flat[o:o + n].view(shape) for o, n, shape in eng.grad_slices]


def _finish_reduction(eng: Engine) -> None:
    """Data parallel: the gradients this autograd node returns must already be the reduced ones.  AccumulateGrad may
    copy them (`.grad` exists: accumulation, zero_grad(set_to_none=False)) or add them on the compute stream right
    after this node returns — both must be ordered after the bucket all-reduces, which run on NCCL's stream.  Waiting
    here costs no overlap: the decoder bucket still reduces underneath the encoder's backward kernels."""
    cb = eng.on_grads_ready
    reducer = getattr(cb, '__self__', None)
    if reducer is not None and hasattr(reducer, 'finish'):
        reducer.finish()


def _runner_for(eng: Engine):
    from .graphs import GraphRunner
    r = getattr(eng, '_graph_runner', None)
    if r is None:
        r = GraphRunner(eng)
        eng._graph_runner = r
    return r


def _engine_for(model) -> Engine:
    eng = getattr(model, '_eb200_engine', None)
    dev = next(model.parameters()).device
    if eng is None or eng.dev != dev:
        if dev.type != 'cuda':
            raise RuntimeError('emsanet_b200 runs on CUDA (sm_100a) only; move the model with .to("cuda") — '
                               'there is no CPU fallback')
        params = dict(model.named_parameters())
        params.update(dict(model.named_buffers()))
        eng = Engine(config_from_model(model), params)
        object.__setattr__(model, '_eb200_engine', eng)
        object.__setattr__(model, '_eb200_bn_modules', None)
    return eng


def run_model(model, rgb: Optional[torch.Tensor], depth: Optional[torch.Tensor]) -> Dict[str, List[torch.Tensor]]:
    """Run the network; returns {'semantic': [main, side...], 'instance': [c, o(, r), sides...], 'scene': [y]}."""
    eng = _engine_for(model)
    training = model.training
    # main.py toggles `track_running_stats` on every module that has it (main.py:486-498); the reference's BatchNorm
    # modules are collected once per model — walking all ~1000 modules costs 0.3 ms per forward, on the critical path
    bns = getattr(model, '_eb200_bn_modules', None)
    if bns is None:
        bns = [m for m in model.modules() if hasattr(m, 'track_running_stats')]
        object.__setattr__(model, '_eb200_bn_modules', bns)
    track = all(m.track_running_stats for m in bns) if bns else bool(getattr(model, 'track_running_stats', True))
    layout: List = []
    if training and torch.is_grad_enabled():
        params = [eng.P[k] for k in eng.grad_keys]
        flat = _EMSANetFunction.apply(eng, layout, rgb, depth, True, track, *params)
    else:
        with torch.no_grad():
            runner = _runner_for(eng)
            if not training and runner.usable(rgb, depth, False, track):
                return runner.forward(rgb, depth, False, track)
            res = eng.forward(rgb, depth, training, track)
            eng.tape, eng.grads = [], None
        return res
    res: Dict[str, List[torch.Tensor]] = {}
    for (task, _), t in zip(layout, flat):
        res.setdefault(task, []).append(t)
    return res


def assemble_outputs(model, res: Dict[str, List[torch.Tensor]], batch, do_postprocessing: bool):
    """Re-nest the flat outputs exactly as the reference decoders return them."""
    cfg: EngineConfig = model._eb200_engine.cfg
    training = model.training
    n_side = len(cfg.decoder_n_channels)
    nt = 3 if cfg.with_orientation else 2

    def semantic():
        o = res['semantic']
        return o[0], tuple(o[1:1 + n_side]) if training else tuple([None] * n_side)

    def instance():
        o = res['instance']
        main = tuple(o[:nt])
        if not training:
            return main, tuple([None] * n_side)
        return main, tuple(tuple(o[nt + i * nt: nt + (i + 1) * nt]) for i in range(n_side))

    outputs = []
    for name, decoder in model.decoders.items():
        if name == 'panoptic_helper':
            (s, ss), (i, is_) = semantic(), instance()
            out = ((s, i), (ss, is_))
        elif name == 'semantic_decoder':
            out = semantic()
        elif name == 'instance_decoder':
            out = instance()
        elif name == 'scene_decoder':
            out = (res['scene'][0], None)
        else:
            raise NotImplementedError(f'decoder {name}')
        if do_postprocessing:
            out = decoder.postprocessing.postprocess(out, batch, is_training=training)
        outputs.append(out)
    if do_postprocessing:
        outputs = dict(ChainMap(*outputs))
    return outputs


def _patched_forward(self, batch, do_postprocessing=False):
    if torch.onnx.is_in_onnx_export():   # custom kernels cannot be traced
        return self._eb200_stock_forward(batch, do_postprocessing=do_postprocessing)
    mods = self.args.input_modalities
    rgb = batch['rgb'] if 'rgb' in mods else None
    depth = batch['depth'] if 'depth' in mods else None
    res = run_model(self, rgb, depth)
    return assemble_outputs(self, res, batch, do_postprocessing)


def patch(model, postprocessing: bool = False):
    """Make `model(batch, do_postprocessing)` run on the emsanet_b200 kernels.  Returns the same object.
    `postprocessing=True` also swaps the decoders' post-processing objects for the GPU mirrors
    (emsanet_b200/postprocessing.py); by default the reference's own post-processing keeps running."""
    config_from_model(model)    # raises NotImplementedError for unsupported variants
    if postprocessing:
        from . import postprocessing as _pp
        _pp.install(model)
    if not hasattr(model, '_eb200_stock_forward'):
        object.__setattr__(model, '_eb200_stock_forward', model.forward)
        object.__setattr__(model, 'forward', types.MethodType(_patched_forward, model))
    return model


def unpatch(model):
    if hasattr(model, '_eb200_stock_forward'):
        object.__setattr__(model, 'forward', model._eb200_stock_forward)
        object.__delattr__(model, '_eb200_stock_forward')
    return model
