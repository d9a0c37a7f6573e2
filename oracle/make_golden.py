"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this
container.  TEST INFRASTRUCTURE ONLY — never imported by the product.

The reference is pure Python; three third-party imports are absent offline and are stubbed
exactly as SURVEY.md §8(c) describes (cityscapesscripts labels).  The model is built from a
hand-made argparse.Namespace so `emsanet.args` (which needs torchmetrics)
is not required.  The oracle's seeded state_dict is loaded with strict=True — that alone pins
the key/shape/dtype inventory of `oracle.emsanet_oracle.param_shapes` against the reference.

Run:  python oracle/make_golden.py          (writes tests/golden/, prints max deviations)
"""
import argparse
import collections
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def install_reference_shim():
    sys.path[:0] = [f'{REF}/lib/nicr-multitask-scene-analysis/src',
                    f'{REF}/lib/nicr-scene-analysis-datasets/src', REF]
    Label = collections.namedtuple(
        'Label', 'name id trainId category categoryId hasInstances ignoreInEval color')
    labels = [Label(f'c{i}', i, i, 'cat', 0, False, False, (i, i, i)) for i in range(40)]
    m0 = types.ModuleType('cityscapesscripts')
    m1 = types.ModuleType('cityscapesscripts.helpers')
    m2 = types.ModuleType('cityscapesscripts.helpers.labels')
    m2.labels, m2.Label, m1.labels, m0.helpers = labels, Label, m2, m1
    sys.modules.update({'cityscapesscripts': m0, 'cityscapesscripts.helpers': m1,
                        'cityscapesscripts.helpers.labels': m2})


def make_args(cfg, h, w, dropout=0.0):
    a = dict(input_modalities=cfg.modalities, input_height=h, input_width=w, tasks=cfg.tasks,
             enable_panoptic=cfg.enable_panoptic, activation='relu', encoder_normalization='batchnorm',
             decoder_normalization='batchnorm', no_pretrained_backbone=True, dropout_p=dropout,
             encoder_fusion='se-add-uni-rgb' if len(cfg.modalities) == 2 else 'none',
             encoder_decoder_skip_downsamplings=(4, 8, 16), context_module='ppm',
             upsampling_context_module='bilinear', upsampling_prediction='learned-3x3-zeropad',
             instance_offset_encoding='tanh', instance_center_encoding='sigmoid',
             instance_offset_distance_threshold=None, instance_center_heatmap_threshold=0.1,
             instance_center_heatmap_nms_kernel_size=17, instance_center_heatmap_apply_foreground_mask=False,
             instance_center_heatmap_top_k=64, he_init=('encoder-fusion',), encoder_decoder_fusion=None,
             no_zero_init_decoder_residuals=False, debug=False)
    for m in ('rgb', 'depth', 'rgbd'):
        a[f'{m}_encoder_backbone'] = cfg.backbone
        a[f'{m}_encoder_backbone_resnet_block'] = 'nonbottleneck1d'
        a[f'{m}_encoder_backbone_pretrained_weights_filepath'] = None
    for d in ('semantic', 'instance', 'normal'):
        a[f'{d}_decoder'] = 'emsanet'
        a[f'{d}_decoder_n_channels'] = cfg.decoder_n_channels
        a[f'{d}_decoder_downsamplings'] = (16, 8, 4)
        a[f'{d}_decoder_block'] = 'nonbottleneck1d'
        a[f'{d}_decoder_block_dropout_p'] = dropout
        a[f'{d}_decoder_n_blocks'] = cfg.decoder_n_blocks
        a[f'{d}_decoder_dropout_p'] = 0.1
        a[f'{d}_decoder_upsampling'] = 'learned-3x3-zeropad'
        a[f'{d}_encoder_decoder_fusion'] = 'add-rgb'
    return argparse.Namespace(**a)


class _Labels(list):
    classes_is_thing = tuple([True] * 40)
    classes_use_orientations = tuple([True] * 40)


def make_dataset_config(cfg):
    # only the attributes EMSANet.__init__ reads (emsanet/model.py:39-43)
    sem = _Labels(range(cfg.semantic_n_classes))
    sem.classes_is_thing = tuple([True] * cfg.semantic_n_classes)
    sem.classes_use_orientations = tuple([True] * cfg.semantic_n_classes)
    return types.SimpleNamespace(semantic_label_list_without_void=sem,
                                 scene_label_list_without_void=list(range(cfg.scene_n_classes)))


def build_reference(cfg, h, w, sd):
    from emsanet.model import EMSANet
    model = EMSANet(make_args(cfg, h, w), make_dataset_config(cfg))
    model.load_state_dict(sd, strict=True)   # pins key/shape inventory (emsanet/weights.py:162)
    assert list(model.state_dict().keys()) == list(sd.keys()), 'state_dict order differs'
    return model


def sample(t, n=64):
    """deterministic strided sample of a tensor (keeps fixtures small)."""
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy().copy()


def digest(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


CASES = {
    # name: (cfg kwargs, N, H, W)
    'full_rgbd_r34': (dict(), 2, 96, 128),
    'rgb_semantic_r34': (dict(modalities=('rgb',), tasks=('semantic',), enable_panoptic=False), 2, 64, 96),
    'full_rgbd_r18_ragged': (dict(backbone='resnet18'), 3, 96, 160),
}


def main():
    from oracle import emsanet_oracle as O
    install_reference_shim()
    os.makedirs(f'{ROOT}/tests/golden', exist_ok=True)
    for name, (kw, n, h, w) in CASES.items():
        cfg = O.OracleConfig(**kw)
        sd = O.make_state_dict(cfg, seed=0)
        rgb, depth = O.make_inputs(n, h, w, seed=1)
        if 'rgb' not in cfg.modalities:
            rgb = None
        if 'depth' not in cfg.modalities:
            depth = None
        ref = build_reference(cfg, h, w, sd)
        batch = {k: v for k, v in (('rgb', rgb), ('depth', depth)) if v is not None}
        fix = {}
        # ---- eval forward
        ref.eval()
        with torch.no_grad():
            r_eval = O.flatten_outputs(ref(batch))
            o_eval = O.flatten_outputs(O.forward(sd, cfg, rgb, depth, False)[0])
        assert len(r_eval) == len(o_eval)
        dev = max(float((a - b).abs().max()) for a, b in zip(r_eval, o_eval))
        print(f'{name}: eval   n_out={len(r_eval):2d} max|ref-oracle|={dev:.3e}')
        for i, t in enumerate(r_eval):
            fix[f'eval_out{i}_sample'] = sample(t)
            fix[f'eval_out{i}_shape'] = np.array(t.shape)
            fix[f'eval_out{i}_sum'] = np.array(t.double().sum().item())
        # ---- train forward + backward (dropout p=0 on both sides)
        ref.train()
        for p in ref.parameters():
            p.grad = None
        r_out = ref(batch)
        r_flat = O.flatten_outputs(r_out)
        O.bench_loss(r_out).backward()
        o_out, o_grads, o_stats = O.forward_backward(sd, cfg, rgb, depth)
        o_flat = O.flatten_outputs(o_out)
        dev = max(float((a - b).abs().max()) for a, b in zip(r_flat, o_flat))
        r_grads = {k: p.grad for k, p in ref.named_parameters()}
        gdev = max(float((r_grads[k] - o_grads[k]).abs().max() / (r_grads[k].abs().max() + 1e-12)) for k in r_grads)
        r_sd = ref.state_dict()
        sdev = max(float((r_sd[k].double() - v.double()).abs().max()) for k, v in o_stats.items())
        print(f'{name}: train  n_out={len(r_flat):2d} max|ref-oracle|={dev:.3e}  grads rel={gdev:.3e}  running-stats={sdev:.3e}')
        for i, t in enumerate(r_flat):
            fix[f'train_out{i}_sample'] = sample(t)
            fix[f'train_out{i}_shape'] = np.array(t.shape)
            fix[f'train_out{i}_sum'] = np.array(t.double().sum().item())
        gkeys = sorted(r_grads.keys())
        fix['grad_keys'] = np.array(gkeys)
        fix['grad_l2'] = np.array([r_grads[k].double().norm().item() for k in gkeys])
        fix['grad_sample'] = np.stack([np.resize(sample(r_grads[k], 8), 8) for k in gkeys])
        skeys = sorted(k for k in o_stats if 'num_batches' not in k)
        fix['stat_keys'] = np.array(skeys)
        fix['stat_sample'] = np.stack([np.resize(sample(r_sd[k], 8), 8) for k in skeys])
        fix['sd_digest'] = np.array(hashlib.sha256(''.join(digest(v) for v in sd.values()).encode()).hexdigest())
        fix['n_state_entries'] = np.array(len(sd))
        fix['meta'] = np.array(repr((kw, n, h, w)))
        np.savez_compressed(f'{ROOT}/tests/golden/{name}.npz', **fix)
        print(f'  wrote tests/golden/{name}.npz  ({len(sd)} state entries)')


if __name__ == '__main__':
    main()
