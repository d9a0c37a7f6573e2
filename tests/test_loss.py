"""Fused semantic cross-entropy.

not-gpu: the oracle (oracle/loss_oracle.py) against the fixtures made from the UNMODIFIED reference class
         (oracle/make_golden_loss.py): loss within 2e-6 relative, gradient within 2e-6, element count identical;
         the host side (autograd.Function, mirror class, install) with the two C-ABI calls replaced by the oracle.
gpu:     the kernels against the oracle (first B200 run: profiles/r2_unverified_kernels_first_run.log).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as L

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'loss')
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and not f.startswith('instance_'))
INSTANCE_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and f.startswith('instance_'))


def _load(name):
    fix = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    meta = json.loads(bytes(fix['meta']).decode())
    kw = dict(meta['inputs'])
    if 'dtype' in kw:
        kw['dtype'] = getattr(torch, kw['dtype'].split('.')[-1])
    logits, target, weights = L.make_inputs(**kw)
    return fix, meta, logits, target, (weights if meta['weighted'] else None)


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_fixture(name):
    fix, meta, logits, target, weights = _load(name)
    loss, n, grad = L.cross_entropy_semantic(logits, target, weights, meta['eps'])
    assert n == int(fix['n_elements'])
    assert abs(loss - float(fix['loss'])) <= 2e-6 * max(1.0, abs(loss))
    assert np.abs(grad[:, ::3, ::2, ::3] * meta['upstream'] - fix['grad_sample']).max() <= 2e-6


def _emulate(monkeypatch):
    from emsanet_b200 import losses

    def fwd(logits, target, weights, eps):
        loss, n, _ = L.cross_entropy_semantic(logits, target, weights, eps)
        return torch.tensor([loss], dtype=torch.float64), torch.tensor([n], dtype=torch.int64)

    def bwd(logits, target, weights, eps, grad_out):
        _, _, g = L.cross_entropy_semantic(logits, target, weights, eps)
        return (torch.from_numpy(g) * grad_out.double()).float()
    monkeypatch.setattr(losses, 'ce_forward', fwd)
    monkeypatch.setattr(losses, 'ce_backward', bwd)
    monkeypatch.setattr(losses, '_check', lambda t, what: t.contiguous())
    return losses


@pytest.mark.parametrize('name', CASES)
def test_host_side_with_emulated_abi(name, monkeypatch):
    losses = _emulate(monkeypatch)
    fix, meta, logits, target, weights = _load(name)
    mod = losses.CrossEntropyLossSemanticB200(weights=weights, label_smoothing=meta['eps'])
    x = logits.clone().requires_grad_(True)
    side = logits[:, :, ::2, ::2].clone().requires_grad_(True)
    out = mod([x, side], [target, target[:, ::2, ::2]])
    assert len(out) == 2
    (loss, n), (loss2, n2) = out
    assert loss.dtype == torch.float32 and loss.ndim == 0 and isinstance(n, int) and n == int(fix['n_elements'])
    assert abs(loss.item() - float(fix['loss'])) <= 2e-6 * max(1.0, abs(loss.item()))
    (loss * meta['upstream'] + loss2 * 0.0).backward()
    assert np.abs(x.grad.numpy()[:, ::3, ::2, ::3] - fix['grad_sample']).max() <= 2e-6
    assert side.grad is not None and float(side.grad.abs().max()) == 0.0


def test_no_cpu_fallback_and_unsupported_variant():
    from emsanet_b200 import _lib, losses
    with pytest.raises(_lib.EB200Error, match='no CPU path'):
        losses.CrossEntropyLossSemanticB200()([torch.zeros(1, 3, 4, 4)], [torch.zeros(1, 4, 4, dtype=torch.uint8)])
    with pytest.raises(NotImplementedError):
        losses.CrossEntropyLossSemanticB200(weights=torch.ones(3), weighted_reduction=True)


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference checkout only exists in the build container')
def test_install_on_the_real_reference_task_helper():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), 'oracle'))
    import make_golden as mg
    mg.install_reference_shim()
    from nicr_mt_scene_analysis.loss.ce import CrossEntropyLossSemantic
    from emsanet_b200 import losses
    import types
    helper = types.SimpleNamespace(_loss=CrossEntropyLossSemantic(weights=torch.ones(40) * 2, label_smoothing=0.1))
    losses.install(helper)
    assert isinstance(helper._loss, losses.CrossEntropyLossSemanticB200)
    assert helper._loss._label_smoothing == pytest.approx(0.1) and float(helper._loss._weights[0]) == 2.0


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_kernels_match_oracle(name):
    from emsanet_b200 import losses
    fix, meta, logits, target, weights = _load(name)
    mod = losses.CrossEntropyLossSemanticB200(weights=None if weights is None else weights.cuda(),
                                              label_smoothing=meta['eps'])
    x = logits.cuda().requires_grad_(True)
    (loss, n), = mod([x], [target.cuda()])
    (loss * meta['upstream']).backward()
    o_loss, o_n, o_grad = L.cross_entropy_semantic(logits, target, weights, meta['eps'])
    assert n == o_n == int(fix['n_elements'])
    assert abs(loss.item() - o_loss) <= 2e-6 * max(1.0, abs(o_loss))
    assert np.abs(x.grad.cpu().double().numpy() - o_grad * meta['upstream']).max() <= 2e-6
    assert np.abs(x.grad.cpu().numpy()[:, ::3, ::2, ::3] - fix['grad_sample']).max() <= 2e-6


# ------------------------------------------------------------------------------------------------ regression losses
def _load_instance(name):
    fix = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    meta = json.loads(bytes(fix['meta']).decode())
    return fix, meta, L.make_instance_inputs(**meta['inputs'])


# tag -> (kind, prediction key, target key, mask key, channel_dim)
PARTS = {'center_mse': (0, 'center', 't_center', 'center_mask', None), 'center_l1': (1, 'center', 't_center', 'center_mask', None),
         'offset': (1, 'offset', 't_offset', 'fg', 1), 'orientation': (2, 'orientation', 't_orientation', 'ofg', 1)}


def _part(d, tag):
    kind, pk, tk, mk, cd = PARTS[tag]
    pred = d[pk][:, 0] if pk == 'center' else d[pk]
    return kind, pred, d[tk], d[mk], cd


@pytest.mark.parametrize('name', INSTANCE_CASES)
def test_regression_oracle_matches_reference_fixture(name):
    fix, meta, d = _load_instance(name)
    for tag in PARTS:
        kind, pred, target, mask, cd = _part(d, tag)
        loss, n, grad = L.masked_loss(kind, pred, target, mask, cd)
        assert n == int(fix[tag + '_n']), tag
        assert abs(loss - float(fix[tag + '_loss'])) <= 2e-6 * max(1.0, abs(loss)), tag
        g = (grad * meta['upstream']).reshape(grad.shape[0], -1)[:, ::3]
        assert np.abs(g - fix[tag + '_grad']).max() <= 2e-6, tag


class _StubHelper:
    """the attributes / methods of InstanceTaskHelper that `instance_losses` uses (MT/task_helper/base.py:95-160)"""
    _disable_multiscale_supervision = False
    _loss_name_instance_center = 'mse'

    def __init__(self, scales):
        self.scales = scales
        import types
        self._loss_orientation = types.SimpleNamespace(_kappa=1.0)

    def collect_predictions_for_loss(self, predictions_post, predictions_post_key, side_outputs_key):
        preds = [predictions_post[predictions_post_key]] + list(predictions_post[side_outputs_key])
        return preds, ['main'] + [f'down_{s}' for s in self.scales[1:]], self.scales[1:]

    def collect_targets_for_loss(self, batch, batch_key, downscales):
        return [batch[batch_key]] + [batch[f'{batch_key}@{s}'] for s in downscales]

    def mark_as_total(self, key):
        return 'total_' + key


def _instance_batch(device):
    ds = [L.make_instance_inputs(2, 32, 48, seed=41), L.make_instance_inputs(2, 16, 24, seed=42, fg_fraction=0.1),
          L.make_instance_inputs(2, 8, 12, seed=43, fg_fraction=0.0)]
    scales = [1, 2, 4]
    keymap = {'instance_center_mask': 'center_mask', 'instance_center': 't_center', 'instance_foreground': 'fg',
              'instance_offset': 't_offset', 'orientation_foreground': 'ofg', 'orientation': 't_orientation'}
    batch = {}
    for s, d in zip(scales, ds):
        for bk, dk in keymap.items():
            batch[bk if s == 1 else f'{bk}@{s}'] = d[dk].to(device)
    preds = [tuple(d[k].to(device).requires_grad_(True) for k in ('center', 'offset', 'orientation')) for d in ds]
    post = {'instance_output': preds[0], 'instance_side_outputs': preds[1:]}
    return ds, scales, batch, preds, post


def _check_instance_losses(losses, device):
    ds, scales, batch, preds, post = _instance_batch(device)
    helper = _StubHelper(scales)
    out = losses.instance_losses(helper, batch, 0, post)
    keys = ['main', 'down_2', 'down_4']
    want, tot = {}, {'center': [0.0, 0], 'offset': [0.0, 0], 'orientation': [0.0, 0]}
    grads = []
    for key, d in zip(keys, ds):
        g = {}
        for name, tag in (('center', 'center_mse'), ('offset', 'offset'), ('orientation', 'orientation')):
            kind, pred, target, mask, cd = _part(d, tag)
            loss, n, grad = L.masked_loss(kind, pred, target, mask, cd)
            n_div = max(n, 1) if name == 'orientation' else n
            want[f'instance_{name}_loss_{key}'] = loss / n_div if n_div else float('nan')
            tot[name][0] += loss
            tot[name][1] += n_div
            g[name] = grad
        grads.append(g)
    for name, (l, n) in tot.items():
        want['total_instance_' + name] = l / n if n else l
    assert set(out) == set(want)
    for k, v in want.items():
        got = float(out[k])
        if np.isnan(v) or np.isinf(v):
            assert not np.isfinite(got), k           # 0 valid pixels: the reference divides by zero as well
        else:
            assert abs(got - v) <= 3e-6 * max(1.0, abs(v)), (k, got, v)
    # gradient of the weighted sum of the three totals, like the training loop forms it (MT/task_helper/base.py)
    total = out['total_instance_center'] * 2.0 + out['total_instance_offset'] * 0.5 + out['total_instance_orientation']
    total.backward()
    for (pc, po, pr), g in zip(preds, grads):
        assert np.abs(pc.grad.cpu().double().numpy()[:, 0] - g['center'] * 2.0 / tot['center'][1]).max() <= 1e-7
        assert np.abs(po.grad.cpu().double().numpy() - g['offset'] * 0.5 / tot['offset'][1]).max() <= 1e-7
        assert np.abs(pr.grad.cpu().double().numpy() - g['orientation'] / tot['orientation'][1]).max() <= 1e-7


def test_instance_losses_host_side_with_emulated_abi(monkeypatch):
    from emsanet_b200 import losses

    def fwd(kind, pred, target, mask, channel_dim, kappa=1.0):
        loss, n, _ = L.masked_loss(kind, pred, target, mask, channel_dim, kappa)
        return torch.tensor([loss], dtype=torch.float64), torch.tensor([n], dtype=torch.int64)

    def bwd(kind, pred, target, mask, channel_dim, kappa, grad_out):
        _, _, g = L.masked_loss(kind, pred, target, mask, channel_dim, kappa)
        return (torch.from_numpy(g) * grad_out.double()).float().reshape(pred.shape)
    monkeypatch.setattr(losses, 'masked_loss_forward', fwd)
    monkeypatch.setattr(losses, 'masked_loss_backward', bwd)
    monkeypatch.setattr(losses, '_check', lambda t, what: t.contiguous())
    _check_instance_losses(losses, 'cpu')
    # the LossBase-level mirrors: same (loss, n_elements) as the reference classes gave for the fixtures
    fix, meta, d = _load_instance('instance_main_scale')
    (loss, n), = losses.MSELossB200()([d['center'][:, 0] * d['center_mask']], [d['t_center']])
    assert n == d['t_center'].numel() and abs(float(loss) - float(fix['center_mse_loss'])) <= 2e-6 * float(loss)
    (loss, n), = losses.L1LossB200()([d['offset'] * d['fg'][:, None]], [d['t_offset']])
    assert n == d['fg'].numel() and abs(float(loss) - float(fix['offset_loss'])) <= 2e-6 * float(loss)
    m = d['ofg'].flatten()
    pr = d['orientation'].permute(0, 2, 3, 1).reshape(-1, 2)[m]
    tg = d['t_orientation'].permute(0, 2, 3, 1).reshape(-1, 2)[m]
    (loss, n), = losses.VonMisesLossBiternionB200()([pr], [tg])
    assert n == int(m.sum()) and abs(float(loss) - float(fix['orientation_loss'])) <= 2e-6 * float(loss)
    with pytest.raises(ValueError):
        losses.VonMisesLossBiternionB200()([d['orientation']], [d['t_orientation']])


@pytest.mark.gpu
@pytest.mark.parametrize('name', INSTANCE_CASES)
def test_regression_kernels_match_oracle(name):
    from emsanet_b200 import losses
    fix, meta, d = _load_instance(name)
    for tag in PARTS:
        kind, pred, target, mask, cd = _part(d, tag)
        x = pred.cuda().contiguous().requires_grad_(True)
        loss, count = losses.fused_masked_loss(kind, x, target.cuda(), mask.cuda(), cd)
        (loss * meta['upstream']).backward()
        o_loss, o_n, o_grad = L.masked_loss(kind, pred, target, mask, cd)
        assert int(count) == o_n == int(fix[tag + '_n']), tag
        assert abs(float(loss) - o_loss) <= 3e-6 * max(1.0, abs(o_loss)), tag
        assert np.abs(x.grad.cpu().double().numpy() - o_grad * meta['upstream']).max() <= 2e-6, tag
        g = x.grad.cpu().numpy().reshape(x.shape[0], -1)[:, ::3]
        assert np.abs(g - fix[tag + '_grad']).max() <= 2e-6, tag


@pytest.mark.gpu
def test_instance_losses_without_host_sync_match_oracle():
    from emsanet_b200 import losses
    _check_instance_losses(losses, 'cuda')


@pytest.mark.gpu
def test_out_of_range_target_is_reported():
    """ADVICE r1 (loss.cu:48): a label above the class range must not be silently treated as void"""
    from emsanet_b200 import _lib, losses
    logits, target, _ = L.make_inputs(1, 5, 8, 8, seed=3)
    target[0, 0, 0] = 9
    with pytest.raises(_lib.EB200Error, match='above the number of classes'):
        losses.CrossEntropyLossSemanticB200()([logits.cuda()], [target.cuda()])
