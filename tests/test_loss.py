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
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz'))


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
