"""The CPU oracle against the golden fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py, run in the build container where /root/reference is mounted).
fp32, same ATen ops in the same order => tolerance 1e-5 relative (observed: bit-exact)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import emsanet_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz'))
RTOL = 1e-5


def _sample(t, n=64):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].numpy()


def _load(name):
    fix = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    kw, n, h, w = ast.literal_eval(str(fix['meta']))
    cfg = O.OracleConfig(**kw)
    sd = O.make_state_dict(cfg, seed=0)
    rgb, depth = O.make_inputs(n, h, w, seed=1)
    if 'rgb' not in cfg.modalities:
        rgb = None
    if 'depth' not in cfg.modalities:
        depth = None
    return fix, cfg, sd, rgb, depth


@pytest.mark.parametrize('name', CASES)
def test_state_dict_inventory(name):
    fix, cfg, sd, _, _ = _load(name)
    assert len(sd) == int(fix['n_state_entries'])  # 1056 for the full model
    assert all(v.dtype in (torch.float32, torch.long) for v in sd.values())


@pytest.mark.parametrize('name', CASES)
def test_eval_forward_matches_reference(name):
    fix, cfg, sd, rgb, depth = _load(name)
    with torch.no_grad():
        outs = O.flatten_outputs(O.forward(sd, cfg, rgb, depth, False)[0])
    for i, t in enumerate(outs):
        assert tuple(fix[f'eval_out{i}_shape']) == tuple(t.shape)
        np.testing.assert_allclose(_sample(t), fix[f'eval_out{i}_sample'], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(t.double().sum().item(), float(fix[f'eval_out{i}_sum']), rtol=1e-6, atol=1e-3)
    assert f'eval_out{len(outs)}_shape' not in fix


@pytest.mark.parametrize('name', CASES)
def test_train_forward_backward_matches_reference(name):
    fix, cfg, sd, rgb, depth = _load(name)
    outs, grads, stats = O.forward_backward(sd, cfg, rgb, depth)
    flat = O.flatten_outputs(outs)
    for i, t in enumerate(flat):
        assert tuple(fix[f'train_out{i}_shape']) == tuple(t.shape)
        np.testing.assert_allclose(_sample(t), fix[f'train_out{i}_sample'], rtol=RTOL, atol=1e-6)
    gkeys = [str(k) for k in fix['grad_keys']]
    assert sorted(grads.keys()) == gkeys
    l2 = np.array([grads[k].double().norm().item() for k in gkeys])
    np.testing.assert_allclose(l2, fix['grad_l2'], rtol=1e-4, atol=1e-9)
    skeys = [str(k) for k in fix['stat_keys']]
    got = np.stack([np.resize(_sample(stats[k], 8), 8) for k in skeys])
    np.testing.assert_allclose(got, fix['stat_sample'], rtol=RTOL, atol=1e-6)


def test_dropout_sites_cover_all_blocks():
    cfg = O.OracleConfig()
    sites = O.dropout_sites(cfg)
    assert len(sites) == 2 * 16 + 2 * 9  # 50 Dropout2d per forward
    sd = O.make_state_dict(cfg)
    for p, c, _ in sites:
        assert sd[p + 'norm2.weight'].shape == (c,)
