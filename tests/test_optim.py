"""Fused optimizer step + weight re-layout (emsanet/optimizer.py:29-59, main.py:597-599).

gpu: FusedSGD against torch.optim.SGD (momentum, nesterov, weight decay — the reference's configuration) over 10 steps:
     fp32 master parameters and momentum buffers BIT-identical; the bf16 tensor-core layouts the step kernel wrote are
     the ones the engine's own re-layout produces from the updated parameters; FusedAdam / FusedAdamW against torch,
     bit-identical too; state_dict round trips with the stock classes; training through the module API with the re-layout
     launch gone from the step.
not gpu: the mapping of `args` to optimizers, refusal of CPU models."""
import argparse
import copy

import pytest
import torch


def _model(seed=0):
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    args = default_args(input_height=64, input_width=96, rgb_encoder_backbone='resnet18',
                        depth_encoder_backbone='resnet18', dropout_p=0.0, semantic_decoder_block_dropout_p=0.0,
                        instance_decoder_block_dropout_p=0.0, no_zero_init_decoder_residuals=True)
    torch.manual_seed(seed)
    return EMSANetB200(args, simple_dataset_config())


def test_get_optimizer_contract_on_cpu():
    from emsanet_b200 import optim
    m = _model()
    a = argparse.Namespace(optimizer='sgd', learning_rate=0.01, weight_decay=1e-4, momentum=0.9)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        optim.get_optimizer(a, m)
    a.optimizer = 'radam'
    with pytest.raises(NotImplementedError):
        optim.get_optimizer(a, m)
    a.optimizer = 'lamb'
    with pytest.raises(ValueError):
        optim.get_optimizer(a, m)


def _random_grads(models, gen, scale=0.1):
    for ps in zip(*[m.parameters() for m in models]):
        g = torch.randn(ps[0].shape, device='cuda', generator=gen) * scale
        for p in ps:
            p.grad = g.clone()


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', [dict(momentum=0.9, weight_decay=1e-4, nesterov=True),      # the reference's SGD
                                 dict(momentum=0.0, weight_decay=0.0, nesterov=False),
                                 dict(momentum=0.8, weight_decay=0.0, nesterov=False)])
def test_fused_sgd_is_bit_identical_to_torch(cfg):
    from emsanet_b200 import ops, optim
    base = _model().cuda()
    m1, m2 = copy.deepcopy(base), copy.deepcopy(base)
    o1 = optim.FusedSGD(m1, lr=0.03, **cfg)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.03, **cfg)
    gen = torch.Generator(device='cuda').manual_seed(1)
    for step in range(10):
        _random_grads((m1, m2), gen)
        if step == 5:                               # an lr scheduler changes the rate between steps
            o1.param_groups[0]['lr'] = o2.param_groups[0]['lr'] = 0.011
        o1.step()
        o2.step()
    for (k, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        assert torch.equal(p1, p2), k
        if cfg['momentum']:
            assert torch.equal(o1.state[p1]['momentum_buffer'], o2.state[p2]['momentum_buffer']), k
    # the bf16 layouts written by the step kernel == the engine's own re-layout of the updated parameters
    eng = m1._eb200_engine
    checked = 0
    for k in eng._pack_owner_keys:
        if 'task_convs' in k or k.endswith('conv1.weight') and eng.P[k].shape[2] == 7:
            continue
        have = eng._packed[k][1]
        want = ops.pack_weight(eng.P[k].detach(), need_bwd=have.bwd is not None)
        assert torch.equal(have.fwd, want.fwd) and (have.bwd is None or torch.equal(have.bwd, want.bwd)), k
        checked += 1
    assert checked > 100
    # ... and the engine does not re-pack on its next forward (nothing changed behind the optimizer's back)
    vers = eng._pack_versions
    eng.refresh_weights()
    assert eng._pack_versions == vers


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['adam', 'adamw'])
def test_fused_adam_matches_torch(name):
    from emsanet_b200 import optim
    base = _model().cuda()
    m1, m2 = copy.deepcopy(base), copy.deepcopy(base)
    cls1, cls2 = (optim.FusedAdam, torch.optim.Adam) if name == 'adam' else (optim.FusedAdamW, torch.optim.AdamW)
    o1 = cls1(m1, lr=2e-3, weight_decay=1e-2)
    o2 = cls2(m2.parameters(), lr=2e-3, weight_decay=1e-2, betas=(0.9, 0.999))
    gen = torch.Generator(device='cuda').manual_seed(2)
    for _ in range(10):
        _random_grads((m1, m2), gen)
        o1.step()
        o2.step()
    # bit-identical as well (measured: scripts/probe_adam.py) once the scalars 1-beta, lr/bias_correction1 and
    # 1-lr*weight_decay are formed in Python doubles like torch.optim forms them
    for (k, p1), p2 in zip(m1.named_parameters(), m2.parameters()):
        assert torch.equal(p1, p2), (k, float((p1 - p2).abs().max()))
        assert torch.equal(o1.state[p1]['exp_avg'], o2.state[p2]['exp_avg']), k
        assert torch.equal(o1.state[p1]['exp_avg_sq'], o2.state[p2]['exp_avg_sq']), k


@pytest.mark.gpu
def test_state_dict_round_trips_with_torch_sgd():
    from emsanet_b200 import optim
    base = _model().cuda()
    m1, m2, m3 = copy.deepcopy(base), copy.deepcopy(base), copy.deepcopy(base)
    kw = dict(lr=0.02, momentum=0.9, weight_decay=1e-4, nesterov=True)
    o1 = optim.FusedSGD(m1, **kw)
    o2 = torch.optim.SGD(m2.parameters(), **kw)
    gen = torch.Generator(device='cuda').manual_seed(3)
    for _ in range(3):
        _random_grads((m1, m2), gen)
        o1.step()
        o2.step()
    # stock -> fused: continue from a checkpoint written by torch.optim.SGD (main.py:459-466)
    m3.load_state_dict(m2.state_dict())
    o3 = optim.FusedSGD(m3, **kw)
    o3.load_state_dict(o2.state_dict())
    # fused -> stock
    o2b = torch.optim.SGD(m2.parameters(), **kw)
    o2b.load_state_dict(copy.deepcopy(o1.state_dict()))   # (torch's load does not copy same-device tensors: no aliasing)
    for _ in range(3):
        _random_grads((m1, m2, m3), gen)
        o1.step()
        o2b.step()
        o3.step()
    for (k, p1), p2, p3 in zip(m1.named_parameters(), m2.parameters(), m3.parameters()):
        assert torch.equal(p1, p2) and torch.equal(p1, p3), k


@pytest.mark.gpu
def test_training_through_the_module_with_the_fused_step():
    """main.py:597-599 with `optimizer = emsanet_b200.optim.get_optimizer(args, model)`: loss goes down, the forward
    graph recorded after the optimizer exists has no re-layout launch, and an edit behind the optimizer's back
    (load_state_dict) is still picked up."""
    from oracle import emsanet_oracle as O
    from emsanet_b200 import optim
    m = _model().cuda().train()
    a = argparse.Namespace(optimizer='sgd', learning_rate=0.02, weight_decay=1e-4, momentum=0.9)
    opt = optim.get_optimizer(a, m)
    batches = [tuple(t.cuda() for t in O.make_inputs(4, 64, 96, seed=50 + s)) for s in range(3)]
    losses = []
    for s in range(12):
        rgb, depth = batches[s % 3]
        out = m({'rgb': rgb, 'depth': depth})
        loss = sum((o.float() ** 2).mean() for o in O.flatten_outputs(out))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.6 * losses[0], losses
    eng = m._eb200_engine
    entry = next(iter(eng._graph_runner.entries.values()))
    ref_model = _model().cuda().train()
    ref_model({'rgb': batches[0][0], 'depth': batches[0][1]})
    ref_entry = next(iter(ref_model._eb200_engine._graph_runner.entries.values()))
    assert entry.fwd_launches == ref_entry.fwd_launches - 1, (entry.fwd_launches, ref_entry.fwd_launches)
    # weights changed by somebody else: the next forward must see them
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    k0 = 'encoder.backbone_rgb.layer1.0.conv1_1.weight'
    sd[k0] = sd[k0] * 0.0
    m.load_state_dict(sd)
    m({'rgb': batches[0][0], 'depth': batches[0][1]})
    assert float(eng._packed[k0][1].fwd.float().abs().max()) == 0.0
