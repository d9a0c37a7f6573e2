"""Pin oracle/loss_oracle.py against the UNMODIFIED reference loss class (/root/reference, build container only) and
write tests/golden/loss/*.npz.  TEST INFRASTRUCTURE ONLY.      python oracle/make_golden_loss.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_oracle as L                     # noqa: E402
from oracle.make_golden import install_reference_shim   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'loss')
# name -> (make_inputs kwargs, use class weights, label smoothing)
CASES = {
    'main_scale_weighted': (dict(n=2, c=40, h=48, w=64, seed=21), True, 0.0),
    'side_scale_smoothing': (dict(n=3, c=40, h=15, w=20, seed=22, void_fraction=0.5), True, 0.1),
    'unweighted_int64_targets': (dict(n=1, c=13, h=30, w=41, seed=23, dtype=torch.int64), False, 0.0),
    'all_void': (dict(n=1, c=5, h=8, w=8, seed=24, void_fraction=1.1), True, 0.05),
}


def main():
    install_reference_shim()
    from nicr_mt_scene_analysis.loss.ce import CrossEntropyLossSemantic
    os.makedirs(OUT, exist_ok=True)
    for name, (kw, weighted, eps) in CASES.items():
        logits, target, weights = L.make_inputs(**kw)
        ref = CrossEntropyLossSemantic(weights=weights if weighted else None, label_smoothing=eps)
        x = logits.clone().requires_grad_(True)
        (loss, n_el), = ref([x], [target])
        (loss * 0.37).backward()                        # an upstream gradient, as the task helper's normalisation gives
        o_loss, o_n, o_grad = L.cross_entropy_semantic(logits, target, weights if weighted else None, eps)
        assert o_n == n_el, (name, o_n, n_el)
        rel = abs(o_loss - float(loss)) / max(1.0, abs(o_loss))
        gerr = float(np.abs(o_grad * 0.37 - x.grad.double().numpy()).max())
        assert rel <= 2e-6 and gerr <= 2e-6, (name, rel, gerr)
        meta = {'inputs': {k: (str(v) if k == 'dtype' else v) for k, v in kw.items()}, 'weighted': weighted, 'eps': eps,
                'upstream': 0.37}
        np.savez_compressed(os.path.join(OUT, name + '.npz'), loss=np.float32(float(loss)), n_elements=np.int64(n_el),
                            grad_sample=x.grad.numpy()[:, ::3, ::2, ::3].copy(),
                            meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        print(f'{name}: loss {float(loss):.4f} n {n_el}  oracle rel dev {rel:.2e}  grad max dev {gerr:.2e}')


INSTANCE_CASES = {
    'instance_main_scale': dict(n=2, h=48, w=64, seed=31),
    'instance_side_scale_sparse': dict(n=3, h=15, w=20, seed=32, fg_fraction=0.05),
    'instance_no_foreground': dict(n=1, h=8, w=10, seed=33, fg_fraction=0.0),
}


def main_instance():
    """the three regression losses exactly as MT/task_helper/instance.py:118-207 composes them, with the UNMODIFIED
    reference loss classes; the oracle's loss, element count and gradient against them; fixtures"""
    from nicr_mt_scene_analysis.loss.l1 import L1Loss
    from nicr_mt_scene_analysis.loss.mse import MSELoss
    from nicr_mt_scene_analysis.loss.vonmises import VonMisesLossBiternion
    for name, kw in INSTANCE_CASES.items():
        d = L.make_instance_inputs(**kw)
        out = {}
        for center_cls, tag in ((MSELoss, 'center_mse'), (L1Loss, 'center_l1')):
            x = d['center'].clone().requires_grad_(True)
            (loss, _), = center_cls(reduction='sum')([x[:, 0] * d['center_mask']], [d['t_center']])   # instance.py:131-137
            n = d['center_mask'].sum().item()                                                          # :138-139
            (loss * 0.41).backward()
            o_loss, o_n, o_grad = L.masked_loss(0 if tag.endswith('mse') else 1, d['center'][:, 0], d['t_center'],
                                                d['center_mask'], None)
            out[tag] = (float(loss), n, x.grad[:, 0].numpy(), o_loss, o_n, o_grad * 0.41)
        x = d['offset'].clone().requires_grad_(True)
        mask = d['fg'].unsqueeze(1).expand_as(x)                                                      # :155-160
        (loss, _), = L1Loss(reduction='sum')([x * mask], [d['t_offset']])
        (loss * 0.41).backward()
        o = L.masked_loss(1, d['offset'], d['t_offset'], d['fg'], 1)
        out['offset'] = (float(loss), d['fg'].sum().item(), x.grad.numpy(), o[0], o[1], o[2] * 0.41)
        x = d['orientation'].clone().requires_grad_(True)
        pred = x.contiguous().permute((0, 2, 3, 1)).reshape(-1, 2)                                     # :186-192
        tgt = d['t_orientation'].permute((0, 2, 3, 1)).reshape(-1, 2)
        m = d['ofg'].flatten()
        (loss, _), = VonMisesLossBiternion()([pred[m, :]], [tgt[m, :]])
        (loss * 0.41).backward()
        o = L.masked_loss(2, d['orientation'], d['t_orientation'], d['ofg'], 1, 1.0)
        g = x.grad.numpy() if x.grad is not None else np.zeros(tuple(x.shape), np.float32)
        out['orientation'] = (float(loss), m.sum().item(), g, o[0], o[1], o[2] * 0.41)
        arrays = {}
        for tag, (loss, n, grad, o_loss, o_n, o_grad) in out.items():
            assert o_n == n, (name, tag, o_n, n)
            rel = abs(o_loss - loss) / max(1.0, abs(o_loss))
            gerr = float(np.abs(o_grad - grad.astype(np.float64)).max())
            assert rel <= 2e-6 and gerr <= 2e-6, (name, tag, rel, gerr)
            arrays[tag + '_loss'], arrays[tag + '_n'] = np.float32(loss), np.int64(n)
            arrays[tag + '_grad'] = grad.reshape(grad.shape[0], -1)[:, ::3].copy()
            print(f'{name}/{tag}: loss {loss:.5f} n {n}  oracle rel dev {rel:.2e}  grad max dev {gerr:.2e}')
        meta = {'inputs': kw, 'upstream': 0.41}
        np.savez_compressed(os.path.join(OUT, name + '.npz'), meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                            **arrays)


if __name__ == '__main__':
    main()
    main_instance()
