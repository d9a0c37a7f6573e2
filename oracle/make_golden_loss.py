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


if __name__ == '__main__':
    main()
