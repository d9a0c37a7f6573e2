import copy, sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_optim import _model, _random_grads
from emsanet_b200 import optim
for name in ('adam', 'adamw'):
    base = _model().cuda()
    m1, m2 = copy.deepcopy(base), copy.deepcopy(base)
    cls1, cls2 = (optim.FusedAdam, torch.optim.Adam) if name == 'adam' else (optim.FusedAdamW, torch.optim.AdamW)
    o1 = cls1(m1, lr=2e-3, weight_decay=1e-2); o2 = cls2(m2.parameters(), lr=2e-3, weight_decay=1e-2, betas=(0.9, 0.999))
    gen = torch.Generator(device='cuda').manual_seed(2)
    for s in range(10):
        _random_grads((m1, m2), gen); o1.step(); o2.step()
        d = max(float((p1 - p2).abs().max()) for p1, p2 in zip(m1.parameters(), m2.parameters()))
        r = max(float(((p1 - p2).abs() / (p2.abs() + 1e-3)).max()) for p1, p2 in zip(m1.parameters(), m2.parameters()))
        dm = max(float((o1.state[p1]['exp_avg'] - o2.state[p2]['exp_avg']).abs().max()) for p1, p2 in zip(m1.parameters(), m2.parameters()))
        dv = max(float(((o1.state[p1]['exp_avg_sq'] - o2.state[p2]['exp_avg_sq']).abs() / (o2.state[p2]['exp_avg_sq'] + 1e-12)).max()) for p1, p2 in zip(m1.parameters(), m2.parameters()))
        print(name, s, 'param abs %.3e rel %.3e  exp_avg abs %.3e  exp_avg_sq rel %.3e' % (d, r, dm, dv))
