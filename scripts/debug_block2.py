import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from emsanet_b200 import ops
from emsanet_b200.engine import Engine, EngineConfig
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
nhwc = lambda x: x.permute(0, 2, 3, 1).contiguous()
nchw = lambda x: x.permute(0, 3, 1, 2).contiguous()
bf = lambda x: x.to(torch.bfloat16).float()
torch.manual_seed(0)
n, c, h, w = 4, 64, 24, 32
p = 'blk.'
sd = {}
for nm, shp in [('conv1_1.weight', (c, c, 3, 1)), ('conv1_2.weight', (c, c, 1, 3)), ('conv2_1.weight', (c, c, 3, 1)), ('conv2_2.weight', (c, c, 1, 3))]:
    sd[p + nm] = bf(torch.randn(shp) * math.sqrt(2.0 / (3 * c)))
for nm in ('conv1_1.bias', 'conv2_1.bias'):
    sd[p + nm] = torch.randn(c) * 0.05
for bn in ('norm1.', 'norm2.'):
    sd[p + bn + 'weight'] = torch.rand(c) + 0.5; sd[p + bn + 'bias'] = torch.randn(c) * 0.1
    sd[p + bn + 'running_mean'] = torch.zeros(c); sd[p + bn + 'running_var'] = torch.ones(c); sd[p + bn + 'num_batches_tracked'] = torch.zeros((), dtype=torch.long)
x = bf(torch.randn(n, c, h, w).clamp_min(0)); dout = bf(torch.randn(n, c, h, w))
use_mask = len(sys.argv) > 1
mask = ((torch.rand(n, c) > 0.2).float() / 0.8) if use_mask else None
# reference with retained intermediates
L = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
xr = x.clone().requires_grad_(True)
T = {}
def rq(t):
    return t + (t.detach().to(torch.bfloat16).float() - t.detach())   # bf16 storage, straight-through
def keep(name, t):
    t = rq(t); t.retain_grad(); T[name] = t; return t
c11 = keep('c11', F.conv2d(xr, L[p+'conv1_1.weight'], L[p+'conv1_1.bias'], 1, (1, 0))); a11 = keep('a11', F.relu(c11))
c12 = keep('c12', F.conv2d(a11, L[p+'conv1_2.weight'], None, 1, (0, 1)))
a12 = keep('a12', F.relu(F.batch_norm(c12, None, None, L[p+'norm1.weight'], L[p+'norm1.bias'], True, 0.1, 1e-5)))
c21 = keep('c21', F.conv2d(a12, L[p+'conv2_1.weight'], L[p+'conv2_1.bias'], 1, (1, 0))); a21 = keep('a21', F.relu(c21))
c22 = keep('c22', F.conv2d(a21, L[p+'conv2_2.weight'], None, 1, (0, 1)))
b2 = F.batch_norm(c22, None, None, L[p+'norm2.weight'], L[p+'norm2.bias'], True, 0.1, 1e-5)
if use_mask: b2 = b2 * mask[:, :, None, None]
out = rq(F.relu(b2 + xr))
out.backward(dout)
# engine with instrumentation: monkeypatch ops to capture intermediates
eng = Engine(EngineConfig(), {k: v.cuda() for k, v in sd.items()})
eng.begin(True, True, {p: mask.cuda()} if use_mask else None); eng.alloc_param_grads()
cap = []
orig_dgrad, orig_bnb = ops.conv2d_dgrad, ops.bn_backward
def dgrad(*a, **k):
    r = orig_dgrad(*a, **k); cap.append(('dgrad', r)); return r
def bnb(*a, **k):
    r = orig_bnb(*a, **k); cap.append(('bn', r[0])); cap.append(('dres', r[1])); return r
ops.conv2d_dgrad, ops.bn_backward = dgrad, bnb
xe = nhwc(x).cuda().to(torch.bfloat16)
o = eng.nbt1d(xe, p, 1)
eng.grads.add(o, nhwc(dout).cuda().to(torch.bfloat16))
eng.run_tape()
print('out', rel(nchw(o).float(), out))
names = ['dc22', 'dz', 'dc21', 'da12', 'dc12', None, 'dc11', 'dx']
refs = {'dc22': T['c22'].grad, 'dc21': T['c21'].grad, 'da12': T['a12'].grad, 'dc12': T['c12'].grad, 'dc11': T['c11'].grad, 'dx': xr.grad}
for (kind, t), nm in zip(cap, names):
    if nm in refs and t is not None:
        print(nm, kind, rel(nchw(t).float(), refs[nm]))
for k in L:
    if L[k].is_floating_point() and L[k].requires_grad and L[k].grad is not None:
        print(k, rel(eng.G[k], L[k].grad))
