import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from oracle import emsanet_oracle as O
from emsanet_b200.engine import Engine, EngineConfig
training = len(sys.argv) > 1 and sys.argv[1] == 'train'
ocfg = O.OracleConfig(emulate_bf16_storage=True)
sd = O.make_state_dict(ocfg, 0)
for k in sd:
    if k.endswith('norm2.weight'): sd[k] = sd[k] * 0.15
rgb, depth = O.make_inputs(4, 96, 128, 1)
taps = {}
with torch.no_grad():
    out, _ = O.forward(sd, ocfg, rgb, depth, training, taps=taps)
eng = Engine(EngineConfig(dropout_p_encoder=0, dropout_p_decoder=0), {k: v.cuda() for k, v in sd.items()})
eng.taps = {}
with torch.no_grad():
    res = eng.forward(rgb.cuda(), depth.cuda(), training)
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
for k in sorted(taps):
    if k.endswith('.out') and k not in eng.taps: continue
for k, v in eng.taps.items():
    if k in taps:
        print(f'{k:70s} {rel(v.permute(0,3,1,2).float(), taps[k]):.4e}')
# instance head internals via oracle recompute
p = 'decoders.panoptic_helper.instance_decoder.'
x = taps[p + 'decoder_modules.2.fused']
hp = p + '_task_head.'
s = F.conv2d(x, sd[hp + 'shared_conv.conv.weight'], None, 1, 1)
if training:
    s = F.batch_norm(s, None, None, sd[hp+'shared_conv.norm.weight'], sd[hp+'shared_conv.norm.bias'], True, 0.1, 1e-5)
else:
    s = F.batch_norm(s, sd[hp+'shared_conv.norm.running_mean'], sd[hp+'shared_conv.norm.running_var'], sd[hp+'shared_conv.norm.weight'], sd[hp+'shared_conv.norm.bias'], False, 0., 1e-5)
s = F.relu(s)
print('shared', rel(eng.taps[hp + 'shared'].permute(0,3,1,2).float(), s))
outs = [F.conv2d(s[:, 32*t:32*t+32], sd[hp+f'task_convs.{t}.weight'], sd[hp+f'task_convs.{t}.bias'], 1, 1) for t in range(3)]
cat = torch.cat(outs, 1)
t8 = eng.taps[hp + 'task8'].permute(0,3,1,2).float().cpu()
print('task8', rel(t8[:, :5], cat), 'pad ch max', t8[:, 5:].abs().max().item())
for c in range(5):
    print('  ch', c, rel(t8[:, c], cat[:, c]))
y = cat
for u in range(2):
    y = F.interpolate(y, scale_factor=2., mode='nearest')
    y = F.conv2d(y, sd[hp+f'upsampling.{u}.conv.weight'], sd[hp+f'upsampling.{u}.conv.bias'], 1, 1, 1, 5)
pa = eng.taps[hp + 'pre_act'].permute(0,3,1,2).float().cpu()
print('pre_act', rel(pa[:, :5], y))
for c in range(5):
    print('  ch', c, rel(pa[:, c], y[:, c]), 'ref absmean', y[:, c].abs().mean().item())
flat = O.flatten_outputs(out)
print('center stats: ref mean', flat[1].mean().item(), 'std', flat[1].std().item())
print('scene ref', flat[-1][0][:5], 'got', res['scene'][0][0][:5].cpu())
