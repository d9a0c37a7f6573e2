"""Which arithmetic (fma vs mul+add in the three `a + alpha*b` steps) makes FusedSGD bit-identical to torch.optim.SGD?
Run once on the GPU box; the matching combination is the kernel's default (flags = 0)."""
import copy
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
from emsanet_b200 import optim

args = default_args(input_height=64, input_width=96, rgb_encoder_backbone='resnet18', depth_encoder_backbone='resnet18')
torch.manual_seed(0)
base = EMSANetB200(args, simple_dataset_config()).cuda()
for flags in range(8):
    m1, m2 = copy.deepcopy(base), copy.deepcopy(base)
    o1 = optim.FusedSGD(m1, lr=0.03, momentum=0.9, weight_decay=1e-4, nesterov=True)
    o1.flags = flags
    o2 = torch.optim.SGD(m2.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4, nesterov=True)
    g = torch.Generator(device='cuda').manual_seed(1)
    for step in range(4):
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p1.shape, device='cuda', generator=g) * 0.1
            p1.grad, p2.grad = gr.clone(), gr.clone()
        o1.step(); o2.step()
    bad = sum(int(not torch.equal(p1, p2)) for p1, p2 in zip(m1.parameters(), m2.parameters()))
    worst = max(float((p1 - p2).abs().max()) for p1, p2 in zip(m1.parameters(), m2.parameters()))
    print(f'flags={flags}: {bad} of {len(list(m1.parameters()))} tensors differ, max |diff| {worst:.3e}')
