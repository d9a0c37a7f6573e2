"""emsanet_b200 — B200-native (sm_100a) forward/backward engine for EMSANet.

The package holds only the hot path: `csrc/` (hand-written CUDA kernels + the C ABI declared in
`include/emsanet_b200.h`) and the host-side mirror of the reference's nn.Module interface
(`patch()` swaps `EMSANet.forward`; state_dict, optimizer and scripts of the reference stay as they are).
There is no CPU or PyTorch fallback: importing the ops without the built library raises.
"""
__version__ = '0.1.0'
