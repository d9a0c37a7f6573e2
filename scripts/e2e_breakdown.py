"""Where the e2e step (module API, host inputs, torch loss) spends GPU time beyond the resident step: kernel table of a few
e2e steps from torch.profiler (CUPTI), grouped into engine kernels / torch kernels / copies.  Usage (GPU box):
    python scripts/e2e_breakdown.py [batch] > gpurun_out/e2e_breakdown.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config  # noqa: E402


def flatten(o):
    if o is None:
        return []
    if isinstance(o, (list, tuple)):
        return [t for v in o for t in flatten(v)]
    return [o]


class MeanSquares(torch.autograd.Function):
    """bench.py's stand-in loss sum_i mean(o_i^2): norm reduction forward, one scaled copy per output backward"""
    @staticmethod
    def forward(ctx, *outs):
        ctx.save_for_backward(*outs)
        return sum(torch.linalg.vector_norm(o).square() / o.numel() for o in outs)

    @staticmethod
    def backward(ctx, g):
        return tuple(o * (g * (2.0 / o.numel())) for o in ctx.saved_tensors)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device('cuda', 0)
    model = EMSANetB200(default_args(input_height=480, input_width=640), simple_dataset_config()).to(dev).train()
    rgb_h = torch.randn(n, 3, 480, 640).pin_memory()
    depth_h = torch.randn(n, 1, 480, 640).pin_memory()
    copy_stream = torch.cuda.Stream()
    staged = {}

    def stage():
        with torch.cuda.stream(copy_stream):
            staged['b'] = {'rgb': rgb_h.to(dev, non_blocking=True), 'depth': depth_h.to(dev, non_blocking=True)}
            staged['e'] = torch.cuda.Event()
            staged['e'].record(copy_stream)

    def step():
        if 'b' not in staged:
            stage()
        torch.cuda.current_stream().wait_event(staged['e'])
        batch = staged.pop('b')
        for t in batch.values():
            t.record_stream(torch.cuda.current_stream())
        out = model(batch)
        stage()
        loss = MeanSquares.apply(*flatten(out))
        for p in model.parameters():
            p.grad = None
        loss.backward()
        return float(loss.item())

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f'e2e step without profiler: {e0.elapsed_time(e1) / 10:.3f} ms')
    steps = 3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, 'device_time_total', None)
        if t is None:
            t = getattr(ev, 'cuda_time_total', 0)
        if t > 0:
            rows.append((t / steps / 1e3, ev.count / steps, ev.key))
    rows.sort(reverse=True)
    eng = sum(r[0] for r in rows if 'eb::' in r[2])
    other = [r for r in rows if not ('eb::' in r[2])]
    print(f'engine kernels: {eng:.2f} ms / step;  everything else: {sum(r[0] for r in other):.2f} ms / step')
    for ms, cnt, key in other[:30]:
        print(f'  {ms:8.3f} ms  n={cnt:6.1f}  {key[:110]}')
    print('top engine kernels:')
    for ms, cnt, key in [r for r in rows if r not in other][:12]:
        print(f'  {ms:8.3f} ms  n={cnt:6.1f}  {key[:110]}')


if __name__ == '__main__':
    main()
