"""Data-parallel gradient exchange check, run under torchrun with >= 2 GPUs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_check.py
Every rank runs forward/backward on its own batch through the nn.Module API with a GradAllReducer attached; afterwards
every `.grad` must be BIT-identical on all ranks (an all-reduce leaves the same bytes everywhere; a bucket that was
missed, reduced too early or reduced twice leaves rank-specific values), differ from the rank's local gradient (the
ranks see different images), and the buckets must tile the flat buffer exactly once.  Modes: graph replay (default),
eager launches, gradient accumulation without zero_grad()."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def flatten(o):
    if o is None:
        return []
    if isinstance(o, (list, tuple)):
        return [t for v in o for t in flatten(v)]
    return [o]


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from emsanet_b200.ddp import GradAllReducer
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    from emsanet_b200.patch import _engine_for
    results = {}
    for mode in ('graph', 'eager', 'accumulate'):
        if mode == 'eager':
            os.environ['EB200_NO_GRAPH'] = '1'
        else:
            os.environ.pop('EB200_NO_GRAPH', None)
        torch.manual_seed(0)
        args = default_args(input_height=128, input_width=160, rgb_encoder_backbone='resnet18',
                            depth_encoder_backbone='resnet18')
        model = EMSANetB200(args, simple_dataset_config()).cuda().train()
        eng = _engine_for(model)
        g = torch.Generator().manual_seed(100 + rank)
        batches = [{'rgb': torch.randn(3, 3, 128, 160, generator=g).cuda(),
                    'depth': torch.randn(3, 1, 128, 160, generator=g).cuda()} for _ in range(3)]
        seen = []
        reducer = GradAllReducer(eng)

        class Spy:                      # records the bucket ranges; the autograd node finds finish() through __self__
            def on_grads_ready(self, flat, lo, hi):
                seen.append((lo, hi))
                return reducer.on_grads_ready(flat, lo, hi)

            def finish(self):
                return reducer.finish()
        eng.on_grads_ready = Spy().on_grads_ready
        for step, batch in enumerate(batches):
            if mode != 'accumulate' or step == 0:
                model.zero_grad(set_to_none=True)
            del seen[:]
            loss = sum((o.float() ** 2).mean() for o in flatten(model(batch)))
            loss.backward()
            torch.cuda.synchronize()
            ranges = sorted(seen)
            total = eng.flat_grad.numel()
            assert ranges[0][0] == 0 and ranges[-1][1] == total and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])), \
                f'{mode}: buckets do not tile the gradient buffer: {ranges} of {total}'
            worst = 0.0
            for k, p in model.named_parameters():
                ref = p.grad.detach().clone()
                dist.broadcast(ref, src=0)
                if not torch.equal(ref, p.grad):
                    worst = max(worst, float((ref - p.grad).abs().max()))
                    raise AssertionError(f'{mode} step {step}: {k} differs between ranks by {worst:.3e}')
        results[mode] = len(ranges)
        # the reduced gradient is the mean over ranks, not the local one: an un-reduced run must differ
        eng.on_grads_ready = None
        model.zero_grad(set_to_none=True)
        loss = sum((o.float() ** 2).mean() for o in flatten(model(batches[-1])))
        loss.backward()
        torch.cuda.synchronize()
        k0, p0 = next((k, p) for k, p in model.named_parameters() if 'layer3' in k and p.dim() == 4)
        local_g = p0.grad.detach().clone()
        ref = local_g.clone()
        dist.broadcast(ref, src=0)
        differs = torch.tensor([0.0 if torch.equal(ref, local_g) else 1.0], device='cuda')
        dist.all_reduce(differs)
        assert float(differs) >= 1.0, f'{mode}: local gradients are identical across ranks — the check is vacuous'
    if rank == 0:
        print('DDP CHECK OK', results, f'world={world}')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
