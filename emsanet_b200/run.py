"""Launcher: run the reference's own scripts UNCHANGED on the emsanet_b200 engine.

    python -m emsanet_b200.run [launcher options] main.py <the script's own arguments ...>
    torchrun --nproc-per-node 8 -m emsanet_b200.run main.py ...            (one process per GPU, data parallel)

`main.py`, `inference_samples.py`, `inference_dataset.py`, `inference_time_whole_model.py` (BASELINE.json north_star:
"so main.py / inference_*.py run unchanged") are executed with `runpy` exactly as `python main.py ...` would run them;
before that the launcher

  1. finds the reference (`--reference-root`, $EMSANET_B200_REFERENCE, /root/reference, <repo>/baseline/_ref — the pip
     install made by scripts/install_reference.sh) and fails loudly if there is none;
  2. appends stand-ins for the three third-party packages missing offline (emsanet_b200/compat/README.md) to sys.path
     and PYTHONPATH (so DataLoader / wandb child processes find them too) — real installations win;
  3. hooks `EMSANet.to(device)` (the call at main.py:82, inference_*.py): once the model sits on a CUDA device,
     `emsanet_b200.patch.patch(model)` swaps its forward for the CUDA engine.  Parameters stay the reference's
     nn.Parameter objects: the optimizer built before `.to()` (main.py:436), checkpoints, `load_weights` keep working;
  4. optional, each one the mirror of a reference component behind the same interface:
       --gpu-postprocessing   decoders' post-processing objects -> emsanet_b200.postprocessing   
       --fused-losses         task helpers' loss objects        -> emsanet_b200.losses           (row 2)
       --fused-optimizer      emsanet.optimizer.get_optimizer   -> emsanet_b200.optim            (row 3)
  5. under torchrun (WORLD_SIZE > 1), makes the script rank-aware without touching it: CUDA device =
     LOCAL_RANK; NCCL process group; parameters and buffers broadcast from rank 0; `RandomSamplerSubset`
     (MT/data/_dataloader.py:71-103) draws ONE permutation (rank 0's seed) and every rank takes its strided share;
     the gradient all-reduce is attached to the engine (emsanet_b200.ddp — gradients come out of `loss.backward()`
     already averaged); wandb is disabled and the results directory suffixed on ranks > 0.

There is no fallback: a model variant the engine does not implement raises NotImplementedError from patch() — unless
`--allow-stock-fallback` is given, in which case that model keeps the reference's own PyTorch forward
and a warning says so.
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
import warnings
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
COMPAT = os.path.join(HERE, 'compat')
SCRIPTS = ('main.py', 'inference_samples.py', 'inference_dataset.py', 'inference_time_whole_model.py')


def find_reference(explicit: Optional[str] = None) -> str:
    """root directory that holds main.py and the emsanet package; raises if none of the known places has it"""
    tried = []
    for root in (explicit, os.environ.get('EMSANET_B200_REFERENCE'), '/root/reference',
                 os.path.join(REPO, 'baseline', '_ref')):
        if not root:
            continue
        tried.append(root)
        if os.path.isfile(os.path.join(root, 'main.py')) and os.path.isdir(os.path.join(root, 'emsanet')):
            return os.path.abspath(root)
    raise FileNotFoundError('emsanet_b200.run: no EMSANet reference found (looked in ' + ', '.join(tried) + '); pass '
                            '--reference-root, set EMSANET_B200_REFERENCE, or run scripts/install_reference.sh')


def reference_paths(root: str) -> List[str]:
    """sys.path entries that make `emsanet`, `nicr_mt_scene_analysis`, `nicr_scene_analysis_datasets` importable"""
    paths = [root]
    for lib in ('nicr-multitask-scene-analysis', 'nicr-scene-analysis-datasets'):
        src = os.path.join(root, 'lib', lib, 'src')
        if os.path.isdir(src):                       # source checkout layout; a pip --target install is flat
            paths.insert(0, src)
    return paths


def setup_paths(root: str) -> None:
    new = [p for p in reference_paths(root) if p not in sys.path]
    sys.path[:0] = new
    if COMPAT not in sys.path:
        sys.path.append(COMPAT)                      # appended: a real torchmetrics / matplotlib / cityscapesscripts wins
    env = os.environ.get('PYTHONPATH', '')
    parts = [p for p in env.split(os.pathsep) if p]
    for p in new + [REPO]:
        if p not in parts:
            parts.insert(0, p)
    if COMPAT not in parts:
        parts.append(COMPAT)
    os.environ['PYTHONPATH'] = os.pathsep.join(parts)


# ------------------------------------------------------------------------------------------------ hooks
def hook_model(opts) -> None:
    """EMSANet.to(cuda) -> patch(model) (+ data-parallel wiring)"""
    import torch
    from emsanet.model import EMSANet
    from . import patch as _patch
    stock_to = EMSANet.to
    if getattr(stock_to, '_eb200_hook', False):
        return

    def to(self, *args, **kwargs):
        out = stock_to(self, *args, **kwargs)
        dev = next(out.parameters()).device
        if dev.type != 'cuda' or hasattr(out, '_eb200_stock_forward'):
            return out
        try:
            _patch.patch(out, postprocessing=opts.gpu_postprocessing)
        except NotImplementedError as e:
            if not opts.allow_stock_fallback:
                raise
            warnings.warn(f'emsanet_b200.run: {e} — this model keeps the reference\'s PyTorch forward')
            return out
        world = int(os.environ.get('WORLD_SIZE', '1'))
        if world > 1:
            import torch.distributed as dist
            from .ddp import GradAllReducer
            with torch.no_grad():                     # every rank starts from rank 0's initialisation
                for t in list(out.parameters()) + list(out.buffers()):
                    dist.broadcast(t.data, src=0)
            eng = _patch._engine_for(out)
            object.__setattr__(out, '_eb200_reducer', GradAllReducer(eng))
        print(f'[emsanet_b200] {type(out).__name__} runs on the sm_100a engine ({dev}, world size {world})', flush=True)
        return out
    to._eb200_hook = True
    EMSANet.to = to


def hook_losses() -> None:
    """task helpers' `initialize(device)` -> also install the fused loss mirrors"""
    from nicr_mt_scene_analysis.task_helper.instance import InstanceTaskHelper
    from nicr_mt_scene_analysis.task_helper.semantic import SemanticTaskHelper
    from . import losses
    for cls in (SemanticTaskHelper, InstanceTaskHelper):
        stock = cls.initialize
        if getattr(stock, '_eb200_hook', False):
            continue

        def initialize(self, device, _stock=stock):
            r = _stock(self, device)
            if str(device).startswith('cuda'):
                try:
                    losses.install(self)
                except NotImplementedError as e:
                    warnings.warn(f'emsanet_b200.run: {e}; keeping the reference loss')
            return r
        initialize._eb200_hook = True
        cls.initialize = initialize


def hook_optimizer() -> None:
    """emsanet.optimizer.get_optimizer(args, parameters) is called BEFORE the model moves to the GPU (main.py:436, :82):
    hand out a stock-looking optimizer object that becomes the fused one at its first step()."""
    import torch
    import emsanet.optimizer as ref_opt
    from . import optim
    stock = ref_opt.get_optimizer
    if getattr(stock, '_eb200_hook', False):
        return

    def get_optimizer(args, parameters):
        params = list(parameters)
        opt = stock(args, params)
        if args.optimizer.lower() not in ('sgd', 'adam', 'adamw'):
            return opt
        stock_step = opt.step
        state = {'fused': None}

        def step(self_, closure=None):
            if state['fused'] is None:
                model = _model_of(params)
                if model is None or next(model.parameters()).device.type != 'cuda':
                    return stock_step(closure)
                fused = optim.get_optimizer(args, model)
                fused.param_groups[0]['lr'] = opt.param_groups[0]['lr']
                state['fused'] = fused
            f = state['fused']
            for k in ('lr', 'momentum', 'weight_decay', 'betas'):     # lr schedulers keep writing into `opt`
                if k in opt.param_groups[0] and k in f.param_groups[0]:
                    f.param_groups[0][k] = opt.param_groups[0][k]
            r = f.step(closure)
            opt.state = f.state                                       # checkpoints written from `opt` carry the state
            return r
        import types
        opt.step = types.MethodType(step, opt)       # a bound method: lr schedulers wrap `optimizer.step.__func__`
        return opt
    get_optimizer._eb200_hook = True
    ref_opt.get_optimizer = get_optimizer
    main_mod = sys.modules.get('__main__')
    if main_mod is not None and getattr(main_mod, 'get_optimizer', None) is stock:
        main_mod.get_optimizer = get_optimizer


_PATCHED_MODELS: List = []


def _model_of(params):
    ids = {id(p) for p in params}
    for m in _PATCHED_MODELS:
        if ids and ids <= {id(p) for p in m.parameters()}:
            return m
    return None


def hook_sampler(rank: int, world: int) -> None:
    """one global permutation, strided shares (MT/data/_dataloader.py:71-115)"""
    import torch
    import torch.distributed as dist
    from nicr_mt_scene_analysis.data import _dataloader as dl
    cls = dl.RandomSamplerSubset
    stock_iter, stock_len = cls.__iter__, cls.__len__
    if getattr(stock_iter, '_eb200_hook', False):
        return

    def __iter__(self):
        # rank 0 draws the seed the stock sampler would draw; all ranks replay the stock iterator with it
        seed = torch.empty((), dtype=torch.int64).random_()
        if dist.is_available() and dist.is_initialized():
            t = seed.to('cuda' if dist.get_backend() == 'nccl' else 'cpu')
            dist.broadcast(t, src=0)
            seed = t.cpu()
        import random
        py_state, torch_state = random.getstate(), torch.random.get_rng_state()
        try:
            torch.manual_seed(int(seed.item()) % (2 ** 63))
            random.seed(int(seed.item()) % (2 ** 32))
            indices = list(stock_iter(self))
        finally:
            random.setstate(py_state)
            torch.random.set_rng_state(torch_state)
        usable = len(indices) - len(indices) % world              # every rank runs the same number of samples
        yield from indices[rank:usable:world]

    def __len__(self):
        return stock_len(self) // world
    __iter__._eb200_hook = True
    cls.__iter__, cls.__len__ = __iter__, __len__


def setup_distributed() -> tuple:
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return 0, 1
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
    if rank > 0:
        os.environ['WANDB_MODE'] = 'disabled'
    return rank, world


def rank_aware_argv(script: str, argv: List[str], rank: int, world: int) -> List[str]:
    """main.py only: wandb off and an own results directory on ranks > 0; `--device cuda` means this rank's GPU"""
    if world <= 1 or os.path.basename(script) != 'main.py':
        return argv
    argv = list(argv)
    if rank > 0:
        if '--wandb-mode' in argv:
            argv[argv.index('--wandb-mode') + 1] = 'disabled'
        else:
            argv += ['--wandb-mode', 'disabled']
        if '--results-basepath' in argv:
            i = argv.index('--results-basepath') + 1
            argv[i] = os.path.join(argv[i], f'rank{rank}')
        else:
            argv += ['--results-basepath', os.path.join('.', 'results', f'rank{rank}')]
    return argv


def parse(argv: List[str]):
    ap = argparse.ArgumentParser(prog='python -m emsanet_b200.run', description=__doc__.split('\n\n')[0],
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--reference-root', default=None)
    ap.add_argument('--gpu-postprocessing', action='store_true')
    ap.add_argument('--fused-losses', action='store_true')
    ap.add_argument('--fused-optimizer', action='store_true')
    ap.add_argument('--allow-stock-fallback', action='store_true')
    ap.add_argument('script', help='main.py | inference_samples.py | inference_dataset.py | inference_time_whole_model.py '
                                   '(or a path to another script that builds an EMSANet)')
    ap.add_argument('script_args', nargs=argparse.REMAINDER)
    return ap.parse_args(argv)


def main(argv: Optional[List[str]] = None) -> None:
    opts = parse(sys.argv[1:] if argv is None else argv)
    root = find_reference(opts.reference_root)
    setup_paths(root)
    script = opts.script if os.path.isfile(opts.script) else os.path.join(root, opts.script)
    if not os.path.isfile(script):
        raise FileNotFoundError(f'emsanet_b200.run: {opts.script} not found (also not under {root})')
    rank, world = setup_distributed()
    import torch
    from emsanet.model import EMSANet
    stock_init = EMSANet.__init__

    def init(self, *a, **k):
        stock_init(self, *a, **k)
        _PATCHED_MODELS.append(self)
    if not getattr(stock_init, '_eb200_hook', False):
        init._eb200_hook = True
        EMSANet.__init__ = init
    hook_model(opts)
    if opts.fused_losses:
        hook_losses()
    if opts.fused_optimizer:
        hook_optimizer()
    if world > 1:
        hook_sampler(rank, world)
    sys.argv = [script] + rank_aware_argv(script, opts.script_args, rank, world)
    try:
        runpy.run_path(script, run_name='__main__')
    finally:
        if world > 1 and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
