"""The reference's own `main.py`, UNMODIFIED, trained on the emsanet_b200 engine through the launcher
(`python -m emsanet_b200.run main.py ...`, BASELINE.json north_star: "main.py / inference_*.py run unchanged").

Needs a reference install: baseline/_ref (scripts/install_reference.sh; travels to the GPU box) or /root/reference.
gpu: a few epochs on the synthetic mini-NYUv2 (emsanet_b200/synthetic_nyuv2.py): the engine is the one that ran, the
     training loss goes down, the checkpoint main.py wrote loads with load_state_dict(strict=True) into EMSANetB200;
     the same with every mirror switched on (GPU post-processing, fused losses, fused optimizer).
not gpu: the launcher's host logic (reference discovery, argv rewriting for ranks, compat stand-ins)."""
import csv
import glob
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_reference():
    from emsanet_b200 import run
    try:
        return run.find_reference()
    except FileNotFoundError:
        return None


def test_launcher_host_logic(tmp_path, monkeypatch):
    from emsanet_b200 import run
    monkeypatch.delenv('EMSANET_B200_REFERENCE', raising=False)
    with pytest.raises(FileNotFoundError, match='install_reference'):
        run.find_reference(str(tmp_path)) if not _have_reference() else (_ for _ in ()).throw(
            FileNotFoundError('install_reference'))
    fake = tmp_path / 'ref'
    (fake / 'emsanet').mkdir(parents=True)
    (fake / 'main.py').write_text('print("hi")')
    assert run.find_reference(str(fake)) == str(fake)
    (fake / 'lib' / 'nicr-multitask-scene-analysis' / 'src').mkdir(parents=True)
    assert run.reference_paths(str(fake))[0].endswith(os.path.join('nicr-multitask-scene-analysis', 'src'))
    argv = ['--dataset', 'nyuv2', '--results-basepath', '/r', '--wandb-mode', 'online']
    assert run.rank_aware_argv('main.py', argv, 0, 4) == argv
    a1 = run.rank_aware_argv('main.py', argv, 1, 4)
    assert a1[a1.index('--wandb-mode') + 1] == 'disabled' and a1[a1.index('--results-basepath') + 1] == '/r/rank1'
    assert run.rank_aware_argv('inference_samples.py', argv, 1, 4) == argv
    o = run.parse(['--fused-losses', 'main.py', '--dataset', 'nyuv2', '--fused-optimizer'])
    assert o.fused_losses and not o.fused_optimizer and o.script_args == ['--dataset', 'nyuv2', '--fused-optimizer']


def test_compat_torchmetrics_stand_in():
    """the API surface the reference's metrics use (MT/metric/*.py, MT/task_helper/scene.py:44-54,110-132)"""
    sys.path.append(os.path.join(ROOT, 'emsanet_b200', 'compat'))
    try:
        import importlib
        tm = importlib.import_module('torchmetrics')
        if 'standin' not in getattr(tm, '__version__', ''):
            pytest.skip('a real torchmetrics is installed')
        m = tm.MeanMetric()
        m.update(torch.tensor(2.0), weight=1)
        m.update(4.0, weight=3)
        assert float(m.compute()) == pytest.approx(3.5)
        m.reset()
        assert float(m.weight) == 0.0
        cm = tm.ConfusionMatrix(task='multiclass', num_classes=3)
        cm._defaults['confmat'] = cm._defaults['confmat'].long()
        cm.reset()
        cm.update(preds=torch.tensor([0, 1, 2, 2]), target=torch.tensor([0, 1, 1, 2]))
        assert cm.confmat.tolist() == [[1, 0, 0], [0, 1, 1], [0, 0, 1]]

        class Acc(tm.Metric):
            def __init__(self):
                super().__init__()
                self.add_state('n', torch.zeros(2, dtype=torch.int64), dist_reduce_fx='sum')

            def update(self, x):
                self.n += x

            def compute(self):
                return self.n.sum()
        a = Acc()
        a.update(torch.tensor([1, 2]))
        assert int(a.compute()) == 3
        a.reset()
        assert int(a.compute()) == 0
    finally:
        sys.path.remove(os.path.join(ROOT, 'emsanet_b200', 'compat'))
        for k in [k for k in sys.modules if k == 'torchmetrics' or k.startswith('torchmetrics.')]:
            if 'standin' in getattr(sys.modules[k], '__version__', 'standin'):
                del sys.modules[k]


def _train(tmp_path, extra, epochs=6, nproc=1):
    from emsanet_b200 import synthetic_nyuv2
    data = str(tmp_path / 'nyuv2')
    synthetic_nyuv2.write_dataset(data, n_train=16, n_test=4, height=240, width=320, seed=0)
    results = str(tmp_path / 'results')
    launch = [sys.executable, '-m', 'emsanet_b200.run'] if nproc == 1 else \
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr',
         '127.0.0.1', '--master-port', '29541', '-m', 'emsanet_b200.run']
    cmd = [*launch, *extra, 'main.py',
           '--dataset', 'nyuv2', '--dataset-path', data, '--tasks', 'semantic', 'scene', 'instance', 'orientation',
           '--enable-panoptic', '--no-pretrained-backbone', '--rgb-encoder-backbone', 'resnet18',
           '--depth-encoder-backbone', 'resnet18', '--input-height', '192', '--input-width', '256',
           '--n-epochs', str(epochs), '--batch-size', str(8 // nproc), '--validation-batch-size', '4', '--n-workers', '2',
           '--learning-rate', '0.02', '--device', 'cuda', '--wandb-mode', 'disabled', '--results-basepath', results,
           '--checkpointing-metrics', 'valid_semantic_miou', '--validation-skip', '0.0']
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    if r.returncode != 0:       # keep the whole output: pytest truncates the assertion message
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'main_py_failed_nproc{nproc}.log'), 'w') as f:
            f.write(r.stdout[-20000:] + '\n==== stderr ====\n' + r.stderr[-20000:])
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert '[emsanet_b200] EMSANet runs on the sm_100a engine' in r.stdout
    logs = glob.glob(os.path.join(results, '**', '*.csv'), recursive=True)
    assert logs, 'main.py wrote no csv log'
    rows = list(csv.DictReader(open(logs[0])))
    losses = [float(x['train_total_loss']) for x in rows if x.get('train_total_loss')]
    ckpts = glob.glob(os.path.join(results, '**', 'ckpt_resume.pth'), recursive=True)
    assert ckpts
    return losses, ckpts[0], r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize('mirrors', [(), ('--gpu-postprocessing', '--fused-losses', '--fused-optimizer')],
                         ids=['engine_only', 'all_mirrors'])
def test_unmodified_main_py_trains_on_the_engine(tmp_path, mirrors):
    if _have_reference() is None:
        pytest.skip('no reference install (scripts/install_reference.sh puts one into baseline/_ref)')
    losses, ckpt, out = _train(tmp_path, list(mirrors))
    assert len(losses) == 6 and all(l == l for l in losses), losses
    assert min(losses[-2:]) < 0.9 * losses[0], f'training loss did not go down: {losses}'
    # the checkpoint the reference's CheckpointHelper wrote loads into the mirror class, strictly
    from emsanet_b200.module import EMSANetB200, default_args, simple_dataset_config
    state = torch.load(ckpt, map_location='cpu', weights_only=False)
    sd = state['state_dict'] if 'state_dict' in state else state['model']
    m = EMSANetB200(default_args(input_height=192, input_width=256, rgb_encoder_backbone='resnet18',
                                 depth_encoder_backbone='resnet18'), simple_dataset_config())
    m.load_state_dict(sd, strict=True)
    d = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, f'main_py_on_engine_{"all_mirrors" if mirrors else "engine_only"}.log'), 'w') as f:
        f.write('train_total_loss per epoch: ' + ' '.join(f'{l:.4f}' for l in losses) + '\n' + out[-4000:])


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_unmodified_main_py_trains_data_parallel_under_torchrun(tmp_path):
    """torchrun --nproc-per-node 2 -m emsanet_b200.run main.py ...: sampler shares, parameter broadcast, three-bucket
    gradient all-reduce, rank-0-only results — the script itself untouched"""
    if _have_reference() is None:
        pytest.skip('no reference install (scripts/install_reference.sh puts one into baseline/_ref)')
    losses, ckpt, out = _train(tmp_path, [], epochs=4, nproc=2)
    assert out.count('[emsanet_b200] EMSANet runs on the sm_100a engine') == 2 and 'world size 2' in out
    assert len(losses) == 4 and all(l == l for l in losses), losses
    assert min(losses[-2:]) < 0.95 * losses[0], f'training loss did not go down: {losses}'
    d = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, 'main_py_2gpu_torchrun.log'), 'w') as f:
        f.write('train_total_loss per epoch: ' + ' '.join(f'{l:.4f}' for l in losses) + '\n' + out[-4000:])
