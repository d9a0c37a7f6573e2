"""Minimal single-process stand-in for the torchmetrics API the reference uses (MT/metric/*.py, MT/task_helper/scene.py,
main.py:25): Metric (add_state / update / compute / reset / forward, states move with .to()), MeanMetric,
ConfusionMatrix(task='multiclass').  No distributed synchronisation.  See ../README.md."""
from typing import Any, Callable, Dict, Optional

import torch

__version__ = '0.0-standin'


class Metric(torch.nn.Module):
    full_state_update: Optional[bool] = None
    higher_is_better: Optional[bool] = None
    is_differentiable: Optional[bool] = None

    def __init__(self, **kwargs: Any) -> None:
        super().__init__()
        self._defaults: Dict[str, Any] = {}
        self._reductions: Dict[str, Any] = {}
        self._update_count = 0

    def add_state(self, name: str, default, dist_reduce_fx: Optional[Callable] = None, persistent: bool = False) -> None:
        if isinstance(default, torch.Tensor):
            self._defaults[name] = default.detach().clone()
            setattr(self, name, default.detach().clone())
        else:
            self._defaults[name] = list(default)
            setattr(self, name, list(default))
        self._reductions[name] = dist_reduce_fx

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        for name, default in self._defaults.items():
            cur = getattr(self, name)
            if isinstance(cur, torch.Tensor):
                setattr(self, name, fn(cur))
                self._defaults[name] = fn(default)
            else:
                setattr(self, name, [fn(t) for t in cur])
        return out

    def reset(self) -> None:
        self._update_count = 0
        for name, default in self._defaults.items():
            cur = getattr(self, name)
            if isinstance(default, torch.Tensor):
                dev = cur.device if isinstance(cur, torch.Tensor) else default.device
                setattr(self, name, default.detach().clone().to(dev))
            else:
                setattr(self, name, [])

    def update(self, *args, **kwargs) -> None:
        raise NotImplementedError

    def compute(self):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        self.update(*args, **kwargs)
        self._update_count += 1
        return self.compute()


class MeanMetric(Metric):
    def __init__(self, nan_strategy: str = 'warn', **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.add_state('mean_value', torch.tensor(0.0, dtype=torch.float64), dist_reduce_fx='sum')
        self.add_state('weight', torch.tensor(0.0, dtype=torch.float64), dist_reduce_fx='sum')

    def update(self, value, weight=1.0) -> None:
        value = torch.as_tensor(value, dtype=torch.float64, device=self.mean_value.device).detach()
        weight = torch.as_tensor(weight, dtype=torch.float64, device=self.mean_value.device)
        weight = torch.broadcast_to(weight, value.shape)
        keep = ~torch.isnan(value)
        self.mean_value = self.mean_value + (value[keep] * weight[keep]).sum()
        self.weight = self.weight + weight[keep].sum()

    def compute(self) -> torch.Tensor:
        return (self.mean_value / self.weight).to(torch.float32)


class MulticlassConfusionMatrix(Metric):
    def __init__(self, num_classes: int, normalize: Optional[str] = None, ignore_index: Optional[int] = None,
                 **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.num_classes, self.normalize, self.ignore_index = num_classes, normalize, ignore_index
        self.add_state('confmat', torch.zeros(num_classes, num_classes, dtype=torch.long), dist_reduce_fx='sum')

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        if preds.ndim == target.ndim + 1:
            preds = preds.argmax(dim=1)
        preds, target = preds.flatten().long(), target.flatten().long()
        if self.ignore_index is not None:
            keep = target != self.ignore_index
            preds, target = preds[keep], target[keep]
        idx = target * self.num_classes + preds
        cm = torch.bincount(idx, minlength=self.num_classes ** 2).reshape(self.num_classes, self.num_classes)
        self.confmat = self.confmat + cm.to(self.confmat.device, self.confmat.dtype)

    def compute(self) -> torch.Tensor:
        cm = self.confmat
        if self.normalize == 'true':
            cm = cm / cm.sum(1, keepdim=True).clamp(min=1)
        elif self.normalize == 'pred':
            cm = cm / cm.sum(0, keepdim=True).clamp(min=1)
        elif self.normalize == 'all':
            cm = cm / cm.sum().clamp(min=1)
        return cm


def ConfusionMatrix(task: str = 'multiclass', num_classes: Optional[int] = None, **kwargs: Any):
    if task != 'multiclass':
        raise NotImplementedError(f"stand-in torchmetrics: ConfusionMatrix(task='{task}')")
    return MulticlassConfusionMatrix(num_classes=num_classes, **kwargs)
