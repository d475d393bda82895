"""Minimal training-step harness for BASELINE config 5: features forward (fused kernel, no_grad) -> Normalize ->
DCNN forward / backward -> optimizer step, optionally under DDP (one process per GPU, NCCL gradient all-reduce).

Re-creates ``Trainer._run_batch`` of the reference (train_classifier.py:945-995) without its bookkeeping: the
transform runs under ``torch.no_grad()`` (:965-967), the model consumes the features in place on the device, the
loss is cross entropy, the optimizer Adam (train_classifier.py:1212-1219).  The reference wraps the model in DDP
twice (:280 -> :340, :1027 -> :340); here it is wrapped once.
"""
from __future__ import annotations

import torch
from torch import nn

from .dcnn import DCNN, DCNNConfig


class TrainStep:
    def __init__(self, transforms: nn.Module, normalize: nn.Module, time_len: int, channels: int = 1,
                 time_dim_add: int = 0, device="cuda", lr: float = 4e-4, weight_decay: float = 1e-3, ddp: bool = False):
        self.transforms, self.normalize = transforms, normalize
        self.device = torch.device(device)
        cfg = DCNNConfig(in_channels=channels, time_len=time_len, time_dim_add=time_dim_add, sync_bn=ddp)
        model = DCNN(cfg).to(self.device)
        if ddp:
            from torch.nn.parallel import DistributedDataParallel as DDP
            model = DDP(model, device_ids=[self.device.index])
        self.model = model
        self.loss_fun = nn.CrossEntropyLoss()
        self.optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)

    def __call__(self, audio: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """One optimisation step on a batch of frames ``[B, 1, N]`` (host or device); returns the detached loss."""
        audio = audio.to(self.device, non_blocking=True)
        labels = labels.to(self.device, non_blocking=True)
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.no_grad():
            feats, _ = self.transforms(audio)
            feats = self.normalize(feats)
        out = self.model(feats)
        loss = self.loss_fun(out, labels)
        loss.backward()
        self.optimizer.step()
        return loss.detach()
