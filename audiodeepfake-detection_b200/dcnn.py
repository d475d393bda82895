"""The consumer of the feature front-end: the reference's DCNN classifier, kept in plain PyTorch.

BASELINE.json north_star: "The DCNN itself stays in PyTorch and consumes the features in place".  This
module exists so the shipped checkpoints (reference models/*.pt) can be loaded and used as known-answer
fixtures for the transform, and so the training-step harness has a model to drive.  Layer layout mirrors
reference models.py:240-317 (module indices inside ``cnn`` / ``dil_conv`` / ``fc`` must line up with the
checkpoint keys); hyper-parameters come from reference scripts/gridsearch_config.py:128-134 and
scripts/start_exps.sh (flattend_size=320, time_dim_add 1 for sym5 / 0 otherwise).

Two stem variants are provided:
  * ``pooled``  -- the class as shipped (MaxPool2d after conv 0, 2, 5).  Matches the sym5 and stft checkpoints.
  * ``strided`` -- the coif4 checkpoint has no module slots for the pools (keys cnn.0..cnn.16, BN at
    2,5,8,11,14).  Reconstructed (NOT in the reference source): pools replaced by stride 2 on the preceding
    conv, paddings (2,0,1,1,1,0).  See SURVEY.md section 8c.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import torch
from torch import nn


@dataclass
class DCNNConfig:
    in_channels: int = 1
    time_len: int = 95            # input_dim[-1] of the reference (T of the feature tensor)
    time_dim_add: int = 0
    channels: tuple = (64, 64, 96, 128, 32)
    kernel1: int = 3
    flattend_size: int = 320
    dropout_cnn: float = 0.6
    dropout_lstm: float = 0.2
    stem: str = "pooled"          # "pooled" | "strided"
    sync_bn: bool = False


def _bn(ch: int, affine: bool, sync: bool) -> nn.Module:
    cls = nn.SyncBatchNorm if sync else nn.BatchNorm2d
    return cls(ch, affine=affine)


class DCNN(nn.Module):
    """Deep CNN with dilated convolutions over a time-as-channel view (reference models.py:240-317)."""

    def __init__(self, cfg: DCNNConfig):
        super().__init__()
        self.cfg = cfg
        c1, c2, c3, c4, c5 = cfg.channels
        pooled = cfg.stem == "pooled"
        # (in, out, kernel, padding, followed-by-downsample, followed-by-BN)
        plan = [
            (cfg.in_channels, c1, cfg.kernel1, 2, True, True),
            (c1, c2, 1, 0, False, True),
            (c2, c3, 3, 1, True, True),
            (c3, c4, 3, 1, False, True),
            (c4, c5, 3, 1, False, True),
            (c5, 64, 3, 1 if pooled else 0, True, False),
        ]
        layers: list[nn.Module] = []
        for cin, cout, k, pad, down, bn in plan:
            stride = 2 if (down and not pooled) else 1
            layers += [nn.Conv2d(cin, cout, k, stride=stride, padding=pad), nn.PReLU()]
            if down and pooled:
                layers.append(nn.MaxPool2d(2, 2))
            if bn:
                layers.append(_bn(cout, False, cfg.sync_bn))
        layers.append(nn.Dropout(cfg.dropout_cnn))
        self.cnn = nn.Sequential(*layers)

        td = cfg.time_len // 8 + cfg.time_dim_add
        self.dil_conv = nn.Sequential(
            _bn(td, True, cfg.sync_bn), nn.Conv2d(td, td, 3, 1, padding=1, dilation=1), nn.PReLU(),
            _bn(td, True, cfg.sync_bn), nn.Conv2d(td, td, 5, 1, padding=2, dilation=2), nn.PReLU(),
            _bn(td, True, cfg.sync_bn), nn.Conv2d(td, td, 7, 1, padding=2, dilation=4), nn.PReLU(),
            nn.Dropout(cfg.dropout_lstm),
        )
        self.fc = nn.Sequential(nn.Flatten(2), nn.Linear(cfg.flattend_size, 2))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # x: logical [B, C, P, T] (a view of [B, C, T, P] memory) -> contiguous [B, C, T, P] for free
        x = self.cnn(x.permute(0, 1, 3, 2))
        x = x.permute(0, 2, 1, 3).contiguous()      # [B, time, channels, packets]
        x = self.dil_conv(x)
        return self.fc(x).mean(1)


def strip_ddp_prefix(state: dict) -> dict:
    """The shipped snapshots were saved through two nested DDP wrappers: keys start with 'module.module.'."""
    return {re.sub(r"^(module\.)+", "", k): v for k, v in state.items()}


def load_reference_checkpoint(path: str, time_len: int, time_dim_add: int, map_location="cpu") -> DCNN:
    """Load one of reference models/*.pt ({'MODEL_STATE', 'EPOCHS_RUN'}) into a matching DCNN (eval mode)."""
    snap = torch.load(path, map_location=map_location, weights_only=True)
    state = strip_ddp_prefix(snap["MODEL_STATE"] if "MODEL_STATE" in snap else snap)
    stem = "pooled" if "cnn.18.weight" in state else "strided"
    model = DCNN(DCNNConfig(time_len=time_len, time_dim_add=time_dim_add, stem=stem))
    model.load_state_dict(state, strict=True)
    return model.eval()
