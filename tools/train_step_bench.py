#!/usr/bin/env python
"""BASELINE config 5: DCNN training step (fused sym5 level-8 features fwd + CNN fwd/bwd + Adam, DDP gradient
all-reduce) at batch 512 per GPU.  Run alone (N=1) or under torchrun; prints one JSON line on rank 0.
The step is DCNN-bound (1.3 GFLOP/frame forward vs 3.6 MFLOP for the transform): a parity / plumbing check of the
"features are consumed in place on the device" claim, not the headline metric."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import audiodeepfake_detection_b200 as afd  # noqa: E402
from audiodeepfake_detection_b200.train_step import TrainStep  # noqa: E402


class Args(dict):
    __getattr__ = dict.get


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    steps, warmup, B = 20, 5, 512
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    args = Args(transform="packets", num_of_scales=256, hop_length=220, log_scale=True, power=2.0, wavelet="sym5",
                loss_less="False", features="none", block_norm=False, mean=[-13.6], std=[4.9])
    tr, norm = afd.get_transforms(args, "none", dev, False)
    step = TrainStep(tr, norm, time_len=95, time_dim_add=1, device=dev, ddp=world > 1)
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(B, 1, 22050, device=dev, generator=g) * 0.1
    y = (torch.arange(B, device=dev) % 2).long()
    # transform alone, for the share of the step it takes
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        tr(x)
    e0.record()
    for _ in range(20):
        tr(x)
    e1.record()
    torch.cuda.synchronize()
    ms_transform = e0.elapsed_time(e1) / 20
    for _ in range(warmup):
        loss = step(x, y)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        loss = step(x, y)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        print(json.dumps({"metric": "dcnn_train_step_frames_per_sec", "value": B * world / (ms * 1e-3), "unit": "frames/s",
                          "n_gpus": world, "ms_per_step": ms, "ms_transform_per_step": ms_transform,
                          "transform_share": ms_transform / ms, "batch_per_gpu": B, "loss": float(loss),
                          "config": "configs[4]: sym5 level-8 features + DCNN fwd/bwd + Adam, DDP" }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
