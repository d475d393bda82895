"""N>1 host logic on CPU: shard bookkeeping and the fingerprint all-reduce over a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiodeepfake_detection_b200.sharding import shard_bounds


@pytest.mark.parametrize("n,world", [(4096, 1), (4096, 8), (1_000_000, 8), (10, 4), (3, 8), (0, 2)])
def test_shard_bounds_partition(n, world):
    spans = [shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(n, world, world)


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from audiodeepfake_detection_b200.fingerprint import FingerprintAccumulator
        from audiodeepfake_detection_b200.sharding import gather_features, local_shard
        from oracle import wpt_oracle

        level = 6
        clips = (np.random.default_rng(11).standard_normal((5, 1, 2048)) * 0.1).astype(np.float32)   # same on all ranks
        mine = local_shard(torch.from_numpy(clips))                     # 3 + 2 clips
        assert mine.shape[0] == (3 if rank == 0 else 2)
        # the kernel's contribution is emulated by the oracle here (no GPU): what is under test is the exchange
        acc = FingerprintAccumulator(level, "cpu")
        sums, count = wpt_oracle.haar_fingerprint_sums(mine.numpy(), level)
        acc.sums += torch.from_numpy(sums)
        acc.count += count
        acc.all_reduce()
        want_sums, want_count = wpt_oracle.haar_fingerprint_sums(clips, level)
        assert int(acc.count.item()) == want_count
        assert np.allclose(acc.mean().numpy(), want_sums / want_count, rtol=1e-12)
        # ragged gather of per-rank feature batches restores the original order
        feats = torch.arange(5 * 6, dtype=torch.float32).reshape(5, 1, 2, 3)
        back = gather_features(local_shard(feats).clone())
        assert torch.equal(back, feats)
        # normalisation statistics of a sharded training set: per-rank moments + counts meet in one all-reduce
        from audiodeepfake_detection_b200.wavelet_math import merge_moments_across_ranks
        vals = torch.arange(1, 11, dtype=torch.float64)                # the "features" of the whole set
        part = vals[:6] if rank == 0 else vals[6:]
        mom = torch.stack([part.sum(), (part * part).sum()]).reshape(1, 2)
        merged, n = merge_moments_across_ranks(mom, part.numel())
        assert n == 10 and torch.allclose(merged, torch.tensor([[vals.sum(), (vals * vals).sum()]], dtype=torch.float64))
        open(os.path.join(tmp, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_fingerprint_allreduce_and_gather_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
