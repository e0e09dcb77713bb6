"""N > 1 host logic on CPU: world_size-2 gloo processes shard a request by image index (seed + k), "generate"
deterministic per-seed images, all-gather once, and every rank must hold the images in global order — the same
images a single rank produces for the whole batch (the reference's seed-offset property, distributed.py:297-305)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))


def _fake_image(seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (8, 8, 3), generator=g, dtype=torch.uint8)


def _worker(rank, world, total, port, out_dir):
    from b200sd import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = sharding.shard_counts(total, world)
    start, count = sharding.shard_bounds(total, world, rank)
    local = torch.stack([_fake_image(1000 + start + k) for k in range(count)]) if count else torch.empty((0, 8, 8, 3), dtype=torch.uint8)
    full = sharding.all_gather_images(local, counts)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("total,world", [(8, 2), (7, 2), (5, 2)])
def test_shard_and_gather_two_ranks(tmp_path, total, world):
    port = 29500 + (os.getpid() + total) % 2000
    mp.spawn(_worker, args=(world, total, port, str(tmp_path)), nprocs=world, join=True)
    want = torch.stack([_fake_image(1000 + k) for k in range(total)])
    for r in range(world):
        assert torch.equal(torch.load(os.path.join(str(tmp_path), f"r{r}.pt")), want)


def test_shard_counts_match_scheduler_split():
    from b200sd import sharding
    assert sharding.shard_counts(32, 8) == [4] * 8
    assert sharding.shard_counts(8, 3) == [3, 3, 2]          # SURVEY.md App. A "remainder"
    assert sharding.shard_counts(30, 8) == [4, 4, 4, 4, 4, 4, 3, 3]
    assert sharding.shard_bounds(30, 8, 6) == (24, 3)
    assert sum(sharding.shard_counts(7, 4)) == 7
