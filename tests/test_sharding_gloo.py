"""CPU tier, world_size 2 over gloo: the N > 1 path.  The hot path has no collective (clips are independent);
what multi-GPU adds is (a) the block partition of the batch, (b) gap draws that do not depend on the number of
ranks and (c) the optional final gather -- all host logic, exercised here with CPU tensors."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

from ml_audio_inpainting_b200 import gaps, sharding   # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.world() == (rank, world, rank)
        L, sr, hop = 80000, 16000, 192
        g = gaps.gap_len_samples(0.1, sr)
        np.random.seed(1234)                                   # every rank draws the WHOLE batch identically
        starts = gaps.draw_starts_exclusive(L, g, n_items)
        lo, hi = sharding.shard_bounds(n_items, rank, world)
        mine = starts[lo:hi]
        f0, f1 = gaps.cnnblstm_frame_range(mine, g, sr, hop)
        local = torch.from_numpy(np.stack([mine, mine + g, f0, f1], 1).astype(np.int64))
        full = sharding.gather_rows(local, n_items)
        assert full.shape == (n_items, 4)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7])
def test_two_rank_sharding_matches_single_process(tmp_path, n_items):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    L, sr, hop = 80000, 16000, 192
    g = gaps.gap_len_samples(0.1, sr)
    np.random.seed(1234)
    starts = gaps.draw_starts_exclusive(L, g, n_items)
    f0, f1 = gaps.cnnblstm_frame_range(starts, g, sr, hop)
    ref = np.stack([starts, starts + g, f0, f1], 1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), ref)
