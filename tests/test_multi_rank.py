"""N>1 host-side logic on CPU (gloo, world_size 2): the frame batch is block-partitioned
over ranks with no data-path collective; the only communication is the
barrier and the max-over-ranks of the timing, exactly as bench.py does under torchrun."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

import cmsisdsp_b200 as cd
from oracle_lib import oracle
from seeded_inputs import cfft_input


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, N, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = cfft_input("q15", N, frames=B, seed=11)               # every rank can regenerate the global batch
    lo, hi = cd.shard_frames(B, world, rank)
    y = oracle().cfft("q15", N, x[lo:hi], 0, 1)               # stand-in for the per-rank GPU work
    np.save(os.path.join(out_dir, f"part{rank}.npy"), y)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)  # "device time" of this rank
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "meta.npy"), np.array([t.item(), cnt.item()]))
    dist.destroy_process_group()


def test_two_rank_partition_covers_batch_exactly(tmp_path):
    B, N, world = 37, 64, 2
    mp.spawn(_worker, args=(world, _free_port(), B, N, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"part{r}.npy") for r in range(world)]
    whole = oracle().cfft("q15", N, cfft_input("q15", N, frames=B, seed=11), 0, 1)
    assert np.array_equal(np.concatenate(parts), whole)
    tmax, total = np.load(tmp_path / "meta.npy")
    assert tmax == world and total == B
