"""The N>1 plumbing on CPU: world_size-2 gloo job, interleaved row bands, gather, de-interleave.

The GPU renderer is replaced by the oracle here (the partition and the collective are what is
under test); bench.py --mode bands runs the same partition code over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
W, H, SPP = 96, 61, 2   # odd height: ranks get bands of different length


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _oracle
    from rtrace_b200 import partition
    scene = _oracle.Scene()
    start, stride, rows = partition.band_spec(H, rank, world)
    cap = partition.band_capacity(H, world)
    band = np.zeros((cap, W, 4), np.uint8)
    band[:rows], ctr = scene.render_rows(W, H, SPP, start, stride, rows, threads=1)
    t = torch.from_numpy(band)
    gathered = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    rays = torch.tensor([ctr.primary_rays + ctr.shadow_rays], dtype=torch.int64)
    dist.all_reduce(rays)   # whole-job ray count, as bench.py sums it
    if rank == 0:
        frame = partition.deinterleave(gathered, H).numpy()
        np.save(out_path, frame)
        np.save(out_path + ".rays.npy", rays.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_gather_reassembles_the_frame(tmp_path, oracle_scene8):
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    frame = np.load(out)
    ref, ctr = oracle_scene8.render(W, H, SPP)
    assert frame.shape == ref.shape and np.array_equal(frame, ref)
    assert int(np.load(out + ".rays.npy")[0]) == ctr.primary_rays + ctr.shadow_rays


def test_partition_arithmetic():
    sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
    from rtrace_b200 import partition
    for h in (1, 2, 7, 61, 2160, 4320):
        for g in (1, 2, 3, 4, 8):
            rows = [partition.band_rows(h, r, g) for r in range(g)]
            assert sum(rows) == h and max(rows) == partition.band_capacity(h, g) and max(rows) - min(rows) <= 1
            covered = sorted(r + k * g for r in range(g) for k in range(rows[r]))
            assert covered == list(range(h))
            bands = [np.full((partition.band_capacity(h, g), 2), -1) for _ in range(g)]
            for r in range(g):
                for k in range(rows[r]):
                    bands[r][k] = r + k * g
            assert (partition.deinterleave(bands, h)[:, 0] == np.arange(h)).all()
            # blocked partition: every row exactly once, counts agree
            for block in (4, 16):
                allrows = []
                for r in range(g):
                    start, stride, blk, count = partition.block_band_spec(h, r, g, block)
                    rows_r = partition.block_band_rows(h, r, g, block)
                    assert len(rows_r) == count and blk == block and (not rows_r or rows_r[0] == start)
                    allrows += rows_r
                assert sorted(allrows) == list(range(h))
