"""World-size-2 test of the multi-GPU path's host logic on CPU (gloo): spp slices tile the frame's
samples, one reduce(sum) to rank 0, one division by spp.  There is no GPU here, so each rank's slice
is rendered by the oracle standing in for the device renderer; what is under test is
vecchio_b200.sharding (the partition and the reduce), which bench.py uses unchanged with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_spp_slices_tile_the_sample_range():
    from vecchio_b200.sharding import spp_slice
    for spp in (8, 1000, 1001, 2000, 10_000, 257):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, n = spp_slice(r, world, spp)
                cover += list(range(b, b + n))
                assert abs(n - spp / world) < 1
            assert cover == list(range(spp))
    with pytest.raises(ValueError):
        spp_slice(2, 2, 100)
    with pytest.raises(ValueError):
        spp_slice(0, 8, 4)


def _worker(rank, world, port, W, spp, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    import vecchio_b200 as vb
    from oracle import pyoracle as po
    from vecchio_b200.sharding import reduce_sums_to_root, spp_slice

    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = vb.Scene("cornell_box")
    cam = scene.next_camera()
    o = po.OracleScene(scene)
    begin, count = spp_slice(rank, world, spp)
    # the slice as per-pixel SUMS, like vk_render_device (the oracle returns sum / spp)
    mean, _, st = o.render(cam, vb.render_params(W, W, spp, 50, seed=9, spp_begin=begin, spp_count=count), threads=2)
    sums = torch.from_numpy(mean.astype(np.float32) * np.float32(spp)).contiguous()
    reduce_sums_to_root(sums, world)
    paths = torch.tensor([st.paths], dtype=torch.int64)
    dist.all_reduce(paths)
    if rank == 0:
        np.save(out_path, np.concatenate([(sums / np.float32(spp)).numpy().ravel(), [float(paths.item())]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reduce_to_the_single_rank_render(tmp_path, vb, po):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    W, spp, world = 24, 10, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank0.npy")
    mp.spawn(_worker, args=(world, port, W, spp, out), nprocs=world, join=True)
    got = np.load(out)
    scene = vb.Scene("cornell_box")
    cam = scene.next_camera()
    whole, _, st = po.OracleScene(scene).render(cam, vb.render_params(W, W, spp, 50, seed=9), threads=2)
    assert got[-1] == st.paths == W * W * spp
    # same samples (the oracle seeds per (seed, pixel, global sample index)); only the fp32 order of the
    # two partial sums differs
    assert np.allclose(got[:-1].reshape(W, W, 3), whole, rtol=2e-6, atol=1e-7)
