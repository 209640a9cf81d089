"""N>1 host logic on CPU: world_size-2 gloo processes split the samples, accumulate with the
CPU oracle standing in for the per-rank renderer, and reduce onto rank 0; the result must equal
the single-rank render of the same sample set."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_split_samples_covers_range():
    from tuturenderer_b200.multigpu import split_samples
    for spp in (0, 1, 7, 16, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            ranges = [split_samples(spp, world, r) for r in range(world)]
            assert ranges[0][0] == 0
            for (b0, c0), (b1, _c1) in zip(ranges, ranges[1:]):
                assert b0 + c0 == b1
            assert ranges[-1][0] + ranges[-1][1] == spp
            counts = [c for _, c in ranges]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        split_samples(4, 2, 2)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from tuturenderer_b200 import api
    from tuturenderer_b200.multigpu import render_distributed
    from oracle import oracle_py
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(24, 24)
    osc = oracle_py.OracleScene(sc)
    spp = 5  # odd: ranks get 3 and 2 samples

    def accumulate(begin, count, accum):
        # oracle returns sum * (1/total_spp); undo the scale to get the raw sums
        img = osc.render_path(count, seed=9, sample_begin=begin, total_spp=1, threads=2)
        accum += torch.from_numpy(img.reshape(-1))

    accum = torch.zeros(24 * 24 * 3, dtype=torch.float32)
    render_distributed(accumulate, accum, spp, rank, world, dist)
    if rank == 0:
        np.save(out_path, (accum / spp).numpy())
    dist.barrier()
    dist.destroy_process_group()


def _worker_bdpt(rank, world, port, out_path):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from tuturenderer_b200 import api
    from tuturenderer_b200.multigpu import render_distributed
    from oracle import oracle_py
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(20, 20)
    sc.bkgcolor = (0.1, 0.2, 0.3)
    osc = oracle_py.OracleScene(sc)
    spp = 5
    bkg = np.array(sc.bkgcolor, np.float32)

    def accumulate(begin, count, accum):
        # the oracle returns bkgcolor + sums / total_spp; with total_spp = 1 the sums are (img - bkg)
        img = osc.render_bdpt(count, seed=9, sample_begin=begin, total_spp=1, threads=1)
        accum += torch.from_numpy((img - bkg).reshape(-1))

    accum = torch.zeros(20 * 20 * 3, dtype=torch.float32)
    render_distributed(accumulate, accum, spp, rank, world, dist)
    if rank == 0:
        np.save(out_path, (accum / spp).numpy().reshape(20, 20, 3) + bkg)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_bdpt_equals_single_rank(api, oracle, tmp_path):
    """BDPT shards the same way (t = 1 splats land in the rank's own full-frame buffer; bkgcolor is
    added once after the reduce)."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "img.npy"
    mp.spawn(_worker_bdpt, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(20, 20)
    sc.bkgcolor = (0.1, 0.2, 0.3)
    want = oracle.OracleScene(sc).render_bdpt(5, seed=9, threads=1)
    assert np.allclose(got, want, rtol=2e-5, atol=2e-6)


def test_two_rank_gloo_render_equals_single_rank(api, oracle, tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "img.npy"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(24, 24)
    want = oracle.OracleScene(sc).render_path(5, seed=9).reshape(-1)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
