"""world_size-2 `gloo` test of the N>1 host logic (polycap_b200/distributed.py): contiguous id shards and the
single SUM all-reduce of the tallies.  The per-rank compute here is the CPU oracle (test infrastructure),
standing in for the GPU session; what is checked is the sharding/reduction, against a single-rank run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def test_shard_range_tiles():
    from polycap_b200.distributed import shard_range
    for n, w in ((10, 3), (7, 8), (1000003, 8), (0, 2)):
        edges = [shard_range(n, r, w, id0=5) for r in range(w)]
        assert edges[0][0] == 5 and edges[-1][1] == 5 + n
        for (a, b), (c, d) in zip(edges, edges[1:]):
            assert b == c and 0 <= (b - a) - (d - c) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def _worker(rank, world, port, n_ids, seed, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import harness as H
    from polycap_b200.distributed import shard_range, allreduce_tallies, efficiencies
    from conftest import optic_of, source_of
    g = load_golden("ellip")
    lo, hi = shard_range(n_ids, rank, world)
    sw, cnt, _, _ = H.Oracle().run_ids(source_of(H, g), optic_of(H, g), seed, lo, hi - lo, threads=2)
    t_sum = torch.from_numpy(sw.copy())
    t_cnt = torch.zeros(16, dtype=torch.int64)
    t_cnt[0], t_cnt[1], t_cnt[2], t_cnt[3] = hi - lo, cnt["n_exit"], cnt["n_not_entered"], cnt["n_not_transmitted"]
    t_cnt[4], t_cnt[5], t_cnt[6] = cnt["n_uncounted"], cnt["sum_irefl_exit"], cnt["sum_irefl_all"]
    allreduce_tallies(t_sum, t_cnt)
    if rank == 0:
        np.savez(out, sw=t_sum.numpy(), cnt=t_cnt.numpy(), eff=efficiencies(t_sum.numpy(), t_cnt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tallies_equal_single_rank(tmp_path, harness, orc):
    from conftest import optic_of, source_of
    n_ids, seed = 4001, 31337
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_ids, seed, out), nprocs=2, join=True)
    got = np.load(out)
    g = load_golden("ellip")
    sw, cnt, _, _ = orc.run_ids(source_of(harness, g), optic_of(harness, g), seed, 0, n_ids, threads=4)
    assert np.allclose(got["sw"], sw, rtol=1e-12)
    assert got["cnt"][0] == n_ids and got["cnt"][1] == cnt["n_exit"] and got["cnt"][2] == cnt["n_not_entered"]
    assert got["cnt"][3] == cnt["n_not_transmitted"] and got["cnt"][6] == cnt["sum_irefl_all"]
    started = cnt["n_exit"] + cnt["n_not_entered"] + cnt["n_not_transmitted"]
    assert np.allclose(got["eff"], sw / started, rtol=1e-12)
