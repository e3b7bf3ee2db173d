"""
Multi-process host logic on CPU (gloo, world_size 2): column sharding and the
all-reduce of ensemble statistics -- the only collective of the path.  The
per-rank inputs are synthetic water-table index tables; the sharded result must
equal the single-process statistics.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydromodel_b200.sharding import column_range, local_moments, reduce_moments
from hydromodel_b200.synthetic import ensemble_member, default_params
from conftest import SITE


def test_column_range_partitions_exactly():
    for n_total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [column_range(n_total, r, world) for r in range(world)]
            assert sum(c for _, c in blocks) == n_total
            assert blocks[0][0] == 0
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        column_range(10, 2, 2)


def test_ensemble_members_are_rank_independent():
    """A column's parameters depend on its global id only, not on how the
    ensemble is split (so 1-GPU and 8-GPU runs integrate the same columns)."""
    base = default_params(SITE)
    a = ensemble_member(base, 700, 1024)[0]["Soil_Properties"]
    b = ensemble_member(base, 700, 65536)[0]["Soil_Properties"]
    assert a == b


def _worker(rank, world, port, table, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = column_range(table.shape[0], rank, world)
    local = table[first:first + count]
    valid = local >= 0
    cnt, s, q = (torch.from_numpy(v) for v in local_moments(local, valid))
    mean, std, n = reduce_moments(cnt, s, q)
    if rank == 0:
        out["mean"], out["std"], out["n"] = mean.numpy().copy(), std.numpy().copy(), n.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_match_single_process():
    rng = np.random.default_rng(5)
    table = rng.integers(100, 200, size=(37, 50)).astype(np.float64)
    table[rng.random(table.shape) < 0.05] = -1.0            # skipped rows
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, table, out), nprocs=2, join=True)
    valid = table >= 0
    cnt = valid.sum(axis=0)
    mean = np.where(valid, table, 0).sum(axis=0) / cnt
    std = np.sqrt(np.maximum(np.where(valid, table ** 2, 0).sum(axis=0) / cnt - mean ** 2, 0))
    assert np.array_equal(out["n"], cnt.astype(float))
    assert np.allclose(out["mean"], mean, rtol=1e-13)
    assert np.allclose(out["std"], std, rtol=1e-9)


def _share_worker(rank, world, port, n_total, out):
    """What EnsembleSimulation.from_global does on the host, per rank: take this rank's block of
    the global ensemble, pack it, and all-reduce a statistic (here: count / sum / sum of squares
    of the porosity profile, the stand-in for theta without a device)."""
    from hydromodel_b200 import pack
    from hydromodel_b200.ensemble import EnsembleSimulation
    from hydromodel_b200.synthetic import ensemble_columns, synthetic_forcing
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = default_params(SITE)

    def members(first, count):
        return ensemble_columns(base, first, count, perturb_theta=True)
    first, count, b, cols, _ = EnsembleSimulation.local_share(None, members, n_total)   # rank/world from the env
    ens = pack.pack_ensemble(b, synthetic_forcing(30), cols)
    phi = ens["stat"][:, pack.S_PHI_NODE, :]
    cnt, s, q = (torch.from_numpy(v) for v in local_moments(phi))
    mean, std, n = reduce_moments(cnt, s, q)
    ids = torch.tensor([float(first), float(count)])
    gathered = [torch.zeros(2) for _ in range(world)]
    dist.all_gather(gathered, ids)
    if rank == 0:
        out["mean"], out["std"], out["n"] = mean.numpy().copy(), std.numpy().copy(), n.numpy().copy()
        out["blocks"] = [tuple(int(v) for v in g.tolist()) for g in gathered]
    dist.barrier()
    dist.destroy_process_group()


def test_from_global_host_sharding_world_size_2():
    from hydromodel_b200 import pack
    from hydromodel_b200.synthetic import ensemble_columns, synthetic_forcing
    n_total = 37
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_share_worker, args=(2, port, n_total, out), nprocs=2, join=True)
    assert out["blocks"] == [(0, 19), (19, 18)]
    b, cols = ensemble_columns(default_params(SITE), 0, n_total, perturb_theta=True)
    phi = pack.pack_ensemble(b, synthetic_forcing(30), cols)["stat"][:, pack.S_PHI_NODE, :]
    assert np.all(out["n"] == n_total)
    assert np.allclose(out["mean"], phi.mean(axis=0), rtol=1e-13)
    assert np.allclose(out["std"], phi.std(axis=0), rtol=1e-6, atol=1e-12)
