"""GPU suite: the sharding drivers with the REAL CUDA path under a process group.

Two ranks share the one GPU of the test box (gloo for the record gather: NCCL refuses two ranks on
one device; the NCCL gather runs in `bench.py --gpus N`).  Under test: `map_sharded` and
`BatchedTransientGridSearch.run(group)` give, on every rank, exactly the records of a single-rank
run -- bit for bit, lnBtSG included (the marginals are accumulated in fixed point, so they do not
depend on the order in which CTAs finish): SURVEY appendix E.6 on hardware.

Exponential window: bit-identical for ANY split of the templates into launches.  Rectangular
window: the host planner picks the tile shape from the number of templates in a launch, and the
FP32 split point of a cell follows its tile, so F_mn is reproducible to the last bit only between
launches of the same size (otherwise to ~1e-7 relative); the test therefore shards 12 templates
in launches of 3 on every rank count."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

N_ATOMS, T_TOTAL = 300, 12
FIELDS = ("maxF", "m_ML", "n_ML", "t0_ML", "tau_ML", "lnBtSG", "t0_MP", "tau_MP", "m_MP", "n_MP", "status",
          "N_t0", "N_tau", "numAtoms", "t0_data")
RANGES = {"F0": [0.0, float(T_TOTAL - 1), 1.0], "F1": [0.0], "F2": [0.0], "Alpha": [1.0], "Delta": [0.5]}


def _make_shard(lo, hi):
    from pyfstat_b200.atoms import synth_atoms

    # seeded by GLOBAL template index: every rank count sees the same templates
    return synth_atoms(hi - lo, N_ATOMS, ("H1", "L1"), seed=4000 + lo)


def _atoms_for_points(points):
    from pyfstat_b200.atoms import AtomBatch

    parts = [_make_shard(int(p["F0"]), int(p["F0"]) + 1) for p in points]
    return AtomBatch(np.concatenate([b.atoms for b in parts]), np.concatenate([b.n_atoms for b in parts]), 1800)


def _run_all(out_dir, tag):
    from pyfstat_b200.batch import map_sharded
    from pyfstat_b200.grid_search import BatchedTransientGridSearch
    from pyfstat_b200.window import canonical_window

    for win, chunk in (("rect", 3), ("exp", 5)):  # exp: ragged launches (5 + 1 | 5 + 1 vs 5 + 5 + 2)
        w = canonical_window(win, 10**9, N_ATOMS)
        rec = map_sharded(_make_shard, T_TOTAL, w, BtSG=True, device=0, chunk=chunk)
        np.save(os.path.join(out_dir, f"{tag}_{win}.npy"), rec)
    s = BatchedTransientGridSearch(_atoms_for_points, RANGES, canonical_window("rect", 10**9, N_ATOMS), BtSG=True,
                                   batch_size=3, device=0)
    data = s.run()
    np.save(os.path.join(out_dir, f"{tag}_grid.npy"), data)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    _run_all(out_dir, f"rank{rank}")
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_drivers_on_gpu_match_single_rank(tmp_path):
    import torch.multiprocessing as mp

    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    _run_all(str(tmp_path), "single")  # no process group in this process: one rank holds everything
    for name in ("rect", "exp"):
        single = np.load(tmp_path / f"single_{name}.npy")
        assert len(single) == T_TOTAL and np.all(single["status"] == 0)
        assert len(set(single["maxF"].tolist())) == T_TOTAL  # templates are distinct
        for rank in range(world):
            got = np.load(tmp_path / f"rank{rank}_{name}.npy")
            for f in FIELDS:
                assert got[f].tobytes() == single[f].tobytes(), (name, rank, f)
    single = np.load(tmp_path / "single_grid.npy")
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}_grid.npy")
        assert got.tobytes() == single.tobytes(), rank
