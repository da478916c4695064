"""CPU suite, part 3: the N>1 path (template sharding + record all_gather) on world_size-2
gloo.  The device computation is replaced by the CPU oracle through map_sharded's `compute`
hook -- what is under test is the host logic: contiguous shards, no collective on the data
path, one all_gather of the records, identical results on every rank and vs a single rank."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, T, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import tcw_oracle as O
    from pyfstat_b200 import _lib
    from pyfstat_b200.atoms import synth_atoms
    from pyfstat_b200.batch import map_sharded
    from pyfstat_b200.window import canonical_window

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    w = canonical_window("rect", 10**9, 40)

    def make_shard(lo, hi):
        # atoms are seeded by template index, so every rank count sees identical templates
        return synth_atoms(hi - lo, 40, ("H1", "L1"), seed=1000 + lo)

    def compute(batch):
        rc, res = O.batch(batch.atoms, batch.n_atoms, batch.TAtom, w, want_btsg=True, num_threads=1)
        assert rc == 0
        out = np.zeros(batch.T, dtype=_lib.RESULT_DTYPE)
        for t in range(batch.T):
            for k in ("maxF", "m_ML", "n_ML", "t0_ML", "tau_ML", "lnBtSG", "t0_MP", "tau_MP", "N_t0", "N_tau"):
                out[k][t] = getattr(res[t], k)
        return out

    rec = map_sharded(make_shard, T, w, BtSG=True, compute=compute, chunk=3)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), rec)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_map_world_size_2_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp

    T, world, port = 7, 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, T, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert len(r0) == T and r0.tobytes() == r1.tobytes()  # every rank holds all records

    # single process, no process group
    sys.path.insert(0, ROOT)
    from oracle import tcw_oracle as O
    from pyfstat_b200.atoms import synth_atoms
    from pyfstat_b200.window import canonical_window

    w = canonical_window("rect", 10**9, 40)
    for t in range(T):
        b = synth_atoms(1, 40, ("H1", "L1"), seed=1000 + t)
        o = O.compute_map(b.template(0), 1800, w)
        assert r0["maxF"][t] == np.float32(o["maxF"])
        assert (r0["m_ML"][t], r0["n_ML"][t]) == (o["m_ML"], o["n_ML"])
        assert r0["lnBtSG"][t] == o["lnBtSG"]
    assert len(set(r0["maxF"].tolist())) == T  # templates are distinct


def test_gather_without_process_group_is_identity():
    from pyfstat_b200 import _lib
    from pyfstat_b200.batch import gather_records

    rec = np.zeros(5, dtype=_lib.RESULT_DTYPE)
    rec["maxF"] = np.arange(5)
    assert gather_records(rec, 5) is rec
    with pytest.raises(ValueError):
        gather_records(rec, 6)
