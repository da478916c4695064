"""Batched TransientGridSearch-compatible driver: same grid, columns, detection
statistic and text format as the reference (grid_based_searches.py:173-217, 471-501,
1023-1042, 1080-1142, 1236-1252), maps evaluated in batches.  CPU variant: device replaced by
the oracle; GPU variant: the real thing, compared with the oracle per template."""

import numpy as np
import pytest

from pyfstat_b200 import _lib
from pyfstat_b200 import grid_search as gs
from pyfstat_b200.atoms import synth_atoms
from pyfstat_b200.window import TransientWindowRange, canonical_window

N_ATOMS = 96
RANGES = {"F0": [30.0, 30.5, 0.25], "F1": [-2e-10, -1e-10, 1e-10], "F2": [0], "Alpha": [1.0], "Delta": [0.5]}


def atoms_for_points(points):
    # deterministic per grid point: seed from the (F0, F1) position
    seeds = [int(round((p["F0"] - 30.0) / 0.25)) * 10 + int(round((p["F1"] + 2e-10) / 1e-10)) for p in points]
    parts = [synth_atoms(1, N_ATOMS, ("H1", "L1"), seed=500 + s) for s in seeds]
    from pyfstat_b200.atoms import batch_from_detector_lists

    return batch_from_detector_lists([b.template(0) for b in parts], 1800)


def check_against_oracle(search, oracle, rtol):
    data = search.data
    assert list(data.dtype.names) == ["F0", "F1", "F2", "Alpha", "Delta", "twoF", "maxTwoF", "lnBtSG",
                                      "t0_ML", "tau_ML", "t0_MP", "tau_MP"]
    assert len(data) == 3 * 2 and search.total_iterations == 6
    # grid order = itertools.product, last key fastest
    assert np.allclose(data["F0"], np.repeat([30.0, 30.25, 30.5], 2))
    w = search.window
    for i in range(len(data)):
        tpl = atoms_for_points(search.input_data[i : i + 1]).template(0)
        o = oracle.compute_map(tpl, 1800, w)
        full = oracle.compute_map(tpl, 1800, TransientWindowRange(type=0), want_btsg=False)
        assert data["maxTwoF"][i] == pytest.approx(2 * o["maxF"], rel=rtol)
        assert data["twoF"][i] == pytest.approx(2 * full["maxF"], rel=rtol)
        assert data["lnBtSG"][i] == pytest.approx(o["lnBtSG"], abs=1e-4)
        assert (data["t0_ML"][i], data["tau_ML"][i]) == (o["t0_ML"], o["tau_ML"])
        assert data["t0_MP"][i] == pytest.approx(o["t0_MP"]) and data["tau_MP"][i] == pytest.approx(o["tau_MP"])
    best = search.get_max_det_stat()
    assert best["lnBtSG"] == data["lnBtSG"].max()


def check_file_format(search, tmp_path):
    out = tmp_path / "grid.txt"
    search.output_file_header = ["date: today", "parameters: {}"]
    search.save_array_to_disk(str(out))
    lines = out.read_text().splitlines()
    assert lines[0] == "# date: today" and lines[2] == "# " + " ".join(search.output_keys)
    back = np.genfromtxt(str(out), names=search.output_keys)
    assert np.allclose(back["maxTwoF"], search.data["maxTwoF"], rtol=1e-8)
    assert np.array_equal(back["t0_ML"], search.data["t0_ML"])
    assert np.allclose(back["F0"], search.data["F0"], rtol=0, atol=1e-12)


def test_get_array_from_tuple_matches_reference_rule():
    assert np.allclose(gs.get_array_from_tuple([1.0, 2.0, 0.25]), [1.0, 1.25, 1.5, 1.75, 2.0])
    assert np.allclose(gs.get_array_from_tuple([5.0]), [5.0])
    assert np.allclose(gs.get_array_from_tuple([1, 2, 3, 4]), [1, 2, 3, 4])


def test_batched_grid_search_host_logic(oracle, monkeypatch, tmp_path):
    """No GPU: map_batch replaced by the oracle (test double), batching across uneven batches."""

    def fake_map_batch(batch, window, BtSG=False, want_fmn=False, **kw):
        res = np.zeros(batch.T, dtype=_lib.RESULT_DTYPE)
        for t in range(batch.T):
            r = oracle.compute_map(batch.template(t), batch.TAtom, window, want_btsg=BtSG, allow_degenerate=True)
            for k in ("maxF", "m_ML", "n_ML", "t0_ML", "tau_ML", "lnBtSG", "t0_MP", "tau_MP", "N_t0", "N_tau"):
                res[k][t] = r[k]
        return res, None

    in_flight = []

    def fake_submit(batch, window, BtSG=False, **kw):
        assert not in_flight, "one batch in flight"
        in_flight.append((batch, window, BtSG))
        return in_flight

    def fake_wait(ticket, **kw):
        batch, window, BtSG = ticket.pop()
        return fake_map_batch(batch, window, BtSG)[0]

    monkeypatch.setattr(gs, "submit_batch", fake_submit)
    monkeypatch.setattr(gs, "wait_batch", fake_wait)
    monkeypatch.setattr(gs, "map_again", lambda window, batch, **kw: fake_map_batch(batch, window)[0])
    w = canonical_window("rect", 10**9, N_ATOMS)
    s = gs.BatchedTransientGridSearch(atoms_for_points, RANGES, w, BtSG=True, batch_size=4)
    s.run()
    check_against_oracle(s, oracle, rtol=1e-7)
    check_file_format(s, tmp_path)
    s2 = gs.BatchedTransientGridSearch(atoms_for_points, RANGES, w, BtSG=False, batch_size=100)
    d2 = s2.run()
    assert "lnBtSG" not in d2.dtype.names and "t0_MP" not in d2.dtype.names
    assert np.array_equal(d2["maxTwoF"], s.data["maxTwoF"])


@pytest.mark.gpu
@pytest.mark.parametrize("win", ["rect", "exp"])
def test_batched_grid_search_on_gpu(oracle, tmp_path, win):
    w = canonical_window(win, 10**9, N_ATOMS)
    s = gs.BatchedTransientGridSearch(atoms_for_points, RANGES, w, BtSG=True, batch_size=4)
    s.run()
    check_against_oracle(s, oracle, rtol=1e-4)
    check_file_format(s, tmp_path)
    assert np.all(s.records["path"] >= 1) and np.all(s.records["status"] == 0)
