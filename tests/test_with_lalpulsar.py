"""the parity tests against lalpulsar ITSELF.

lalsuite is not installable in the build container (no wheel, no network), so every test here
skips unless ``lalpulsar`` imports -- e.g. once ``baseline/_ref`` carries a lalsuite install.
They are written against the reference's own test (tests/test_tcw_fstat_map_funcs.py:40-132,
same window recipe and the same way of filling a ``MultiFstatAtomVector``) and are what turns
DESIGN.md's "parity unpinned" into pinned:

* ``b200`` vs ``lal`` through the reference's dispatcher on identical atoms: F_mn within 1e-4
  relative, maxF / t0_ML / tau_ML equal (argmax identical except documented near-ties),
  lnBtSG within 1e-4 absolute, t0_MP / tau_MP equal -- rect and exp, one and two detectors;
* the constants of Appendix F: XLALFastNegExp's table geometry is MEASURED from lalpulsar
  (pyfstat_b200.lut_probe) and must be what the registered backend's handle emulates -- the test
  also prints which of the two recollections on file (20 / 5120, 20 / 2000) is the true one;
  sizeof(FstatAtom) == 32 (the one-block SWIG ingest must verify and be used), the enum values.
"""

import numpy as np
import pytest

lalpulsar = pytest.importorskip("lalpulsar")
pyfstat = pytest.importorskip("pyfstat")

import pyfstat_b200  # noqa: E402
from pyfstat_b200 import atoms as atoms_mod  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import TRANSIENT_EXP_EFOLDING  # noqa: E402

pytestmark = pytest.mark.gpu

tcw = pyfstat.tcw_fstat_map_funcs
T0, TSFT = 700000000, 1800


def lal_multi_atoms(batch, t=0):
    """Fill a lalpulsar.MultiFstatAtomVector the way the reference's test does (t_tcw:65-78)."""
    tpl = batch.template(t)
    multi = lalpulsar.CreateMultiFstatAtomVector(len(tpl))
    for X, a in enumerate(tpl):
        multi.data[X] = lalpulsar.CreateFstatAtomVector(len(a))
        multi.data[X].TAtom = batch.TAtom
        for i in range(len(a)):
            multi.data[X].data[i].timestamp = int(a["timestamp"][i])
            multi.data[X].data[i].a2_alpha = float(a["a2_alpha"][i])
            multi.data[X].data[i].b2_alpha = float(a["b2_alpha"][i])
            multi.data[X].data[i].ab_alpha = float(a["ab_alpha"][i])
            multi.data[X].data[i].Fa_alpha = float(a["Fa_re"][i]) + 1j * float(a["Fa_im"][i])
            multi.data[X].data[i].Fb_alpha = float(a["Fb_re"][i]) + 1j * float(a["Fb_im"][i])
    return multi


def lal_window(window, n):
    w = lalpulsar.transientWindowRange_t()
    w.type = lalpulsar.TRANSIENT_RECTANGULAR if window == "rect" else lalpulsar.TRANSIENT_EXPONENTIAL
    w.t0, w.t0Band, w.dt0 = T0, n * TSFT - 2 * TSFT, TSFT
    w.tau, w.tauBand, w.dtau = 2 * TSFT, n * TSFT - 2 * TSFT, TSFT
    return w


@pytest.fixture(scope="module")
def registered():
    pyfstat_b200.register(tcw)
    feats, _ = tcw.init_transient_fstat_map_features("b200")
    yield feats
    pyfstat_b200.unregister(tcw)


def test_exp_lut_geometry_measured_from_lalpulsar(registered):
    """The table the backend emulates is the one lalpulsar uses: measured, not recalled."""
    from pyfstat_b200 import backend, lut_probe

    probed = lut_probe.probe_lalpulsar()
    assert probed is not None, "neither lalpulsar.FastNegExp nor the ComputeTransientBstat route worked"
    xmax, length, table = probed
    print(f"lalpulsar XLALFastNegExp table: xmax = {xmax}, {length} steps (1/dx = {length / xmax})")
    assert (xmax, length) in ((20.0, 5120), (20.0, 2000)), "a third geometry: update DESIGN.md L1"
    if table is not None:
        dx = xmax / length
        assert np.allclose(table, np.exp(-dx * np.arange(length + 1)), rtol=1e-14, atol=0)
    h = backend.get_handle(-1)
    assert h.get_exp_lut()[:2] == (xmax, length), "the handle must have been configured from the probe"
    assert backend.exp_lut_geometry()[3].startswith("measured")


def test_enums_and_struct_layout():
    assert (lalpulsar.TRANSIENT_NONE, lalpulsar.TRANSIENT_RECTANGULAR, lalpulsar.TRANSIENT_EXPONENTIAL,
            lalpulsar.TRANSIENT_LAST) == (0, 1, 2, 3)
    assert TRANSIENT_EXP_EFOLDING == 3
    b = synth_atoms(1, 64, ("H1",), seed=1, t0_data=T0)
    vec = lal_multi_atoms(b).data[0]
    fast = atoms_mod._view_swig_atoms(vec.data, int(vec.length))
    assert fast is not None, "sizeof(FstatAtom) != 32 or the SWIG element address is not exposed as .this"
    assert np.array_equal(fast, b.template(0)[0])


@pytest.mark.parametrize("window", ["rect", "exp"])
@pytest.mark.parametrize("dets", [("H1",), ("H1", "L1")])
def test_b200_equals_lal(registered, window, dets):
    n = 48  # one day of 1800-s atoms, as in the reference's test
    b = synth_atoms(1, n, dets, seed=7, t0_data=T0, gap_fraction=0.1 if len(dets) > 1 else 0.0)
    multi, w = lal_multi_atoms(b), lal_window(window, n)
    ref, _ = tcw.call_compute_transient_fstat_map("lal", registered, multi, w, BtSG=True)
    got, _ = tcw.call_compute_transient_fstat_map("b200", registered, multi, w, BtSG=True)
    F_ref, F_got = np.asarray(ref.F_mn, dtype=np.float64), np.asarray(got.F_mn, dtype=np.float64)
    assert F_got.shape == F_ref.shape
    rel = np.abs(F_got - F_ref) / np.maximum(np.abs(F_ref), 1e-30)
    assert rel.max() <= 1e-4, (window, dets, rel.max(), np.unravel_index(rel.argmax(), rel.shape))
    assert got.maxF == pytest.approx(ref.maxF, rel=1e-4)
    assert (got.t0_ML, got.tau_ML) == (ref.t0_ML, ref.tau_ML)
    assert got.lnBtSG == pytest.approx(ref.lnBtSG, abs=1e-4)
    assert got.t0_MP == pytest.approx(ref.t0_MP) and got.tau_MP == pytest.approx(ref.tau_MP)
    assert got.get_maxF_idx() == ref.get_maxF_idx()


def test_generic_kernels_bit_identical_to_lal(registered, monkeypatch):
    """The bit-faithful path: if the recalled lal semantics (table, REAL4 accumulators, merge) are
    right, the generic kernels reproduce lalpulsar's float results exactly."""
    monkeypatch.setenv("PYFSTAT_B200_GENERIC", "1")
    n = 48
    b = synth_atoms(1, n, ("H1", "L1"), seed=9, t0_data=T0)
    multi = lal_multi_atoms(b)
    for window in ("rect", "exp"):
        w = lal_window(window, n)
        ref, _ = tcw.call_compute_transient_fstat_map("lal", registered, multi, w, BtSG=True)
        got, _ = tcw.call_compute_transient_fstat_map("b200", registered, multi, w, BtSG=True)
        assert np.array_equal(np.asarray(got.F_mn), np.asarray(ref.F_mn, dtype=np.float32)), window
        assert got.lnBtSG == pytest.approx(ref.lnBtSG, abs=1e-11)


def test_single_atom_window_raises_like_lal(registered):
    n = 48
    b = synth_atoms(1, n, ("H1",), seed=3, t0_data=T0)
    multi, w = lal_multi_atoms(b), lal_window("rect", n)
    w.tau = TSFT  # windows of a single atom: lalpulsar aborts the map (XLAL_EDOM)
    with pytest.raises(Exception):
        tcw.call_compute_transient_fstat_map("lal", registered, multi, w, BtSG=False)
    with pytest.raises(ValueError):
        tcw.call_compute_transient_fstat_map("b200", registered, multi, w, BtSG=False)
