"""CPU suite, part 1: pins the oracle (oracle/tcw_oracle.c).

* against the reference's OWN kernels compiled for the host (oracle/_ref, live, build
  container only) and against the committed golden fixtures those produced
  (tests/golden/*.npz, everywhere);
* against the reference's own Python class pyTransientFstatMap (lnBtSG / MP / text format),
  through the fixtures and, where /root/reference exists, live;
* known answers and index-range properties.
"""

import math

import numpy as np
import pytest
from conftest import Win, golden_cases, load_golden
from hypothesis import given, settings
from hypothesis import strategies as st

from pyfstat_b200.atoms import synth_atoms
from pyfstat_b200.window import TransientWindowRange, canonical_window


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_kernel_fixture(oracle, name):
    """F_mn of the restatement (semantics 'pycuda': float exp, float products, no degenerate
    abort) is bit-identical to what the reference's own .cu kernels produced."""
    z, batch, w = load_golden(name)
    r = oracle.compute_map(batch.template(0), batch.TAtom, w, semantics=oracle.SEM_PYCUDA, exact_exp=True,
                           allow_degenerate=True, want_btsg=False)
    F = r["F_mn"].astype(np.float32)
    assert F.shape == z["F_ref"].shape
    assert np.array_equal(F, z["F_ref"]), f"max |dF| = {np.abs(F - z['F_ref']).max()}"
    assert np.array_equal(oracle.merged_to_matrix(r["merged"]), z["merged_matrix"])
    assert (r["m_ML"], r["n_ML"]) == tuple(int(x) for x in z["argmax"])
    assert np.float32(r["maxF"]) == z["maxF"]


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_bstat_matches_reference_python_fixture(oracle, name):
    """exact-exp lnBtSG / t0_MP / tau_MP equal the reference's numpy ports (tcw:196-287) run
    on the same map.  The reference sums float32 exp terms; the oracle sums in double."""
    z, batch, w = load_golden(name)
    b = oracle.bstat(z["F_ref"].astype(np.float64), float(z["maxF"]), w, use_lut=False)
    assert b["lnBtSG"] == pytest.approx(float(z["lnBtSG_numpy"]), abs=2e-5)
    assert b["t0_MP"] == pytest.approx(float(z["t0_MP"]), abs=1e-6)
    assert b["tau_MP"] == pytest.approx(float(z["tau_MP"]), abs=1e-6)


def test_oracle_matches_reference_kernels_live(oracle):
    """Live comparison with oracle/_ref on more shapes (build container only)."""
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = int(rng.integers(8, 120))
        dets = ("H1", "L1") if trial % 2 else ("H1",)
        b = synth_atoms(1, n, dets, seed=100 + trial, gap_fraction=0.1 if trial % 3 == 0 else 0.0)
        wtype = 1 + trial % 2
        dt0 = int(rng.integers(1, 4)) * 900
        dtau = int(rng.integers(1, 4)) * 900
        N_t0 = int(rng.integers(1, 12))
        N_tau = int(rng.integers(N_t0, 30))  # Rect.cu:48 only writes rows m < N_tauRange
        w = TransientWindowRange(wtype, 10**9 + int(rng.integers(0, 5000)), (N_t0 - 1) * dt0, dt0,
                                 int(rng.integers(1800, 9000)), (N_tau - 1) * dtau, dtau)
        r = oracle.compute_map(b.template(0), 1800, w, semantics=oracle.SEM_PYCUDA, exact_exp=True,
                               allow_degenerate=True, want_btsg=False)
        Fref = oracle.ref_kernel_map(oracle.merged_to_matrix(r["merged"]), 1800, int(r["t0_data"]), w)
        assert np.array_equal(r["F_mn"].astype(np.float32), Fref), (trial, w)


def test_reference_rect_kernel_row_guard_quirk(oracle):
    """Rect.cu:48 guards rows with `m < N_tauRange`, so with
    N_t0Range > N_tauRange the reference kernel never writes rows m >= N_tauRange.  The
    oracle (like lalpulsar) computes them; the rows the reference does write agree."""
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference)")
    b = synth_atoms(1, 60, ("H1",), seed=3)
    w = TransientWindowRange(1, 10**9, 40 * 1800, 1800, 3600, 9 * 1800, 1800)  # 41 x 10
    r = oracle.compute_map(b.template(0), 1800, w, semantics=oracle.SEM_PYCUDA, allow_degenerate=True,
                           want_btsg=False)
    Fref = oracle.ref_kernel_map(oracle.merged_to_matrix(r["merged"]), 1800, 10**9, w)
    assert np.isnan(Fref[10:]).all() and not np.isnan(Fref[:10]).any()
    assert np.array_equal(r["F_mn"].astype(np.float32)[:10], Fref[:10])
    assert not np.isnan(r["F_mn"]).any()


def test_reference_python_class_live(oracle, ref_tcw):
    """lnBtSG / MP estimators vs the real pyTransientFstatMap on a random map, and the
    text format round trip (tcw:159-184, 289-317)."""
    rng = np.random.default_rng(1)
    F = (0.5 * rng.chisquare(4, size=(23, 31))).astype(np.float32)
    F[7, 11] = 14.5
    w = TransientWindowRange(1, 10**9, 22 * 1800, 1800, 3600, 30 * 1800, 1800)
    fm = ref_tcw.pyTransientFstatMap(N_t0Range=23, N_tauRange=31)
    fm.F_mn = F
    fm.maxF = F.max()
    b = oracle.bstat(F.astype(np.float64), float(F.max()), w, use_lut=False)
    assert b["lnBtSG"] == pytest.approx(fm.get_lnBtSG(), abs=2e-5)
    assert b["t0_MP"] == pytest.approx(fm.get_t0_max_posterior(w), abs=1e-6)
    assert b["tau_MP"] == pytest.approx(fm.get_tau_max_posterior(w), abs=1e-6)
    assert (b["m_MP"], b["n_MP"]) == (7, 11)


# ---- XLALFastNegExp restatement ------------------------------------------------------------


def test_fast_neg_exp_table(oracle, explut):
    """Both table geometries on file (20 / 5120; the other recollection: 20 / 2000)."""
    xmax, length = explut
    dx = xmax / length
    lut = oracle.exp_lut()
    assert oracle.get_exp_lut() == (xmax, length)
    assert len(lut) == length + 1 and lut[0] == 1.0
    assert lut[length] == pytest.approx(math.exp(-xmax), rel=1e-14)
    assert oracle.fast_neg_exp(xmax * 1.000005) == 0.0
    assert oracle.fast_neg_exp(-0.5) == pytest.approx(math.exp(0.5), rel=1e-15)
    # nearest-point lookup: x in [i*dx - dx/2, i*dx + dx/2) -> LUT[i]
    assert oracle.fast_neg_exp(1.49 * dx) == lut[1]
    assert oracle.fast_neg_exp(1.51 * dx) == lut[2]
    xs = np.linspace(0, xmax, 4001)
    err = max(abs(oracle.fast_neg_exp(x) - math.exp(-x)) / math.exp(-x) for x in xs)
    assert 0.2 * dx < err < 0.51 * dx  # half a table step


def test_lut_probe_measures_the_geometry(oracle, explut):
    """pyfstat_b200.lut_probe recovers (xmax, length, entries) from the step function alone --
    the same routine that measures lalpulsar.FastNegExp wherever lalpulsar is importable."""
    from pyfstat_b200 import lut_probe

    xmax, length, table = lut_probe.measure_exp_lut(oracle.fast_neg_exp)
    assert (xmax, length) == explut
    assert np.array_equal(table, oracle.exp_lut())
    with pytest.raises(ValueError):
        lut_probe.measure_exp_lut(lambda x: math.exp(-x))  # smooth: no cut-off / no steps
    with pytest.raises(ValueError):
        lut_probe.measure_exp_lut(lambda x: 0.0 if x > 20 else math.exp(-math.floor(x * 10) / 10))  # floor, not nearest
    assert lut_probe.parse_geometry("20:2000") == (20.0, 2000)
    assert lut_probe.probe_lalpulsar() is None or len(lut_probe.probe_lalpulsar()) == 3


# ---- index ranges: the uint32 arithmetic of Rect.cu:21-31, 54-69 / Exp.cu:27-65 -----------

U32 = 0xFFFFFFFF


def model_index_range(wtype, t0_m, tau_n, t0_data, TAtom, numAtoms):
    """Python-integer model of the reference's C arithmetic (wrap + signed reinterpretation)."""
    def to_i32(u):
        return u - (1 << 32) if u & 0x80000000 else u

    half = TAtom // 2
    q = (((t0_m - t0_data + half) & U32) // TAtom) & U32
    i = max(to_i32(q), 0)
    i_t0 = min(i, numAtoms - 1)
    t1 = (t0_m + (3 if wtype == 2 else 1) * tau_n) & U32
    q = ((((t1 - t0_data + half) & U32) // TAtom) - 1) & U32
    i = max(to_i32(q), 0)
    i_t1 = min(i, numAtoms - 1)
    return i_t0, i_t1


@settings(max_examples=400, deadline=None)
@given(
    wtype=st.sampled_from([1, 2]),
    t0_data=st.integers(0, U32),
    off=st.integers(-10_000_000, 400_000_000),
    tau_n=st.integers(0, U32),
    TAtom=st.one_of(st.integers(1, 7200), st.sampled_from([1, 2, 1800, 1801, 65536, 2**31, U32])),
    numAtoms=st.integers(1, 200_000),
)
def test_index_range_matches_integer_model(oracle, wtype, t0_data, off, tau_n, TAtom, numAtoms):
    t0_m = (t0_data + off) & U32
    assert oracle.index_range(wtype, t0_m, tau_n, t0_data, TAtom, numAtoms) == model_index_range(
        wtype, t0_m, tau_n, t0_data, TAtom, numAtoms
    )


def test_index_range_wraparound_cases(oracle):
    # t0 before the data: the unsigned subtraction wraps and clamps to numAtoms-1, not 0 (A.2)
    assert oracle.index_range(1, 10**9 - 1000, 3600, 10**9, 1800, 48)[0] == 47
    # t0 within half an atom before the data start rounds to index 0
    assert oracle.index_range(1, 10**9 - 900, 3600, 10**9, 1800, 48)[0] == 0
    # quotient 0 for t1: "- 1" wraps to -1 -> clamped to 0
    assert oracle.index_range(1, 10**9, 0, 10**9, 1800, 48) == (0, 0)
    # window beyond the data end clamps to numAtoms-1
    assert oracle.index_range(1, 10**9, 10**8, 10**9, 1800, 48) == (0, 47)
    # exponential window: t1 = t0 + 3 tau
    assert oracle.index_range(2, 10**9, 3600, 10**9, 1800, 48) == (0, 5)


# ---- known answers ----------------------------------------------------


def test_single_cell_map_lnBtSG_closed_form(oracle):
    b = synth_atoms(1, 48, ("H1", "L1"), seed=2)
    w = TransientWindowRange(1, 10**9 + 5 * 1800, 0, 1800, 20 * 1800, 0, 1800)
    r = oracle.compute_map(b.template(0), 1800, w)
    assert r["F_mn"].shape == (1, 1)
    assert r["lnBtSG"] == pytest.approx(math.log(70.0) + r["maxF"], abs=1e-12)
    assert r["t0_MP"] == w.t0 and r["tau_MP"] == w.tau
    assert (r["t0_ML"], r["tau_ML"]) == (w.t0, w.tau)


def test_rect_last_tau_of_first_row_is_full_span_F(oracle):
    """t_tcw:124-127 analogue: F_mn[0,-1] of the canonical rect map == F over all data, which
    is also what TRANSIENT_NONE computes (tcw:742-749) -- without mutating the input."""
    b = synth_atoms(1, 96, ("H1", "L1"), seed=4)
    w = canonical_window("rect", 10**9, 96)
    r = oracle.compute_map(b.template(0), 1800, w)
    wn = TransientWindowRange(0, 123, 456, 789, 1011, 1213, 1415)
    before = (wn.type, wn.t0, wn.t0Band, wn.dt0, wn.tau, wn.tauBand, wn.dtau)
    rn = oracle.compute_map(b.template(0), 1800, wn)
    assert (wn.type, wn.t0, wn.t0Band, wn.dt0, wn.tau, wn.tauBand, wn.dtau) == before
    assert rn["F_mn"].shape == (1, 1)
    assert rn["F_mn"][0, 0] == r["F_mn"][0, -1]
    m = r["merged"]
    F_direct = oracle.fstat_from_sums(*(np.float32(sum(np.float32(v) for v in m[c])) for c in oracle.CHANNELS))
    assert rn["maxF"] == pytest.approx(F_direct, rel=2e-6)


def test_running_sums_equal_vanilla_sums(oracle):
    """Rect.cu:75-91 extends float32 running sums; recomputing every cell from scratch (the
    'vanilla' method) gives identical floats on sane windows -- which is why the generic CUDA
    kernel may sum per cell."""
    b = synth_atoms(1, 80, ("H1",), seed=6)
    w = canonical_window("rect", 10**9, 80)
    a = oracle.compute_map(b.template(0), 1800, w, want_btsg=False)
    v = oracle.compute_map(b.template(0), 1800, w, want_btsg=False, rect_vanilla=True)
    assert np.array_equal(a["F_mn"], v["F_mn"])


def test_degenerate_single_atom_window(oracle):
    """lal aborts when any cell has i_t1 == i_t0 (A.4-2); pycuda semantics silently proceed."""
    b = synth_atoms(1, 48, ("H1",), seed=7)
    w = canonical_window("rect", 10**9, 48)
    w.t0Band = 48 * 1800 - 1800  # t0 reaches the last atom
    r = oracle.compute_map(b.template(0), 1800, w)
    assert r["status"] == oracle.ERR_DEGENERATE
    r2 = oracle.compute_map(b.template(0), 1800, w, allow_degenerate=True)
    assert r2["status"] == 0
    # a single H1 atom has a2*b2 == ab^2: ill-conditioned -> fallback F = 2 (Rect.cu:112)
    assert r2["F_mn"][-1, 0] == 2.0


def test_constant_atoms_closed_form(oracle):
    """Identical atoms with non-degenerate (a2,b2,ab): every window sum is L * atom, so
    F = L * F_1 exactly representable relations hold: F(2L) = 2 F(L)."""
    n = 64
    from pyfstat_b200.atoms import ATOM_DTYPE

    a = np.zeros(n, dtype=ATOM_DTYPE)
    a["timestamp"] = 10**9 + 1800 * np.arange(n)
    a["a2_alpha"], a["b2_alpha"], a["ab_alpha"] = 0.5, 0.25, 0.125
    a["Fa_re"], a["Fa_im"], a["Fb_re"], a["Fb_im"] = 1.0, 0.5, -0.25, 0.75
    w = TransientWindowRange(1, 10**9, 0, 1800, 2 * 1800, 32 * 1800, 1800)
    r = oracle.compute_map([a], 1800, w, want_btsg=False)
    F = r["F_mn"][0]
    assert F[2] == pytest.approx(2 * F[0], rel=1e-6)  # tau = 4 atoms vs 2 atoms
    assert F[6] == pytest.approx(4 * F[0], rel=1e-6)
    assert np.all(np.diff(F) > 0)


def test_tie_rule_first_occurrence(oracle):
    """All-zero atoms: every cell is the F=2 fallback; argmax must be the first cell."""
    from pyfstat_b200.atoms import ATOM_DTYPE

    a = np.zeros(32, dtype=ATOM_DTYPE)
    a["timestamp"] = 10**9 + 1800 * np.arange(32)
    w = canonical_window("rect", 10**9, 32)
    r = oracle.compute_map([a], 1800, w)
    assert np.all(r["F_mn"] == 2.0)
    assert (r["m_ML"], r["n_ML"], r["maxF"]) == (0, 0, 2.0)
    assert (r["m_MP"], r["n_MP"]) == (0, 0)
    assert r["lnBtSG"] == pytest.approx(math.log(70.0) + 2.0, abs=1e-12)


def test_unknown_window_type_rejected(oracle):
    b = synth_atoms(1, 16, ("H1",), seed=1)
    with pytest.raises(ValueError):
        oracle.compute_map(b.template(0), 1800, TransientWindowRange(3, 10**9, 0, 1800, 3600, 0, 1800))


def test_lut_vs_exact_matters_at_1e4(oracle):
    """Why the backend emulates the lookup table: on a 3-day H1+L1 exp map the LUT-weighted F
    differs from the exact-exp F by far more than the 1e-4 parity target."""
    b = synth_atoms(1, 144, ("H1", "L1"), seed=8)
    w = canonical_window("exp", 10**9, 144)
    lut = oracle.compute_map(b.template(0), 1800, w, want_btsg=False)["F_mn"]
    exact = oracle.compute_map(b.template(0), 1800, w, want_btsg=False, exact_exp=True)["F_mn"]
    rel = np.abs(lut - exact) / np.abs(exact)
    assert 1e-4 < np.median(rel) < 2e-2


def test_lut_probe_through_bstat_route(oracle, explut):
    """When lalpulsar does not export FastNegExp, the table geometry is recovered from its
    ComputeTransientBstat on a 1 x 2 map (lut_probe.neg_exp_through_bstat) -- exercised here with a
    stand-in module whose Bstat is the oracle's restatement."""
    import types

    from pyfstat_b200 import lut_probe

    class Map:
        def __init__(self):
            self.F_mn = types.SimpleNamespace(data=np.zeros((1, 2)))
            self.maxF = 0.0

    win = TransientWindowRange(1, 0, 0, 1, 1, 1, 1)
    fake = types.SimpleNamespace(
        TRANSIENT_RECTANGULAR=1,
        CreateTransientFstatMap=lambda n_t0, n_tau: Map(),
        transientWindowRange_t=lambda: types.SimpleNamespace(),
        ComputeTransientBstat=lambda wr, fm: oracle.bstat(fm.F_mn.data, fm.maxF, win, use_lut=True)["lnBtSG"],
    )
    xmax, length, table = lut_probe.probe_lalpulsar(fake)
    assert (xmax, length) == explut and table is None
    fake.FastNegExp = oracle.fast_neg_exp  # the direct route wins when it exists, and returns the entries
    xmax, length, table = lut_probe.probe_lalpulsar(fake)
    assert (xmax, length) == explut and np.array_equal(table, oracle.exp_lut())
    assert lut_probe.probe_lalpulsar(types.SimpleNamespace()) is None  # neither route: keep the default
