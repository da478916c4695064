"""CPU suite, part 2: the C-ABI library (loads, exports, host-only entry points), the
host-side mirror of the reference interface, and registration into the REAL reference
registry/dispatcher (no compute calls: there is no GPU here and no CPU fallback)."""

import ctypes
import math
import os
import re

import numpy as np
import pytest
from conftest import ROOT
from hypothesis import given, settings
from hypothesis import strategies as st

import pyfstat_b200
from pyfstat_b200 import _lib, backend
from pyfstat_b200.atoms import ATOM_DTYPE, AtomBatch, from_multi_fstat_atoms, synth_atoms
from pyfstat_b200.batch import shard_range
from pyfstat_b200.window import TransientWindowRange, canonical_window

U32 = 0xFFFFFFFF


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


# ---- C ABI ------------------------------------------------------------------------------


def test_library_exports_every_declared_symbol():
    """Every function declared in include/tcw_b200.h is exported by libtcw_b200.so."""
    header = open(os.path.join(ROOT, "include", "tcw_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tcw_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    L = _lib.load_library()
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in tcw_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert L.tcw_abi_version() == _lib.TCW_ABI_VERSION


def test_struct_layouts_match_header():
    assert ATOM_DTYPE.itemsize == 32
    assert ctypes.sizeof(_lib.CWindowRange) == 28
    assert ctypes.sizeof(_lib.CResult) == _lib.RESULT_DTYPE.itemsize == 80
    for name, _ in _lib.CResult._fields_:
        assert getattr(_lib.CResult, name).offset == _lib.RESULT_DTYPE.fields[name][1], name


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_fails_loudly():
    with pytest.raises(_lib.TcwError) as e:
        _lib.Handle(0)
    assert "no CPU fallback" in str(e.value)
    assert not backend.backend_available()
    b = synth_atoms(1, 16, ("H1",), seed=1)
    with pytest.raises(_lib.TcwError):
        pyfstat_b200.b200_compute_transient_fstat_map(b, canonical_window("rect", 10**9, 16), False)


def test_map_dims_host_helper():
    L = _lib.load_library()
    a, b = ctypes.c_uint32(), ctypes.c_uint32()
    w = _lib.c_window(canonical_window("rect", 10**9, 1440))
    assert L.tcw_map_dims(ctypes.byref(w), ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (1439, 1441)
    w = _lib.CWindowRange(3, 0, 0, 1, 0, 0, 1)
    assert L.tcw_map_dims(ctypes.byref(w), ctypes.byref(a), ctypes.byref(b)) == _lib.E_WINDOW
    w = _lib.CWindowRange(1, 0, 10, 0, 0, 10, 1)
    assert L.tcw_map_dims(ctypes.byref(w), ctypes.byref(a), ctypes.byref(b)) == _lib.E_INVALID
    w = _lib.CWindowRange(0, 5, 6, 7, 8, 9, 10)
    assert L.tcw_map_dims(ctypes.byref(w), ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (1, 1)


@settings(max_examples=600, deadline=None)
@given(
    wtype=st.sampled_from([1, 2]),
    t0_data=st.integers(0, U32),
    off=st.integers(-10_000_000, 400_000_000),
    tau_n=st.integers(0, U32),
    TAtom=st.one_of(st.integers(1, 7200), st.sampled_from([1, 2, 3, 1800, 1801, 4096, 65536, 2**31 - 1, 2**31, 2**31 + 1, U32])),
    numAtoms=st.integers(1, 200_000),
)
def test_kernel_index_arithmetic_bit_exact(oracle, wtype, t0_data, off, tau_n, TAtom, numAtoms):
    """The kernels' index math (magic-number division included; host build of the very same
    inline functions) equals the oracle's plain C division for arbitrary uint32 inputs."""
    t0_m = (t0_data + off) & U32
    assert _lib.cell_index_range(wtype, t0_m, tau_n, t0_data, TAtom, numAtoms) == oracle.index_range(
        wtype, t0_m, tau_n, t0_data, TAtom, numAtoms
    )


def test_magic_division_boundaries(oracle):
    """Dividend sweep around multiples of the divisor, for awkward divisors."""
    for TAtom in (1, 2, 3, 7, 900, 1800, 1801, 2**16, 2**16 + 1, 2**31 - 1, 2**31, 2**32 - 1):
        for k in (0, 1, 2, 1000, (2**32 - 1) // TAtom):
            for d in (-2, -1, 0, 1, 2):
                x = k * TAtom + d
                if not 0 <= x <= U32:
                    continue
                # choose t0_m so that t0_m - t0_data + TAtom//2 == x (mod 2^32)
                t0_m = (x - TAtom // 2) & U32
                got = _lib.cell_index_range(1, t0_m, 0, 0, TAtom, 2**31)[0]
                q = x // TAtom
                want = min(max(q - (1 << 32) if q & 0x80000000 else q, 0), 2**31 - 1)
                assert got == want, (TAtom, x)


# ---- host-side mirror of the reference interface --------------------------------------------


class FakeAtom:
    def __init__(self, rec):
        self.timestamp = int(rec["timestamp"])
        self.a2_alpha = float(rec["a2_alpha"])
        self.b2_alpha = float(rec["b2_alpha"])
        self.ab_alpha = float(rec["ab_alpha"])
        self.Fa_alpha = complex(rec["Fa_re"], rec["Fa_im"])
        self.Fb_alpha = complex(rec["Fb_re"], rec["Fb_im"])


class FakeVec:
    def __init__(self, arr, TAtom):
        self.length = len(arr)
        self.TAtom = TAtom
        self.data = [FakeAtom(r) for r in arr]


class FakeMulti:
    """Duck type of lalpulsar.MultiFstatAtomVector as the reference reads it (tcw:607-632)."""

    def __init__(self, batch, t=0):
        self.data = [FakeVec(a, batch.TAtom) for a in batch.template(t)]
        self.length = len(self.data)


def test_adapter_accepts_lalpulsar_duck_type():
    b = synth_atoms(1, 20, ("H1", "L1"), seed=3, gap_fraction=0.2)
    got = from_multi_fstat_atoms(FakeMulti(b))
    assert got.T == 1 and got.numDet == 2 and got.TAtom == 1800
    for X in range(2):
        assert np.array_equal(got.template(0)[X], b.template(0)[X])
    assert from_multi_fstat_atoms(b) is b
    vec = FakeMulti(b)
    vec.data[1].TAtom = 900
    with pytest.raises(ValueError):
        from_multi_fstat_atoms(vec)


def test_atom_batch_validation():
    b = synth_atoms(3, 10, ("H1",), seed=1)
    assert (b.T, b.numDet, b.stride) == (3, 1, 10) and len(b[1:3]) == 2
    with pytest.raises(TypeError):
        AtomBatch(np.zeros((1, 1, 4)), np.ones((1, 1)), 1800)
    with pytest.raises(ValueError):
        AtomBatch(b.atoms, np.zeros((3, 1), dtype=np.uint32), 1800)
    with pytest.raises(ValueError):
        AtomBatch(b.atoms, b.n_atoms, 0)


def test_synthetic_atoms_recipe():
    """deterministic per (seed, template), rank-1 per-atom covariance, E[2F]=4."""
    a = synth_atoms(4, 64, ("H1", "L1"), seed=10)
    b = synth_atoms(2, 64, ("H1", "L1"), seed=12)
    assert np.array_equal(a.atoms[2:], b.atoms)  # template t uses default_rng(seed + t)
    r = a.atoms[0, 0]
    assert np.allclose(r["a2_alpha"] * r["b2_alpha"], r["ab_alpha"] ** 2, rtol=1e-5, atol=1e-12)
    assert np.array_equal(r["timestamp"], 10**9 + 1800 * np.arange(64))
    g = synth_atoms(1, 200, ("H1", "L1"), seed=1, gap_fraction=0.1)
    assert 150 < g.n_atoms[0, 0] < 200 and not np.array_equal(
        g.template(0)[0]["timestamp"][: 100], g.template(0)[1]["timestamp"][: 100]
    )


def test_window_range_mirror():
    w = canonical_window("exp", 10**9, 1440)
    assert w.dims() == (1439, 1441) and w.type == pyfstat_b200.TRANSIENT_EXPONENTIAL
    with pytest.raises(ValueError) as e:
        TransientWindowRange(3).check_type()
    assert "Unknown window-type (3)" in str(e.value)  # message of tcw:691-697
    with pytest.raises(ValueError):
        TransientWindowRange.from_any(TransientWindowRange(1, -5, 0, 1, 0, 0, 1))
    src = TransientWindowRange(1, 1, 2, 3, 4, 5, 6)
    cp = TransientWindowRange.from_any(src)
    cp.t0 = 99
    assert src.t0 == 1


def test_shard_range_partitions_templates():
    for T in (0, 1, 7, 8, 100, 100000):
        for G in (1, 2, 3, 8):
            parts = [shard_range(T, r, G) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == T
            assert all(parts[i][1] == parts[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


# ---- registration into the registry / dispatcher ---------------------------------------------


def _check_registration(tcw):
    assert "b200" not in tcw.fstatmap_versions
    with pytest.raises(ValueError):
        tcw.init_transient_fstat_map_features("b200")  # stock behaviour (tcw:485-490)
    pyfstat_b200.register(tcw)
    try:
        assert set(tcw.fstatmap_versions) >= {"lal", "b200"}
        feats = tcw._get_transient_fstat_map_features()
        assert set(feats) >= {"lal", "pycuda", "b200"}
        assert feats["b200"] == backend.backend_available()
        # the other backends keep working exactly as before
        f2, ctx = tcw.init_transient_fstat_map_features("lal")
        assert ctx is None and "b200" in f2
        with pytest.raises(ValueError):
            tcw.init_transient_fstat_map_features("nonsense")
        if not feats["b200"]:
            with pytest.raises(RuntimeError):
                tcw.init_transient_fstat_map_features("b200")
            with pytest.raises(Exception) as e:
                tcw.call_compute_transient_fstat_map("b200", feats, None, None, False)
            assert "not available" in str(e.value)  # tcw:536-539
        pyfstat_b200.register(tcw)  # idempotent
    finally:
        pyfstat_b200.unregister(tcw)
    assert "b200" not in tcw.fstatmap_versions
    with pytest.raises(ValueError):
        tcw.init_transient_fstat_map_features("b200")


def test_register_into_real_reference_module(ref_tcw):
    _check_registration(ref_tcw)


def test_register_into_stand_in_module():
    import fake_tcw

    _check_registration(fake_tcw)


class OracleBackedHandle:
    """TEST-ONLY stand-in for _lib.Handle that answers map_batch with the CPU oracle, so that
    the Python plumbing (dispatcher -> registered callable -> result class) can be driven
    through the REAL reference dispatcher in a container without a GPU."""

    device_name = "oracle (test double)"
    _h = 1
    device_index = 0

    def __init__(self, oracle):
        self.O = oracle
        self.calls = []
        self.uploads = 0
        self.generation = 0
        self._last_batch = None

    def get_exp_lut(self):
        x, n = self.O.get_exp_lut()
        return x, n, True

    def _compute(self, batch, window, flags, raise_on_degenerate):
        self.calls.append(flags)
        w = TransientWindowRange.from_any(window)
        res = np.zeros(batch.T, dtype=_lib.RESULT_DTYPE)
        Fs = []
        for t in range(batch.T):
            r = self.O.compute_map(batch.template(t), batch.TAtom, w, exact_exp=bool(flags & _lib.EXP_EXACT),
                                   allow_degenerate=True, want_btsg=bool(flags & _lib.WANT_BTSG))
            for k in ("maxF", "m_ML", "n_ML", "t0_ML", "tau_ML", "m_MP", "n_MP", "N_t0", "N_tau", "numAtoms",
                      "t0_data", "lnBtSG", "t0_MP", "tau_MP"):
                res[k][t] = r[k]
            if r["status"] and raise_on_degenerate and not (flags & _lib.ALLOW_DEGENERATE):
                raise _lib.DegenerateWindowError(_lib.E_DEGENERATE, "degenerate")
            Fs.append(r["F_mn"].astype(np.float32))
        return res, (np.stack(Fs) if flags & _lib.WANT_FMN else None)

    def map_batch(self, batch, window, flags=0, raise_on_degenerate=True):
        self.generation += 1
        self.uploads += 1
        self._last_batch = batch
        return self._compute(batch, window, flags, raise_on_degenerate)

    # resident API (what the result object uses for everything after the first call)
    def upload(self, batch):
        self.generation += 1
        self.uploads += 1
        self._last_batch = batch

    def map_resident(self, window, flags=0):
        self.generation += 1
        self._resident = self._compute(self._last_batch, window, flags, False)

    def fetch_results(self, raise_on_degenerate=True):
        return self._resident[0]

    def fetch_fmn(self, t, N_t0, N_tau):
        F = self._resident[1][t]
        assert F.shape == (N_t0, N_tau)
        return F.copy()

    def map_batch_windows(self, batch, windows, flags=0, raise_on_degenerate=True):
        if isinstance(windows, np.ndarray):  # (T, 7) uint32 rows of transientWindowRange_t
            windows = [TransientWindowRange(*(int(v) for v in row)) for row in windows]
        recs = [self.map_batch(batch[t], w, flags, raise_on_degenerate)[0] for t, w in enumerate(windows)]
        return np.concatenate(recs), None

    def close(self):
        pass


def test_plugin_through_real_reference_dispatcher(ref_tcw, oracle, monkeypatch, tmp_path):
    """register() + the reference's own call_compute_transient_fstat_map + the reference's
    own pyTransientFstatMap base class, with the device replaced by a test double."""
    fake = OracleBackedHandle(oracle)
    monkeypatch.setattr(backend, "get_handle", lambda device=-1: fake)
    monkeypatch.setattr(backend, "backend_present", lambda: True)
    monkeypatch.setattr(backend, "select_device", lambda name=None: 0)
    pyfstat_b200.register(ref_tcw)
    try:
        feats, ctx = ref_tcw.init_transient_fstat_map_features("b200")
        assert feats["b200"] and ctx is None
        b = synth_atoms(1, 48, ("H1", "L1"), seed=21)
        w = canonical_window("rect", 10**9, 48)
        before = (w.type, w.t0, w.t0Band, w.dt0, w.tau, w.tauBand, w.dtau)
        fm, timing = ref_tcw.call_compute_transient_fstat_map("b200", feats, FakeMulti(b), w, BtSG=True)
        assert (w.type, w.t0, w.t0Band, w.dt0, w.tau, w.tauBand, w.dtau) == before
        assert isinstance(fm, ref_tcw.pyTransientFstatMap) and timing >= 0
        o = oracle.compute_map(b.template(0), 1800, w)
        assert fm.maxF == np.float32(o["maxF"]) and (fm.t0_ML, fm.tau_ML) == (o["t0_ML"], o["tau_ML"])
        assert fm.lnBtSG == o["lnBtSG"] and fm.t0_MP == o["t0_MP"] and fm.tau_MP == o["tau_MP"]
        n_calls = len(fake.calls)
        # what the callers read (core.py:1460,1465,1527; grid_based_searches.py:1128-1133)
        assert fm.get_maxF_idx() == (o["m_ML"], o["n_ML"])
        assert fm.get_lnBtSG() == o["lnBtSG"]
        assert fm.get_t0_max_posterior(w) == pytest.approx(o["t0_MP"])
        assert fm.get_tau_max_posterior(w) == pytest.approx(o["tau_MP"])
        assert len(fake.calls) == n_calls, "fused results must not trigger another device call"
        assert fm._F_mn is None, "F_mn must stay lazy until it is read"
        # single cell as core.py:1541 reads it: one 1x1 map, the full map is NOT materialised
        F = fm.F_mn
        assert F.shape == (47, 49) and F.dtype == np.float32 and len(F) == 47
        assert F[fm.get_maxF_idx()] == np.float32(fm.maxF)
        assert len(fake.calls) == n_calls + 1 and not (fake.calls[-1] & _lib.WANT_FMN)
        assert fm._F_mn is None
        with pytest.raises(IndexError):
            F[47, 0]
        # full map as tcw:311 / the tests read it: materialised once, then a plain ndarray
        Fa = np.asarray(F)
        assert len(fake.calls) == n_calls + 2 and fake.calls[-1] & _lib.WANT_FMN
        assert isinstance(fm.F_mn, np.ndarray) and np.shares_memory(fm.F_mn, Fa)
        assert F[3, 4] == Fa[3, 4] and F[-1, -1] == Fa[46, 48] and F[3].shape == (49,)
        assert float(F.max()) == fm.maxF and (2 * F)[1, 2] == 2 * Fa[1, 2]
        assert len(fake.calls) == n_calls + 2
        assert fm.F_at(3, 4) == Fa[3, 4]
        F = Fa
        # the reference's own writer/reader on our object (tcw:289-317, 159-184)
        path = tmp_path / "map.dat"
        fm.write_F_mn_to_file(str(path), w, header=["hello"])
        rd = ref_tcw.pyTransientFstatMap(from_file=str(path))
        assert rd.F_mn.shape == F.shape and np.allclose(rd.F_mn, F, rtol=1e-7)
        assert (rd.t0_ML, rd.tau_ML) == (fm.t0_ML, fm.tau_ML)
        # BtSG=False: nan until asked (tcw:142-144), then computed on demand
        fm2, _ = ref_tcw.call_compute_transient_fstat_map("b200", feats, FakeMulti(b), w, BtSG=False)
        assert math.isnan(fm2.lnBtSG) and math.isnan(fm2.t0_MP)
        n_calls, n_up = len(fake.calls), fake.uploads
        assert fm2.get_lnBtSG() == o["lnBtSG"]
        # ONE pass serves both lnBtSG and the F_mn read that follows, on the resident atoms
        assert np.array_equal(np.asarray(fm2.F_mn), F)
        assert fm2.get_t0_max_posterior(w) == pytest.approx(o["t0_MP"])
        assert len(fake.calls) == n_calls + 1 and fake.uploads == n_up
        # unknown window type: ValueError like tcw:691-697
        with pytest.raises(ValueError):
            ref_tcw.call_compute_transient_fstat_map(
                "b200", feats, FakeMulti(b), TransientWindowRange(3, 0, 0, 1, 0, 0, 1), False
            )
        # TRANSIENT_NONE: full-span rect, caller's object untouched
        wn = TransientWindowRange(0, 1, 2, 3, 4, 5, 6)
        fm3, _ = ref_tcw.call_compute_transient_fstat_map("b200", feats, FakeMulti(b), wn, False)
        assert (wn.type, wn.t0, wn.tau) == (0, 1, 4) and fm3.F_mn.shape == (1, 1)
        assert fm3.F_mn[0, 0] == F[0, -1]
    finally:
        pyfstat_b200.unregister(ref_tcw)


def test_semicoherent_and_bsgl_helpers(oracle, monkeypatch):
    """per-segment 2F (core.py:2282-2289), per-detector sums with the NaN
    rule (core.py:2236-2260), per-detector 2F at the multi-detector argmax (core.py:1527-1541)
    and the cumulative 2F (core.py:1648-1665) -- host logic against the oracle (test double)."""
    from pyfstat_b200 import batch as B
    from pyfstat_b200 import semicoherent as SC

    fake = OracleBackedHandle(oracle)
    monkeypatch.setattr(backend, "get_handle", lambda device=-1: fake)
    monkeypatch.setattr(SC, "get_handle", lambda device=-1: fake)
    monkeypatch.setattr(B, "get_handle", lambda device=-1: fake)
    n, TAtom, t0 = 96, 1800, 10**9
    b = synth_atoms(2, n, ("H1", "L1"), seed=31)
    nsegs = 8
    tb = np.linspace(t0, t0 + n * TAtom, nsegs + 1)
    w = SC.semicoherent_window_range(tb, tb[1] - tb[0])
    assert (w.type, w.t0, w.t0Band, w.dt0, w.tau, w.tauBand, w.dtau) == (1, t0, (nsegs - 1) * 12 * TAtom, 12 * TAtom,
                                                                       12 * TAtom, 0, 1)
    assert w.dims() == (nsegs, 1)
    twoF = SC.per_segment_twoF(b, w)
    assert twoF.shape == (2, nsegs)
    for t in range(2):
        o = oracle.compute_map(b.template(t), TAtom, w)
        assert np.array_equal(twoF[t], 2.0 * o["F_mn"][:, 0].astype(np.float32).astype(np.float64))
    twoFX, per_seg = SC.single_IFO_twoFs(b, w)
    assert twoFX.shape == (2, 2) and per_seg.shape == (2, 2, nsegs)
    o = oracle.compute_map([b.template(1)[1]], TAtom, w)
    assert np.array_equal(per_seg[1, 1], 2.0 * o["F_mn"][:, 0].astype(np.float32).astype(np.float64))
    assert twoFX[1, 1] == per_seg[1, 1].sum()
    with pytest.raises(ValueError):
        SC.single_detector_batch(b, 2)
    # transient BSGL ingredient: per-detector 2F at the multi-detector argmax cell
    wt = canonical_window("rect", t0, n)
    rec, _ = B.map_batch(b, wt)
    tx = SC.twoFX_at_maxTwoF(b, wt, rec)
    for t in range(2):
        for X in range(2):
            oX = oracle.compute_map([b.template(t)[X]], TAtom, wt, allow_degenerate=True)
            assert tx[t, X] == 2.0 * float(np.float32(oX["F_mn"][int(rec["m_ML"][t]), int(rec["n_ML"][t])]))
    # cumulative 2F: equally spaced durations -> one 1 x N_tau map; otherwise one call each
    durs = SC.cumulative_durations(t0, t0 + n * TAtom, TAtom, 48)
    assert durs[0] == 2 * TAtom and durs[-1] == n * TAtom
    n_calls = len(fake.calls)
    cum = SC.twoF_cumulative(b, t0, durs)
    assert cum.shape == (2, 48) and len(fake.calls) == n_calls + 1
    for k in (0, 17, 47):
        o1 = oracle.compute_map(b.template(0), TAtom, TransientWindowRange(1, t0, 0, 1, int(durs[k]), 0, 1))
        assert cum[0, k] == 2.0 * float(np.float32(o1["maxF"]))
    cum2 = SC.twoF_cumulative(b, t0, [2 * TAtom, 5 * TAtom, 6 * TAtom])
    assert cum2.shape == (2, 3) and cum2[0, 0] == cum[0, 0]

    # install(): the two SemiCoherentSearch methods, attributes as in the reference
    class FakeResults:
        numDetectors = 2

        def __init__(self, multi):
            self.multiFatoms = [multi]

    class SemiCoherentSearch:
        singleFstats = True

        def _get_per_segment_twoF(self):
            raise AssertionError("must be replaced")

        def get_semicoherent_single_IFO_twoFs(self, record_segments=False):
            raise AssertionError("must be replaced")

    class FakeCore:
        pass

    FakeCore.SemiCoherentSearch = SemiCoherentSearch
    SC.install(FakeCore)
    s = SemiCoherentSearch()
    s.FstatResults = FakeResults(FakeMulti(b[0]))
    s.semicoherentWindowRange = w
    s.twoFX = np.zeros(3)
    s.twoFX_per_segment = np.zeros((3, nsegs))
    assert np.array_equal(s._get_per_segment_twoF(), twoF[0])
    out = s.get_semicoherent_single_IFO_twoFs(record_segments=True)
    assert out[0] == twoFX[0, 0] and out[1] == twoFX[0, 1] and out[2] == 0
    assert np.array_equal(s.twoFX_per_segment[1], per_seg[0, 1])
    SC.uninstall(FakeCore)
    with pytest.raises(AssertionError):
        SemiCoherentSearch()._get_per_segment_twoF()


class _SwigThis:
    def __init__(self, addr):
        self._a = addr

    def __int__(self):
        return self._a


class _SwigAtom(FakeAtom):
    """Element wrapper as SWIG hands them out: attributes + ``.this`` = C address of the struct."""

    def __init__(self, rec, addr, reads):
        super().__init__(rec)
        self.this = _SwigThis(addr)
        reads[0] += 1


class _SwigArray:
    """View of a contiguous C array of 32-byte FstatAtom structs (here: a numpy buffer)."""

    def __init__(self, arr, stride=32, corrupt=None):
        self.arr = np.ascontiguousarray(arr)
        self.stride, self.corrupt, self.reads = stride, corrupt, [0]

    def __getitem__(self, i):
        rec = self.arr[i]
        if self.corrupt is not None and i == self.corrupt:
            rec = rec.copy()
            rec["ab_alpha"] += 1.0  # the attribute path disagrees with the raw bytes
        return _SwigAtom(rec, self.arr.ctypes.data + i * self.stride, self.reads)


def test_zero_copy_swig_ingest_is_verified(monkeypatch):
    """one block read of lal's FstatAtom[] instead of 6N attribute reads -- used
    only when the layout assumption verifies at run time, otherwise the per-atom loop."""
    b = synth_atoms(1, 500, ("H1", "L1"), seed=41)

    def multi(**kw):
        m = FakeMulti(b)
        for X, v in enumerate(m.data):
            v.data = _SwigArray(b.template(0)[X], **kw)
        return m

    m = multi()
    got = from_multi_fstat_atoms(m)
    for X in range(2):
        assert np.array_equal(got.template(0)[X], b.template(0)[X])
        assert m.data[X].data.reads[0] <= 6, "block read: only the verification atoms are touched"
        assert not np.shares_memory(got.atoms, m.data[X].data.arr)
    # wrong stride (not the 32-byte record) or bytes that disagree with the attributes: slow path, same result
    for kw in ({"stride": 40}, {"corrupt": 499}, {"corrupt": 1}):
        m = multi(**kw)
        got = from_multi_fstat_atoms(m)
        want = [a.copy() for a in b.template(0)]
        if "corrupt" in kw:
            for a in want:
                a["ab_alpha"][kw["corrupt"]] += 1.0
        for X in range(2):
            assert np.array_equal(got.template(0)[X], want[X])
            assert m.data[X].data.reads[0] >= 500
    # switch
    monkeypatch.setenv("PYFSTAT_B200_ZERO_COPY", "0")
    m = multi()
    from_multi_fstat_atoms(m)
    assert m.data[0].data.reads[0] >= 500
