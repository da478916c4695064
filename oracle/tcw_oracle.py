"""ctypes bindings of the CPU oracle (oracle/tcw_oracle.c) and of the host-compiled reference
kernels (oracle/_ref/libtcw_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
legs, never by the product package.  See the header of tcw_oracle.c for what is and is not
pinned ("parity unpinned" against lalpulsar itself).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libtcw_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtcw_ref.so")

ATOM_DTYPE = np.dtype(
    [
        ("timestamp", "<u4"),
        ("a2_alpha", "<f4"),
        ("b2_alpha", "<f4"),
        ("ab_alpha", "<f4"),
        ("Fa_re", "<f4"),
        ("Fa_im", "<f4"),
        ("Fb_re", "<f4"),
        ("Fb_im", "<f4"),
    ]
)
CHANNELS = ("a2_alpha", "b2_alpha", "ab_alpha", "Fa_re", "Fa_im", "Fb_re", "Fb_im")

SEM_LAL = 0
SEM_PYCUDA = 1
ERR_DEGENERATE = -5


class WindowRange(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("type", "t0", "t0Band", "dt0", "tau", "tauBand", "dtau")]


class OracleResult(C.Structure):
    _fields_ = [
        ("lnBtSG", C.c_double),
        ("t0_MP", C.c_double),
        ("tau_MP", C.c_double),
        ("maxF", C.c_double),
        ("m_ML", C.c_uint32),
        ("n_ML", C.c_uint32),
        ("t0_ML", C.c_uint32),
        ("tau_ML", C.c_uint32),
        ("m_MP", C.c_uint32),
        ("n_MP", C.c_uint32),
        ("N_t0", C.c_uint32),
        ("N_tau", C.c_uint32),
        ("numAtoms", C.c_uint32),
        ("t0_data", C.c_uint32),
        ("status", C.c_int32),
    ]


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(HERE, "tcw_oracle.c")
    ):
        subprocess.run(["make", "-C", HERE, "libtcw_oracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_SO)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.oracle_fast_neg_exp.restype = C.c_double
        L.oracle_fast_neg_exp.argtypes = [C.c_double]
        L.oracle_fstat_from_sums.restype = C.c_float
        L.oracle_fstat_from_sums.argtypes = [C.c_float] * 7
        L.oracle_index_range.restype = None
        L.oracle_set_exp_lut.argtypes = [C.c_double, C.c_uint32]
        _lib = L
    return _lib


def ref_lib():
    """The reference's own kernels compiled for the host, or None where not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = C.CDLL(REF_SO)
        _ref.ref_rect_rows_needed.restype = C.c_uint
    return _ref


def _win(w) -> WindowRange:
    return WindowRange(
        int(w.type), int(w.t0), int(w.t0Band), int(w.dt0), int(w.tau), int(w.tauBand), int(w.dtau)
    )


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def fast_neg_exp(x: float) -> float:
    return lib().oracle_fast_neg_exp(float(x))


def set_exp_lut(xmax: float = 20.0, length: int = 5120) -> None:
    """Geometry of the emulated XLALFastNegExp table (default: SURVEY A.4-1, 20 / 5120)."""
    rc = lib().oracle_set_exp_lut(float(xmax), int(length))
    if rc:
        raise ValueError("oracle_set_exp_lut failed")


def get_exp_lut():
    x, n = C.c_double(), C.c_uint32()
    lib().oracle_get_exp_lut(C.byref(x), C.byref(n))
    return x.value, n.value


def exp_lut() -> np.ndarray:
    _, length = get_exp_lut()
    out = np.zeros(length + 1, dtype=np.float64)
    n = lib().oracle_exp_lut(_ptr(out), length + 1)
    assert n == length + 1
    return out


def fstat_from_sums(A, B, Cc, Far, Fai, Fbr, Fbi) -> float:
    return lib().oracle_fstat_from_sums(*(np.float32(v).item() for v in (A, B, Cc, Far, Fai, Fbr, Fbi)))


def index_range(wtype, t0_m, tau_n, t0_data, TAtom, numAtoms):
    i0, i1, t1 = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().oracle_index_range(
        C.c_uint32(wtype),
        C.c_uint32(t0_m & 0xFFFFFFFF),
        C.c_uint32(tau_n & 0xFFFFFFFF),
        C.c_uint32(t0_data),
        C.c_uint32(TAtom),
        C.c_uint32(numAtoms),
        C.byref(i0),
        C.byref(i1),
        C.byref(t1),
    )
    return i0.value, i1.value


def _pack(det_arrays):
    numDet = len(det_arrays)
    stride = max(1, max(len(a) for a in det_arrays))
    atoms = np.zeros((numDet, stride), dtype=ATOM_DTYPE)
    n_atoms = np.zeros(numDet, dtype=np.uint32)
    for X, a in enumerate(det_arrays):
        atoms[X, : len(a)] = a
        n_atoms[X] = len(a)
    return atoms, n_atoms, stride


def merge_binned(det_arrays, TAtom: int) -> np.ndarray:
    """XLALmergeMultiFstatAtomsBinned restatement -> merged ATOM_DTYPE array."""
    atoms, n_atoms, stride = _pack(det_arrays)
    tmin = min(int(a["timestamp"][0]) for a in det_arrays if len(a))
    tmax = max(int(a["timestamp"][-1]) for a in det_arrays if len(a))
    cap = (tmax - tmin) // TAtom + 2
    out = np.zeros(cap, dtype=ATOM_DTYPE)
    N = C.c_uint32()
    rc = lib().oracle_merge_binned(
        _ptr(atoms), _ptr(n_atoms), len(det_arrays), C.c_uint32(stride), C.c_uint32(TAtom), _ptr(out),
        C.c_uint32(cap), C.byref(N),
    )
    if rc:
        raise ValueError(f"oracle_merge_binned failed: {rc}")
    return out[: N.value].copy()


def merged_to_matrix(merged: np.ndarray) -> np.ndarray:
    """[numAtoms x 7] float32, the reference's atomsInputMatrix (tcw:711-721)."""
    return np.ascontiguousarray(np.column_stack([merged[c] for c in CHANNELS]).astype(np.float32))


def compute_map(
    det_arrays,
    TAtom: int,
    window,
    *,
    semantics: int = SEM_LAL,
    exact_exp: bool = False,
    allow_degenerate: bool = False,
    want_btsg: bool = True,
    rect_vanilla: bool = False,
    want_fmn: bool = True,
):
    """One template through the oracle.  Returns a dict with F_mn (float64 [N_t0,N_tau]) and
    the result fields; ``status == ERR_DEGENERATE`` flags lal's single-atom abort."""
    merged = merge_binned(det_arrays, TAtom)
    w = _win(window)
    if w.type == 0:  # TRANSIENT_NONE -> rect over all data, on a copy (tcw:742-749)
        w = WindowRange(1, int(merged["timestamp"][0]), 0, TAtom, len(merged) * TAtom, 0, TAtom)
    N_t0, N_tau = C.c_uint32(), C.c_uint32()
    rc = lib().oracle_map_dims(C.byref(w), C.byref(N_t0), C.byref(N_tau))
    if rc == -2:
        raise ValueError("Unknown window-type")
    if rc:
        raise ValueError(f"oracle_map_dims failed: {rc}")
    F = np.zeros((N_t0.value, N_tau.value), dtype=np.float64) if (want_fmn or want_btsg) else None
    res = OracleResult()
    rc = lib().oracle_map(
        _ptr(merged), C.c_uint32(len(merged)), C.c_uint32(TAtom), C.byref(w), int(semantics),
        int(exact_exp), int(allow_degenerate), int(rect_vanilla), _ptr(F) if F is not None else None,
        C.byref(res),
    )
    if rc not in (0, ERR_DEGENERATE):
        raise ValueError(f"oracle_map failed: {rc}")
    status = res.status
    if want_btsg:
        lib().oracle_bstat(
            _ptr(F), N_t0, N_tau, C.c_double(res.maxF), C.byref(w), int(not exact_exp), C.byref(res)
        )
    out = {f: getattr(res, f) for f, _ in OracleResult._fields_}
    out["status"] = status
    out["F_mn"] = F
    out["merged"] = merged
    return out


def bstat(F_mn: np.ndarray, maxF: float, window, use_lut: bool):
    """lnBtSG, t0_MP, tau_MP, m_MP, n_MP of a given map (lal flavour if use_lut)."""
    F = np.ascontiguousarray(F_mn, dtype=np.float64)
    res = OracleResult()
    w = _win(window)
    rc = lib().oracle_bstat(
        _ptr(F), C.c_uint32(F.shape[0]), C.c_uint32(F.shape[1]), C.c_double(float(maxF)), C.byref(w),
        int(use_lut), C.byref(res),
    )
    if rc:
        raise ValueError(f"oracle_bstat failed: {rc}")
    return {k: getattr(res, k) for k in ("lnBtSG", "t0_MP", "tau_MP", "m_MP", "n_MP")}


def batch(atoms: np.ndarray, n_atoms: np.ndarray, TAtom: int, window, *, semantics=SEM_LAL,
          exact_exp=False, allow_degenerate=False, want_btsg=True, num_threads=1):
    """T templates (atoms [T,numDet,stride]) through the oracle, OpenMP over templates."""
    T, numDet, stride = atoms.shape
    atoms = np.ascontiguousarray(atoms)
    n_atoms = np.ascontiguousarray(n_atoms, dtype=np.uint32)
    res = (OracleResult * T)()
    w = _win(window)
    rc = lib().oracle_batch(
        _ptr(atoms), _ptr(n_atoms), C.c_uint32(stride), C.c_uint32(TAtom), T, numDet, C.byref(w),
        int(semantics), int(exact_exp), int(allow_degenerate), int(want_btsg), int(num_threads), res,
    )
    return rc, res


# ---- the reference's own kernels on the CPU (oracle/_ref) --------------------------------


def ref_kernel_map(matrix: np.ndarray, TAtom: int, t0_data: int, window) -> np.ndarray:
    """Run the reference's pycuda kernel (host-compiled) with the launch geometry of its
    wrapper.  ``matrix``: [numAtoms x 7] float32 (tcw:711-721).  Returns F_mn float32;
    cells the reference kernel never writes (Rect.cu:48 guard) are NaN."""
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libtcw_ref.so not built (needs /root/reference)")
    matrix = np.ascontiguousarray(matrix, dtype=np.float32)
    N_t0 = int(window.t0Band) // int(window.dt0) + 1
    N_tau = int(window.tauBand) // int(window.dtau) + 1
    args = [
        _ptr(matrix), C.c_uint(matrix.shape[0]), C.c_uint(TAtom), C.c_uint(t0_data),
        C.c_uint(int(window.t0)), C.c_uint(int(window.dt0)), C.c_uint(int(window.tau)),
        C.c_uint(int(window.dtau)), C.c_uint(N_t0), C.c_uint(N_tau),
    ]
    if int(window.type) == 1:
        rows = max(int(R.ref_rect_rows_needed(C.c_uint(N_t0))), N_t0)
        F = np.full((rows, N_tau), np.nan, dtype=np.float32)
        R.ref_rect(*args, _ptr(F))
        return F[:N_t0].copy()
    if int(window.type) == 2:
        F = np.full((N_t0, N_tau), np.nan, dtype=np.float32)
        R.ref_exp(*args, _ptr(F))
        return F
    raise ValueError("ref kernels exist for rect and exp windows only")
