"""ctypes binding of ``libtcw_b200.so`` (C ABI declared in ``include/tcw_b200.h``).

Thin by design: argument marshalling only.  There is no CPU fallback anywhere in this
package -- if the library cannot be loaded or no CUDA device is usable, the calls raise.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .atoms import ATOM_DTYPE, AtomBatch
from .window import TransientWindowRange

PKG = os.path.dirname(os.path.abspath(__file__))
# $PYFSTAT_B200_LIB: another build of the same library (development: kernel variants side by side)
LIB_PATH = os.environ.get("PYFSTAT_B200_LIB") or os.path.join(PKG, "libtcw_b200.so")

TCW_ABI_VERSION = 3

# flags (include/tcw_b200.h)
WANT_FMN = 0x1
WANT_BTSG = 0x2
EXP_EXACT = 0x4
ALLOW_DEGENERATE = 0x8
FORCE_GENERIC = 0x10
BTSG_TABLE = 0x20
EXP_DIRECT = 0x40

# default geometry of the emulated XLALFastNegExp table (TCW_EXPLUT_DEFAULT_* in the header)
EXPLUT_DEFAULT = (20.0, 5120)

# error codes
E_INVALID, E_WINDOW, E_CUDA, E_NOMEM, E_DEGENERATE, E_STATE = -1, -2, -3, -4, -5, -6

EXPORTED_SYMBOLS = (
    "tcw_abi_version", "tcw_create", "tcw_destroy", "tcw_last_error", "tcw_device_name",
    "tcw_map_dims", "tcw_map_batch", "tcw_map_batch_windows", "tcw_submit", "tcw_wait", "tcw_upload_atoms", "tcw_map_resident", "tcw_fetch_results",
    "tcw_fetch_fmn", "tcw_fetch_merged", "tcw_synchronize", "tcw_timer_start", "tcw_timer_stop",
    "tcw_last_stage_ms", "tcw_last_exp_stage_ms", "tcw_launch_count", "tcw_flush_l2", "tcw_microbench", "tcw_microbench_ffma2",
    "tcw_host_alloc",
    "tcw_host_free", "tcw_cell_index_range", "tcw_set_exp_lut", "tcw_get_exp_lut",
    "tcw_device_count", "tcw_device_name_of", "tcw_results_device",
)


class CWindowRange(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("type", "t0", "t0Band", "dt0", "tau", "tauBand", "dtau")]


class CResult(C.Structure):
    _fields_ = [
        ("lnBtSG", C.c_double),
        ("t0_MP", C.c_double),
        ("tau_MP", C.c_double),
        ("maxF", C.c_float),
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
        ("path", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


RESULT_DTYPE = np.dtype(
    [
        ("lnBtSG", "<f8"), ("t0_MP", "<f8"), ("tau_MP", "<f8"), ("maxF", "<f4"),
        ("m_ML", "<u4"), ("n_ML", "<u4"), ("t0_ML", "<u4"), ("tau_ML", "<u4"),
        ("m_MP", "<u4"), ("n_MP", "<u4"), ("N_t0", "<u4"), ("N_tau", "<u4"),
        ("numAtoms", "<u4"), ("t0_data", "<u4"), ("status", "<i4"), ("path", "<u4"),
        ("reserved", "<u4"),
    ],
    align=True,
)
assert RESULT_DTYPE.itemsize == C.sizeof(CResult), (RESULT_DTYPE.itemsize, C.sizeof(CResult))


class TcwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[tcw_b200 error {code}] {msg}")
        self.code = code


class DegenerateWindowError(TcwError, ValueError):
    """A (t0,tau) cell has a single-atom window; lalpulsar aborts the map there (XLAL_EDOM)."""


_lib = None


def load_library(build_if_missing: bool = True):
    """Load the CUDA library, building it in-tree when missing and nvcc is available.
    Raises if that fails -- by design nothing else can compute a map."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _build

        if not os.environ.get("PYFSTAT_B200_LIB") and _build.needs_build() and _build.find_nvcc() is not None:
            _build.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing and could not be built (nvcc not found). "
            "pyfstat_b200 has no CPU fallback: build it with `python -m pyfstat_b200.build`."
        )
    L = C.CDLL(LIB_PATH)
    L.tcw_abi_version.restype = C.c_int
    if L.tcw_abi_version() != TCW_ABI_VERSION:
        raise ImportError("libtcw_b200.so ABI version mismatch; rebuild with pyfstat_b200.build")
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    L.tcw_create.argtypes = [i32, C.POINTER(vp)]
    L.tcw_destroy.argtypes = [vp]
    L.tcw_last_error.argtypes = [vp]
    L.tcw_last_error.restype = C.c_char_p
    L.tcw_device_name.argtypes = [vp, C.c_char_p, i32]
    L.tcw_map_dims.argtypes = [C.POINTER(CWindowRange), C.POINTER(u32), C.POINTER(u32)]
    L.tcw_map_batch.argtypes = [vp, vp, vp, u32, u32, i32, i32, C.POINTER(CWindowRange), u32, vp, vp]
    L.tcw_map_batch_windows.argtypes = [vp, vp, vp, u32, u32, i32, i32, vp, u32, vp, vp]
    L.tcw_upload_atoms.argtypes = [vp, vp, vp, u32, u32, i32, i32]
    L.tcw_map_resident.argtypes = [vp, C.POINTER(CWindowRange), u32]
    L.tcw_submit.argtypes = [vp, vp, vp, u32, u32, i32, i32, C.POINTER(CWindowRange), u32]
    L.tcw_wait.argtypes = [vp, vp, vp]
    L.tcw_fetch_results.argtypes = [vp, vp]
    L.tcw_fetch_fmn.argtypes = [vp, i32, vp]
    L.tcw_results_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.tcw_fetch_merged.argtypes = [vp, i32, vp, u32]
    L.tcw_synchronize.argtypes = [vp]
    L.tcw_timer_start.argtypes = [vp]
    L.tcw_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.tcw_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tcw_last_exp_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tcw_launch_count.argtypes = [vp]
    L.tcw_launch_count.restype = C.c_uint64
    L.tcw_flush_l2.argtypes = [vp]
    L.tcw_microbench.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tcw_microbench_ffma2.argtypes = [vp, C.POINTER(C.c_double)]
    L.tcw_host_alloc.argtypes = [C.c_uint64]
    L.tcw_host_alloc.restype = vp
    L.tcw_host_free.argtypes = [vp]
    L.tcw_host_free.restype = None
    L.tcw_cell_index_range.argtypes = [u32, u32, u32, u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    L.tcw_device_count.argtypes = []
    L.tcw_device_count.restype = i32
    L.tcw_device_name_of.argtypes = [i32, C.c_char_p, i32]
    L.tcw_set_exp_lut.argtypes = [vp, C.c_double, u32, vp]
    L.tcw_get_exp_lut.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u32), C.POINTER(i32)]
    _lib = L
    return L


def device_names(build_if_missing: bool = False):
    """Names of the CUDA devices, without opening a context or a handle (tcw:419-432)."""
    L = load_library(build_if_missing)
    out = []
    for d in range(L.tcw_device_count()):
        buf = C.create_string_buffer(256)
        if L.tcw_device_name_of(d, buf, 256) == 0:
            out.append(buf.value.decode())
    return out


def c_window(w) -> CWindowRange:
    w = TransientWindowRange.from_any(w)
    return CWindowRange(w.type, w.t0, w.t0Band, w.dt0, w.tau, w.tauBand, w.dtau)


def cell_index_range(window_type, t0_m, tau_n, t0_data, TAtom, numAtoms):
    """Host-only: (i_t0, i_t1) exactly as the kernels compute them."""
    L = load_library()
    a, b = C.c_uint32(), C.c_uint32()
    rc = L.tcw_cell_index_range(
        window_type, t0_m & 0xFFFFFFFF, tau_n & 0xFFFFFFFF, t0_data, TAtom, numAtoms, C.byref(a), C.byref(b)
    )
    if rc:
        raise TcwError(rc, "tcw_cell_index_range failed")
    return a.value, b.value


class PinnedBuffer:
    """cudaHostAlloc'ed buffer exposing the Python buffer protocol through a ctypes array.

    Lifetime: ``array`` (and therefore every ``np.frombuffer`` view of it, whose ``.base`` chain
    holds the ctypes array) keeps this object -- the owner of the allocation -- alive through
    ``array._owner``; ``cudaFreeHost`` runs only once the last view is gone."""

    def __init__(self, nbytes: int):
        L = load_library()
        self._ptr = L.tcw_host_alloc(nbytes)
        if not self._ptr:
            raise MemoryError(f"cudaHostAlloc({nbytes}) failed")
        self.nbytes = nbytes
        self.array = (C.c_ubyte * nbytes).from_address(self._ptr)
        self.array._owner = self  # the exported buffer owns the allocation (cycle: freed by the gc)

    def __del__(self):
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr and _lib is not None:
            _lib.tcw_host_free(ptr)


class DeviceBytes:
    """A span of device memory for zero-copy hand-over to torch (``__cuda_array_interface__`` v2)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self.nbytes = nbytes


class Handle:
    """RAII wrapper of ``tcw_handle`` (one per device and process is enough: calls are
    serialised on the handle's own stream)."""

    def __init__(self, device: int = -1):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.tcw_create(device, C.byref(h))
        if rc:
            raise TcwError(rc, (self.L.tcw_last_error(None) or b"").decode())
        self._h = h
        self._keep = None  # keeps the uploaded host arrays alive
        self._last_batch = None  # the batch whose atoms are resident on the device
        self.device_index = device
        self.generation = 0  # bumped by every map: tells whether the device F_mn is still a given map's

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.L.tcw_destroy(h)

    detach = close  # so the handle can stand in for a pycuda context (core.py:499-516)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow_degenerate_status: bool = False):
        if rc == 0:
            return
        msg = (self.L.tcw_last_error(self._h) or b"").decode()
        if rc == E_DEGENERATE:
            if allow_degenerate_status:
                return
            raise DegenerateWindowError(rc, msg)
        if rc == E_WINDOW:
            raise ValueError(msg)
        if rc == E_NOMEM:
            raise MemoryError(msg)
        raise TcwError(rc, msg)

    def set_exp_lut(self, xmax: float = EXPLUT_DEFAULT[0], length: int = EXPLUT_DEFAULT[1], table=None):
        """``tcw_set_exp_lut``: geometry (and optionally the ``length + 1`` entries) of the emulated
        XLALFastNegExp table."""
        tab = None
        if table is not None:
            tab = np.ascontiguousarray(table, dtype=np.float64)
            if tab.shape != (int(length) + 1,):
                raise ValueError("table needs length + 1 entries")
        self._check(self.L.tcw_set_exp_lut(self._h, float(xmax), int(length),
                                           tab.ctypes.data if tab is not None else None))

    def get_exp_lut(self):
        """``(xmax, length, canonical)`` of the table this handle emulates."""
        x, n, c = C.c_double(), C.c_uint32(), C.c_int()
        self._check(self.L.tcw_get_exp_lut(self._h, C.byref(x), C.byref(n), C.byref(c)))
        return x.value, n.value, bool(c.value)

    @property
    def device_name(self) -> str:
        buf = C.create_string_buffer(256)
        self._check(self.L.tcw_device_name(self._h, buf, 256))
        return buf.value.decode()

    # ---- one-shot host API -----------------------------------------------------------
    def map_batch(self, batch: AtomBatch, window, flags: int = 0, *, raise_on_degenerate: bool = True):
        """``tcw_map_batch``: returns ``(results: np.ndarray[RESULT_DTYPE], F_mn or None)``."""
        w = TransientWindowRange.from_any(window)
        w.check_type()
        cw = c_window(w)
        N_t0, N_tau = w.dims()
        results = np.zeros(batch.T, dtype=RESULT_DTYPE)
        F = None
        if flags & WANT_FMN:
            F = np.empty((batch.T, N_t0, N_tau), dtype=np.float32)
        rc = self.L.tcw_map_batch(
            self._h, batch.atoms.ctypes.data, batch.n_atoms.ctypes.data, batch.stride, batch.TAtom,
            batch.T, batch.numDet, C.byref(cw), flags, F.ctypes.data if F is not None else None,
            results.ctypes.data,
        )
        self._last_batch = None
        self.generation += 1
        self._check(rc, allow_degenerate_status=not raise_on_degenerate)
        self._T, self._last_batch = batch.T, batch  # the atoms stay resident: see batch.map_again
        return results, F

    def map_batch_windows(self, batch: AtomBatch, windows, flags: int = 0, *, raise_on_degenerate: bool = True):
        """``tcw_map_batch_windows``: one window range per template (same type and map shape).

        ``windows``: a sequence of window-range objects, or -- the fast path for a sampler step,
        no per-object Python work -- a ``(T, 7)`` uint32 array with the columns of
        ``transientWindowRange_t`` (type, t0, t0Band, dt0, tau, tauBand, dtau)."""
        if isinstance(windows, np.ndarray):
            wa = np.ascontiguousarray(windows, dtype=np.uint32)
            if wa.shape != (batch.T, 7):
                raise ValueError("need one window range (7 uint32 columns) per template")
            if (wa[:, 0] >= 3).any():
                TransientWindowRange(int(wa[:, 0].max())).check_type()  # raises like tcw:691-697
            w0 = TransientWindowRange(*(int(v) for v in wa[0]))
            N_t0, N_tau = w0.dims()
            cws = wa.ctypes.data_as(C.c_void_p)
        else:
            ws = [TransientWindowRange.from_any(w) for w in windows]
            if len(ws) != batch.T:
                raise ValueError("need one window range per template")
            for w in ws:
                w.check_type()
            N_t0, N_tau = ws[0].dims()
            cws = C.cast((CWindowRange * batch.T)(
                *[CWindowRange(w.type, w.t0, w.t0Band, w.dt0, w.tau, w.tauBand, w.dtau) for w in ws]), C.c_void_p)
        results = np.zeros(batch.T, dtype=RESULT_DTYPE)
        F = np.empty((batch.T, N_t0, N_tau), dtype=np.float32) if flags & WANT_FMN else None
        self._last_batch = None
        self.generation += 1
        rc = self.L.tcw_map_batch_windows(
            self._h, batch.atoms.ctypes.data, batch.n_atoms.ctypes.data, batch.stride, batch.TAtom,
            batch.T, batch.numDet, cws, flags, F.ctypes.data if F is not None else None,
            results.ctypes.data,
        )
        self._check(rc, allow_degenerate_status=not raise_on_degenerate)
        self._T, self._last_batch = batch.T, batch
        return results, F

    # ---- asynchronous split of map_batch ---------------------------------------------
    def submit(self, batch: AtomBatch, window, flags: int = 0):
        """``tcw_submit``: enqueue copies + kernels for ``batch`` and return at once; the host is
        free (e.g. to produce the next batch's atoms) until :meth:`wait`.  One batch in flight."""
        w = TransientWindowRange.from_any(window)
        w.check_type()
        cw = c_window(w)
        self._last_batch = None
        self.generation += 1
        self._check(self.L.tcw_submit(
            self._h, batch.atoms.ctypes.data, batch.n_atoms.ctypes.data, batch.stride, batch.TAtom,
            batch.T, batch.numDet, C.byref(cw), flags))
        self._pending = (batch, w, flags)  # keeps the host buffers alive until wait()

    def wait(self, *, raise_on_degenerate: bool = True):
        """``tcw_wait``: ``(results, F_mn or None)`` of the submitted batch."""
        if getattr(self, "_pending", None) is None:
            raise TcwError(E_STATE, "wait() without a submitted batch")
        batch, w, flags = self._pending
        self._pending = None
        N_t0, N_tau = w.dims()
        results = np.zeros(batch.T, dtype=RESULT_DTYPE)
        F = np.empty((batch.T, N_t0, N_tau), dtype=np.float32) if flags & WANT_FMN else None
        rc = self.L.tcw_wait(self._h, F.ctypes.data if F is not None else None, results.ctypes.data)
        self._check(rc, allow_degenerate_status=not raise_on_degenerate)
        self._T, self._last_batch = batch.T, batch
        return results, F

    # ---- resident API ----------------------------------------------------------------
    def upload(self, batch: AtomBatch):
        self._keep = batch
        self._last_batch = None
        self.generation += 1
        self._check(
            self.L.tcw_upload_atoms(
                self._h, batch.atoms.ctypes.data, batch.n_atoms.ctypes.data, batch.stride, batch.TAtom,
                batch.T, batch.numDet,
            )
        )
        self._T, self._last_batch = batch.T, batch

    def map_resident(self, window, flags: int = 0):
        cw = c_window(window)
        self.generation += 1
        self._check(self.L.tcw_map_resident(self._h, C.byref(cw), flags))

    def fetch_results(self, raise_on_degenerate: bool = True) -> np.ndarray:
        results = np.zeros(self._T, dtype=RESULT_DTYPE)
        self._check(self.L.tcw_fetch_results(self._h, results.ctypes.data), not raise_on_degenerate)
        return results

    def results_device(self):
        """The last map's result records where they lie in device memory, as an object exposing
        ``__cuda_array_interface__`` (uint8, ``T * itemsize`` bytes): ``torch.as_tensor(view, device="cuda")``
        wraps it without a copy.  Valid until the next map / upload call on this handle."""
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.tcw_results_device(self._h, C.byref(p), C.byref(n)))
        return DeviceBytes(p.value, int(n.value) * RESULT_DTYPE.itemsize)

    def fetch_fmn(self, t: int, N_t0: int, N_tau: int) -> np.ndarray:
        F = np.empty((N_t0, N_tau), dtype=np.float32)
        self._check(self.L.tcw_fetch_fmn(self._h, t, F.ctypes.data))
        return F

    def fetch_merged(self, t: int, numAtoms: int) -> np.ndarray:
        out = np.empty((7, numAtoms), dtype=np.float32)
        self._check(self.L.tcw_fetch_merged(self._h, t, out.ctypes.data, numAtoms))
        return out

    def synchronize(self):
        self._check(self.L.tcw_synchronize(self._h))

    def timer_start(self):
        self._check(self.L.tcw_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self.L.tcw_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def last_stage_ms(self):
        ms = (C.c_float * 5)()
        self._check(self.L.tcw_last_stage_ms(self._h, ms))
        return dict(zip(("prep", "table", "map", "btsg", "finalize"), (float(x) for x in ms)))

    def last_exp_stage_ms(self):
        """The map stage of the last call split for the exponential window's recurrence path: operand
        preparation, tensor-core pass, walk (all zero when another path ran)."""
        ms = (C.c_float * 3)()
        self._check(self.L.tcw_last_exp_stage_ms(self._h, ms))
        return dict(zip(("operands", "tensor", "walk"), (float(x) for x in ms)))

    @property
    def launch_count(self) -> int:
        return int(self.L.tcw_launch_count(self._h))

    def flush_l2(self):
        self._check(self.L.tcw_flush_l2(self._h))

    def microbench(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.L.tcw_microbench(self._h, C.byref(a), C.byref(b)))
        c = C.c_double()
        self._check(self.L.tcw_microbench_ffma2(self._h, C.byref(c)))
        return {"ffma_tflops": a.value, "dadd_tflops": b.value, "ffma2_tflops": c.value}


def pinned_atoms_alloc():
    """Allocator for :func:`pyfstat_b200.atoms.synth_atoms` placing the batch in pinned memory.
    The returned buffers own their allocation (see :class:`PinnedBuffer`): dropping the allocator
    does not free memory that an :class:`AtomBatch` still views."""

    def alloc(nbytes):
        return PinnedBuffer(nbytes).array

    return alloc
