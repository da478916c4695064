"""The drop-in ``tCWFstatMapVersion`` backend: registration into PyFstat's own registry and
the ``pyTransientFstatMap``-compatible result type.

Boundary (all citations into the PyFstat tree, ``tcw`` = pyfstat/tcw_fstat_map_funcs.py):

* registry ``tcw.fstatmap_versions[name] = callable(multiFstatAtoms, windowRange, BtSG)``
  (tcw:320-327, looked up at tcw:530-533);
* feature gating: ``features[name]`` must be truthy (tcw:531-543) and
  ``init_transient_fstat_map_features(name, cudaDeviceName)`` must accept the name -- the
  stock function raises ``ValueError`` for anything but "lal"/"pycuda" (tcw:485-490) and is
  called through the module attribute by ``ComputeFstat`` (pyfstat/core.py:893-899), so
  :func:`register` wraps both module attributes;
* context lifetime: ``ComputeFstat`` only installs a finalizer calling
  ``gpu_context.detach()`` when "cuda" is in the version name (core.py:493-505).  The default
  name ``"b200"`` has no "cuda" in it: ``gpu_context`` is ``None`` and the device handle is a
  process-global closed at exit.

No CPU fallback: every map is computed by the CUDA library; if it cannot be loaded or no
device is usable, ``features[name]`` is False and calls raise.
"""

from __future__ import annotations

import atexit
import logging
import os
import sys
import threading

import numpy as np

from . import _lib, lut_probe
from .atoms import AtomBatch, from_multi_fstat_atoms
from .window import TRANSIENT_NONE, TransientWindowRange

logger = logging.getLogger(__name__)

BACKEND_NAME = "b200"

_handles = {}
_handles_lock = threading.Lock()
_probed_lut = None  # (xmax, length, table) measured from lalpulsar, False if unavailable


def resolve_device(device: int = -1) -> int:
    """``-1`` -> ``$CUDA_DEVICE`` (as the reference, tcw:434-437, 466-469), default 0."""
    if device is None or device < 0:
        return int(os.environ.get("CUDA_DEVICE", "0"))
    return int(device)


def exp_lut_geometry():
    """Geometry ``(xmax, length, table | None, source)`` of the XLALFastNegExp table new handles
    emulate, in this order: ``$PYFSTAT_B200_EXPLUT="xmax:length"``; measured from lalpulsar
    itself when it is importable (:mod:`pyfstat_b200.lut_probe`); the library default
    (``_lib.EXPLUT_DEFAULT``)."""
    global _probed_lut
    env = os.environ.get("PYFSTAT_B200_EXPLUT")
    if env:
        xmax, length = lut_probe.parse_geometry(env)
        return xmax, length, None, "$PYFSTAT_B200_EXPLUT"
    if _probed_lut is None:
        _probed_lut = lut_probe.probe_lalpulsar() or False
        if _probed_lut:
            logger.info("XLALFastNegExp table measured from lalpulsar: xmax = %g, %d steps",
                        _probed_lut[0], _probed_lut[1])
    if _probed_lut:
        return _probed_lut[0], _probed_lut[1], _probed_lut[2], "measured from lalpulsar.FastNegExp"
    return _lib.EXPLUT_DEFAULT[0], _lib.EXPLUT_DEFAULT[1], None, "library default (SURVEY A.4-1)"


def get_handle(device: int = -1) -> "_lib.Handle":
    """Process-global handle per device index (calls are serialised on its stream)."""
    idx = resolve_device(device)
    with _handles_lock:
        h = _handles.get(idx)
        if h is None or h._h is None:
            h = _lib.Handle(idx)
            xmax, length, table, _src = exp_lut_geometry()
            if table is not None or (xmax, length) != h.get_exp_lut()[:2]:
                h.set_exp_lut(xmax, length, table)
            _handles[idx] = h
        return h


@atexit.register
def _close_handles():
    for h in list(_handles.values()):
        try:
            h.close()
        except Exception:  # pragma: no cover
            pass
    _handles.clear()


def default_flags() -> int:
    """Backend knobs the caller API has no slot for come from the environment:
    ``PYFSTAT_B200_EXP=exact`` selects exact exp() instead of lalpulsar's lookup table;
    ``PYFSTAT_B200_ALLOW_DEGENERATE=1`` gives pycuda semantics for single-atom windows;
    ``PYFSTAT_B200_GENERIC=1`` forces the generic (bit-faithful) kernels;
    ``PYFSTAT_B200_EXP_DIRECT=1`` keeps the exponential window on the tiled direct sum (no recurrence /
    tensor-core pass)."""
    flags = 0
    if os.environ.get("PYFSTAT_B200_EXP_DIRECT", "0") not in ("0", ""):
        flags |= _lib.EXP_DIRECT
    if os.environ.get("PYFSTAT_B200_EXP", "lal").lower() == "exact":
        flags |= _lib.EXP_EXACT
    if os.environ.get("PYFSTAT_B200_ALLOW_DEGENERATE", "0") not in ("0", ""):
        flags |= _lib.ALLOW_DEGENERATE
    if os.environ.get("PYFSTAT_B200_GENERIC", "0") not in ("0", ""):
        flags |= _lib.FORCE_GENERIC
    return flags


def _base_class(tcw_module=None):
    """``pyTransientFstatMap`` of the given (or an already imported) PyFstat module, so
    ``isinstance(result, pyTransientFstatMap)`` holds; ``object`` if PyFstat is absent."""
    if tcw_module is None:
        tcw_module = sys.modules.get("pyfstat.tcw_fstat_map_funcs")
    return getattr(tcw_module, "pyTransientFstatMap", object) if tcw_module else object


class LazyFmn:
    """What ``FstatMap.F_mn`` returns before the map has been copied to the host.

    ``F_mn[m, n]`` with two integers -- the one access pattern of the transient BSGL
    (``FXstatMap.F_mn[idx_maxTwoF]``, core.py:1541) -- evaluates that single cell on the GPU
    (a 1x1 map) instead of computing, copying and caching the whole ``[N_t0, N_tau]`` array.
    Any other use (slicing, ``np.asarray``, iteration, arithmetic, attributes such as ``.max()``)
    materialises the array once and behaves like it.
    """

    def __init__(self, owner):
        self._owner = owner

    def _array(self):
        return self._owner._materialise()

    @property
    def shape(self):
        return self._owner.shape

    @property
    def dtype(self):
        return np.dtype(np.float32)

    @property
    def ndim(self):
        return 2

    def __len__(self):
        return self._owner.shape[0]

    def __getitem__(self, idx):
        if (
            isinstance(idx, tuple)
            and len(idx) == 2
            and all(isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)) for i in idx)
            and self._owner._F_mn is None
        ):
            N_t0, N_tau = self._owner.shape
            m, n = int(idx[0]), int(idx[1])
            if not (-N_t0 <= m < N_t0 and -N_tau <= n < N_tau):
                raise IndexError(f"index {idx} is out of bounds for F_mn of shape {(N_t0, N_tau)}")
            return np.float32(self._owner.F_at(m % N_t0, n % N_tau))
        return self._array()[idx]

    def __array__(self, dtype=None, copy=None):
        a = self._array()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __iter__(self):
        return iter(self._array())

    def __getattr__(self, name):  # everything else: behave like the ndarray
        return getattr(self._array(), name)

    def __repr__(self):
        return f"LazyFmn(shape={self.shape}, materialised={self._owner._F_mn is not None})"


def _lazy_binop(name):
    def op(self, *args):
        return getattr(self._array(), name)(*args)

    op.__name__ = name
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__neg__", "__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__"):
    setattr(LazyFmn, _n, _lazy_binop(_n))
LazyFmn.__hash__ = None


_class_cache = {}


def fstat_map_class(tcw_module=None):
    """Result type: subclass of the reference's ``pyTransientFstatMap`` (tcw:50-317) whose
    reductions come fused from the GPU and whose ``F_mn`` is materialised lazily."""
    base = _base_class(tcw_module)
    cls = _class_cache.get(base)
    if cls is not None:
        return cls

    class B200TransientFstatMap(base):  # type: ignore[misc,valid-type]
        """F(t0,tau) map computed on a B200.

        Same attributes as ``pyTransientFstatMap`` -- ``F_mn`` (F, not 2F; ``[m over t0, n
        over tau]``), ``maxF``, ``t0_ML``, ``tau_ML``, ``lnBtSG``, ``t0_MP``, ``tau_MP``
        (``nan`` unless computed with ``BtSG=True``, tcw:142-144).  ``F_mn`` is only computed,
        copied to the host and cached when it is first read; ``get_maxF_idx()`` and the
        ``get_*`` estimators return the fused GPU results without touching it.

        Nothing is computed twice: whichever of ``F_mn`` / ``get_lnBtSG()`` is asked for first
        runs ONE pass that produces both (``F_mn`` stays on the device until it is read), on the
        atoms still resident from the original call when no other map ran in between.
        """

        def __init__(self, record, batch, window, flags, handle_device=-1):
            # deliberately NOT calling base.__init__: it would allocate a dense F_mn
            self._rec = record
            self._batch = batch
            self._window = window
            self._flags = flags & ~(_lib.WANT_FMN | _lib.WANT_BTSG)
            self._device = handle_device
            self._F_mn = None
            self._fmn_gen = None  # handle generation whose device F_mn is this map's
            self.maxF = float(record["maxF"])
            self.t0_ML = int(record["t0_ML"])
            self.tau_ML = int(record["tau_ML"])
            self.lnBtSG = float(record["lnBtSG"])
            self.t0_MP = float(record["t0_MP"])
            self.tau_MP = float(record["tau_MP"])
            self._have_btsg = not np.isnan(self.lnBtSG)

        # ---- one place that talks to the GPU -----------------------------------------
        def _run(self, window, flags):
            """Records of ``window`` over this map's atoms: on the atoms still resident on the
            device when the handle's last batch is ours (no host-to-device copy), else through
            ``tcw_map_batch``.  ``WANT_FMN`` leaves F_mn on the device (see ``_materialise``)."""
            h = get_handle(self._device)
            if getattr(h, "_last_batch", None) is self._batch:
                h.map_resident(window, flags)
                res = h.fetch_results(raise_on_degenerate=False)
            else:
                h.upload(self._batch)
                h.map_resident(window, flags)
                res = h.fetch_results(raise_on_degenerate=False)
            return h, res

        def _full_pass(self):
            """F_mn (left on the device) and the lnBtSG record in ONE pass."""
            h, res = self._run(self._window, self._flags | _lib.WANT_FMN | _lib.WANT_BTSG)
            self._fmn_gen = h.generation
            self._rec = res[0]
            self.lnBtSG = float(res["lnBtSG"][0])
            self._have_btsg = True
            return h

        # ---- lazy F_mn --------------------------------------------------------------
        @property
        def F_mn(self):
            """The ``[N_t0, N_tau]`` float32 array once materialised; before that a
            :class:`LazyFmn` that serves single-cell reads without materialising."""
            return self._F_mn if self._F_mn is not None else LazyFmn(self)

        def _materialise(self):
            if self._F_mn is None:
                h = get_handle(self._device)
                if self._fmn_gen is None or h.generation != self._fmn_gen:
                    h = self._full_pass()
                self._F_mn = h.fetch_fmn(0, *self.shape)
            return self._F_mn

        @F_mn.setter
        def F_mn(self, value):
            self._F_mn = value

        @property
        def shape(self):
            return int(self._rec["N_t0"]), int(self._rec["N_tau"])

        def F_at(self, m: int, n: int) -> float:
            """Single cell without materialising the map: a 1x1 map at (t0_m, tau_n)."""
            if self._F_mn is not None:
                return float(self._F_mn[m, n])
            w = self._window
            one = TransientWindowRange(w.type, w.t0 + m * w.dt0, 0, w.dt0, w.tau + n * w.dtau, 0, w.dtau)
            _, res = self._run(one, self._flags | _lib.ALLOW_DEGENERATE)
            return float(res["maxF"][0])

        # ---- fused reductions (override the numpy versions, tcw:186-287) ------------
        def get_maxF_idx(self):
            return int(self._rec["m_ML"]), int(self._rec["n_ML"])

        def _ensure_btsg(self):
            if not self._have_btsg:
                self._full_pass()

        def get_lnBtSG(self):
            self._ensure_btsg()
            self.lnBtSG = float(self._rec["lnBtSG"])
            return self.lnBtSG

        def _mp_window(self, windowRange):
            # TRANSIENT_NONE: the reference's pycuda path mutates the caller's range into the rect
            # window spanning the data (tcw:742-749); ours is never mutated, so substitute here
            return self._window if int(windowRange.type) == TRANSIENT_NONE else windowRange

        def get_t0_max_posterior(self, windowRange):
            self._ensure_btsg()
            windowRange = self._mp_window(windowRange)
            N_t0 = int(self._rec["N_t0"])
            dx = windowRange.t0Band / N_t0  # consistent with LAL (tcw:248-251)
            self.t0_MP = windowRange.t0 + (int(self._rec["m_MP"]) + 0.5) * dx
            return self.t0_MP

        def get_tau_max_posterior(self, windowRange):
            self._ensure_btsg()
            windowRange = self._mp_window(windowRange)
            N_tau = int(self._rec["N_tau"])
            dy = windowRange.tauBand / N_tau  # tcw:283-286
            self.tau_MP = windowRange.tau + (int(self._rec["n_MP"]) + 0.5) * dy
            return self.tau_MP

        if base is object:
            # PyFstat absent: the text format of tcw:289-317 / 159-184 -- optional "# " header
            # lines, a column line, then one row "t0 tau 2F" per cell, t0 outer / tau inner
            def write_F_mn_to_file(self, tCWfile, windowRange, header=[]):
                F = np.asarray(self.F_mn, dtype=np.float64)
                windowRange = self._mp_window(windowRange)
                t0s = windowRange.t0 + windowRange.dt0 * np.arange(F.shape[0], dtype=np.int64)
                taus = windowRange.tau + windowRange.dtau * np.arange(F.shape[1], dtype=np.int64)
                table = np.column_stack([np.repeat(t0s, F.shape[1]), np.tile(taus, F.shape[0]), 2.0 * F.ravel()])
                head = "\n".join([str(line) for line in header] + ["t0[s]     tau[s]     2F"])
                np.savetxt(tCWfile, table, fmt="  %10d %10d %- 11.8g", header=head, comments="# ")

            @classmethod
            def read_from_file(cls, tCWfile):
                """Map object holding the F_mn of a file written by ``write_F_mn_to_file`` (2F is
                halved back to F; rows and columns from the distinct t0 / tau values)."""
                cols = np.loadtxt(tCWfile, comments="#", ndmin=2)
                t0s, taus = np.unique(cols[:, 0]), np.unique(cols[:, 1])
                self = cls.__new__(cls)
                self._F_mn = (0.5 * cols[:, 2]).reshape(len(t0s), len(taus))
                idx = np.unravel_index(np.argmax(self._F_mn), self._F_mn.shape)
                self.maxF = float(self._F_mn[idx])
                self.t0_ML, self.tau_ML = float(t0s[idx[0]]), float(taus[idx[1]])
                self.lnBtSG = self.t0_MP = self.tau_MP = float("nan")
                self._have_btsg, self._fmn_gen = False, None
                self._rec = {"N_t0": len(t0s), "N_tau": len(taus), "m_ML": idx[0], "n_ML": idx[1]}
                return self

    B200TransientFstatMap.__qualname__ = "B200TransientFstatMap"
    _class_cache[base] = B200TransientFstatMap
    return B200TransientFstatMap


def b200_compute_transient_fstat_map(multiFstatAtoms, windowRange, BtSG=False, *, tcw_module=None,
                                     device: int = -1, flags: int | None = None):
    """The registered callable: same signature and semantics as
    ``lalpulsar_compute_transient_fstat_map`` / ``pycuda_compute_transient_fstat_map``
    (tcw:547-587, 656-834).

    ``multiFstatAtoms``: ``lalpulsar.MultiFstatAtomVector`` (duck-typed) or an
    :class:`~pyfstat_b200.atoms.AtomBatch` with one template.  ``windowRange``:
    ``lalpulsar.transientWindowRange_t`` (duck-typed); never mutated.  Raises ``ValueError``
    for an unknown window type (tcw:691-697) and -- following lalpulsar, not pycuda -- for a
    degenerate single-atom window.
    """
    window = TransientWindowRange.from_any(windowRange)
    window.check_type()
    batch = from_multi_fstat_atoms(multiFstatAtoms)
    if batch.T != 1:
        raise ValueError("the registered backend computes one template per call; use map_batch()")
    if flags is None:
        flags = default_flags()
    h = get_handle(device)
    res, _ = h.map_batch(batch, window, flags | (_lib.WANT_BTSG if BtSG else 0))
    device = h.device_index
    if window.type == TRANSIENT_NONE:  # describe the substituted window for lazy F_mn / MP
        window = TransientWindowRange(
            1, int(res["t0_data"][0]), 0, batch.TAtom, int(res["numAtoms"][0]) * batch.TAtom, 0, batch.TAtom
        )
    cls = fstat_map_class(tcw_module)
    return cls(res[0], batch, window, flags, device)


def backend_available() -> bool:
    """True if the CUDA library loads and a device can be opened (creates the handle)."""
    try:
        get_handle(-1)
        return True
    except Exception as e:  # noqa: BLE001
        logger.debug("b200 backend unavailable: %s", e)
        return False


def backend_present() -> bool:
    """Cheap, side-effect-free feature probe: the library file exists (no nvcc build is
    attempted) and the driver reports a CUDA device.  No context, stream or handle is created,
    so a process that only ever uses ``lal`` / ``pycuda`` pays nothing for this backend."""
    try:
        if not os.path.exists(_lib.LIB_PATH):
            return False
        return _lib.load_library(build_if_missing=False).tcw_device_count() > 0
    except Exception as e:  # noqa: BLE001
        logger.debug("b200 backend not present: %s", e)
        return False


def select_device(cudaDeviceName=None) -> int:
    """Device choice of ``init_transient_fstat_map_features`` (tcw:419-476): all devices are
    enumerated; ``cudaDeviceName`` is matched PARTIALLY against their names (spaces and
    underscores as dashes), first match wins (warning if several) and is written to
    ``$CUDA_DEVICE``; otherwise ``$CUDA_DEVICE`` or 0.  Raises the reference's RuntimeErrors."""
    names = [n.replace(" ", "-").replace("_", "-") for n in _lib.device_names()]
    logger.info("Found %d CUDA device(s).", len(names))
    for n, name in enumerate(names):
        logger.info("device %d: model: %s", n, name)
    devnum = int(os.environ.get("CUDA_DEVICE", "0"))
    if cudaDeviceName:
        matches = [i for i, name in enumerate(names) if cudaDeviceName in name]
        if not matches:
            raise RuntimeError(
                'Requested CUDA device "{}" not found. Available devices: [{}]'.format(cudaDeviceName, ",".join(names))
            )
        devnum = matches[0]
        if len(matches) > 1:
            logger.warning('Found %d CUDA devices matching name "%s". Choosing first one with index %d.',
                           len(matches), cudaDeviceName, devnum)
        os.environ["CUDA_DEVICE"] = str(devnum)
    if devnum >= len(names):
        raise RuntimeError(
            "Requested CUDA device number {} exceeds number of available devices!"
            " Please change through environment variable $CUDA_DEVICE.".format(devnum)
        )
    logger.info("Choosing CUDA device %d, of %d devices present: %s", devnum, len(names), names[devnum])
    return devnum


def register(tcw_module=None, name: str = BACKEND_NAME):
    """Register the backend in PyFstat's registry under ``name`` -- no PyFstat source change.

    ``tcw_module``: the ``pyfstat.tcw_fstat_map_funcs`` module object (default: import it).
    After this, ``ComputeFstat(tCWFstatMapVersion=name)``,
    ``TransientGridSearch(tCWFstatMapVersion=name)`` and
    ``MCMCTransientSearch(tCWFstatMapVersion=name)`` work unchanged.
    """
    if tcw_module is None:
        import importlib

        tcw_module = importlib.import_module("pyfstat.tcw_fstat_map_funcs")
    if getattr(tcw_module, "_b200_registered", None) == name:
        return tcw_module
    if "cuda" in name:
        logger.warning(
            "backend name %r contains 'cuda': ComputeFstat will install a finalizer that calls "
            "gpu_context.detach() (core.py:493-505); the returned context supports that.", name
        )

    tcw_module.fstatmap_versions[name] = (
        lambda multiFstatAtoms, windowRange, BtSG: b200_compute_transient_fstat_map(
            multiFstatAtoms, windowRange, BtSG, tcw_module=tcw_module
        )
    )

    orig_features = tcw_module._get_transient_fstat_map_features
    orig_init = tcw_module.init_transient_fstat_map_features

    def _get_transient_fstat_map_features():
        features = orig_features()
        features[name] = backend_present()  # import-probing only, like the stock entries (tcw:351-358)
        return features

    def init_transient_fstat_map_features(feature="lal", cudaDeviceName=None):
        if feature != name:
            # other backends: report ours lazily -- no handle, no context, no build
            features, ctx = orig_init(feature, cudaDeviceName)
            features[name] = backend_present()
            return features, ctx
        features = _get_transient_fstat_map_features()
        if not features[name]:
            raise RuntimeError(f"{name} use was requested, but no CUDA device / library is usable.")
        devnum = select_device(cudaDeviceName)
        h = get_handle(devnum)  # opens the device (raises if it is not an sm_100 part)
        logger.info("Transient F-stat maps on CUDA device %d: %s (backend %r).", devnum, h.device_name, name)
        # the handle is process-global and closed at exit; a context object is only handed
        # out when the caller will detach() it ("cuda" in the name, core.py:493-505)
        return features, (_DetachableContext() if "cuda" in name else None)

    _get_transient_fstat_map_features._b200_wrapped = orig_features
    init_transient_fstat_map_features._b200_wrapped = orig_init
    tcw_module._get_transient_fstat_map_features = _get_transient_fstat_map_features
    tcw_module.init_transient_fstat_map_features = init_transient_fstat_map_features
    tcw_module._b200_registered = name
    return tcw_module


class _DetachableContext:
    """Stands in for a pycuda context where PyFstat wants to ``detach()`` one."""

    def detach(self):
        pass


def unregister(tcw_module, name: str = BACKEND_NAME):
    """Undo :func:`register` (used by the tests)."""
    if getattr(tcw_module, "_b200_registered", None) != name:
        return
    tcw_module.fstatmap_versions.pop(name, None)
    tcw_module._get_transient_fstat_map_features = tcw_module._get_transient_fstat_map_features._b200_wrapped
    tcw_module.init_transient_fstat_map_features = tcw_module.init_transient_fstat_map_features._b200_wrapped
    del tcw_module._b200_registered
