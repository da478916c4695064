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

from . import _lib
from .atoms import AtomBatch, from_multi_fstat_atoms
from .window import TRANSIENT_NONE, TransientWindowRange

logger = logging.getLogger(__name__)

BACKEND_NAME = "b200"

_handles = {}
_handles_lock = threading.Lock()


def get_handle(device: int = -1) -> "_lib.Handle":
    """Process-global handle per device (calls are serialised on its stream)."""
    with _handles_lock:
        h = _handles.get(device)
        if h is None or h._h is None:
            h = _lib.Handle(device)
            _handles[device] = h
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
    ``PYFSTAT_B200_GENERIC=1`` forces the generic (bit-faithful) kernels."""
    flags = 0
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
        """

        def __init__(self, record, batch, window, flags, handle_device=-1):
            # deliberately NOT calling base.__init__: it would allocate a dense F_mn
            self._rec = record
            self._batch = batch
            self._window = window
            self._flags = flags & ~(_lib.WANT_FMN | _lib.WANT_BTSG)
            self._device = handle_device
            self._F_mn = None
            self.maxF = float(record["maxF"])
            self.t0_ML = int(record["t0_ML"])
            self.tau_ML = int(record["tau_ML"])
            self.lnBtSG = float(record["lnBtSG"])
            self.t0_MP = float(record["t0_MP"])
            self.tau_MP = float(record["tau_MP"])
            self._have_btsg = not np.isnan(self.lnBtSG)

        # ---- lazy F_mn --------------------------------------------------------------
        @property
        def F_mn(self):
            """The ``[N_t0, N_tau]`` float32 array once materialised; before that a
            :class:`LazyFmn` that serves single-cell reads without materialising."""
            return self._F_mn if self._F_mn is not None else LazyFmn(self)

        def _materialise(self):
            if self._F_mn is None:
                h = get_handle(self._device)
                _, F = h.map_batch(
                    self._batch, self._window, self._flags | _lib.WANT_FMN, raise_on_degenerate=False
                )
                self._F_mn = F[0]
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
            h = get_handle(self._device)
            res, _ = h.map_batch(self._batch, one, self._flags | _lib.ALLOW_DEGENERATE)
            return float(res["maxF"][0])

        # ---- fused reductions (override the numpy versions, tcw:186-287) ------------
        def get_maxF_idx(self):
            return int(self._rec["m_ML"]), int(self._rec["n_ML"])

        def _ensure_btsg(self):
            if not self._have_btsg:
                h = get_handle(self._device)
                res, _ = h.map_batch(
                    self._batch, self._window, self._flags | _lib.WANT_BTSG, raise_on_degenerate=False
                )
                self._rec = res[0]
                self.lnBtSG = float(res["lnBtSG"][0])
                self._t0_MP_gpu = float(res["t0_MP"][0])
                self._tau_MP_gpu = float(res["tau_MP"][0])
                self._have_btsg = True
            else:
                self._t0_MP_gpu = float(self._rec["t0_MP"])
                self._tau_MP_gpu = float(self._rec["tau_MP"])

        def get_lnBtSG(self):
            self._ensure_btsg()
            self.lnBtSG = float(self._rec["lnBtSG"])
            return self.lnBtSG

        def get_t0_max_posterior(self, windowRange):
            self._ensure_btsg()
            N_t0 = int(self._rec["N_t0"])
            dx = windowRange.t0Band / N_t0  # consistent with LAL (tcw:248-251)
            self.t0_MP = windowRange.t0 + (int(self._rec["m_MP"]) + 0.5) * dx
            return self.t0_MP

        def get_tau_max_posterior(self, windowRange):
            self._ensure_btsg()
            N_tau = int(self._rec["N_tau"])
            dy = windowRange.tauBand / N_tau  # tcw:283-286
            self.tau_MP = windowRange.tau + (int(self._rec["n_MP"]) + 0.5) * dy
            return self.tau_MP

        if base is object:
            # PyFstat absent: provide the text writer with the reference's format
            # (tcw:289-317: columns t0[s] tau[s] 2F, "  %10d %10d %- 11.8g")
            def write_F_mn_to_file(self, tCWfile, windowRange, header=[]):
                with open(tCWfile, "w") as tfp:
                    for hline in header:
                        tfp.write("# {:s}\n".format(hline))
                    tfp.write("# t0[s]     tau[s]     2F\n")
                    for m, F_m in enumerate(self.F_mn):
                        this_t0 = windowRange.t0 + m * windowRange.dt0
                        for n, this_F in enumerate(F_m):
                            this_tau = windowRange.tau + n * windowRange.dtau
                            tfp.write("  %10d %10d %- 11.8g\n" % (this_t0, this_tau, 2.0 * this_F))

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
    if window.type == TRANSIENT_NONE:  # describe the substituted window for lazy F_mn / MP
        window = TransientWindowRange(
            1, int(res["t0_data"][0]), 0, batch.TAtom, int(res["numAtoms"][0]) * batch.TAtom, 0, batch.TAtom
        )
    cls = fstat_map_class(tcw_module)
    return cls(res[0], batch, window, flags, device)


def backend_available() -> bool:
    """True if the CUDA library loads and a device can be opened."""
    try:
        get_handle(-1)
        return True
    except Exception as e:  # noqa: BLE001
        logger.debug("b200 backend unavailable: %s", e)
        return False


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
        features[name] = backend_available()
        return features

    def init_transient_fstat_map_features(feature="lal", cudaDeviceName=None):
        if feature != name:
            features, ctx = orig_init(feature, cudaDeviceName)
            features[name] = backend_available()
            return features, ctx
        features = _get_transient_fstat_map_features()
        if not features[name]:
            raise RuntimeError(f"{name} use was requested, but no CUDA device / library is usable.")
        h = get_handle(-1)
        devname = h.device_name.replace(" ", "-").replace("_", "-")
        if cudaDeviceName and cudaDeviceName not in devname:  # partial match as tcw:440-454
            raise RuntimeError(
                'Requested CUDA device "{}" not found. Available devices: [{}]'.format(cudaDeviceName, devname)
            )
        logger.info("Transient F-stat maps on CUDA device %s (backend %r).", devname, name)
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
