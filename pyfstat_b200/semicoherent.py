"""Callers of the transient map that BYPASS PyFstat's backend registry.

The reference computes three more quantities with ``lalpulsar.ComputeTransientFstatMap``
directly, i.e. not through ``tcw.call_compute_transient_fstat_map``:

* semi-coherent per-segment 2F: a rectangular map with ``N_t0 = nsegs`` rows and ONE column,
  ``t0 = tboundaries[0]``, ``dt0 = tau = Tcoh`` (``core.py:2005-2019``, read as
  ``2 * F_mn[:, 0]`` at ``core.py:2282-2289``);
* the same per detector for the semi-coherent BSGL (``core.py:2236-2260``: NaNs treated as 0);
* the per-detector 2F at the (t0, tau) cell that maximises the multi-detector F, for the
  transient BSGL (``core.py:1503-1545``: today one full map per detector, read at one index).

Plus ``calculate_twoF_cumulative`` (``core.py:1648-1665``), a Python loop over durations that
is one ``1 x N_tau`` rectangular map.  Here all of them are batched calls into the same C ABI
(``tcw_map_batch`` / ``tcw_map_batch_windows``), many templates at a time.
:func:`install` swaps the two ``SemiCoherentSearch`` methods for these, without a source change.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .atoms import AtomBatch, from_multi_fstat_atoms
from .backend import default_flags, get_handle
from .window import TRANSIENT_RECTANGULAR, TransientWindowRange


def semicoherent_window_range(tboundaries, Tcoh) -> TransientWindowRange:
    """``SemiCoherentSearch.init_semicoherent_parameters`` window (core.py:2005-2019):
    ``[t0, t0 + t0Band]`` in steps of ``Tcoh``, a single duration ``tau = Tcoh``."""
    return TransientWindowRange(
        TRANSIENT_RECTANGULAR,
        int(tboundaries[0]),
        int(tboundaries[-1] - tboundaries[0] - Tcoh),
        int(Tcoh),
        int(Tcoh),
        0,
        1,  # "Irrelevant" (core.py:2019), but must be != 0
    )


def single_detector_batch(batch: AtomBatch, X: int) -> AtomBatch:
    """Atoms of detector ``X`` only, every template (``extract_singleIFOmultiFatoms_from_multiAtoms``,
    utils/atoms.py:8-41, for a whole batch)."""
    if not 0 <= X < batch.numDet:
        raise ValueError(f"Detector index {X} is out of range for atoms of {batch.numDet} detectors.")
    return AtomBatch(batch.atoms[:, X : X + 1, :], batch.n_atoms[:, X : X + 1], batch.TAtom)


def per_segment_twoF(batch: AtomBatch, window, *, device: int = -1, flags: int | None = None) -> np.ndarray:
    """``2 * F_mn[:, 0]`` of the semi-coherent window for every template: ``[T, nsegs]`` float64
    (``SemiCoherentSearch._get_per_segment_twoF``, core.py:2282-2289)."""
    w = TransientWindowRange.from_any(window)
    if flags is None:
        flags = default_flags()
    _, F = get_handle(device).map_batch(batch, w, flags | _lib.WANT_FMN)
    return 2.0 * F[:, :, 0].astype(np.float64)


def single_IFO_twoFs(batch: AtomBatch, window, *, device: int = -1, flags: int | None = None):
    """Semi-coherent single-detector statistics (core.py:2236-2260).

    Returns ``(twoFX [T, numDet], twoFX_per_segment [T, numDet, nsegs])``; as in the reference a
    NaN in a detector's per-segment values is treated as zero and the sum re-computed.
    """
    per_det = []
    for X in range(batch.numDet):
        per_det.append(per_segment_twoF(single_detector_batch(batch, X), window, device=device, flags=flags))
    per_seg = np.stack(per_det, axis=1)
    twoFX = per_seg.sum(axis=2)
    bad = np.isnan(twoFX)
    if bad.any():
        twoFX[bad] = np.nan_to_num(per_seg, nan=0.0).sum(axis=2)[bad]
        per_seg = np.where(bad[:, :, None], np.nan_to_num(per_seg, nan=0.0), per_seg)
    return twoFX, per_seg


def twoFX_at_maxTwoF(batch: AtomBatch, window, records, *, device: int = -1, flags: int | None = None) -> np.ndarray:
    """Per-detector ``2F`` at each template's multi-detector argmax cell: ``[T, numDet]``.

    ``ComputeFstat.get_transient_log10BSGL`` (core.py:1527-1541) computes one FULL map per
    detector and reads it at ``idx_maxTwoF``; here each (template, detector) is a 1x1 map at
    ``(t0_ML, tau_ML)`` -- one ``tcw_map_batch_windows`` call per detector for the whole batch.
    ``records`` are the multi-detector results of :func:`pyfstat_b200.batch.map_batch`.
    """
    w = TransientWindowRange.from_any(window)
    if len(records) != batch.T:
        raise ValueError("need one result record per template")
    if flags is None:
        flags = default_flags()
    wins = np.zeros((batch.T, 7), dtype=np.uint32)  # one transientWindowRange_t row per template
    wins[:, 0], wins[:, 1], wins[:, 3], wins[:, 4], wins[:, 6] = w.type, records["t0_ML"], w.dt0, records["tau_ML"], w.dtau
    out = np.empty((batch.T, batch.numDet), dtype=np.float64)
    h = get_handle(device)
    for X in range(batch.numDet):
        # a single detector's atoms may leave one atom in the window: F = 2 fallback as pycuda
        rec, _ = h.map_batch_windows(single_detector_batch(batch, X), wins, flags | _lib.ALLOW_DEGENERATE)
        out[:, X] = 2.0 * rec["maxF"].astype(np.float64)
    return out


def cumulative_durations(tstart, tend, Tsft, num_segments) -> np.ndarray:
    """``ComputeFstat._set_up_cumulative_times`` (core.py:1566-1571): first duration ``2 Tsft``,
    last one the whole span."""
    return np.linspace(2 * Tsft, tend - tstart, num_segments)


def twoF_cumulative(batch: AtomBatch, tstart, durations, *, device: int = -1, flags: int | None = None) -> np.ndarray:
    """``2F`` over ``[tstart, tstart + duration]`` for every duration: ``[T, len(durations)]``.

    ``calculate_twoF_cumulative`` (core.py:1648-1665) loops ``get_fullycoherent_detstat`` over
    the durations; each is the rectangular-window F at ``t0 = int(tstart)``,
    ``tau = int(duration)`` (core.py:1447-1449).  When the truncated durations are equally
    spaced this is ONE ``1 x N_tau`` map; otherwise one 1x1 map per duration.
    """
    taus = np.asarray([int(d) for d in durations], dtype=np.int64)
    if len(taus) == 0:
        return np.zeros((batch.T, 0))
    if flags is None:
        flags = default_flags()
    h = get_handle(device)
    steps = np.diff(taus)
    if len(taus) == 1 or (steps.min() == steps.max() and steps[0] > 0):
        dtau = int(steps[0]) if len(taus) > 1 else 1
        w = TransientWindowRange(TRANSIENT_RECTANGULAR, int(tstart), 0, dtau, int(taus[0]), int(taus[-1] - taus[0]), dtau)
        _, F = h.map_batch(batch, w, flags | _lib.WANT_FMN)
        return 2.0 * F[:, 0, :].astype(np.float64)
    out = np.empty((batch.T, len(taus)), dtype=np.float64)
    for k, tau in enumerate(taus):
        w = TransientWindowRange(TRANSIENT_RECTANGULAR, int(tstart), 0, 1, int(tau), 0, 1)
        rec, _ = h.map_batch(batch, w, flags)
        out[:, k] = 2.0 * rec["maxF"].astype(np.float64)
    return out


def install(core_module=None):
    """Route ``SemiCoherentSearch``'s two direct lalpulsar calls through this backend.

    Replaces ``_get_per_segment_twoF`` (core.py:2282-2289) and
    ``get_semicoherent_single_IFO_twoFs`` (core.py:2214-2260) on the class; attributes read and
    written are the reference's (``FstatResults.multiFatoms[0]``, ``semicoherentWindowRange``,
    ``singleFstats``, ``twoFX``, ``twoFX_per_segment``).  Returns the patched class.
    """
    if core_module is None:
        import importlib

        core_module = importlib.import_module("pyfstat.core")
    cls = core_module.SemiCoherentSearch
    if getattr(cls, "_b200_installed", False):
        return cls

    def _get_per_segment_twoF(self):
        batch = from_multi_fstat_atoms(self.FstatResults.multiFatoms[0])
        return per_segment_twoF(batch, self.semicoherentWindowRange)[0]

    def get_semicoherent_single_IFO_twoFs(self, record_segments=False):
        if not self.singleFstats:
            raise RuntimeError("This function is available only if singleFstats or BSGL options were set.")
        batch = from_multi_fstat_atoms(self.FstatResults.multiFatoms[0])
        twoFX, per_seg = single_IFO_twoFs(batch, self.semicoherentWindowRange)
        numDet = self.FstatResults.numDetectors
        for X in range(numDet):
            self.twoFX[X] = twoFX[0, X]
            if record_segments:
                # the reference assigns every detector row from the last computed detector
                # (core.py:2256-2259 broadcasts); here each detector keeps its own segments
                self.twoFX_per_segment[X, :] = per_seg[0, X]
        return self.twoFX

    cls._b200_orig = (cls._get_per_segment_twoF, cls.get_semicoherent_single_IFO_twoFs)
    cls._get_per_segment_twoF = _get_per_segment_twoF
    cls.get_semicoherent_single_IFO_twoFs = get_semicoherent_single_IFO_twoFs
    cls._b200_installed = True
    return cls


def uninstall(core_module):
    cls = core_module.SemiCoherentSearch
    if getattr(cls, "_b200_installed", False):
        cls._get_per_segment_twoF, cls.get_semicoherent_single_IFO_twoFs = cls._b200_orig
        del cls._b200_orig
        cls._b200_installed = False
