"""F-stat atoms on the host: the 32-byte record layout, the adapter from lalpulsar's
``MultiFstatAtomVector`` (duck-typed) and the synthetic atom generator used by the tests
and by ``bench.py``.

The record mirrors lalpulsar's ``FstatAtom`` in the field order the reference reads it
(``pyfstat/tcw_fstat_map_funcs.py:610-617``): ``timestamp`` u32; ``a2_alpha``, ``b2_alpha``,
``ab_alpha`` f32; ``Fa_alpha``, ``Fb_alpha`` complex64 (stored as re/im pairs).  It is the
``tcw_atom`` struct of ``include/tcw_b200.h``.
"""

from __future__ import annotations

import numpy as np

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
assert ATOM_DTYPE.itemsize == 32

# order of the 7 summed quantities = columns of the reference's atomsInputMatrix (tcw:711-721)
CHANNELS = ("a2_alpha", "b2_alpha", "ab_alpha", "Fa_re", "Fa_im", "Fb_re", "Fb_im")


class AtomBatch:
    """A batch of ``T`` templates x ``numDet`` detector atom vectors in one contiguous
    (optionally pinned) host array, ready for a single host-to-device copy.

    ``atoms[t, X, :n_atoms[t, X]]`` are the valid atoms of template ``t``, detector ``X``.
    """

    def __init__(self, atoms: np.ndarray, n_atoms: np.ndarray, TAtom: int):
        atoms = np.asarray(atoms)
        if atoms.dtype != ATOM_DTYPE:
            raise TypeError(f"atoms must have dtype ATOM_DTYPE, got {atoms.dtype}")
        if atoms.ndim != 3:
            raise ValueError("atoms must have shape (T, numDet, stride)")
        if not atoms.flags["C_CONTIGUOUS"]:
            atoms = np.ascontiguousarray(atoms)
        n_atoms = np.ascontiguousarray(n_atoms, dtype=np.uint32)
        if n_atoms.shape != atoms.shape[:2]:
            raise ValueError("n_atoms must have shape (T, numDet)")
        if n_atoms.max(initial=0) > atoms.shape[2]:
            raise ValueError("each detector vector needs 0 <= n_atoms <= stride")
        if n_atoms.size and n_atoms.max(axis=1).min() < 1:
            raise ValueError("every template needs at least one atom in some detector")
        if int(TAtom) <= 0:
            raise ValueError("TAtom must be a positive integer")
        self.atoms = atoms
        self.n_atoms = n_atoms
        self.TAtom = int(TAtom)

    @property
    def T(self) -> int:
        return self.atoms.shape[0]

    @property
    def numDet(self) -> int:
        return self.atoms.shape[1]

    @property
    def stride(self) -> int:
        return self.atoms.shape[2]

    @property
    def nbytes(self) -> int:
        return self.atoms.nbytes + self.n_atoms.nbytes

    def __len__(self) -> int:
        return self.T

    def __getitem__(self, sl) -> "AtomBatch":
        if isinstance(sl, int):
            sl = slice(sl, sl + 1)
        return AtomBatch(self.atoms[sl], self.n_atoms[sl], self.TAtom)

    def template(self, t: int):
        """List of per-detector atom arrays (valid part only) of template ``t``."""
        return [self.atoms[t, X, : self.n_atoms[t, X]] for X in range(self.numDet)]


def batch_from_detector_lists(templates, TAtom: int) -> AtomBatch:
    """Build an :class:`AtomBatch` from ``templates[t][X]`` = 1-D ATOM_DTYPE arrays."""
    T = len(templates)
    numDet = len(templates[0])
    stride = max(1, max(len(a) for tpl in templates for a in tpl))
    atoms = np.zeros((T, numDet, stride), dtype=ATOM_DTYPE)
    n_atoms = np.zeros((T, numDet), dtype=np.uint32)
    for t, tpl in enumerate(templates):
        if len(tpl) != numDet:
            raise ValueError("all templates need the same number of detectors")
        for X, a in enumerate(tpl):
            atoms[t, X, : len(a)] = a
            n_atoms[t, X] = len(a)
    return AtomBatch(atoms, n_atoms, TAtom)


def from_multi_fstat_atoms(multiFstatAtoms) -> AtomBatch:
    """Adapter for the registered backend's first argument.

    Accepts what the reference passes (``pyfstat/core.py:1454``): a
    ``lalpulsar.MultiFstatAtomVector`` -- duck-typed as ``.length``, ``.data[X].length``,
    ``.data[X].TAtom``, ``.data[X].data[i].{timestamp,a2_alpha,b2_alpha,ab_alpha,Fa_alpha,
    Fb_alpha}`` (tcw:607-632, tests/test_tcw_fstat_map_funcs.py:65-78) -- or an
    :class:`AtomBatch` with ``T == 1``, or a list of per-detector ATOM_DTYPE arrays carrying
    a ``TAtom`` via ``(arrays, TAtom)``.
    """
    if isinstance(multiFstatAtoms, AtomBatch):
        return multiFstatAtoms
    if isinstance(multiFstatAtoms, tuple) and len(multiFstatAtoms) == 2:
        arrays, TAtom = multiFstatAtoms
        return batch_from_detector_lists([list(arrays)], TAtom)
    numDet = int(multiFstatAtoms.length)
    if numDet < 1:
        raise ValueError("multiFstatAtoms holds no detector")
    TAtom = int(multiFstatAtoms.data[0].TAtom)
    per_det = []
    for X in range(numDet):
        vec = multiFstatAtoms.data[X]
        if int(vec.TAtom) != TAtom:
            raise ValueError("all detectors must share TAtom (XLALmergeMultiFstatAtomsBinned)")
        n = int(vec.length)
        out = _view_swig_atoms(vec.data, n) if n >= 2 and _zero_copy_enabled() else None
        if out is None:
            out = np.zeros(n, dtype=ATOM_DTYPE)
            data = vec.data
            # per-atom attribute reads, like reshape_FstatAtomsVector (tcw:627-632)
            for i in range(n):
                out[i] = _read_atom(data[i])
        per_det.append(out)
    return batch_from_detector_lists([per_det], TAtom)


def _read_atom(atom):
    """One ``FstatAtom`` through its attributes (tcw:610-617), as an ATOM_DTYPE scalar tuple."""
    Fa = complex(atom.Fa_alpha)
    Fb = complex(atom.Fb_alpha)
    return (atom.timestamp, atom.a2_alpha, atom.b2_alpha, atom.ab_alpha, Fa.real, Fa.imag, Fb.real, Fb.imag)


def _zero_copy_enabled() -> bool:
    import os

    return os.environ.get("PYFSTAT_B200_ZERO_COPY", "1") not in ("0", "")


def _view_swig_atoms(data, n: int):
    """read lal's ``FstatAtom[n]`` as ONE block instead of ``6 n`` SWIG attribute reads.

    lalpulsar's ``FstatAtom`` is ``{UINT4 timestamp; REAL4 a2_alpha, b2_alpha, ab_alpha; COMPLEX8
    Fa_alpha, Fb_alpha}`` = the 32-byte ``tcw_atom`` record, and a SWIG element wrapper exposes
    its C address as ``int(element.this)``.  Neither fact can be checked against lalpulsar in
    the build container, so the view is VERIFIED at run time: the address stride between
    elements 0 and 1 must be 32 bytes, and the first, second and last records must equal what
    the attribute path reads.  Any doubt (no ``.this``, other stride, mismatch, exception)
    returns None and the caller falls back to the per-atom loop.  The result is a copy; the
    caller's memory is never aliased or written.
    """
    import ctypes

    try:
        first, second = data[0], data[1]
        a0, a1 = int(first.this), int(second.this)
        if a0 <= 0 or a1 - a0 != ATOM_DTYPE.itemsize:
            return None
        raw = (ctypes.c_char * (n * ATOM_DTYPE.itemsize)).from_address(a0)
        out = np.frombuffer(raw, dtype=ATOM_DTYPE, count=n).copy()
        for i in (0, 1, n - 1):
            ref = np.zeros(1, dtype=ATOM_DTYPE)
            ref[0] = _read_atom(data[i])
            if out[i].tobytes() != ref[0].tobytes():
                return None
        return out
    except Exception:  # noqa: BLE001 -- anything unexpected: use the verified slow path
        return None


# --------------------------------------------------------------------------------------
# synthetic atoms: lalpulsar is not available, so atoms are synthesised
# --------------------------------------------------------------------------------------

SIDEREAL_DAY = 86164.0905
_DET_PHASES = {  # detector-specific phases/amplitudes of the toy antenna patterns
    "H1": (0.3, 1.1, 0.9, 0.55, 2.0, 0.4),
    "L1": (1.9, 0.2, 0.8, 0.65, 0.7, 2.6),
    "V1": (4.1, 3.0, 0.7, 0.60, 5.2, 1.3),
}


def _antenna_patterns(det: str, t: np.ndarray):
    """Smooth toy antenna-pattern functions a(t), b(t): sidereal + half-sidereal sinusoids."""
    p1, p2, amp_a, amp_b, p3, p4 = _DET_PHASES[det]
    w = 2.0 * np.pi / SIDEREAL_DAY
    a = amp_a * (0.35 + 0.5 * np.sin(w * t + p1) + 0.3 * np.sin(2 * w * t + p2))
    b = amp_b * (0.25 + 0.45 * np.cos(w * t + p3) + 0.35 * np.sin(2 * w * t + p4))
    return a, b


def synth_atoms(
    T: int,
    n_per_det: int,
    detectors=("H1",),
    *,
    seed: int = 0,
    t0_data: int = 1_000_000_000,
    TAtom: int = 1800,
    inject=None,
    gap_fraction: float = 0.0,
    pinned_alloc=None,
) -> AtomBatch:
    """Synthetic F-stat atoms for ``T`` templates (SURVEY 8d recipe).

    Per detector: ``a2=a^2, b2=b^2, ab=a*b`` from smooth toy antenna patterns (noise-weighted
    atoms are scale-free for F), ``Fa = a*z/sqrt(2)``, ``Fb = b*z/sqrt(2)`` with ``z`` a unit
    complex normal per atom (the rank-1 per-atom covariance 1/2 [[a2,ab],[ab,b2]], giving
    E[2F] = 4 in noise).  Template ``t`` uses ``default_rng(seed + t)``.

    ``inject``: optional dict ``{"type": "rect"|"exp", "t0": gps, "tau": s, "c": (c1, c2)}``
    adding ``M . c * g(t)`` with complex amplitudes c and window g (t_tcw:56 uses
    t0 = tau = Tspan/4).  ``gap_fraction``: fraction of atoms dropped per detector
    (different gaps per detector, same for all templates, as for real SFT sets).
    ``pinned_alloc(nbytes) -> buffer`` lets the caller place the batch in pinned memory.
    """
    detectors = tuple(detectors)
    numDet = len(detectors)
    ts_full = t0_data + TAtom * np.arange(n_per_det, dtype=np.int64)
    gap_rng = np.random.default_rng(987654321 + seed)
    keep = []
    for X in range(numDet):
        k = np.ones(n_per_det, dtype=bool)
        if gap_fraction > 0:
            k = gap_rng.random(n_per_det) >= gap_fraction
            k[0] = X == 0 or k[0]  # keep the very first atom in detector 0
            if not k.any():
                k[0] = True
        keep.append(np.flatnonzero(k))
    stride = max(len(k) for k in keep)
    shape = (T, numDet, stride)
    if pinned_alloc is not None:
        buf = pinned_alloc(int(np.prod(shape)) * ATOM_DTYPE.itemsize)
        atoms = np.frombuffer(buf, dtype=ATOM_DTYPE, count=int(np.prod(shape))).reshape(shape)
        atoms[...] = np.zeros((), dtype=ATOM_DTYPE)
    else:
        atoms = np.zeros(shape, dtype=ATOM_DTYPE)
    n_atoms = np.zeros((T, numDet), dtype=np.uint32)
    pat = []
    for X, det in enumerate(detectors):
        t = ts_full[keep[X]].astype(np.float64)
        a, b = _antenna_patterns(det, t)
        sig = None
        if inject is not None:
            dt = t - float(inject["t0"])
            if inject["type"] == "rect":
                g = ((dt >= 0) & (dt < float(inject["tau"]))).astype(np.float64)
            else:
                g = np.where(dt >= 0, np.exp(-np.clip(dt, 0, None) / float(inject["tau"])), 0.0)
                g = np.where(dt <= 3.0 * float(inject["tau"]), g, 0.0)
            c1, c2 = (complex(c) for c in inject["c"])
            sig = ((a * a * c1 + a * b * c2) * g, (a * b * c1 + b * b * c2) * g)
        pat.append((t, a, b, sig))
    inv_sqrt2 = 1.0 / np.sqrt(2.0)
    for tpl in range(T):
        rng = np.random.default_rng(seed + tpl)
        for X in range(numDet):
            t, a, b, sig = pat[X]
            n = len(t)
            z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * inv_sqrt2
            Fa = a * z
            Fb = b * z
            if sig is not None:
                Fa = Fa + sig[0]
                Fb = Fb + sig[1]
            rec = atoms[tpl, X, :n]
            rec["timestamp"] = t.astype(np.uint32)
            rec["a2_alpha"] = (a * a).astype(np.float32)
            rec["b2_alpha"] = (b * b).astype(np.float32)
            rec["ab_alpha"] = (a * b).astype(np.float32)
            rec["Fa_re"] = Fa.real.astype(np.float32)
            rec["Fa_im"] = Fa.imag.astype(np.float32)
            rec["Fb_re"] = Fb.real.astype(np.float32)
            rec["Fb_im"] = Fb.imag.astype(np.float32)
            n_atoms[tpl, X] = n
    return AtomBatch(atoms, n_atoms, TAtom)
