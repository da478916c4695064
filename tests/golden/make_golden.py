"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

Two pieces of real reference code are executed here and their outputs committed as fixtures,
because /root/reference does not exist on the GPU box:

1. the reference's own map kernels (pyCUDAkernels/*.cu) compiled for the host by
   oracle/Makefile into oracle/_ref/libtcw_ref.so and launched with the geometry of their
   host wrappers (tcw_fstat_map_funcs.py:878-898, 959-979) -> ``F_ref`` (float32);
2. the reference's own Python class ``pyTransientFstatMap`` (tcw_fstat_map_funcs.py:50-317),
   loaded standalone from /root/reference/pyfstat/tcw_fstat_map_funcs.py, fed with that F_ref:
   ``get_lnBtSG()``, ``get_t0_max_posterior()``, ``get_tau_max_posterior()``,
   ``get_maxF_idx()`` and ``write_F_mn_to_file()``.

Inputs are the seeded synthetic atoms of pyfstat_b200.atoms.synth_atoms (stored in the
fixture as well, so the fixture is self-contained).

    python tests/golden/make_golden.py
"""

import importlib.util
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import tcw_oracle as O  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import TransientWindowRange, canonical_window  # noqa: E402

REF_TCW = "/root/reference/pyfstat/tcw_fstat_map_funcs.py"


def load_ref_tcw():
    spec = importlib.util.spec_from_file_location("ref_tcw_fstat_map_funcs", REF_TCW)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = {
    # name: (detectors, n_per_det, window builder, seed, gap_fraction, inject)
    "rect_H1_1day": (("H1",), 48, lambda: canonical_window("rect", 700000000, 48), 1, 0.0, None, 700000000),
    "exp_H1_1day": (("H1",), 48, lambda: canonical_window("exp", 700000000, 48), 1, 0.0, None, 700000000),
    "rect_H1L1_3day_signal": (
        ("H1", "L1"), 144, lambda: canonical_window("rect", 10**9, 144), 7, 0.0,
        {"type": "rect", "t0": 10**9 + 36 * 1800, "tau": 36 * 1800, "c": (0.9 + 0.3j, -0.4 + 0.5j)}, 10**9,
    ),
    "exp_H1L1_3day_signal": (
        ("H1", "L1"), 144, lambda: canonical_window("exp", 10**9, 144), 8, 0.0,
        {"type": "exp", "t0": 10**9 + 36 * 1800, "tau": 36 * 1800, "c": (0.9 + 0.3j, -0.4 + 0.5j)}, 10**9,
    ),
    "rect_H1L1_gapped": (("H1", "L1"), 120, lambda: canonical_window("rect", 10**9, 120), 9, 0.12, None, 10**9),
    "exp_H1L1_gapped": (("H1", "L1"), 120, lambda: canonical_window("exp", 10**9, 120), 10, 0.12, None, 10**9),
    # unaligned / coarse steps (dt0 != dtau, odd offsets)
    "rect_H1_unaligned": (
        ("H1",), 96,
        lambda: TransientWindowRange(1, 10**9 + 700, 20 * 2700, 2700, 5000, 80 * 1800, 4500), 11, 0.0, None, 10**9,
    ),
    "exp_H1_unaligned": (
        ("H1",), 96,
        lambda: TransientWindowRange(2, 10**9 + 700, 60 * 1800, 2700, 5000, 20 * 1800, 4500), 12, 0.0, None, 10**9,
    ),
}


def main():
    O.build()
    assert O.ref_lib() is not None, "oracle/_ref/libtcw_ref.so missing (needs /root/reference)"
    tcw = load_ref_tcw()
    for name, (dets, n, mkwin, seed, gap, inject, t0d) in CASES.items():
        b = synth_atoms(1, n, dets, seed=seed, t0_data=t0d, gap_fraction=gap, inject=inject)
        w = mkwin()
        merged = O.merge_binned(b.template(0), b.TAtom)
        matrix = O.merged_to_matrix(merged)
        F_ref = O.ref_kernel_map(matrix, b.TAtom, int(merged["timestamp"][0]), w)
        assert not np.isnan(F_ref).any(), name
        # the reference's own result class on the reference kernel's output
        fm = tcw.pyTransientFstatMap(N_t0Range=F_ref.shape[0], N_tauRange=F_ref.shape[1])
        fm.F_mn = F_ref
        fm.maxF = F_ref.max()
        idx = fm.get_maxF_idx()
        lnB = fm.get_lnBtSG()
        t0_MP = fm.get_t0_max_posterior(w)
        tau_MP = fm.get_tau_max_posterior(w)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "map.dat")
            fm.write_F_mn_to_file(path, w, header=["golden fixture", name])
            text = open(path).read()
            rd = tcw.pyTransientFstatMap(from_file=path)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            atoms=b.atoms, n_atoms=b.n_atoms, TAtom=b.TAtom,
            window=np.array([w.type, w.t0, w.t0Band, w.dt0, w.tau, w.tauBand, w.dtau], dtype=np.uint32),
            merged_matrix=matrix, t0_data=int(merged["timestamp"][0]),
            F_ref=F_ref, maxF=np.float32(fm.maxF), argmax=np.array(idx, dtype=np.int64),
            lnBtSG_numpy=np.float64(lnB), t0_MP=np.float64(t0_MP), tau_MP=np.float64(tau_MP),
            text_head="\n".join(text.splitlines()[:12]), text_nlines=len(text.splitlines()),
            reread_maxF=np.float64(rd.maxF), reread_t0_ML=np.float64(rd.t0_ML), reread_tau_ML=np.float64(rd.tau_ML),
        )
        print(f"{name}: F_ref {F_ref.shape} maxF={fm.maxF:.6f} idx={idx} lnBtSG={lnB:.6f} t0_MP={t0_MP} tau_MP={tau_MP}")


if __name__ == "__main__":
    main()
