"""Error statistics of the tiled kernels against the CPU oracle (lal semantics) at BASELINE sizes
(development aid, run under gpurun).  One JSON line per case; rectangular-window cases carry rows
stratified by the condition number of the window-summed antenna-pattern matrix (the documented
exception of DESIGN.md section 2 lives in the high-cond strata), and every case is run under both
exp-table geometries on file when the table matters (exp window; lnBtSG always)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import tcw_oracle as O  # noqa: E402
from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

COND_EDGES = (1.0, 10.0, 1e2, 1e3, 2e3, 5e3, 1e4, np.inf)


def cond_map(merged, N_t0, N_tau, tau0_atoms=2):
    """Condition number per cell of the canonical rect map (dt0 = dtau = TAtom, aligned), float64."""
    P = {c: np.concatenate([[0.0], np.cumsum(merged[c].astype(np.float64))]) for c in ("a2_alpha", "b2_alpha", "ab_alpha")}
    N = len(merged)
    m = np.arange(N_t0)[:, None]
    n = np.arange(N_tau)[None, :]
    i0 = np.minimum(m, N - 1)
    i1 = np.minimum(m + tau0_atoms + n - 1, N - 1)
    A = P["a2_alpha"][i1 + 1] - P["a2_alpha"][i0]
    B = P["b2_alpha"][i1 + 1] - P["b2_alpha"][i0]
    C = P["ab_alpha"][i1 + 1] - P["ab_alpha"][i0]
    d = np.sqrt((A - B) ** 2 + 4 * C * C)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(A + B - d > 0, (A + B + d) / (A + B - d), np.inf)


h = L.Handle(0)
CASES = (("rect", 1440, ("H1",), "configs[0]"), ("rect", 1440, ("H1", "L1"), ""), ("rect", 2880, ("H1", "L1"), "configs[2]"),
         ("exp", 1440, ("H1", "L1"), "configs[1]"), ("exp", 720, ("H1",), ""))
for win, n, dets, label in CASES:
    for geom in ((20.0, 5120), (20.0, 2000)):
        if win == "rect" and geom != (20.0, 5120) and len(dets) > 1:
            continue  # the table only enters lnBtSG there: one geometry is enough
        h.set_exp_lut(*geom)
        O.set_exp_lut(*geom)
        b = synth_atoms(1, n, dets, seed=171)
        w = canonical_window(win, 10**9, n)
        t0 = time.time()
        o = O.compute_map(b.template(0), b.TAtom, w)
        dt = time.time() - t0
        Fo = o["F_mn"]
        # exponential window: the default path (FP64 recurrence + FP16 tensor-core pass) and the tiled direct sum
        for path_name, path_flags in ((("recurrence+tensor", 0), ("direct_sum", L.EXP_DIRECT)) if win == "exp" else (("tiled", 0),)):
            res, F = h.map_batch(b, w, L.WANT_FMN | L.WANT_BTSG | path_flags)
            rel = np.abs(F[0] - Fo) / np.maximum(np.abs(Fo), 1e-30)
            row = dict(
                window=win, atoms=n, detectors="+".join(dets), baseline=label, exp_lut=f"{geom[0]:g}:{geom[1]}", kernels=path_name,
                cells=int(rel.size), oracle_s=round(dt, 2),
                rel_median=float(np.median(rel)), rel_p99=float(np.quantile(rel, 0.99)), rel_p9999=float(np.quantile(rel, 0.9999)),
                rel_max=float(rel.max()), n_gt_1e4=int((rel > 1e-4).sum()), n_gt_1e5=int((rel > 1e-5).sum()),
                argmax_equal=bool((int(res["m_ML"][0]), int(res["n_ML"][0])) == (o["m_ML"], o["n_ML"])),
                maxF_rel=float(abs(float(res["maxF"][0]) - o["maxF"]) / o["maxF"]),
                lnBtSG_abs=float(abs(float(res["lnBtSG"][0]) - o["lnBtSG"])),
                MP_equal=bool((int(res["m_MP"][0]), int(res["n_MP"][0])) == (o["m_MP"], o["n_MP"])),
            )
            if win == "rect":
                cond = cond_map(o["merged"], *Fo.shape)
                fallback = (F[0] == 2.0) != (Fo == 2.0)
                strata = []
                for lo, hi in zip(COND_EDGES[:-1], COND_EDGES[1:]):
                    sel = (cond >= lo) & (cond < hi) & ~fallback
                    if sel.any():
                        strata.append(dict(cond=f"[{lo:g}, {hi:g})", cells=int(sel.sum()), rel_median=float(np.median(rel[sel])),
                                           rel_max=float(rel[sel].max()), n_gt_1e4=int((rel[sel] > 1e-4).sum()),
                                           max_rel_over_cond=float((rel[sel] / cond[sel]).max())))
                row["cond_strata"] = strata
                row["cells_flipped_across_the_cut"] = int(fallback.sum())
            print(json.dumps(row), flush=True)
h.close()
