"""Quick GPU parity + timing probe (development aid, run under gpurun).

Compares the CUDA paths (generic and tiled) with the CPU oracle on small seeded cases and
prints kernel timings for the BASELINE shapes.  The formal parity tests live in tests/.
"""

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


def compare(h, batch, w, flags, label, exact=False):
    res, F = h.map_batch(batch, w, flags | L.WANT_FMN | L.WANT_BTSG, raise_on_degenerate=False)
    out = []
    for t in range(batch.T):
        o = O.compute_map(batch.template(t), batch.TAtom, w, exact_exp=exact, allow_degenerate=True)
        Fo = o["F_mn"]
        Fg = F[t].astype(np.float64)
        rel = np.abs(Fg - Fo) / np.maximum(np.abs(Fo), 1e-30)
        out.append(
            dict(
                label=label, t=t, path=int(res["path"][t]), shape=list(Fo.shape),
                bit_equal=bool(np.array_equal(F[t], Fo.astype(np.float32))),
                max_rel=float(rel.max()), n_gt_1e4=int((rel > 1e-4).sum()),
                maxF=(float(res["maxF"][t]), o["maxF"]),
                argmax=((int(res["m_ML"][t]), int(res["n_ML"][t])), (o["m_ML"], o["n_ML"])),
                lnBtSG=(float(res["lnBtSG"][t]), o["lnBtSG"]),
                MP=((int(res["m_MP"][t]), int(res["n_MP"][t])), (o["m_MP"], o["n_MP"])),
                status=(int(res["status"][t]), o["status"]),
            )
        )
    return out


def main():
    h = L.Handle(0)
    print("device:", h.device_name)
    print("microbench:", h.microbench())
    rows = []
    for dets in (("H1",), ("H1", "L1")):
        for N in (48, 200):
            b = synth_atoms(2, N, dets, seed=11 * N + len(dets))
            for win in ("rect", "exp"):
                w = canonical_window(win, 10**9, N)
                rows += compare(h, b, w, L.FORCE_GENERIC, f"generic {win} N={N} {dets}")
                rows += compare(h, b, w, 0, f"fast {win} N={N} {dets}")
    # gapped data
    b = synth_atoms(1, 300, ("H1", "L1"), seed=5, gap_fraction=0.1)
    for win in ("rect", "exp"):
        w = canonical_window(win, 10**9, 300)
        rows += compare(h, b, w, L.FORCE_GENERIC, f"generic {win} gapped")
        rows += compare(h, b, w, 0, f"fast {win} gapped")
    for r in rows:
        print(json.dumps(r))

    # timings on the BASELINE shapes (kernel-only, resident)
    for win, N, dets, T in (("rect", 1440, ("H1",), 32), ("rect", 2880, ("H1", "L1"), 32),
                            ("exp", 1440, ("H1", "L1"), 16)):
        b = synth_atoms(T, N, dets, seed=3)
        w = canonical_window(win, 10**9, N)
        h.upload(b)
        for flags, name in ((0, "max only"), (L.WANT_FMN, "F_mn"), (L.WANT_BTSG, "BtSG"),
                            (L.WANT_FMN | L.WANT_BTSG, "F_mn+BtSG")):
            for rep in range(3):
                h.timer_start()
                h.map_resident(w, flags)
                ms = h.timer_stop()
            st = h.last_stage_ms()
            cells = (N - 1) * (N + 1) * T
            print(f"{win} N={N} T={T} {name}: {ms:.3f} ms  {cells / ms * 1e3:.3e} cells/s  stages={st}")
    h.close()


if __name__ == "__main__":
    main()
