"""Single-template exp-window latency per stage (development aid)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyfstat_b200 import _lib as L
from pyfstat_b200.atoms import synth_atoms
from pyfstat_b200.window import canonical_window
h = L.Handle(0)
for N in (1440, 2880, 5760):
    b = synth_atoms(1, N, ("H1", "L1"), seed=3)
    w = canonical_window("exp", 10**9, N)
    h.upload(b)
    for name, fl in (("lut", L.WANT_BTSG), ("exact", L.WANT_BTSG | L.EXP_EXACT)):
        rows = []
        for i in range(8):
            h.synchronize(); h.timer_start(); h.map_resident(w, fl); t = h.timer_stop()
            if i >= 3:
                rows.append((t, h.last_stage_ms(), h.last_exp_stage_ms()))
        print(N, name, "total %.3f ms" % statistics.mean(r[0] for r in rows), {k: round(statistics.mean(r[1][k] for r in rows), 3) for k in rows[0][1]},
              {k: round(statistics.mean(r[2][k] for r in rows), 3) for k in rows[0][2]}, flush=True)
h.close()
