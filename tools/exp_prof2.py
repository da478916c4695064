"""exp walker experiments: python tools/exp_prof2.py N T"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyfstat_b200 import _lib as L
from pyfstat_b200.atoms import synth_atoms
from pyfstat_b200.window import canonical_window
N, T = int(sys.argv[1]), int(sys.argv[2])
h = L.Handle(0)
b = synth_atoms(T, N, ("H1", "L1"), seed=3)
w = canonical_window("exp", 10**9, N)
h.upload(b)
for name, fl in (("exact max-only", L.EXP_EXACT), ("exact btsg", L.EXP_EXACT | L.WANT_BTSG), ("exact fmn", L.EXP_EXACT | L.WANT_FMN),
                 ("lut max-only", 0), ("lut btsg", L.WANT_BTSG)):
    ms = []
    for i in range(5):
        h.flush_l2(); h.synchronize()
        h.map_resident(w, fl)
        h.synchronize()
        if i >= 2:
            ms.append((h.last_stage_ms()["map"], h.last_exp_stage_ms()))
    print(name, "map %.3f" % statistics.mean(m[0] for m in ms), {k: round(statistics.mean(m[1][k] for m in ms), 3) for k in ms[0][1]}, flush=True)
h.close()
