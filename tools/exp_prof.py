"""One resident exp-window map (target for ncu): python tools/exp_prof.py N T [exact|direct]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

N, T = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else ""
fl = L.WANT_BTSG | (L.EXP_EXACT if "exact" in mode else 0) | (L.EXP_DIRECT if "direct" in mode else 0)
h = L.Handle(0)
b = synth_atoms(T, N, ("H1", "L1"), seed=3)
w = canonical_window("exp", 10**9, N)
h.upload(b)
for _ in range(2):
    h.map_resident(w, fl)
h.synchronize()
print(h.last_stage_ms(), h.last_exp_stage_ms())
h.close()
