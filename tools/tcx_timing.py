"""Role wait cycles of the tensor-core pass (development build -DTCX_TIMING):
   nvcc ... -DTCX_TIMING -o gpurun_out/libtcw_timing.so ; PYFSTAT_B200_LIB=... python tools/tcx_timing.py N T"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200.atoms import synth_atoms  # noqa: E402
from pyfstat_b200.window import canonical_window  # noqa: E402

N, T = int(sys.argv[1]), int(sys.argv[2])
h = L.Handle(0)
lib = L.load_library()
b = synth_atoms(T, N, ("H1", "L1"), seed=3)
w = canonical_window("exp", 10**9, N)
h.upload(b)
out = (ctypes.c_ulonglong * 16)()
for _ in range(2):
    h.map_resident(w, L.WANT_BTSG)
h.synchronize()
lib.tcw_debug_tcx_timing(out, 1)
h.map_resident(w, L.WANT_BTSG)
h.synchronize()
lib.tcw_debug_tcx_timing(out, 0)
v = [int(x) for x in out]
n_cta = max(v[9], 1)
print("stage_ms", h.last_stage_ms())
print("CTAs", n_cta)
print("producer: total %.0f  wait empty %.1f %%" % (v[0] / n_cta, 100.0 * v[1] / max(v[0], 1)))
print("mma     : total %.0f  decode %.1f %%  wait tmem_empty %.1f %%  wait full %.1f %%  MMAs/CTA %.0f  cycles/MMA %.1f"
      % (v[2] / n_cta, 100.0 * v[3] / max(v[2], 1), 100.0 * v[4] / max(v[2], 1), 100.0 * v[5] / max(v[2], 1), v[6] / n_cta,
         v[2] / max(v[6], 1)))
print("epilogue: total %.0f  wait tmem_full %.1f %%" % (v[7] / n_cta, 100.0 * v[8] / max(v[7], 1)))
h.close()
