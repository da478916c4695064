"""MCMC-step timing through the sampler-facing helper (development aid, run under gpurun)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pyfstat_b200 import _lib as L  # noqa: E402
from pyfstat_b200 import backend  # noqa: E402
from pyfstat_b200.atoms import AtomBatch, synth_atoms  # noqa: E402
from pyfstat_b200.mcmc import transient_detstat_batch  # noqa: E402

n, T = 1440, 256
b0 = synth_atoms(T, n, ("H1", "L1"), seed=1)
pin = L.PinnedBuffer(b0.atoms.nbytes)
arr = np.frombuffer(pin.array, dtype=b0.atoms.dtype, count=b0.atoms.size).reshape(b0.atoms.shape)
arr[...] = b0.atoms
b = AtomBatch(arr, b0.n_atoms, b0.TAtom)
rng = np.random.default_rng(3)
ts = 10**9 + rng.uniform(0, 0.5 * n * 1800, T)
du = rng.uniform(4 * 1800, 0.45 * n * 1800, T)
for win in ("rect", "exp"):
    for i in range(5):
        transient_detstat_batch(b, ts, ts + du, win)
    t0 = time.perf_counter()
    for i in range(20):
        transient_detstat_batch(b, ts, ts + du, win)
    st = backend.get_handle().last_stage_ms()
    print(win, "pinned" if arr is not None else "pageable", "ms/step %.3f" % ((time.perf_counter() - t0) / 20 * 1e3),
          "map stage %.3f ms" % st["map"])
