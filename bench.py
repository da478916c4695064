#!/usr/bin/env python
"""bench.py -- transient F-stat (t0,tau) map throughput on B200 (cells/s, templates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path (merge + scan, map kernel, lnBtSG pass, finalize) over one
batch of synthetic templates per GPU.

Headline workload (`exp120`) = BASELINE.json configs[3]'s shape, the largest single-GPU
configuration: exponential window, 120 days of 1800-s atoms, H1+L1 (5759 x 5761 map, 8.5e10 atom
visits per template), 8 templates per GPU per step, lnBtSG on.  The other BASELINE configs are
measured in the same line under "configs" (N = 1 only):
  configs[0]  rect, 30 d, H1, ONE template per call through the registered plugin callable (latency)
  configs[1]  exp, 30 d, H1+L1, batched (device-timed + end to end + roofline)
  configs[2]  rect, 60 d, H1+L1: device-timed with F_mn materialised / lnBtSG / fused-only, and the
              10^4-point BatchedTransientGridSearch end to end
  configs[4]  MCMC: 256 walkers per step through the ptemcee-style pool mapper

Timing: CUDA events on the library's own stream around every step (L2 flushed between steps,
outside the events), summed over the K steps, max over ranks.  `e2e` is the same batch through
the public sharding driver `pyfstat_b200.batch.map_sharded` -> C ABI `tcw_map_batch` with PINNED
HOST atoms: H2D + kernels + D2H of the records (+ the NCCL all_gather of the records at N > 1)
inside the timed region (wall clock around the synchronous call).

Multi-GPU: templates are seeded by GLOBAL index and sharded in contiguous blocks (map_sharded).
The main line is weak scaling (8 templates per GPU per step); "strong" pushes a FIXED set of 64
templates through the same driver at every N.  "records_crc" are checksums of gathered result
records (global templates 0..7 / all 64): identical at every N = rank-count invariance on hardware.

`--impl reference`: the reference's CPU path for the same workload -- the C restatement of
lalpulsar's XLALComputeTransientFstatMap/-Bstat (oracle/, kind "port": lalpulsar itself is not
installable here) -- on all host threads, each step a bounded row-subsample of the map.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TATOM = 1800
T0_DATA = 1_000_000_000
STRONG_TEMPLATES = 64  # fixed template set of the strong-scaling leg (divisible by 1, 2, 4, 8)

# name: (window, atoms per detector, detectors, templates per GPU per step, BASELINE config number)
WORKLOADS = {
    "exp120": ("exp", 5760, ("H1", "L1"), 8, 4),
    "exp30": ("exp", 1440, ("H1", "L1"), 128, 2),
    "rect30": ("rect", 1440, ("H1",), 256, 1),
    "rect60": ("rect", 2880, ("H1", "L1"), 64, 3),
}


def exp_atom_visits(n_atoms, N_t0, N_tau, tau0_atoms=2, ef=3):
    """V = sum over cells of (i_t1 - i_t0 + 1) for the canonical exp window."""
    m = np.arange(N_t0, dtype=np.int64)[:, None]
    n = np.arange(N_tau, dtype=np.int64)[None, :]
    i1 = np.minimum(m + ef * (tau0_atoms + n) - 1, n_atoms - 1)
    return int((i1 - m + 1).sum())


def workload_spec(name):
    from pyfstat_b200.window import canonical_window

    win, n, dets, T, cfg = WORKLOADS[name]
    w = canonical_window(win, T0_DATA, n, TATOM)
    N_t0, N_tau = w.dims()
    cells = N_t0 * N_tau
    spec = dict(name=name, window=win, n=n, dets=dets, T=T, cfg=cfg, w=w, N_t0=N_t0, N_tau=N_tau, cells=cells)
    if win == "exp":
        V = exp_atom_visits(n, N_t0, N_tau)
        spec["visits"] = V
        spec["alg_flop"] = 14 * V + 40 * cells  # 7 FMA per atom visit + ~40 flop epilogue
    else:
        spec["alg_bytes"] = 4 * cells + 32 * n * len(dets) + 80  # F_mn store + atoms + record
    return spec


def measured_traffic(name, T):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if it was taken
    at this launch shape (profiles/r02_traffic.json, else r01_traffic.json); else None."""
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", fn))).get(name)
            if d and d["templates_per_launch"] == T:
                return d["dram_bytes_read"] + d["dram_bytes_write"]
        except (OSError, ValueError, KeyError):
            pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def measured_tensor_peaks():
    """Dense 16-bit tensor-core peaks (TFLOP/s): (burst, sustained, source)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), \
            "MEASURED_PEAKS.json bf16_tflops (cuBLAS bf16 8192^3: burst / sustained; kind::f16 runs at the bf16 rate)"
    return 2250.0, 2250.0, "nominal dense bf16 of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.mark_at = 0

    def wait_first(self, timeout=6.0):
        """nvidia-smi needs up to a second to produce its first line (longer with 8 ranks starting one each):
        wait for it BEFORE the timed region, so that the region itself is sampled."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)
        return self

    def mark(self):
        """Samples taken before this point are not part of the record."""
        self.mark_at = len(self.rows)

    def count(self):
        return len(self.rows) - self.mark_at

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows[self.mark_at:]:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


BAD_REASONS = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def workload_config(spec):
    return {
        "workload": (f"BASELINE configs[{spec['cfg'] - 1}] shape: {spec['window']} window, "
                     f"{spec['n'] * TATOM // 86400} d of {TATOM}-s atoms, {'+'.join(spec['dets'])}, "
                     f"map {spec['N_t0']}x{spec['N_tau']} (dt0=dtau=TAtom), lnBtSG on, F_mn not copied to host"),
        "name": spec["name"],
        "templates_per_gpu_per_step": spec["T"],
        "cells_per_template": spec["cells"],
        "l2": "L2 flushed (256 MiB memset) between timed steps, outside the CUDA events",
        "exp_mode": "lal_lut",
        "exp_path": ("FP64 recurrence down the rows + tcgen05 FP16 pass for the lookup table's deviation from the exact "
                     "exponential (default on canonical grids; TCW_EXP_DIRECT = tiled direct sum)"),
    }


# ---------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ---------------------------------------------------------------------------------------
def cpu_sample_window(spec, row_stride):
    """Row-subsampled window range: the same cells as rows m = 0, row_stride, ... of the map."""
    from pyfstat_b200.window import TransientWindowRange

    w = spec["w"]
    N_rows = (spec["N_t0"] - 1) // row_stride + 1
    ws = TransientWindowRange(w.type, w.t0, (N_rows - 1) * row_stride * w.dt0, row_stride * w.dt0, w.tau, w.tauBand, w.dtau)
    return ws, N_rows * spec["N_tau"]


def cpu_row_stride(spec, seconds_per_template):
    """Row stride that keeps one template of the CPU sample near `seconds_per_template` on one core
    (~1.3e8 exp atom visits/s resp. ~3.5e7 rect cells/s measured in round 1)."""
    if spec["window"] == "exp":
        stride = int(round(spec["visits"] / (1.3e8 * seconds_per_template)))
    else:
        stride = int(round(spec["cells"] / (3.5e7 * seconds_per_template)))
    return max(1, stride)


def run_cpu_reference(args, spec):
    """--impl reference: all host threads, OpenMP over templates, K timed steps."""
    from oracle import tcw_oracle as O
    from pyfstat_b200.atoms import synth_atoms

    O.build()
    threads = os.cpu_count() or 1
    row_stride = cpu_row_stride(spec, 3.0)
    per_thread = 1 if spec["window"] == "exp" else 4
    T = threads * per_thread
    ws, cells_per_tpl = cpu_sample_window(spec, row_stride)
    batch = synth_atoms(T, spec["n"], spec["dets"], seed=1000 * spec["cfg"], t0_data=T0_DATA, TAtom=TATOM)
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc, _ = O.batch(batch.atoms, batch.n_atoms, TATOM, ws, want_btsg=True, num_threads=threads)
        dt = time.perf_counter() - t0
        assert rc == 0, rc
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = T * cells_per_tpl * len(times) / total
    sample = (f"{T} templates/step on {threads} OpenMP threads (one template per thread at a time, single-threaded "
              f"per template like XLALComputeTransientFstatMap); every {row_stride}th t0 row of the "
              f"{spec['N_t0']}x{spec['N_tau']} map = {cells_per_tpl} cells/template; lal semantics "
              f"(lookup-table weights, REAL4 sums, lnBtSG)")
    line = {
        "impl": "reference",
        "metric": "transient F-stat (t0,tau) map cells/s",
        "value": value,
        "unit": "cells/s",
        "templates_per_s": value / spec["cells"],
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(spec),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "os_cpu_count": os.cpu_count(),
                         "note": "lal-equivalent C restatement (oracle/tcw_oracle.c), not lalpulsar"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_single_thread(spec):
    """cpu_baseline of the GPU arm: the oracle on ONE thread, bounded sample (~10-20 s)."""
    from oracle import tcw_oracle as O
    from pyfstat_b200.atoms import synth_atoms

    O.build()
    row_stride = cpu_row_stride(spec, 12.0)
    T = 1 if spec["window"] == "exp" else 100
    ws, cells_per_tpl = cpu_sample_window(spec, row_stride)
    batch = synth_atoms(T, spec["n"], spec["dets"], seed=1000 * spec["cfg"], t0_data=T0_DATA, TAtom=TATOM)
    t0 = time.perf_counter()
    rc, _ = O.batch(batch.atoms, batch.n_atoms, TATOM, ws, want_btsg=True, num_threads=1)
    dt = time.perf_counter() - t0
    assert rc == 0
    return {
        "value": T * cells_per_tpl / dt,
        "unit": "cells/s",
        "cores": 1,
        "kind": "port",
        "sample": (f"{T} template(s) of the workload, every {row_stride}th t0 row ({cells_per_tpl} cells each), "
                   f"{dt:.1f} s on one thread; lal semantics incl. lnBtSG"),
        "cpu": cpu_model(),
        "os_cpu_count": os.cpu_count(),
        "note": "lal-equivalent C restatement (oracle/tcw_oracle.c), not lalpulsar",
    }


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def timed_steps(h, w, flags, steps, warmup):
    """Returns (per-step ms list, per-step stage dicts) for K timed steps after W warm-ups."""
    ms, stages = [], []
    for i in range(warmup + steps):
        h.flush_l2()
        h.synchronize()
        h.timer_start()
        h.map_resident(w, flags)
        t = h.timer_stop()
        if i >= warmup:
            ms.append(t)
            st = h.last_stage_ms()
            st.update({"map_" + k: v for k, v in h.last_exp_stage_ms().items()})
            stages.append(st)
    return ms, stages


def exp_split(stages):
    return {k: statistics.mean(s["map_" + k] for s in stages) for k in ("operands", "tensor", "walk")}


def exp_variants(h, L, spec, batch, peaks, steps=2, warmup=1):
    """The same batch through the other exponential-window paths: exact exponentials (the reference's
    pycuda semantics: pure FP64 recurrence) and the tiled direct sum (TCW_EXP_DIRECT, round 1's FFMA2 kernel)."""
    out = {}
    T = batch.T
    h.upload(batch)
    for name, fl in (("exact_exp_recurrence", L.WANT_BTSG | L.EXP_EXACT), ("lut_direct_sum", L.WANT_BTSG | L.EXP_DIRECT)):
        ms, stages = timed_steps(h, spec["w"], fl, steps, warmup)
        map_ms = statistics.mean(s["map"] for s in stages)
        out[name] = {"cells_per_s": T * spec["cells"] / (statistics.mean(ms) * 1e-3), "ms_per_step": statistics.mean(ms),
                     "ms_per_template": statistics.mean(ms) / T, "steps": steps, "stage_ms": mean_stage(stages)}
        if name == "lut_direct_sum":
            out[name]["roofline"] = exp_roofline(spec, T, map_ms, peaks)
    return out


def exp_dispatch_ladder(h, L, spec, batch):
    """What a grid that is NOT canonical costs (exp window, 30-d data): the host planner's fallbacks, so that the
    cliffs between them are on record.  `path` is the result record's path field (2 recurrence + tensor cores,
    1 tiled direct sum, 0 generic kernels)."""
    from pyfstat_b200.window import TransientWindowRange

    n, TA = spec["n"], TATOM
    ladder = {
        # rows two atoms apart (dt0 = 2 TAtom): the recurrence path on the refined grid (every second row emitted) ...
        "dt0_2TAtom_recurrence": (TransientWindowRange(2, T0_DATA, (n - 2) * TA, 2 * TA, 2 * TA, n * TA, TA), 0, 16),
        # ... and the same grid through the tiled direct sum (rows fetch their own records)
        "dt0_2TAtom_tiled": (TransientWindowRange(2, T0_DATA, (n - 2) * TA, 2 * TA, 2 * TA, n * TA, TA), L.EXP_DIRECT, 16),
        # dt0 = dtau = 8 TAtom (a coarse grid over the same data: beyond the tiled kernels' 4 atoms per row)
        "dt0_8TAtom_recurrence": (TransientWindowRange(2, T0_DATA, (n - 8) * TA, 8 * TA, 8 * TA, n * TA, 8 * TA), 0, 16),
        # a grid offset from the atom grid by a third of an atom, dt0 = TAtom: still canonical (one row class)
        "offset_grid_recurrence": (TransientWindowRange(2, T0_DATA + TA // 3, (n - 3) * TA, TA, 2 * TA, n * TA, TA), 0, 16),
        # dt0 = 1.5 TAtom: two row classes, tiled direct sum
        "dt0_1p5TAtom_tiled": (TransientWindowRange(2, T0_DATA, (n - 2) * TA, 3 * TA // 2, 2 * TA, n * TA, TA), 0, 16),
        # the generic kernels (one thread per cell, the reference's sequential sums: the parity anchor)
        "generic_kernel": (spec["w"], L.FORCE_GENERIC, 2),
    }
    out = {}
    for name, (w, fl, T) in ladder.items():
        sub = batch[:T]
        h.upload(sub)
        h.map_resident(w, L.WANT_BTSG | L.ALLOW_DEGENERATE | fl)  # warm-up (weight tables, allocations)
        h.synchronize()
        h.flush_l2()
        h.synchronize()
        h.timer_start()
        h.map_resident(w, L.WANT_BTSG | L.ALLOW_DEGENERATE | fl)
        ms = h.timer_stop()
        rec = h.fetch_results()
        cells = int(rec["N_t0"][0]) * int(rec["N_tau"][0]) if "N_t0" in rec.dtype.names else None
        out[name] = {"templates": T, "ms_per_template": ms / T, "path": int(rec["path"][0]),
                     "cells_per_s": (cells * T / (ms * 1e-3)) if cells else None}
    h.upload(batch)
    return out


def pinned_batch(L, T, spec, seed):
    from pyfstat_b200.atoms import synth_atoms

    return synth_atoms(T, spec["n"], spec["dets"], seed=seed, t0_data=T0_DATA, TAtom=TATOM,
                       pinned_alloc=L.pinned_atoms_alloc())


def records_crc(rec):
    """Checksum of the fields of result records that callers read."""
    keys = ("maxF", "m_ML", "n_ML", "t0_ML", "tau_ML", "lnBtSG", "t0_MP", "tau_MP", "m_MP", "n_MP", "status")
    crc = 0
    for k in keys:
        crc = zlib.crc32(np.ascontiguousarray(rec[k]).tobytes(), crc)
    return f"{crc:08x}"


def mean_stage(stages):
    return {k: statistics.mean(s[k] for s in stages) for k in stages[0]}


def exp_rec_roofline(spec, T, xs, hbm_peak, hbm_src):
    """Rooflines of the exponential window's default path (tcw_exp_rec.cuh): the tensor-core pass for the
    lookup table's deviation from the exact exponential (dominant kernel) and the FP64 walk."""
    burst, sustained, src = measured_tensor_peaks()
    tc_flop = 14 * spec["visits"]  # 7 channel MACs per atom visit; the kernel issues that plus ~5 % k-range padding
    achieved = T * tc_flop / (xs["tensor"] * 1e-3) / 1e12
    walk_bytes = (20 + 4) * spec["cells"]  # correction sums in (3 x FP32 + 4 x FP16), F_mn out (the lnBtSG pass re-reads it)
    walk = T * walk_bytes / (xs["walk"] * 1e-3) / 1e9
    return {
        "bound": "tensor", "kernel": "tcw_exptc_map_kernel<FP16>", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
        "frac": achieved / burst, "frac_of_sustained_peak": achieved / sustained, "peak_sustained": sustained,
        "traffic": measured_traffic(spec["name"], T), "peak_source": src,
        "algorithmic_flop_per_template": tc_flop, "atom_visits_per_template": spec["visits"],
        "launch_ms": xs["tensor"],
        "note": ("algorithmic flop = 2 x 7 channels x atom visits of the map (the direct sum's MAC count) = what the kernel "
                 "executes up to ~5 % k-range padding (7 x 64 = 448 MMA columns per tile and k step, none idle); the tensor pipe "
                 "runs at the rate cuBLAS sustains on this board (power cap), DESIGN.md section 5"),
        "walk_kernel": {"bound": "hbm", "kernel": "tcw_exp_walk_kernel", "achieved": walk, "peak": hbm_peak, "unit": "GB/s",
                        "frac": walk / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes_per_template": walk_bytes,
                        "launch_ms": xs["walk"]},
        "operand_prep_ms": xs["operands"],
    }


def exp_roofline(spec, T, map_ms, peaks):
    achieved = T * spec["alg_flop"] / (map_ms * 1e-3) / 1e12
    return {
        "bound": "fp32", "kernel": "tcw_exp_map_canon_kernel (TCW_EXP_DIRECT)", "achieved": achieved,
        "peak": peaks["ffma_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["ffma_tflops"],
        "frac_of_nominal_74.4": achieved / 74.4,
        "traffic": measured_traffic(spec["name"], T),
        "peak_source": ("FFMA microbenchmark run by this bench on this GPU (tcw_microbench; nominal 74.4 = "
                        "148 SM x 128 lanes x 2 x 1.965 GHz); MEASURED_PEAKS.json carries no FP32 SIMT peak. "
                        "Not HBM- or tensor-bound: SURVEY 8(d) puts the exponential window on the FP32 FMA pipe"),
        "algorithmic_flop_per_template": spec["alg_flop"],
        "atom_visits_per_template": spec["visits"],
        "launch_ms": map_ms,
    }


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from pyfstat_b200 import _lib as L
    from pyfstat_b200 import backend
    from pyfstat_b200.batch import map_sharded, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything libraries print (e.g. NCCL's version
    # banner) goes to stderr until the result is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyfstat_b200 has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    spec = workload_spec(args.workload)
    if args.templates:
        spec["T"] = args.templates
    T = spec["T"]
    T_total = world * T
    seed0 = 1000 * spec["cfg"]
    h = backend.get_handle(local_rank)  # THE handle of this process: the sharding drivers use it too
    flags = L.WANT_BTSG
    # templates are seeded by GLOBAL template index and sharded in contiguous blocks: rank r owns
    # [r T, (r+1) T) of the N T templates of a step (pyfstat_b200.batch.shard_range)
    lo, hi = shard_range(T_total, rank, world)
    batch = pinned_batch(L, hi - lo, spec, seed0 + lo)
    h.upload(batch)
    peaks = h.microbench() if rank == 0 else None

    def measure():
        sampler = ClockSampler(local_rank).start().wait_first()
        barrier()
        sampler.mark()
        launches0 = h.launch_count
        ms, stages = timed_steps(h, spec["w"], flags, args.steps, args.warmup)
        launches = h.launch_count - launches0
        in_region = sampler.count()
        # a default run's timed region lasts ~0.15 s, one or two 100-ms samples: keep the SAME steps running
        # (untimed) until the record has at least 4 samples of this load
        t0 = time.time()
        while sampler.proc is not None and sampler.count() < 4 and time.time() - t0 < 3.0:
            h.map_resident(spec["w"], flags)
            h.synchronize()
        barrier()
        clocks = sampler.stop()
        clocks["samples_in_timed_region"] = in_region
        return ms, stages, launches, clocks

    ms, stages, launches, clocks = measure()
    remeasured = False
    if BAD_REASONS & set(clocks["reasons"]):
        ms, stages, launches, clocks = measure()
        remeasured = True
    # launches counted over warm-up + timed steps; scale to the timed steps only
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    total_ms_max = max_over_ranks(sum(ms))
    if world > 1:
        ll = torch.tensor([launches_timed], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches_all = int(ll.item())
    else:
        launches_all = launches_timed
    cells_all = T_total * spec["cells"] * args.steps
    value = cells_all / (total_ms_max * 1e-3)

    # ---- e2e: the public sharding driver, pinned host atoms in, gathered records out ----
    def e2e_loop(batch_local, T_all, n_timed, n_warm):
        """A JOB of n_timed steps through the public sharding driver: ONE map_sharded call over n_timed x T_all
        templates, each rank walking its contiguous block one step's worth at a time (chunk: a pinned H2D copy, the
        kernels and a D2H read of the records per step) and one all_gather of all records at the end -- the
        shape of BatchedTransientGridSearch.run(group).  (A collective per step makes every step wait for the
        slowest of N ranks: measured 7.5x instead of 7.9x at N = 8.)  The pre-staged pinned batch stands for
        every step's atoms."""
        T_loc = len(batch_local)
        assert T_loc * world == T_all

        def make_shard(a, b):
            assert b - a == T_loc
            return batch_local

        def job(n):
            return map_sharded(make_shard, T_all * n, spec["w"], BtSG=True, device=local_rank, chunk=T_loc)

        if n_warm:
            job(n_warm)
        barrier()
        t0 = time.perf_counter()
        rec = job(n_timed)
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt), rec

    e2e_total, rec = e2e_loop(batch, T_total, args.steps, args.warmup)
    e2e_value = cells_all / e2e_total
    assert len(rec) == T_total * args.steps and np.all(rec["status"] == 0) and np.all(np.isfinite(rec["lnBtSG"]))
    crc_first = records_crc(rec[:T])  # global templates 0..T-1 exist at every N

    # ---- strong scaling: a FIXED template set through the same driver ----
    strong = None
    if not args.no_strong:
        Ts = STRONG_TEMPLATES
        slo, shi = shard_range(Ts, rank, world)
        sbatch = pinned_batch(L, shi - slo, spec, seed0 + 500 + slo)
        n_timed = max(2, min(args.steps, 3))
        s_total, srec = e2e_loop(sbatch, Ts, n_timed, 1)
        assert len(srec) == Ts * n_timed and np.all(srec["status"] == 0)
        # rank r's block of the job holds its (shi - slo) templates n_timed times: the first pass of every rank,
        # in rank order, is the fixed set in global template order
        per = len(srec) // world
        srec = np.concatenate([srec[r * per:r * per + per // n_timed] for r in range(world)])
        strong = {
            "templates_total": Ts, "steps": n_timed, "value": Ts * spec["cells"] * n_timed / s_total,
            "unit": "cells/s", "ms_per_step": 1e3 * s_total / n_timed, "records_crc": records_crc(srec),
            "api": "pyfstat_b200.batch.map_sharded (contiguous template blocks per rank, one all_gather of the job's records)",
            "note": "end to end (pinned host atoms in, gathered records out); records_crc must be identical at every N",
        }
        del sbatch

    # ---- roofline of the dominant kernel (map kernel), from the live stage events ----
    map_ms = statistics.mean(s["map"] for s in stages)
    hbm_peak, hbm_src = measured_peaks()
    if rank == 0:
        if spec["window"] == "exp":
            roofline = exp_rec_roofline(spec, hi - lo, exp_split(stages), hbm_peak, hbm_src)
        else:
            achieved = T * spec["alg_bytes"] / (map_ms * 1e-3) / 1e9
            roofline = {
                "bound": "hbm", "kernel": "tcw_rect_map_kernel", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": measured_traffic(spec["name"], T),
                "peak_source": hbm_src, "algorithmic_bytes_per_template": spec["alg_bytes"], "launch_ms": map_ms,
            }
        xmax, length, canonical = h.get_exp_lut()
        line = {
            "metric": "transient F-stat (t0,tau) map cells/s",
            "value": value,
            "unit": "cells/s",
            "templates_per_s": value / spec["cells"],
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 (f64 recurrences / prefix sums / table indices, f16 tensor-core correction pass with f32 "
                     "accumulation, fixed-point lnBtSG marginals)",
            "data": "synthetic",
            "config": workload_config(spec),
            "clocks": dict(clocks, remeasured=remeasured),
            "e2e": {"value": e2e_value, "unit": "cells/s", "templates_per_s": e2e_value / spec["cells"],
                    "h2d_bytes_per_step": int(batch.nbytes), "d2h_bytes_per_step": int((hi - lo) * L.RESULT_DTYPE.itemsize),
                    "ms_per_step": 1e3 * e2e_total / args.steps,
                    "api": "pyfstat_b200.batch.map_sharded(chunk = one step's templates) -> tcw_map_batch (C ABI) per step: "
                           "pinned host atoms in, records out; ONE job of `steps` steps per rank"
                           + (", one NCCL all_gather of all records at its end" if world > 1 else "")},
            "gpu_launches": launches_all,
            "roofline": roofline,
            "stage_ms": mean_stage(stages),
            "microbench": peaks,
            "exp_lut": {"xmax": xmax, "length": length, "canonical": canonical, "source": backend.exp_lut_geometry()[3]},
            "sharding": {"driver": "pyfstat_b200.batch.map_sharded", "templates_per_step_all_ranks": T_total,
                         "records_crc_templates_0_to_%d" % (T - 1): crc_first,
                         "note": "templates seeded by global index; the checksum of the first block's gathered "
                                 "records must be identical at every N (rank-count invariance)"},
        }
        if strong:
            line["strong"] = strong
        if spec["window"] == "exp" and not args.no_secondary:
            line["other_paths"] = exp_variants(h, L, spec, batch, peaks)
        if not args.no_secondary and world == 1 and args.workload == "exp120":
            line["configs"] = secondary_configs(h, L, hbm_peak, hbm_src, peaks, local_rank)
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_single_thread(spec)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
# the other BASELINE configs (N = 1), each with its own clocks record
# ---------------------------------------------------------------------------------------
def with_clocks(fn, gpu_index):
    sampler = ClockSampler(gpu_index).start().wait_first()
    sampler.mark()
    out = fn()
    out["clocks"] = sampler.stop()
    return out


def secondary_configs(h, L, hbm_peak, hbm_src, peaks, gpu_index):
    return {
        "0_rect30_H1_single_call": with_clocks(lambda: sec_single_call(h, L), gpu_index),
        "1_exp30": with_clocks(lambda: sec_exp30(h, L, peaks), gpu_index),
        "2_rect60": with_clocks(lambda: sec_rect60(h, L, hbm_peak, hbm_src), gpu_index),
        "2_rect60_grid_1e4": with_clocks(lambda: sec_grid(h, L), gpu_index),
        "3_exp120": "the headline record of this line",
        "4_mcmc_256_walkers": with_clocks(lambda: sec_mcmc(h, L), gpu_index),
    }


def sec_single_call(h, L, calls=200):
    """configs[0]: ONE template per call through the registered plugin callable (what PyFstat's
    dispatcher invokes per Doppler point): host atoms in, FstatMap fields out."""
    import pyfstat_b200
    from pyfstat_b200.atoms import synth_atoms

    out = {"api": "pyfstat_b200.b200_compute_transient_fstat_map(multiFstatAtoms, windowRange, BtSG) -> maxF, "
                  "get_maxF_idx(), lnBtSG (one tcw_map_batch per call: H2D + kernels + D2H)"}
    for key, win, n, dets in (("rect30_H1", "rect", 1440, ("H1",)), ("exp30_H1L1", "exp", 1440, ("H1", "L1"))):
        spec = workload_spec("rect30" if win == "rect" else "exp30")
        b = synth_atoms(1, n, dets, seed=1000 * spec["cfg"], t0_data=T0_DATA, TAtom=TATOM)
        for btsg in (False, True):
            ts = []
            for i in range(calls + 20):
                t0 = time.perf_counter()
                fm = pyfstat_b200.b200_compute_transient_fstat_map(b, spec["w"], btsg, device=h.device_index)
                _ = fm.maxF, fm.get_maxF_idx(), fm.lnBtSG
                ts.append(time.perf_counter() - t0)
            ts = ts[20:]
            out[f"{key}_BtSG_{btsg}"] = {
                "calls": calls, "median_ms_per_call": 1e3 * statistics.median(ts), "mean_ms_per_call": 1e3 * statistics.mean(ts),
                "p95_ms_per_call": 1e3 * sorted(ts)[int(0.95 * len(ts))], "templates_per_s": 1.0 / statistics.mean(ts),
                "cells_per_s": spec["cells"] / statistics.mean(ts),
                "h2d_bytes_per_call": int(b.nbytes), "d2h_bytes_per_call": int(L.RESULT_DTYPE.itemsize),
            }
    return out


def sec_exp30(h, L, peaks, steps=10, warmup=3):
    """configs[1]: exp window, 30 d, H1+L1, 128 templates per step."""
    spec = workload_spec("exp30")
    T = spec["T"]
    batch = pinned_batch(L, T, spec, 1000 * spec["cfg"])
    h.upload(batch)
    ms, stages = timed_steps(h, spec["w"], L.WANT_BTSG, steps, warmup)
    map_ms = statistics.mean(s["map"] for s in stages)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        h.map_batch(batch, spec["w"], L.WANT_BTSG)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return {
        "value": T * spec["cells"] / (statistics.mean(ms) * 1e-3), "unit": "cells/s",
        "templates_per_s": T / (statistics.mean(ms) * 1e-3), "ms_per_step": statistics.mean(ms), "steps": steps,
        "e2e": {"value": T * spec["cells"] / statistics.mean(times), "unit": "cells/s",
                "h2d_bytes_per_step": int(batch.nbytes), "d2h_bytes_per_step": int(T * L.RESULT_DTYPE.itemsize),
                "api": "tcw_map_batch (C ABI), pinned host atoms in, records out"},
        "roofline": exp_rec_roofline(spec, T, exp_split(stages), *measured_peaks()), "stage_ms": mean_stage(stages),
        "other_paths": exp_variants(h, L, spec, batch, peaks), "dispatch_ladder": exp_dispatch_ladder(h, L, spec, batch),
        "config": workload_config(spec),
    }


def sec_rect60(h, L, hbm_peak, hbm_src, steps=20, warmup=3):
    """configs[2] shape (60 d, H1+L1, rect), 64 templates per step: F_mn MATERIALISED in HBM (the
    output-bound case whose roofline is HBM bandwidth), lnBtSG, and fused max/argmax only."""
    spec = workload_spec("rect60")
    T = spec["T"]
    batch = pinned_batch(L, T, spec, 1000 * spec["cfg"])
    h.upload(batch)
    out = {"steps": steps}
    for name, flags in (("fmn", L.WANT_FMN), ("fmn_btsg", L.WANT_FMN | L.WANT_BTSG), ("btsg", L.WANT_BTSG),
                        ("btsg_exact", L.WANT_BTSG | L.EXP_EXACT), ("fused_max_only", 0)):
        ms, stages = timed_steps(h, spec["w"], flags, steps, warmup)
        map_ms = statistics.mean(s["map"] for s in stages)
        step_ms = statistics.mean(ms)
        gbs = T * spec["alg_bytes"] / (map_ms * 1e-3) / 1e9
        out[name] = {"cells_per_s": T * spec["cells"] / (step_ms * 1e-3), "ms_per_step": step_ms,
                     "map_kernel_ms": map_ms, "stage_ms": mean_stage(stages)}
        if flags & L.WANT_FMN:
            out[name]["roofline"] = {"bound": "hbm", "kernel": "tcw_rect_map_kernel", "achieved": gbs, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": gbs / hbm_peak,
                                     "traffic": measured_traffic("rect60", T) if name == "fmn" else None,
                                     "peak_source": hbm_src, "algorithmic_bytes_per_template": spec["alg_bytes"]}
        if flags & L.WANT_BTSG:
            pass_ms = statistics.mean(s["btsg"] for s in stages)
            rd = T * 4 * spec["cells"] / (pass_ms * 1e-3) / 1e9
            out[name]["btsg_pass"] = {"bound": "hbm", "kernel": "tcw_btsg_kernel", "achieved": rd, "peak": hbm_peak,
                                      "unit": "GB/s", "frac": rd / hbm_peak, "launch_ms": pass_ms,
                                      "algorithmic_bytes_per_template": 4 * spec["cells"]}
    # end to end through tcw_map_batch (pinned host atoms -> records), lnBtSG on, no F_mn copy
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        h.map_batch(batch, spec["w"], L.WANT_BTSG)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    out["e2e_btsg"] = {"cells_per_s": T * spec["cells"] / statistics.mean(times), "ms_per_step": 1e3 * statistics.mean(times),
                       "h2d_bytes_per_step": int(batch.nbytes), "d2h_bytes_per_step": int(T * L.RESULT_DTYPE.itemsize)}
    out["config"] = workload_config(spec)
    return out


def sec_grid(h, L, n_points=10_000, batch_size=256):
    """configs[2]: TransientGridSearch over a 10^4-point F0/F1 grid, 60 d H1+L1, rect window, lnBtSG,
    through the batched driver.  Atom production is outside the path (in PyFstat:
    lalpulsar.ComputeFstat on the CPU): the driver's callback hands out pre-staged pinned batches."""
    from pyfstat_b200.grid_search import BatchedTransientGridSearch

    spec = workload_spec("rect60")
    ring = [pinned_batch(L, batch_size, spec, 7000 + 1000 * k) for k in range(2)]
    calls = [0]

    def atoms_for_points(points):
        b = ring[calls[0] % len(ring)]
        calls[0] += 1
        return b[: len(points)]

    n_f0 = 100  # the Doppler values are labels only (the atoms are synthetic): integer-valued grid, exact point count
    ranges = {"F0": [0.0, float(n_f0 - 1), 1.0], "F1": [0.0, float(n_points // n_f0 - 1), 1.0],
              "F2": [0], "Alpha": [1.0], "Delta": [0.5]}
    best = None
    for rep in range(2):  # first repetition warms up (allocations, first-touch)
        s = BatchedTransientGridSearch(atoms_for_points, ranges, spec["w"], BtSG=True, batch_size=batch_size,
                                       device=h.device_index)
        t0 = time.perf_counter()
        data = s.run()
        best = time.perf_counter() - t0
    assert len(data) == s.total_iterations and np.all(np.isfinite(data["lnBtSG"]))
    n = s.total_iterations
    return {"grid_points": n, "batch_size": batch_size, "wall_s": best, "templates_per_s": n / best,
            "cells_per_s": n * spec["cells"] / best, "map_call_s": s.timingFstatMap,
            "h2d_bytes_total": int(n * ring[0].nbytes // batch_size), "d2h_bytes_total": int(2 * n * L.RESULT_DTYPE.itemsize),
            "api": "pyfstat_b200.grid_search.BatchedTransientGridSearch.run(): tcw_submit/tcw_wait per batch + the "
                   "full-span twoF map on the resident atoms; output table as TransientGridSearch writes it",
            "note": "atoms handed out from a ring of 2 pre-staged pinned batches (synthetic; atom production is outside the path)"}


def sec_mcmc(h, L, steps=200, warmup=10):
    """configs[4]: one sampler step = 256 walkers, each a 1x1 map with its own (tstart, duration) on
    30 d of H1+L1 atoms, through the ptemcee-style pool mapper (TransientWalkerPool.map -> one
    tcw_map_batch_windows call).  Per-step latency and templates/s for both windows."""
    from pyfstat_b200.mcmc import TransientWalkerPool

    T, n = 256, 1440
    spec = workload_spec("exp30")
    batch = pinned_batch(L, T, spec, 5000)
    rng = np.random.default_rng(5)
    thetas = np.column_stack([T0_DATA + rng.uniform(0, 0.5 * n * TATOM, T), rng.uniform(4 * TATOM, 0.45 * n * TATOM, T)])

    class Search:  # the members of MCMCTransientSearch its _logl uses (mcmc_based_searches.py:3479-3516)
        maxStartTime = T0_DATA + n * TATOM
        likelihooddetstatmultiplier = 0.5
        likelihoodcoef = 0.0
        BtSG = False

        def _set_point_for_evaluation(self, theta):
            return {"tstart": theta[0], "tend": theta[0] + theta[1]}

        def _logl(self, theta, search):
            raise AssertionError("evaluated by the pool")

    class Evaluator:  # ptemcee's LikePriorEvaluator surface
        def __init__(self, s):
            self.logl, self.logp, self.loglargs, self.logpargs = s._logl, (lambda t: 0.0), (None,), ()

    out = {"walkers_per_step": T, "atoms_per_detector": n, "steps": steps,
           "api": "pyfstat_b200.mcmc.TransientWalkerPool.map (ptemcee pool surface) -> transient_detstat_batch -> "
                  "tcw_map_batch_windows (C ABI), pinned host atoms in, (logl, logp) per walker out"}
    for name in ("rect", "exp"):
        s = Search()
        s.transientWindowType = name
        pool = TransientWalkerPool(s, lambda pts: batch, device=h.device_index)
        ev = Evaluator(s)
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(ev, thetas)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        assert len(res) == T and all(np.isfinite(r[0]) for r in res)
        out[name] = {"median_ms_per_step": 1e3 * statistics.median(times), "mean_ms_per_step": 1e3 * statistics.mean(times),
                     "templates_per_s": T / statistics.mean(times), "h2d_bytes_per_step": int(batch.nbytes)}
    out["projected_s_for_2000_steps"] = {k: 2000 * out[k]["mean_ms_per_step"] * 1e-3 for k in ("rect", "exp")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="exp120", choices=sorted(WORKLOADS))
    ap.add_argument("--templates", type=int, default=0, help="templates per GPU per step (default per workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the single-thread cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-set strong-scaling leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return  # rank 0 alone runs the CPU arm
        run_cpu_reference(args, workload_spec(args.workload))
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
