#!/usr/bin/env python
"""bench.py -- transient F-stat (t0,tau) map throughput on B200 (cells/s, templates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path (merge + scan, map kernel, lnBtSG pass, finalize) over
one batch of synthetic templates per GPU.  Default workload = BASELINE.json configs[1]:
exponential window, 30 days of 1800-s atoms, H1+L1 (1439 x 1441 map per template), batched
over templates; `--workload rect60` is configs[2]'s shape (60 d, H1+L1, rect) and is also
measured briefly in every default run (key "rect") because the two windows sit on different
rooflines (exp: FP32 FMA pipe; rect with F_mn materialised: HBM writes).

Timing: CUDA events on the library's own stream around every step (L2 flushed between steps,
outside the events), summed over the K steps, max over ranks.  `e2e` times the same batch
through the C-ABI call `tcw_map_batch` with PINNED HOST atoms: H2D copy + kernels + D2H of the
result records inside the timed region (wall clock around the synchronous call).

`--impl reference`: the reference's CPU path for the same workload -- the C restatement of
lalpulsar's XLALComputeTransientFstatMap/-Bstat (oracle/, kind "port": lalpulsar itself is
not installable here) -- on all host threads, each step a bounded row-subsample of the map.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TATOM = 1800
T0_DATA = 1_000_000_000

# name: (window, atoms per detector, detectors, templates per GPU per step, BASELINE config idx)
WORKLOADS = {
    "exp30": ("exp", 1440, ("H1", "L1"), 128, 2),
    "rect30": ("rect", 1440, ("H1",), 256, 1),
    "rect60": ("rect", 2880, ("H1", "L1"), 64, 3),
    "exp120": ("exp", 5760, ("H1", "L1"), 4, 4),
}


def exp_atom_visits(n_atoms, N_t0, N_tau, tau0_atoms=2, ef=3):
    """V = sum over cells of (i_t1 - i_t0 + 1) for the canonical exp window."""
    m = np.arange(N_t0)[:, None]
    n = np.arange(N_tau)[None, :]
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
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if it was
    taken at this launch shape (profiles/r01_traffic.json); else None."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        d = json.load(open(p)).get(name)
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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

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
        for r in self.rows:
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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


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


def run_cpu_reference(args, spec):
    """--impl reference: all host threads, OpenMP over templates, K timed steps."""
    from oracle import tcw_oracle as O
    from pyfstat_b200.atoms import synth_atoms

    O.build()
    threads = os.cpu_count() or 1
    row_stride = 16 if spec["window"] == "exp" else 1
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
    """cpu_baseline of the GPU arm: the oracle on ONE thread, bounded sample (~10-30 s)."""
    from oracle import tcw_oracle as O
    from pyfstat_b200.atoms import synth_atoms

    O.build()
    row_stride = 1
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
            stages.append(h.last_stage_ms())
    return ms, stages


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from pyfstat_b200 import _lib as L
    from pyfstat_b200.atoms import synth_atoms
    from pyfstat_b200.batch import gather_records

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

    spec = workload_spec(args.workload)
    if args.templates:
        spec["T"] = args.templates
    h = L.Handle(local_rank)
    T = spec["T"]
    flags = L.WANT_BTSG
    # templates are seeded by GLOBAL template index: every rank count sees the same templates
    alloc = L.pinned_atoms_alloc()
    batch = synth_atoms(T, spec["n"], spec["dets"], seed=1000 * spec["cfg"] + rank * T, t0_data=T0_DATA,
                        TAtom=TATOM, pinned_alloc=alloc)
    h.upload(batch)
    peaks = h.microbench() if rank == 0 else None

    def measure():
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        launches0 = h.launch_count
        ms, stages = timed_steps(h, spec["w"], flags, args.steps, args.warmup)
        launches = h.launch_count - launches0
        barrier()
        clocks = sampler.stop()
        return ms, stages, launches, clocks

    ms, stages, launches, clocks = measure()
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])
    remeasured = False
    if bad:
        ms, stages, launches, clocks = measure()
        remeasured = True
    # launches counted over warm-up + timed steps; scale to the timed steps only
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    total_ms = sum(ms)
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms_max = float(tt.item())
        ll = torch.tensor([launches_timed], dtype=torch.int64, device="cuda")
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        launches_all = int(ll.item())
    else:
        total_ms_max, launches_all = total_ms, launches_timed
    cells_all = world * T * spec["cells"] * args.steps
    value = cells_all / (total_ms_max * 1e-3)

    # ---- e2e: C-ABI call with pinned host atoms, H2D + kernels + D2H (+ record gather) ----
    e2e_times = []
    barrier()
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res, _ = h.map_batch(batch, spec["w"], flags)
        if world > 1:
            gather_records(res, world * T)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_times.append(dt)
    barrier()
    e2e_total = sum(e2e_times)
    if world > 1:
        tt = torch.tensor([e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_total = float(tt.item())
    e2e_value = cells_all / e2e_total
    assert np.all(res["status"] == 0) and np.all(np.isfinite(res["lnBtSG"]))

    # ---- roofline of the dominant kernel (map kernel), from the live stage events ----
    map_ms = statistics.mean(s["map"] for s in stages)
    hbm_peak, hbm_src = measured_peaks()
    if rank == 0:
        if spec["window"] == "exp":
            achieved = T * spec["alg_flop"] / (map_ms * 1e-3) / 1e12
            roofline = {
                "bound": "fp32", "kernel": "tcw_exp_map_kernel", "achieved": achieved,
                "peak": peaks["ffma_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["ffma_tflops"],
                "traffic": measured_traffic(spec["name"], T),
                "peak_source": ("FFMA microbenchmark run by this bench on this GPU (tcw_microbench; nominal 74.4 = "
                                "148 SM x 128 lanes x 2 x 1.965 GHz); MEASURED_PEAKS.json carries no FP32 SIMT peak. "
                                "Not HBM- or tensor-bound: SURVEY 8(d) puts the exponential window on the FP32 FMA pipe"),
                "algorithmic_flop_per_template": spec["alg_flop"],
                "atom_visits_per_template": spec["visits"],
                "launch_ms": map_ms,
            }
        else:
            achieved = T * spec["alg_bytes"] / (map_ms * 1e-3) / 1e9
            roofline = {
                "bound": "hbm", "kernel": "tcw_rect_map_kernel", "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": hbm_src,
                "algorithmic_bytes_per_template": spec["alg_bytes"], "launch_ms": map_ms,
                "note": "F_mn goes to an L2-sized scratch for the lnBtSG pass; see key 'rect' for the materialised case",
            }
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
            "dtype": "f32 (f64 prefix sums / lnBtSG sums)",
            "data": "synthetic",
            "config": workload_config(spec),
            "clocks": dict(clocks, remeasured=remeasured),
            "e2e": {"value": e2e_value, "unit": "cells/s", "templates_per_s": e2e_value / spec["cells"],
                    "h2d_bytes_per_step": int(batch.nbytes), "d2h_bytes_per_step": int(T * L.RESULT_DTYPE.itemsize),
                    "api": "tcw_map_batch (C ABI) with pinned host atoms" + (" + NCCL all_gather of records" if world > 1 else "")},
            "gpu_launches": launches_all,
            "roofline": roofline,
            "stage_ms": {k: statistics.mean(s[k] for s in stages) for k in stages[0]},
            "microbench": peaks,
        }
        if not args.no_secondary and args.workload == "exp30":
            line["rect"] = secondary_rect(h, L, hbm_peak, hbm_src)
            line["mcmc"] = secondary_mcmc(h, L)
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_single_thread(spec)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def secondary_rect(h, L, hbm_peak, hbm_src):
    """configs[2] shape (60 d, H1+L1, rect) with F_mn MATERIALISED in HBM: the output-bound
    case whose roofline is HBM write bandwidth."""
    from pyfstat_b200.atoms import synth_atoms

    spec = workload_spec("rect60")
    T = spec["T"]
    alloc = L.pinned_atoms_alloc()
    batch = synth_atoms(T, spec["n"], spec["dets"], seed=3000, t0_data=T0_DATA, TAtom=TATOM, pinned_alloc=alloc)
    h.upload(batch)
    out = {}
    for name, flags in (("fmn", L.WANT_FMN), ("fmn_btsg", L.WANT_FMN | L.WANT_BTSG), ("fused_max_only", 0)):
        ms, stages = timed_steps(h, spec["w"], flags, 5, 3)
        map_ms = statistics.mean(s["map"] for s in stages)
        step_ms = statistics.mean(ms)
        gbs = T * spec["alg_bytes"] / (map_ms * 1e-3) / 1e9
        out[name] = {
            "cells_per_s": T * spec["cells"] / (step_ms * 1e-3),
            "ms_per_step": step_ms,
            "map_kernel_ms": map_ms,
            "stage_ms": {k: statistics.mean(s[k] for s in stages) for k in stages[0]},
        }
        if flags & L.WANT_FMN:
            out[name]["roofline"] = {"bound": "hbm", "kernel": "tcw_rect_map_kernel", "achieved": gbs, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": gbs / hbm_peak,
                                     "traffic": measured_traffic("rect60", T) if name == "fmn" else None,
                                     "peak_source": hbm_src,
                                     "algorithmic_bytes_per_template": spec["alg_bytes"]}
    # end to end through tcw_map_batch (pinned host atoms -> records), lnBtSG on, no F_mn copy
    times = []
    for i in range(8):
        t0 = time.perf_counter()
        h.map_batch(batch, spec["w"], L.WANT_BTSG)
        if i >= 3:
            times.append(time.perf_counter() - t0)
    out["e2e_btsg"] = {"cells_per_s": T * spec["cells"] / statistics.mean(times), "ms_per_step": 1e3 * statistics.mean(times),
                       "h2d_bytes_per_step": int(batch.nbytes), "d2h_bytes_per_step": int(T * L.RESULT_DTYPE.itemsize)}
    out["config"] = workload_config(spec)
    return out


def secondary_mcmc(h, L):
    """BASELINE configs[4] shape: one sampler step = 256 walkers, each a 1x1 map with its own
    (tstart, duration) on 30 d of H1+L1 atoms, through tcw_map_batch_windows (host atoms in,
    records out).  Reports per-step latency and templates/s for both windows."""
    from pyfstat_b200.atoms import synth_atoms

    T, n = 256, 1440
    alloc = L.pinned_atoms_alloc()
    batch = synth_atoms(T, n, ("H1", "L1"), seed=5000, t0_data=T0_DATA, TAtom=TATOM, pinned_alloc=alloc)
    rng = np.random.default_rng(5)
    tstart = T0_DATA + rng.uniform(0, 0.5 * n * TATOM, T)
    dur = rng.uniform(4 * TATOM, 0.45 * n * TATOM, T)
    from pyfstat_b200 import backend
    from pyfstat_b200.mcmc import transient_detstat_batch

    out = {"walkers_per_step": T, "atoms_per_detector": n,
           "api": "pyfstat_b200.mcmc.transient_detstat_batch -> tcw_map_batch_windows (C ABI), pinned host atoms in, "
                  "one detection statistic per walker out"}
    saved = backend._handles.get(-1)
    backend._handles[-1] = h  # the sampler-facing helper runs on this bench's handle
    try:
        for name in ("rect", "exp"):
            times = []
            for i in range(13):
                t0 = time.perf_counter()
                transient_detstat_batch(batch, tstart, tstart + dur, name)
                if i >= 3:
                    times.append(time.perf_counter() - t0)
            out[name] = {"ms_per_step": 1e3 * statistics.mean(times), "templates_per_s": T / statistics.mean(times)}
    finally:
        if saved is None:
            backend._handles.pop(-1, None)
        else:
            backend._handles[-1] = saved
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="exp30", choices=sorted(WORKLOADS))
    ap.add_argument("--templates", type=int, default=0, help="templates per GPU per step (default per workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the single-thread cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the brief rect60 measurement")
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
