#!/usr/bin/env python
"""Headline benchmark of the multigrid solve path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one V-cycle (Multigrid::vCycle, multigrid.cpp:62-110) on the synthetic workload the metric is
quoted on: a 4M-node jittered point cloud on the unit square (2000 x 2000 lattice side), manufactured-solution
Dirichlet Poisson problem, PHS r^3 + polyDeg-4 RBF-FD on the finest level, polyDeg 3 on the coarse levels,
nu=5 SOR sweeps, omega=1.4 (gen_mg_param, testing_functions.cpp:372-380).  The hierarchy is coarsened 4x per
level down to a ~16x16 cloud (the reference's own coarsest cloud has 170 nodes), because the reference's
"coarse solve" is only 2*nu SOR sweeps (multigrid.cpp:92-95).

Reported on ONE JSON line:
  value / ms_per_step  V-cycles per second, throughput mode (multicolour SOR, reordered row sums), operators resident in HBM
  e2e                  same metric through the C-ABI with HOST buffers: every step uploads source_ and values_ from pinned
                       host memory, runs one V-cycle and downloads values_ and the residual
  lexicographic        the reference-faithful mode (dependency-DAG lexicographic SOR, reference-order row sums), reported separately
  solve                cycles and seconds to reduce |b-Ax|_1/|b|_1 below 1e-8 (loop shape of FractionalStepSim.cpp:139-142)
  roofline             the dominant kernel (finest-level SOR sweep): algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline         the reference's own V-cycle on the host (below) on a bounded sample of the same workload, single thread like the reference

`--impl reference` times the reference's own CPU implementation of the path on the stated workload: Multigrid::vCycle /
Grid::sor / Grid::residual from the reference's grid.cpp and multigrid.cpp as compiled into oracle/_ref/libref.so (kind
"reference"), one thread because the reference is strictly serial.  Its operators come from the CPU oracle's set-up on all host
threads, untimed -- bit-identical to the reference's own assembly, whose brute-force kNN is O(N^2) and would not finish at this
size.  Without libref.so (or with --ref-port) the oracle's restatement of the V-cycle is timed instead (kind "port").  The line
carries "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1e-8


def ncu_traffic(kernel, side, fine_poly):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel` on this workload, read from the committed ncu
    summary index (profiles/ncu_traffic.json, keyed kernel|side|polyDeg); None when no capture of this kernel/workload exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get("%s|%d|%d" % (kernel, side, fine_poly), {}).get("dram_bytes")
    except (OSError, ValueError):
        return None


def golden_history(side, fine_poly, key):
    """residual history of the first cycles of this workload from the CPU oracle alone (tests/make_bench_golden.py)"""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_%d_p%d.json" % (side, fine_poly))) as f:
            return json.load(f).get(key)
    except (OSError, ValueError):
        return None


def level_sides(side, levels=None):
    out = [side]
    while (levels is None and out[-1] > 16) or (levels is not None and len(out) < levels):
        out.append((out[-1] + 1) // 2)
    return out[::-1]


def algorithmic_bytes_per_cycle(sides, fine_poly, coarse_poly=3, nu=5):
    """sum_l>=1 [(2nu+1) B_A(l) + B_R(l) + B_P(l)] + B_A(finest) + 2nu B_A(0), fp64 values + int32 columns."""
    from meshlessmultigridpoisson_b200.problems import stencil_size

    n_i = stencil_size(fine_poly)          # Multigrid uses the finest grid's polyDeg for every P and R (multigrid.cpp:22,25)
    L = len(sides)
    total = 0
    for l, s in enumerate(sides):
        N = s * s
        k = stencil_size(fine_poly if l == L - 1 else coarse_poly)
        b_a = N * (12 * k + 24)
        if l >= 1:
            nc = sides[l - 1] ** 2
            total += (2 * nu + 1) * b_a + (nc * (12 * n_i + 8) + 8 * N) + (N * (12 * n_i + 16) + 8 * nc)
        else:
            total += 2 * nu * b_a
        if l == L - 1:
            total += b_a
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_ready(self, limit=8.0):
        """nvidia-smi takes up to a few seconds to print its first line on a multi-GPU box; the timed region may be 40 ms."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < limit:
            time.sleep(0.02)

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l.split(", ") for (_, l) in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_cpu_oracle(side, levels_below, fine_poly, cycles, threads_setup):
    """Bounded sample on the host: same generator/seeds, smaller finest lattice; V-cycle loop single threaded like the
    reference (std::clock pair, testing_functions.cpp:340-344)."""
    import oracle

    os.environ["OMP_NUM_THREADS"] = str(threads_setup)
    sides = level_sides(side)
    t0 = time.time()
    mg = oracle.make_hierarchy(sides, kind=oracle.KIND_DIRICHLET, fine_poly=fine_poly)
    kind = "port"
    try:                                          # the cycles run the reference's own code when oracle/_ref/libref.so is there (see reference_arm)
        if oracle.ReferenceHierarchy.available():
            mg, kind = oracle.ReferenceHierarchy.from_oracle(mg), "reference"
    except (OSError, AttributeError, oracle.OracleError):
        kind = "port"
    setup_s = time.time() - t0
    mg.vcycle(1)                                  # warm caches / page in
    secs = mg.time_vcycles(cycles)
    return dict(sides=sides, setup_s=setup_s, s_per_cycle=secs / cycles, history=mg.history().tolist(), kind=kind)


def reference_arm(args):
    """The reference's own CPU implementation of the path (the oracle port, pinned bit-identical to the reference sources),
    MEASURED on the stated workload: set-up with every host thread (untimed; the reference's O(N^2) kNN is replaced by the
    bit-identical cell search), then `while (mg.residual() >= 1e-8) mg.vCycle();` (loop shape of FractionalStepSim.cpp:139-142)
    from a zero guess on ONE thread, because the reference is strictly serial.  The W warm-up and K timed steps are cycles
    W+1 .. W+K of that solve (a V-cycle costs the same whatever the iterate), each timed around exactly the call the
    reference times (testing_functions.cpp:340-344)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle

    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    sides = level_sides(args.side, args.levels)
    t0 = time.time()
    mg = oracle.make_hierarchy(sides, kind=oracle.KIND_DIRICHLET, fine_poly=args.fine_poly)
    kind = "port"
    if oracle.ReferenceHierarchy.available() and not args.ref_port:
        # oracle/_ref/libref.so = the reference's own grid.cpp / multigrid.cpp compiled where they lie: the cycles that are timed run
        # the reference's code; only the set-up (brute-force kNN, O(N^2) in the reference) is the oracle's, whose matrices are pinned
        # bit-identical to the reference's (tests/test_oracle_cpu.py)
        try:
            mg, kind = oracle.ReferenceHierarchy.from_oracle(mg), "reference"      # the oracle hierarchy is released with the last reference to it
        except (OSError, AttributeError, oracle.OracleError) as e:               # a libref.so that does not load or predates the raw entry points
            print("bench.py: reference objects unavailable (%s); timing the oracle port" % e, file=sys.stderr)
    setup_s = time.time() - t0
    try:                                           # the timed loop is one thread on one core (`taskset -c`), the last one this process may use
        os.sched_setaffinity(0, {max(os.sched_getaffinity(0))})
    except (AttributeError, OSError):
        pass
    per_cycle, t_solve0 = [], time.perf_counter()
    budget_s = max(0.0, args.ref_budget_s - setup_s)
    r = mg.residual()
    while r >= TOL and len(per_cycle) < 400:
        per_cycle.append(mg.time_vcycles(1))
        r = mg.residual()
        if time.perf_counter() - t_solve0 > budget_s and len(per_cycle) >= args.warmup + args.steps:
            break                                  # bounded run: the K timed cycles are complete, the solve is reported as unfinished
    solve_s = time.perf_counter() - t_solve0
    timed = per_cycle[args.warmup: args.warmup + args.steps]
    if len(timed) < args.steps:                    # converged before W+K cycles (small workloads): time what is left of K on the converged iterate
        timed += [mg.time_vcycles(1) for _ in range(args.steps - len(timed))]
    s_step = sum(timed) / len(timed)
    v = 1.0 / s_step
    hist = mg.history()
    rate = float((hist[min(len(hist) - 1, 20)] / hist[min(len(hist) - 1, 5)]) ** (1.0 / max(1, min(len(hist) - 1, 20) - min(len(hist) - 1, 5)))) if len(hist) > 6 else None
    what = ("the reference's OWN Multigrid::vCycle / Grid::sor / Grid::residual (oracle/_ref/libref.so: its grid.cpp and multigrid.cpp compiled "
            "where they lie, Eigen-subset shim, g++ -O2 -ffp-contract=off) on operators assembled by the CPU oracle's threaded set-up (bit-identical to "
            "the reference's own assembly, whose brute-force kNN is O(N^2))" if kind == "reference" else
            "CPU oracle (the reference's grid.cpp / multigrid.cpp restated, pinned bit-identical to the reference sources; g++ -O2 -ffp-contract=off)")
    sample = ("%s, MEASURED on the stated %dx%d hierarchy %s: lexicographic SOR omega=1.4, V-cycle loop on 1 thread because the "
              "reference is serial, pinned to one core (%d host cores present); set-up %.0f s on %d threads, untimed; steps = cycles %d..%d of the solve from a zero guess"
              % (what, args.side, args.side, sides, cores, setup_s, cores, args.warmup + 1, args.warmup + args.steps))
    cfg = workload_config(args, sides)
    cfg["smoother"] = "lexicographic SOR omega=1.4 (the reference's own smoother) -- the GPU arm's throughput mode is multicolour SOR omega=%g; compare `solve`" % args.mc_omega
    line = {
        "impl": "reference", "metric": "vcycles_per_s", "value": v, "unit": "V-cycles/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * s_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "solve": {"tol": TOL, "cycles": len(per_cycle), "seconds": solve_s, "final_residual": r, "converged": bool(r < TOL),
                  "mode": "lexicographic omega=1.4, 1 thread", "convergence_per_cycle": rate},
        "cpu_baseline": {"value": v, "unit": "V-cycles/s", "cores": 1, "kind": kind, "sample": sample, "host_cores": cores, "setup_s": setup_s},
        "e2e": {"value": v, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_config(args, sides):
    return {
        "workload": "2D unit-square jittered cloud %dx%d = %d nodes, manufactured Dirichlet Poisson, %d-level V-cycle (sides %s), "
                    "fine polyDeg %d / coarse polyDeg 3, nu=5, omega=1.4" % (args.side, args.side, args.side ** 2, len(sides), sides, args.fine_poly),
        "smoother": args.smoother + (" (omega=%g)" % args.mc_omega if args.smoother == "multicolour" else " (omega=1.4)"), "l2": "inputs larger than L2 (finest operator alone is >1 GB vs 126 MB L2)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=2000, help="finest lattice side (2000 -> 4M nodes)")
    ap.add_argument("--levels", type=int, default=None, help="number of levels (default: coarsen until the side is <= 16)")
    ap.add_argument("--fine-poly", type=int, default=4)
    ap.add_argument("--smoother", default="multicolour", choices=["multicolour", "lexicographic"])
    ap.add_argument("--mc-omega", type=float, default=0.8, help="relaxation factor of the multicolour (throughput) mode")
    ap.add_argument("--cpu-side", type=int, default=500, help="finest lattice side of the bounded CPU sample")
    ap.add_argument("--cpu-cycles", type=int, default=3)
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port even when oracle/_ref/libref.so (the reference's own sources) is there")
    ap.add_argument("--ref-budget-s", type=float, default=540.0, help="--impl reference: wall budget of the whole arm (set-up included); the solve stops there once "
                                                                       "the W+K cycles are done and is then reported as not converged")
    ap.add_argument("--partition-threshold", type=int, default=200000, help="multi-GPU: levels with fewer rows are replicated")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-lex", action="store_true")
    ap.add_argument("--skip-solve", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from meshlessmultigridpoisson_b200 import build, capi
    from meshlessmultigridpoisson_b200.problems import make_hierarchy

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solve path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build()
    capi.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sides = level_sides(args.side, args.levels)
    t0 = time.time()
    mg = make_hierarchy(sides, "dirichlet", args.fine_poly, device=local)
    mg.sync()
    setup_s = time.time() - t0
    if world > 1:
        # strong scaling: ONE problem, its large levels cut into contiguous row blocks across the ranks (NCCL halo exchange);
        # every rank assembled the full hierarchy (set-up is replicated this round), coarse levels stay replicated
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        mg.set_partition_threshold(args.partition_threshold)
        mg.init_comm(rank, world, uid[0])
    fast = args.smoother == "multicolour"
    mg.set_smoother(capi.MULTICOLOUR if fast else capi.LEXICOGRAPHIC)
    mg.set_arithmetic(capi.ARITH_FAST if fast else capi.ARITH_REFERENCE_ORDER)
    if fast:
        mg.set_omega(args.mc_omega)     # multicolour ordering is unstable at the reference's omega=1.4
    fine = mg.grid(-1)
    A = fine.A_size

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps, CUDA events on the solver's stream
    sampler = ClockSampler(local)
    sampler.wait_ready()
    mg.vCycle(args.warmup)
    barrier()
    mg.enable_timers(True)
    mg.reset_timers()
    launches0 = mg.launch_count()
    c0 = sampler.mark()
    ms = mg.time_vcycles(args.steps)
    c1 = sampler.mark()
    barrier()
    clocks = sampler.stop(c0, c1)
    gpu_launches = mg.launch_count() - launches0
    tm_fine = mg.timers(len(sides) - 1)
    tm_all = mg.timers(-1)
    mg.enable_timers(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step                  # one problem, however many GPUs share it (strong scaling)

    # ---- roofline of the dominant kernel: finest-level SOR sweep
    peak, peak_src = measured_peaks()
    sor = tm_fine["sor"]
    achieved = sor["bytes"] / (sor["ms"] * 1e-3) / 1e9 if sor["ms"] > 0 else 0.0
    shares = {k: round(v["ms"] / max(1e-9, sum(x["ms"] for x in tm_all.values())), 4) for k, v in tm_all.items()}
    traffic, per_launch = None, None
    kernel = capi.last_kernel(0)                                      # the instantiation the last smoothing call (finest level, up-sweep) launched
    if fast:
        per_launch = sor["bytes"] // max(1, 2 * args.steps)           # one launch = all nu sweeps of one smoothing call on the finest level (this rank's rows)
        if world == 1:
            traffic = ncu_traffic(kernel, args.side, args.fine_poly)  # from the committed ncu capture of this kernel on this workload, else null
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel + (" (finest level, this rank's row block: all colours of all nu sweeps in one launch; rows next to a cut are stored into the "
                                    "neighbour rank's vectors over NVLink peer memory; achieved = this rank's algorithmic bytes / CUDA-event time of init + sweep "
                                    "+ halo collection)" if world > 1 else
                                    " (finest level: all colours of all nu sweeps of one smoothing call in one cooperative launch over the colour-major packed "
                                    "operator, fed by cp.async.bulk through a shared-memory ring; achieved = algorithmic bytes of those launches / their "
                                    "CUDA-event time)") if fast else kernel + " (finest level)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch, "bytes_per_sweep": sor["bytes"] // max(1, 2 * 5 * args.steps),
                "sor_share_of_step": shares.get("sor"), "class_shares": shares,
                "per_class_GBps": {k: (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None) for k, v in tm_fine.items() if k != "other"},
                "whole_cycle_GBps": algorithmic_bytes_per_cycle(sides, args.fine_poly) / (ms_per_step * 1e-3) / 1e9}

    # ---- in-run correctness: the first cycles of the timed configuration against the CPU oracle's golden history of this workload
    gold = golden_history(args.side, args.fine_poly, "multicolour_omega0.8" if fast and args.mc_omega == 0.8 else "lexicographic_omega1.4" if not fast else "-")
    hist = mg.residuals_
    check = {"golden": None}
    if gold is not None and args.levels is None:
        m = min(len(gold), len(hist), 6)
        dev = float(np.max(np.abs(hist[:m] - np.array(gold[:m])) / np.array(gold[:m])))
        check = {"golden": "tests/golden/bench_%d_p%d.json (CPU oracle, own operators)" % (args.side, args.fine_poly), "cycles_compared": m, "max_rel_dev": dev,
                 "tol": 1e-6, "ok": bool(dev < 1e-6)}
        if not check["ok"]:
            raise SystemExit("bench.py: residual history of the timed configuration deviates from the oracle's golden history: %r vs %r" % (hist[:m].tolist(), gold[:m]))

    # ---- end to end through the C-ABI with host buffers (pinned), copies inside the timed region.  A rank of a partitioned
    # problem moves only what it works on: source_ of its row block, values_ of its block + halo up, values_ of its block down.
    own_lo, own_hi, need_lo, need_hi = mg.owned_range(-1) if world > 1 else (0, A, 0, A)
    src = torch.empty(own_hi - own_lo, dtype=torch.float64).pin_memory().numpy()
    val = torch.empty(need_hi - need_lo, dtype=torch.float64).pin_memory().numpy()
    whole = (need_lo, need_hi) == (own_lo, own_hi)      # unpartitioned: the block IS the vector, values_ comes straight back into `val`
    out = val if whole else torch.empty(own_hi - own_lo, dtype=torch.float64).pin_memory().numpy()
    if world > 1:
        mg.gather_values()                        # the host copy below must be the complete, current vector
    src[:] = fine.source_[own_lo:own_hi]
    val[:] = fine.values_[need_lo:need_hi]

    def e2e_step():
        fine.write_source_range(own_lo, src)      # H2D
        fine.write_values_range(need_lo, val)     # H2D
        mg.vCycle(1)
        fine.read_values_range(own_lo, out)       # D2H straight into the pinned buffer
        if not whole:
            val[own_lo - need_lo: own_hi - need_lo] = out
        return mg.residuals_[-1:]                 # D2H of the step's residual entry

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": 1.0 / e2e_s, "unit": "V-cycles/s", "h2d_bytes_per_step": 8 * (src.size + val.size), "d2h_bytes_per_step": 8 * out.size + 8,
           "note": "per rank: source_ of the rank's row block + values_ of block and halo up, values_ of the block and the residual entry down"}

    line = {
        "metric": "vcycles_per_s", "value": value, "unit": "V-cycles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, sides), "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline,
        "setup_s": setup_s, "check": check,
    }
    if world > 1:
        line["comm"] = dict(mg.comm_stats(), parallelism="row-block partition of levels >= %d rows; smoother halos travel inside the sweep kernel as stores "
                                                         "into the neighbour's vectors over NVLink peer memory (NCCL send/recv per colour phase if IPC is unavailable); "
                                                         "NCCL send/recv halos for residual / restriction / prolongation, allreduce for the norm; smaller levels "
                                                         "replicated" % args.partition_threshold)

    # ---- solve to 1e-8 from a zero guess (collective: every rank takes part when the problem is partitioned)
    if not args.skip_solve:
        fine.values_ = np.zeros(A)
        barrier()
        t0 = time.perf_counter()
        n, r = mg.solve(TOL, 400)
        mg.sync()
        barrier()
        line["solve"] = {"tol": TOL, "cycles": n, "seconds": time.perf_counter() - t0, "final_residual": r, "mode": args.smoother}
    if rank == 0:
        # ---- the reference-faithful mode, reported separately
        if fast and not args.skip_lex and world == 1:
            mg.set_smoother(capi.LEXICOGRAPHIC)
            mg.set_arithmetic(capi.ARITH_REFERENCE_ORDER)
            mg.set_omega(1.4)
            fine.values_ = np.zeros(A)
            mg.vCycle(1)
            lex_ms = mg.time_vcycles(2) / 2
            nl, _ = fine.lex_levels() if args.side <= 1000 else (None, None)
            line["lexicographic"] = {"value": 1e3 / lex_ms, "unit": "V-cycles/s", "ms_per_step": lex_ms, "dag_levels_finest": nl,
                                     "note": "dependency-DAG sweep, reference-order row sums; bounded by DAG depth x L2 latency, not HBM"}
        # ---- CPU baseline beside it (bounded sample)
        if not args.skip_cpu and world == 1:      # the CPU baseline is reported at N=1 only (rank 0 would keep the other ranks waiting)
            cores = os.cpu_count() or 1
            r = run_cpu_oracle(args.cpu_side, None, args.fine_poly, args.cpu_cycles, cores)
            scale = algorithmic_bytes_per_cycle(r["sides"], args.fine_poly) / algorithmic_bytes_per_cycle(sides, args.fine_poly)
            s_full = r["s_per_cycle"] / scale
            line["cpu_baseline"] = {
                "value": 1.0 / s_full, "unit": "V-cycles/s", "cores": 1, "kind": r["kind"], "host_cores": cores,
                "sample": "%s, lexicographic SOR, %d V-cycles on a %dx%d-side hierarchy %s (%.3f s/cycle measured), scaled by algorithmic "
                          "bytes x%.4g to the %dx%d workload; set-up used %d threads (%.1f s, untimed)"
                          % ("the reference's own Multigrid::vCycle (oracle/_ref/libref.so) on oracle-assembled operators" if r["kind"] == "reference" else "CPU oracle",
                             args.cpu_cycles, args.cpu_side, args.cpu_side, r["sides"], r["s_per_cycle"], 1 / scale, args.side, args.side, cores, r["setup_s"]),
            }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
