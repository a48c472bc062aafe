#!/usr/bin/env python
"""bench.py -- dual-simplex iterations/sec of the B200 engine on BASELINE.json's workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|small]

A *step* is one factorization cycle of the hot path: `cycle` dual simplex iterations plus the
refactorization + recompute (computePrimals/computeDuals) that ends the cycle.  `cycle` is the
engine's default refactorization interval max(ClpSimplex::defaultFactorizationFrequency, m/5)
= 2000 for m = 10 000 (an eta costs one 8m-byte panel column per solve here, a refactorization
O(k^3) flops, so the optimum interval is longer than the reference's 275).  W warm-up steps run
untimed, then exactly K steps are timed with CUDA events on the engine's stream (max over ranks).

Workload (N = 1): BASELINE.json configs[1] -- synthetic random LP m=10k n=100k 1% nnz, fp64,
dual steepest edge, no presolve / scaling / perturbation.  The timed window starts from a
mid-solve basis (tests/golden/c2_status_it12000.npz: the basis the CPU oracle reaches after
12 000 iterations) so that the nucleus of the basis has a representative size; inputs are
larger than L2 (CSC copy of A = 120 MB + factors), no L2 flush is needed.

For N > 1 the same LP is solved with column-sharded pricing (one process per GPU, one NCCL
all-gather of the tableau-row shards per pricing pass): strong scaling.

--impl reference times the reference's CPU implementation of the path.  coin-or/Clp cannot be
built here (its CoinUtils dependency is absent), so the arm runs the CPU restatement in oracle/
("kind": "port") on all host cores; a step is a bounded sample of 25 iterations of the same
window.
"""
from __future__ import annotations

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

WORKLOADS = {
    # name: (m, n, density, seed)
    "c2": (10000, 100000, 0.01, 20260923),
    "c3": (50000, 500000, 0.01, 20260924),
    "small": (1000, 10000, 0.01, 20260923),
}
REF_ITERS_PER_STEP = 25


def clp_default_frequency(m):
    # ClpSimplex::defaultFactorizationFrequency (src/ClpSimplex.cpp:11401-11431)
    return min(10000, 75 + m // 50 if m < 10000 else 75 + 200 + (m - 10000) // 150)


def default_cycle(m):
    # Engine::setupDevice (clp_b200/csrc/engine.cu): max(Clp default, m/5), capped at 2048
    return max(8, min(2048, max(clp_default_frequency(m), m // 5)))


def build_workload(name):
    from clp_b200 import generators as G

    m, n, dens, seed = WORKLOADS[name]
    lp = G.random_sparse_lp(m, n, dens, seed, name=f"rand-{m}x{n}")
    status, start = None, "all-slack basis"
    fx = os.path.join(ROOT, "tests", "golden", "c2_status_it12000.npz")
    if name == "c2" and os.path.exists(fx):
        z = np.load(fx)
        status = z["status"].astype(np.uint8)
        start = f"basis of the CPU oracle after {int(z['iterations'])} iterations (tests/golden/c2_status_it12000.npz)"
    return lp, status, start


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, of measured)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def run_reference(args, lp, status, start, cycle):
    """CPU arm: the oracle port on all host cores; rank 0 only."""
    from oracle.oracle import OracleSimplex

    cores = os.cpu_count() or 1
    o = OracleSimplex(lp)
    if status is not None:
        o.set_status(status)
    o.set_option("threads", cores)
    o.set_option("warmupIterations", args.warmup * REF_ITERS_PER_STEP)
    o.set_option("maximumIterations", (args.warmup + args.steps) * REF_ITERS_PER_STEP)
    o.dual()
    sec, its = o.timed_window()
    value = its / sec if sec > 0 else 0.0
    sample = (f"{its} iterations ({args.steps} steps x {REF_ITERS_PER_STEP}) of the same window, after "
              f"{args.warmup * REF_ITERS_PER_STEP} warm-up iterations; oracle/ port of Clp's dual path, "
              f"{cores} threads in price, LU solves serial")
    return {
        "metric": "dual_simplex_iterations_per_sec", "value": value, "unit": "iterations/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * sec / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": lp.name, "m": lp.m, "n": lp.n, "nnz": lp.nnz, "start": start,
                   "step": f"{REF_ITERS_PER_STEP} iterations (bounded sample of a {cycle}-iteration cycle)"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def _claim_stdout():
    """Rank 0 must print exactly ONE JSON line: park the real stdout and point fd 1 at stderr, so that
    library banners (NCCL prints its version to stdout) cannot precede or follow the line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    lp, status, start = (None, None, None)
    m = WORKLOADS[args.workload][0]
    cycle = default_cycle(m)

    if args.impl == "reference":
        if rank != 0:
            return 0
        lp, status, start = build_workload(args.workload)
        print(json.dumps(run_reference(args, lp, status, start, cycle)), file=out, flush=True)
        return 0

    os.environ.setdefault("NCCL_DEBUG", "WARN")  # NCCL's version banner would go to stdout
    import torch

    import clp_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (clp_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lp, status, start = build_workload(args.workload)

    def new_model(**params):
        s = clp_b200.ClpSimplex()
        s.loadLP(lp)
        if status is not None:
            s.copyinStatus(status)
        for k, v in params.items():
            s.setParameter(k, v)
        if world > 1:
            from clp_b200.sharding import broadcast_unique_id

            uid = clp_b200.ClpSimplex.ncclUniqueId() if rank == 0 else np.zeros(128, dtype=np.uint8)
            uid = broadcast_unique_id(uid, src=0)
            s.initSharding(rank, world, uid)
        return s

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = args.warmup, args.steps
    # ---------------- device-timed run: inputs resident in HBM before the window opens
    s = new_model(batch=args.batch, warmupIterations=W * cycle, maximumIterations=(W + K) * cycle,
                  factorizationFrequency=cycle)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    st = s.dual()
    barrier()
    clocks = sampler.stop()
    ms, its = s.timedWindow()
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = its / (ms / 1000.0) if ms > 0 else 0.0
    steps_done = its / cycle
    launches = s.kernelLaunches()
    status_after = st
    nucleus = s.nucleusSize()

    result = {
        "metric": "dual_simplex_iterations_per_sec", "value": value, "unit": "iterations/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / max(1e-9, steps_done),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": lp.name, "m": lp.m, "n": lp.n, "nnz": lp.nnz, "start": start,
                   "step": f"{cycle} iterations + 1 refactorization", "pricing": "dual steepest edge",
                   "presolve": "off", "scaling": "off", "perturbation": "off",
                   "l2": "inputs larger than L2 (CSC copy 120 MB + factors), no flush",
                   "parallelism": "single GPU" if world == 1 else f"column-sharded pricing x{world}, replicated factors",
                   "timed_iterations": its, "status_after_window": status_after, "nucleus_size": nucleus},
        "clocks": clocks, "gpu_launches": int(launches),
    }

    if rank == 0 and world == 1:
        # ---------------- per-kernel timing (CUDA events around single kernels, no graph replay)
        # one full factorization cycle, so that the eta-panel costs (which grow with the number of
        # updates since the last refactorization) are averaged the way the timed window sees them
        p = new_model(batch=16, timing=1, maximumIterations=cycle, factorizationFrequency=cycle)
        p.dual()
        ph = p.phaseTimes()
        ns = max(1.0, ph["samples"])
        k = p.nucleusSize()
        ldk = (k + 7) // 8 * 8
        peak, peak_src = measured_peak_gbs()
        nb = max(1, int((np.asarray(status) == 1)[: lp.n].sum())) if status is not None else 0
        kern = {
            # algorithmic bytes per launch (DESIGN.md "Roofline accounting")
            "price": (12.0 * lp.nnz + 4.0 * (lp.n + 1) + 1.0 * lp.n + 8.0 * lp.m + 8.0 * lp.n, ph["priceKernel"] / ns),
            # mean over the timed launches: the nucleus size changes when the accuracy gate forces a
            # refactorization inside the cycle
            "ftran_gemv": (ph["ftranGemvBytes"] / ns, ph["ftranGemv"] / ns),
            "btran_gemv": (ph["btranGemvBytes"] / ns, ph["btranGemv"] / ns),
        }
        dom = max(kern, key=lambda q: kern[q][1])
        b, t_ms = kern[dom]
        ach = b / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (same
        # workload, same nucleus size); null when the nucleus differs from the captured one
        traffic, traffic_src, capture = None, None, None
        tp = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
        if os.path.exists(tp):
            rec = json.load(open(tp)).get(dom)
            if rec:
                kc = int(rec.get("nucleus_size", 0))
                cap = {"price": kern["price"][0], "ftran_gemv": 8.0 * kc * ((kc + 7) // 8 * 8) + 48.0 * kc,
                       "btran_gemv": 8.0 * kc * ((kc + 7) // 8 * 8) + 16.0 * kc}[dom]
                # reported as roofline.traffic only when this run's launches have the size of the
                # captured one; otherwise the capture is quoted separately (traffic_capture)
                if abs(b - cap) <= 0.005 * cap:
                    traffic, traffic_src = rec["traffic"], rec["source"]
                capture = {"nucleus_size": kc, "algorithmic_bytes": cap, "traffic": rec["traffic"],
                           "traffic_over_algorithmic": rec["traffic"] / cap, "source": rec["source"]}
        result["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                              "frac": ach / peak, "frac_of_nominal_8TBps": ach / 8000.0, "traffic": traffic,
                              "traffic_source": traffic_src,
                              "traffic_capture": capture,
                              "peak_source": peak_src,
                              "bytes_per_launch": b, "ms_per_launch": t_ms, "nucleus_size": k,
                              "all": {q: {"bytes": v[0], "ms": v[1], "GBps": (v[0] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0)}
                                      for q, v in kern.items()},
                              "phase_us_per_iteration": {q: 1000.0 * ph[q] / ns for q in
                                                         ("chuzr", "btran", "price", "chuzc", "dualUpdate", "ftran", "update")},
                              "refactor_ms_total": ph["refactor"], "basic_structurals_at_start": nb}
        # ---------------- end to end through the C ABI with host buffers
        t0 = time.perf_counter()
        e = new_model(batch=args.batch, maximumIterations=min(K, 4) * cycle, factorizationFrequency=cycle)
        e.dual()
        x = e.primalColumnSolution(); e.dualRowSolution(); e.statusArray(); e.objectiveValue()
        wall = time.perf_counter() - t0
        nm = lp.n + lp.m
        h2d = (4 * (lp.n + 1) + 12 * lp.nnz) + (4 * (lp.m + 1) + 12 * lp.nnz) + 8 * 7 * nm + nm + 12 * lp.m
        d2h = 8 * 2 * nm + 8 * lp.m + nm + 4 * lp.m
        result["e2e"] = {"value": e.numberIterations() / wall, "unit": "iterations/s",
                         "h2d_bytes_per_step": int(h2d / max(1, min(K, 4))), "d2h_bytes_per_step": int(d2h / max(1, min(K, 4))),
                         "includes": "Clpb_loadProblem (pageable host arrays -> HBM), basis hand-over, "
                                     "Clpb_dual for min(K,4) steps, solution read-back", "wall_s": wall,
                         "iterations": e.numberIterations()}
        # ---------------- CPU baseline: the oracle port on the host cores, bounded sample
        from oracle.oracle import OracleSimplex

        cores = os.cpu_count() or 1
        o = OracleSimplex(lp)
        if status is not None:
            o.set_status(status)
        o.set_option("threads", cores)
        o.set_option("maximumSeconds", args.cpu_seconds)
        o.set_option("warmupIterations", 5)
        o.dual()
        sec, cits = o.timed_window()
        result["cpu_baseline"] = {"value": cits / sec if sec > 0 else 0.0, "unit": "iterations/s", "cores": cores,
                                  "kind": "port",
                                  "sample": f"first {cits} iterations ({sec:.1f} s) of the same window on the host; "
                                            "oracle/ restatement of Clp's dual path (coin-or/Clp itself cannot be "
                                            "built: CoinUtils absent)"}
    if rank == 0:
        print(json.dumps(result), file=out, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
