#!/usr/bin/env python
"""Benchmark of the exact simplex hot path (BASELINE.json metric: exact simplex pivots/sec).

One "step" = one complete exact solve (time-to-optimal) of the workload's LP: every pivot runs the
full hot path (pricing, pivot column, ratio test, rank-1 carry update, steepest-edge update).
`value` = pivots / device time with the problem resident in HBM; `e2e` = the same solve through the
C ABI from host buffers (create, upload, solve, result download inside the timed region).

  python bench.py --gpus N --steps K --warmup W [--workload sparse4k] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (m, n_struct, K bounding rows, dense, rule)
    "sparse4k": dict(m=4096, n_struct=8192, k_bounding=90, dense=False, nnz_per_col=8),      # config 4
    "dense16k": dict(m=16384, n_struct=32768, k_bounding=160, dense=True, nnz_per_col=0),    # config 5
    "sparse1k": dict(m=1024, n_struct=2048, k_bounding=60, dense=False, nnz_per_col=8),      # quick check
}


def make_problem(name, seed):
    from relp_b200.generators import bounded_lp
    w = WORKLOADS[name]
    return bounded_lp(w["m"], w["n_struct"], k_bounding=w["k_bounding"], nnz_per_col=w["nnz_per_col"],
                      dense=w["dense"], seed=seed)


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region: NVML in-process (the same counters
    nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.* prints; querying only these keeps
    the driver-lock contention with the solve low), falling back to an `nvidia-smi -lms` subprocess."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.samples = []          # (time, sm_mhz, sm_max_mhz, [reasons])
        self.period = float(os.environ.get("BENCH_SAMPLER_MS", "250")) / 1e3
        self.t0 = self.t1 = None
        self.max_mhz = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if self.device < len(ids) and ids[self.device].strip().isdigit():
                return int(ids[self.device])
        return self.device

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = pynvml
            self._nvml_sample()    # fail here rather than in the thread
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 str(int(self.period * 1e3)), "-i", str(self._physical_index())],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._smi_loop, daemon=True)
            self.th.start()
            # nvidia-smi takes a second or two to attach (driver locks held meanwhile): wait for its first
            # sample so that its start-up does not fall into the warm-up / timed region
            t_end = time.perf_counter() + 15.0
            while not self.samples and time.perf_counter() < t_end and self.proc.poll() is None:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _nvml_sample(self, with_reasons=True):
        # NVML queries contend with the CUDA driver for the device lock (sporadic stalls of the solve were
        # measured with three queries every 100 ms): the max clock is read once, the SM clock every period,
        # the throttle reasons every 4th period and at the end of the timed region
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        if self.max_mhz is None:
            self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        reasons = []
        if with_reasons:
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            reasons = [nm for nm, bit in self.REASONS if mask & bit]
        self.samples.append((time.perf_counter(), float(sm), self.max_mhz, reasons))

    def _nvml_loop(self):
        k = 0
        while not self.stop_flag:
            try:
                self._nvml_sample(with_reasons=(k % 4 == 0))
            except Exception:
                pass
            k += 1
            time.sleep(self.period)

    def _smi_loop(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            rs = [nm for (nm, _), v in zip(self.REASONS, f[5:9]) if v.lower().startswith("active")]
            self.samples.append((time.perf_counter(), sm, mx, rs))

    def window(self, t0, t1):
        """keep the samples taken inside the timed region [t0, t1] (the sampler is started before the warm-up);
        called right after the timed region: one last sample with the throttle reasons is taken here"""
        self.t0, self.t1 = t0, t1
        if self.nvml is not None:
            try:
                self._nvml_sample(with_reasons=True)
            except Exception:
                pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        inside = [x for x in self.samples if self.t0 is None or self.t0 <= x[0] <= self.t1 + self.period]
        if not inside:      # region shorter than one sampling period: take the samples nearest to it
            inside = self.samples[-2:]
        sm = sorted(x[1] for x in inside)
        mx = [x[2] for x in inside]
        reasons = sorted({r for x in inside for r in x[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi -lms"}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(problem, rule, budget_s=15.0, threads=0):
    """The oracle port (C++ restatement of relp's Carry<RationalBig, BasisInverseRows> path; the Rust reference
    cannot be built in this image) timed on this box's host cores on a bounded sample: the first P pivots of
    the same workload and trace, rule initialisation included.  threads = 0: all hardware threads (the column
    loops of pricing / steepest edge are independent; relp itself is single-threaded)."""
    from oracle import fast_oracle
    if not fast_oracle.available():
        raise RuntimeError("oracle/_build/libfast_oracle.so is missing: run __graft_entry__.build()")
    used = fast_oracle.set_threads(threads)
    out = fast_oracle.timed_sample(problem, rule, budget_s)
    out["cores"] = used
    out["host_cores"] = os.cpu_count()
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the box's host cores with all the host
    threads it can use, same config / metric.  One step = the first P pivots of the workload's LP (rule
    initialisation included), P fixed during the warm-up so that a step takes a few seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fast_oracle
    prob = make_problem(args.workload, 0)
    threads = fast_oracle.set_threads(0)
    per_step_s = max(2.0, min(8.0, 150.0 / max(args.steps + args.warmup, 1)))
    # size the prefix: the pivots the port completes within the per-step budget (early pivots are the cheapest
    # ones, so a prefix OVERSTATES the reference's pivots/s over the whole LP); the timed steps repeat exactly
    # this prefix
    fast_oracle.set_time_limit(per_step_s)
    try:
        probe = fast_oracle.solve_problem(prob, args.rule)
    finally:
        fast_oracle.set_time_limit(0)
    P = max(1, len(probe.trace))
    total_p, total_t, whole = 0, 0.0, False
    for step in range(args.warmup + args.steps):
        r = fast_oracle.solve_problem(prob, args.rule, max_pivots=P)
        if step >= args.warmup:
            total_p += len(r.trace)
            total_t += r.seconds
        whole = r.status != "pivot_limit"
    value = total_p / max(total_t, 1e-9)
    sample = (f"{'all' if whole else 'first'} {P if not whole else len(r.trace)} pivots of the same LP and trace per step "
              f"(incl. rule initialisation), C++ big-rational restatement of Carry<RationalBig, BasisInverseRows> "
              f"(oracle/fast_oracle.cpp), {threads} OpenMP threads over the independent column loops")
    line = {
        "impl": "reference", "metric": "exact simplex pivots/sec", "value": value, "unit": "pivots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "exact rational (arbitrary precision)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "pivots/s", "cores": threads, "host_cores": os.cpu_count(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pivots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    w = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: synthetic bounded integer LP m={w['m']} n={w['n_struct']}+{w['m']} "
                        f"slacks, K={w['k_bounding']} bounding rows, coefficients in [-100,100], "
                        f"{'dense' if w['dense'] else '8 nnz/col sparse'} never-binding rows, seed=0",
            "pivot_rule": args.rule, "step": "one exact solve to optimality (time-to-optimal)",
            "parallelism": "single GPU" if world == 1 else
            f"the same LP with its carry row-sharded and its pricing column-sharded over {world} GPUs (one process "
            f"per GPU, NCCL: all-gather of pricing / ratio-test candidates and work-vector partials, pivot row "
            f"replicated by the owner)",
            "l2": ("inputs larger than L2: the int8 constraint block (537 MB) is streamed by every pricing / steepest-"
                   "edge dot and the active carry block exceeds 126 MB from 8 limbs on; no flush needed"
                   if w["dense"] else
                   "every solve re-creates the context and re-uploads the problem, so no step starts with a warm L2; "
                   "the active carry block is smaller than L2 (stated, not flushed)")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="dense16k", choices=sorted(WORKLOADS))
    ap.add_argument("--rule", default="steepest_edge")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import relp_b200
    from relp_b200 import _lib
    _lib.load()   # fails loudly when the CUDA library is missing: there is no fallback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob = make_problem(args.workload, 0)     # N > 1: the SAME LP, its carry row-sharded over the ranks
    # pinned host buffers for the end-to-end leg (every step uploads the whole problem from them)
    names = ["colptr", "rowidx", "vals", "cost", "rhs"]
    for name in names:
        t = torch.from_numpy(getattr(prob, name)).pin_memory()
        setattr(prob, name, t.numpy())
        setattr(prob, "_pin_" + name, t)
    h2d = sum(getattr(prob, n).nbytes for n in names)
    if prob.dense_block is not None:
        t = torch.from_numpy(prob.dense_block).pin_memory()
        prob.dense_block = t.numpy()
        prob._pin_dense = t
        h2d += prob.dense_block.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def share_id():
        if world == 1:
            return None
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(relp_b200.solver.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    def step():
        nid = share_id()
        return relp_b200.solve_relaxation(prob, rule=args.rule, device=local, profile=prof_level, rank=rank,
                                          world=world, nccl_id=nid)

    prof_level = int(os.environ.get("BENCH_PROFILE", "1"))   # timed steps: CUDA events around K1 only (the roofline kernel)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        g = step()
    barrier()
    t0 = time.perf_counter()
    pivots = 0
    dev_ms = 0.0
    e2e_s = 0.0
    launches = 0
    NW = _lib.RG_NWIDTHS
    k1 = {key: [0.0] * NW for key in ("ms", "n", "bytes", "imads")}
    for _ in range(args.steps):
        g = step()
        assert g.status == "optimal"
        pivots += g.pivots
        dev_ms += g.device_ms
        e2e_s += g.seconds_total
        launches += g.stats["kernel_launches"]
        for k in range(NW):
            k1["ms"][k] += g.stats["k1_ms_at_limbs"][k]
            k1["n"][k] += g.stats["k1_launches_at_limbs"][k]
            k1["bytes"][k] += g.stats["k1_bytes_at_limbs"][k]
            k1["imads"][k] += g.stats["k1_imads_at_limbs"][k]
    barrier()
    wall = time.perf_counter() - t0
    sampler.window(t0, t0 + wall)
    clocks = sampler.stop()
    d2h = g.stats["limbs"] * 8 * (prob.m + 2) + 4 * prob.m

    stats = torch.tensor([dev_ms, e2e_s, float(pivots), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, e2e_max = mx[0].item(), mx[1].item()
        pivots_all, launches_all = float(pivots), sm[3].item()   # one job: every rank walks the same pivots
    else:
        dev_ms_max, e2e_max, pivots_all, launches_all = dev_ms, e2e_s, float(pivots), float(launches)

    # one extra (untimed) solve with CUDA events around every phase of the iteration: the phase table
    prof_level = 2
    gp = step()
    assert gp.trace == g.trace
    phase = list(gp.stats["phase_ms"])

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        # integer-pipe peak of the K1 instruction mix (IMAD.WIDE carry chains), measured live on this GPU
        import ctypes as C
        v = C.c_double()
        lib = _lib.load()
        imad_peak = None
        if lib.rg_measure_imad_peak(local, 1.0, C.byref(v)) == 0 and v.value > 0:
            imad_peak = v.value

        # K1 (the rank-1 Bareiss pivot) exactly as the TIMED steps ran it: per limb width, the algorithmic bytes
        # and multiply-adds of every launch (tracked per pivot from the list length and the exact-division
        # width in use) over its CUDA-event time.  Roofline time = max(bytes / HBM peak, IMAD / IMAD peak).
        by_limbs = {}
        widths = g.stats["limb_widths"]
        for k in range(NW):
            if k1["n"][k]:
                t = k1["ms"][k] * 1e-3
                gbs = k1["bytes"][k] / t / 1e9
                gim = k1["imads"][k] / t / 1e9
                ent = {"launches": int(k1["n"][k]), "avg_ms": k1["ms"][k] / k1["n"][k],
                       "alg_bytes_per_launch": k1["bytes"][k] / k1["n"][k],
                       "alg_imads_per_launch": k1["imads"][k] / k1["n"][k],
                       "GB/s": gbs, "hbm_frac": gbs / peak, "GIMAD/s": gim}
                if imad_peak:
                    ent["imad_frac"] = gim * 1e9 / imad_peak
                    ent["roofline_frac"] = max(ent["hbm_frac"], ent["imad_frac"])
                else:
                    ent["roofline_frac"] = ent["hbm_frac"]
                by_limbs[str(widths[k])] = ent
        dom = max(range(NW), key=lambda k: k1["ms"][k])
        Ldom = widths[dom]
        d = by_limbs[str(Ldom)]
        imad_bound = imad_peak is not None and d.get("imad_frac", 0) >= d["hbm_frac"]
        traffic = None
        traffic_src = None
        tp = os.path.join(ROOT, "profiles", "r2_k1_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            ent = tj.get(f"{args.workload}:L{Ldom}:world{world}")
            if ent:
                traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
                traffic_src = ent
        mode = "active-column mode (packed block)" if g.stats.get("active_columns", 0) else "dense carry"
        roof = {"bound": "imad" if imad_bound else "hbm",
                "kernel": (f"k_update_items<L={Ldom}> + k_bn_rows + cost-row k_update" if mode.startswith("active") and Ldom >= 8
                           else f"k_update<L={Ldom}>") + f" {mode}: rank-1 Bareiss pivot of the carry, as run by the timed steps",
                "achieved": d["GIMAD/s"] if imad_bound else d["GB/s"],
                "peak": imad_peak / 1e9 if imad_bound else peak,
                "unit": "GIMAD.WIDE/s" if imad_bound else "GB/s",
                "frac": d["roofline_frac"],
                "hbm": {"achieved": d["GB/s"], "peak": peak, "unit": "GB/s", "frac": d["hbm_frac"],
                        "peak_source": peak_src},
                "imad": {"achieved": d["GIMAD/s"], "peak": imad_peak / 1e9 if imad_peak else None,
                         "unit": "GIMAD.WIDE/s", "frac": d.get("imad_frac"),
                         "peak_source": "rg_measure_imad_peak: IMAD.WIDE carry-chain mix on registers, all SMs, "
                                        "measured in this run"},
                "bound_note": "roofline = the slower of limb bytes at HBM bandwidth and limb multiply-adds at "
                              "integer-pipe peak (north star); 'imad' = integer multiply pipe",
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": d["alg_bytes_per_launch"],
                "algorithmic_imads_per_launch": d["alg_imads_per_launch"],
                "by_limbs": by_limbs, "active_columns_final": g.stats.get("active_columns", 0),
                "share_of_step": sum(k1["ms"]) / dev_ms if dev_ms else None}
        line = {
            "metric": "exact simplex pivots/sec", "value": pivots_all / (dev_ms_max * 1e-3), "unit": "pivots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "int (two's complement multi-limb u64, 2-16 limbs)",
            "data": "synthetic", "config": workload_config(args, world),
            "time_to_optimal_ms": dev_ms_max / args.steps, "pivots_per_solve": pivots / args.steps,
            "limb_histogram": dict(zip(map(str, g.stats["limb_widths"]), g.stats["pivots_at_limbs"])),
            "promotions": g.stats["promotions"], "demotions": g.stats["demotions"],
            "e2e": {"value": pivots_all / e2e_max, "unit": "pivots/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roof, "wall_s": wall,
            "phase_ms_per_step": dict(zip(["column+ratio", "work_vector", "scalars", "k1_update", "se_update",
                                           "price+select", "se_update:finalize+side_stream_wait",
                                           "se_update:gamma_recurrence"], phase[:8])),
            "phase_note": "one extra untimed solve with CUDA events around every phase (ms per solve)",
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(prob, args.rule)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
