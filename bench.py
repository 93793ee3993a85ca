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


def cpu_baseline(problem, rule, budget_s=15.0):
    """The oracle port timed on this box's host cores (1 thread: relp is single-threaded) on a
    bounded sample: the first P pivots of the same workload and trace."""
    try:
        from oracle import fast_oracle
        if fast_oracle.available():
            return fast_oracle.timed_sample(problem, rule, budget_s)
    except ImportError:
        pass
    from oracle import relp_oracle as ro
    from tests.common import provider_from_problem
    provider = provider_from_problem(problem)
    t0 = time.perf_counter()
    probe = 3
    trace = ro.Trace(limit=probe)
    try:
        ro.solve_relaxation(provider, rule, trace)
    except ro.PivotLimit:
        pass
    t_probe = time.perf_counter() - t0
    done = len(trace.pivots)
    per = max(t_probe / max(done, 1), 1e-6)
    P = int(max(probe, min(10000, budget_s / per)))
    t0 = time.perf_counter()
    trace = ro.Trace(limit=P)
    try:
        ro.solve_relaxation(provider, rule, trace)
    except ro.PivotLimit:
        pass
    dt = time.perf_counter() - t0
    n = len(trace.pivots)
    return {"value": n / dt, "unit": "pivots/s", "cores": 1, "kind": "port",
            "sample": f"first {n} pivots of the same LP and trace (incl. rule initialisation), "
                      f"Python fractions.Fraction oracle, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Rust reference cannot be built in
    this image) on the box's host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = make_problem(args.workload, 0)
    total_p, total_t, last = 0.0, 0.0, None
    for step in range(args.warmup + args.steps):
        # bounded sample per step: the whole run stays within a few minutes for any --steps
        # (the sample is a PREFIX of the trace and early pivots are cheap -- small numbers -- so a short sample
        # overstates the reference: 69 pivots/s over the first 80 pivots vs 13 over the whole LP; the budget is
        # therefore kept as large as a few-minute run allows: ~17 s per step up to K = 14, less beyond)
        budget = min(10.0, 140.0 / max(args.steps, 1)) if step >= args.warmup else 1.0
        t0 = time.perf_counter()
        last = cpu_baseline(prob, args.rule, budget_s=budget)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            total_t += dt
            total_p += last["value"]
    value = total_p / max(args.steps, 1)
    line = {
        "impl": "reference", "metric": "exact simplex pivots/sec", "value": value, "unit": "pivots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "exact rational (arbitrary precision)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": dict(last, value=value),
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
            f"carry row-sharded over {world} GPUs (one process per GPU, NCCL: all-gather of ratio-test candidates "
            f"and work-vector partials, all-reduce of the pivot row); pricing and rule update replicated",
            "l2": "carry (>= 268 MB at 2 limbs) exceeds the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="sparse4k", choices=sorted(WORKLOADS))
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
    # pinned host buffers for the end-to-end leg
    for name in ("colptr", "rowidx", "vals", "cost", "rhs"):
        t = torch.from_numpy(getattr(prob, name)).pin_memory()
        setattr(prob, name, t.numpy())
        setattr(prob, "_pin_" + name, t)
    h2d = sum(getattr(prob, n).nbytes for n in ("colptr", "rowidx", "vals", "cost", "rhs"))

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

    prof_level = 1      # timed steps: CUDA events around K1 only (the roofline kernel)

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
    k1_ms = [0.0] * 5
    k1_n = [0] * 5
    phase = [0.0] * 8
    for _ in range(args.steps):
        g = step()
        assert g.status == "optimal"
        pivots += g.pivots
        dev_ms += g.device_ms
        e2e_s += g.seconds_total
        launches += g.stats["kernel_launches"]
        for k in range(5):
            k1_ms[k] += g.stats["k1_ms_at_limbs"][k]
            k1_n[k] += g.stats["k1_launches_at_limbs"][k]
        for k in range(8):
            phase[k] += g.stats["phase_ms"][k]
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

    # one extra (untimed) solve with the active-column mode switched off: the dense rank-1 kernel is the one
    # whose algorithmic bytes are 16 L (m+1)^2; its CUDA-event times give the dense roofline
    # one extra (untimed) solve with events around every phase of the iteration: the phase table
    prof_level = 2
    gp = step()
    assert gp.trace == g.trace
    phase = list(gp.stats["phase_ms"])
    gd = None
    if world == 1:
        gd = relp_b200.solve_relaxation(prob, rule=args.rule, device=local, profile=True, dense_carry=True)
        assert gd.trace == g.trace and gd.objective == g.objective   # both carry modes walk the same pivots

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        w = WORKLOADS[args.workload]
        rows_local = -(-w["m"] // world) + 1
        entries = (w["m"] + 1) * rows_local     # carry entries one dense K1 launch of one rank covers

        def k1_table(ms, cnt, ent):
            out = {}
            for k in range(5):
                if cnt[k]:
                    L = 1 << k
                    avg = ms[k] / cnt[k]
                    gbs = 16.0 * L * ent / (avg * 1e-3) / 1e9
                    out[str(L)] = {"launches": cnt[k], "avg_ms": avg, "GB/s": gbs, "frac": gbs / peak}
            return out

        dense_tab = k1_table(gd.stats["k1_ms_at_limbs"], gd.stats["k1_launches_at_limbs"], entries) if gd else {}
        dom = max(range(5), key=lambda k: k1_ms[k])
        Ldom = 1 << dom
        # active-column mode touches (rows) x (non-trivial columns) entries; report its achieved bandwidth on
        # that footprint with the final list length as the (upper-bound) column count
        nk = g.stats.get("active_columns", 0) or 1
        list_entries = rows_local * nk
        list_tab = k1_table(k1_ms, k1_n, list_entries)
        if dense_tab and str(Ldom) in dense_tab:
            ach = dense_tab[str(Ldom)]["GB/s"]
            kern = f"k_update<L={Ldom}> dense mode (rank-1 Bareiss pivot of the whole carry)"
            alg = 16 * Ldom * entries
        else:
            ach = list_tab[str(Ldom)]["GB/s"]
            kern = f"k_update<L={Ldom}> active-column mode"
            alg = 16 * Ldom * list_entries
        # DRAM traffic of one launch from the committed `ncu --set full` capture of exactly this kernel and
        # workload (profiles/r1_ncu_summaries.md, prof_k1_dense_v3: dram read 1.080858 GB + write 10.13 MB)
        traffic = None
        if args.workload == "sparse4k" and world == 1 and Ldom == 8 and kern.endswith("whole carry)"):
            traffic = 1080858000 + 10126848
        roof = {"bound": "hbm", "kernel": kern, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_note": "ncu dram__bytes_read+write per launch (one capture, profiles/); below the "
                                "algorithmic bytes because unchanged zero entries are not written back: "
                                "traffic / time is the honest HBM utilisation",
                "traffic_frac": (traffic / (dense_tab[str(Ldom)]["avg_ms"] * 1e-3) / 1e9 / peak) if traffic else None,
                "algorithmic_bytes_per_launch": alg,
                "dense_mode_by_limbs": dense_tab,
                "dense_mode_note": "one untimed solve with dense_carry=1; zero entries are read but neither "
                                   "multiplied nor written back, so achieved can exceed the copy peak",
                "active_column_mode_by_limbs": list_tab, "active_columns_final": nk,
                "share_of_step": sum(k1_ms) / dev_ms if dev_ms else None}
        line = {
            "metric": "exact simplex pivots/sec", "value": pivots_all / (dev_ms_max * 1e-3), "unit": "pivots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "int (two's complement multi-limb u64, 2-16 limbs)",
            "data": "synthetic", "config": workload_config(args, world),
            "time_to_optimal_ms": dev_ms_max / args.steps, "pivots_per_solve": pivots / args.steps,
            "limb_histogram": g.stats["pivots_at_limbs"],
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
