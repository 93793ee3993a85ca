"""Which part of bench.py's harness perturbs the solve: torch's CUDA context, or the NVML sampler thread?"""
import os, sys, time, threading
sys.path.insert(0, '.')
import relp_b200, bench
prob = bench.make_problem("sparse4k", 0)

def run(tag, n=8):
    ts, tt = [], []
    for _ in range(n):
        g = relp_b200.solve_relaxation(prob, rule="steepest_edge", profile=1)
        ts.append(round(g.device_ms, 1)); tt.append(round(g.seconds_total * 1e3, 1))
    print(tag, "device ms", ts, "total ms", tt, flush=True)

run("plain          ")
stop = False
def poll(period):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        try: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        time.sleep(period)
th = threading.Thread(target=poll, args=(0.25,), daemon=True); th.start(); time.sleep(0.5)
run("nvml 250ms     ")
stop = True; th.join(); stop = False
import torch
torch.cuda.set_device(0); x = torch.zeros(4, device="cuda"); torch.cuda.synchronize()
run("torch          ")
th = threading.Thread(target=poll, args=(0.25,), daemon=True); th.start(); time.sleep(0.5)
run("torch+nvml     ")
stop = True
