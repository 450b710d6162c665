"""Clock / power behaviour of the search kernel under sustained load (run under gpurun)."""
import json, os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-hummingbird-eval_b200"))
sys.path.insert(0, ROOT)
from hbird_b200 import ops
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
DEV = torch.device("cuda", 0)

def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                    pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)

def run(fn, seconds=2.0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    t = threading.Thread(target=sample, args=(stop, out)); t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < seconds:
        for _ in range(5): fn()
        n += 5
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); t.join()
    ms = e0.elapsed_time(e1) / n
    clk = sorted(c for c, _, _ in out); pw = sorted(p for _, p, _ in out)
    reasons = 0
    for _, _, r in out: reasons |= r
    return ms, clk[len(clk) // 2], pw[len(pw) // 2], hex(reasons)

res = {}
for (name, Q, N, d) in [("cfg2", 12544, 1024000, 384), ("d768_1M", 21904, 1024000, 768)]:
    g = torch.Generator(device="cuda").manual_seed(5)
    feats = torch.randn((N, d), generator=g, device=DEV)
    bank = ops.MemoryBank(d, 1, 1, N, 0, True)
    bank.append_soft(feats, torch.ones((N, 1), device=DEV), normalise=True); bank.finalize()
    q = torch.randn((Q, d), generator=g, device=DEV) * 3
    a = torch.randn((8192, 8192), device=DEV, dtype=torch.bfloat16); b = torch.randn((8192, 8192), device=DEV, dtype=torch.bfloat16)
    ms, clk, pw, rs = run(lambda: torch.matmul(a, b))
    res[f"{name}_cublas8192"] = dict(ms=ms, tflops=2 * 8192**3 / ms / 1e9, clk=clk, watts=pw, reasons=rs)
    for cg in (1, 2):
        for ab in (0, 2, 1):
            bank.configure_search(cta_group=cg); bank.tune_search(4 if cg == 2 else 0, ab)
            ms, clk, pw, rs = run(lambda: bank.search(q, 30, 64))
            tf = 2.0 * Q * N * d / ms / 1e9
            res[f"{name}_cg{cg}_ab{ab}"] = dict(ms=ms, tflops=tf, clk=clk, watts=pw, reasons=rs, mac_per_clk_sm=tf * 1e12 / 2 / (clk * 1e6) / 148)
    bank.close(); del feats
for k, v in res.items(): print(k, json.dumps(v))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "power_probe.json"), "w"), indent=1)
