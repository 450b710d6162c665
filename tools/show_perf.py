import json, sys
r = json.load(open('/root/repo/gpurun_out/probe_perf.json'))
for k, v in r.items():
    if k.startswith('perf_'):
        print(f"{k:36s} ms={v['ms']:8.3f} frac={v['frac_sustained']:.3f} tflops={v['tflops']:.0f}")
