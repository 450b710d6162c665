"""Summarise an ncu report of search_topk_kernel: headline metrics + stall samples per source line."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
def get(name):
    return vals[hdr.index(name)] if name in hdr else None
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]:
    print(f"{k:75s} {get(k)}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] in ("Line #", "#", "Address")]
if not hi:
    print(rows[:3]); sys.exit()
h = rows[hi[0]]
ci = {n: i for i, n in enumerate(h)}
print(h[:6])
tot = 0
lines = []
for r in rows[hi[0] + 1:]:
    if len(r) < len(h): continue
    try: n = int(r[ci["# Samples"]] or 0)
    except ValueError: continue
    tot += n
    lines.append((n, r))
lines.sort(key=lambda t: -t[0])
print("total samples", tot)
for n, r in lines[:40]:
    st = {k[6:]: int(r[ci[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k and (r[ci[k]] or "0") != "0"}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:7d} {100*n/tot:5.1f}%  L{r[0]:>4s} {r[1][:90]:90s} {st}")
