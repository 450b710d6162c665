"""Summarise ncu --set full reports (gpurun_out/prof_*.ncu-rep) into one JSON kept under profiles/.
Usage: ncu_summary.py out.json name=report.ncu-rep ..."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_per_sm",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_memory_wavefronts",
}
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"kernel": vals[hdr.index("Kernel Name")][:120], "report": rep.split("/")[-1]}
    for k, short in KEYS.items():
        if k in hdr:
            i = hdr.index(k)
            try:
                rec[short] = float(vals[i].replace(",", ""))
            except ValueError:
                rec[short] = vals[i]
            rec[short + "_unit"] = units[i]
    out[name] = rec
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, v in out.items():
    print(k, {a: b for a, b in v.items() if not a.endswith("_unit")})
