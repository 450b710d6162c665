"""Per-CUDA-source-line stall samples from an ncu report (cuda,sass correlated view)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, h, ci, recs = None, None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; ci = {n: i for i, n in enumerate(h)}; continue
    if h is None or len(r) < len(h) or r[0] == "" or r[0] in ("Function Name",): continue
    try: n = int(r[ci["# Samples"]] or 0)
    except ValueError: continue
    recs.append((n, cur_file, r))
tot = sum(n for n, _, _ in recs)
print("total samples", tot)
recs.sort(key=lambda t: -t[0])
for n, f, r in recs[:top]:
    st = {k[6:]: int(r[ci[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k and (r[ci[k]] or "0") not in ("0", "")}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:7d} {100*n/tot:5.1f}% {f}:{r[0]:>4s} inst={r[ci['Instructions Executed']]:>10s} {r[1].strip()[:78]:78s} {st}")
