"""Per-CUDA-source-line instruction counts and stall samples of an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py <report.ncu-rep> <trips (warp-level) per launch> [top N]"""
import csv, io, subprocess, sys
rep, trips = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
lines, fname = [], ""
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; iline, isrc, iexe, ismp = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"); continue
    if len(r) < 10 or not r[iline] or not r[iline].isdigit(): continue
    try: lines.append((fname + ":" + r[iline], r[isrc], int(r[iexe]), int(r[ismp])))
    except ValueError: pass
tot = sum(x[2] for x in lines); ts = sum(x[3] for x in lines)
print(f"total warp instructions {tot:.3e} = {tot / trips:.1f} per trip; samples {ts}")
key = 3 if "--by-samples" in sys.argv else 2
for ln, s, e, m in sorted(lines, key=lambda x: -x[key])[:top]:
    print(f"{ln:>22} {e / trips:8.1f}/trip {e / tot * 100:5.1f}%  samples {m / ts * 100:5.1f}%  {s.strip()[:110]}")
