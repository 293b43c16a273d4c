"""python scripts/ncu_source_lines.py REPORT.ncu-rep LAUNCH [TOP]: stall samples and executed warp-instructions of one profiled
launch aggregated by CUDA source line (needs a capture taken with --import-source on from a -lineinfo build).  Finds scalar
bookkeeping that a latency-bound warp cannot hide (this is how the run-time modulo in the decode exchange was found)."""
import csv, subprocess, sys

rep, launch = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
path, func, hdr, lines = None, None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr is not None and r[0].isdigit() and len(r) > i_e:
        try:
            lines.append((path, int(r[0]), r[1].strip(), int(r[i_s] or 0), int(r[i_e] or 0)))
        except ValueError:
            pass
print(func)
ts, te = sum(l[3] for l in lines), sum(l[4] for l in lines)
print(f"{len(lines)} source lines, {ts} samples, {te} warp-instructions executed")
for title, key in (("by stall samples", 3), ("by executed warp-instructions", 4)):
    print(f"--- {title}")
    for l in sorted(lines, key=lambda l: -l[key])[:top]:
        print(f"{l[3]:6d} samples {l[4]:9d} exec  {l[0]}:{l[1]}  {l[2][:110]}")
