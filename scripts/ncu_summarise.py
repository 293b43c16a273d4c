"""Turn the ncu files a GPU call brought back (gpurun_out/) into the tracked evidence under profiles/:
  <round>_ncu_full_summary.csv   selected metrics of every kernel in the `ncu --set full` captures
  <round>_traffic.json           DRAM read+write bytes per launch per kernel (bench.py reports it as roofline.traffic)
  <round>_launches_bench_l8.summary.txt   per-kernel shares from the launch list of the bench command
Runs here (no GPU): `ncu -i file.ncu-rep --page raw --csv` only reads the report.
  python scripts/ncu_summarise.py r2 [gpurun_out]"""
import csv
import glob
import gzip
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short_name(full):
    m = re.search(r"(\w+)(<[^>]*>)?\(", full)
    return (m.group(1) + (m.group(2) or "")) if m else full


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    head, units, body = rows[0], rows[1], rows[2:]
    return head, units, body


def full_summary(rnd, src):
    reps = sorted(glob.glob(os.path.join(src, f"{rnd}_prof_*.ncu-rep")))
    if not reps:
        print("no", f"{rnd}_prof_*.ncu-rep", "under", src)
        return
    traffic = {}
    with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_full_summary.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["capture"] + COLS)
        for rep in reps:
            head, units, body = raw_rows(rep)
            idx = {c: head.index(c) for c in COLS if c in head}
            for r in body:
                cell = lambda c: (f"{r[idx[c]]} {units[idx[c]]}".strip() if c in idx else "")  # noqa: E731
                w.writerow([os.path.basename(rep).replace(".ncu-rep", "")] + [cell(c) for c in COLS])
                if "dram__bytes_read.sum" in idx:
                    b = sum(float(r[idx[c]].replace(",", "")) * UNIT.get(units[idx[c]], 1.0)
                            for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    t = traffic.setdefault(short_name(r[idx["Kernel Name"]]), {"sum": 0.0, "n": 0})
                    t["sum"] += b
                    t["n"] += 1
    js = {k: {"dram_bytes_per_launch_avg": v["sum"] / v["n"], "launches_captured": v["n"]} for k, v in traffic.items()}
    json.dump(js, open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w"), indent=1)
    print("wrote", f"profiles/{rnd}_ncu_full_summary.csv", f"profiles/{rnd}_traffic.json", list(js))


def launch_shares(rnd, src, tail=0.25):
    cands = glob.glob(os.path.join(src, f"{rnd}_launches_*.csv*"))
    for path in cands:
        text = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
        text = text[text.index('"ID"'):]
        agg, total = {}, 0.0
        recs = [r for r in csv.DictReader(io.StringIO(text)) if r.get("Metric Name") == "gpu__time_duration.sum"]
        # `bench.py --steps 1 --warmup 3` runs four generate() passes after the weight init: the last quarter of the
        # list is (a little more than) the one timed pass
        for r in recs[int(len(recs) * (1.0 - tail)):]:
            v = float(r["Metric Value"].replace(",", ""))
            us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v * 1e3 if r["Metric Unit"] in ("ms", "msecond") else v
            key = (short_name(r["Kernel Name"])[:64], r["Grid Size"])
            a = agg.setdefault(key, [0, 0.0])
            a[0] += 1
            a[1] += us
            total += us
        base = os.path.basename(path).split(".csv")[0]
        with open(os.path.join(ROOT, "profiles", base + ".summary.txt"), "w") as fh:
            fh.write(f"last {tail:.0%} of {len(recs)} launches in {os.path.basename(path)} (cold-cache, serialised: compare SHARES)\n")
            fh.write(f"{'kernel':64s} {'grid':>16s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}\n")
            for (k, g), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"{k:64s} {g:>16s} {n:8d} {us / 1e3:10.3f} {us / total * 100:6.1f}% {us / n:9.1f}\n")
        dst = os.path.join(ROOT, "profiles", os.path.basename(path) if path.endswith(".gz") else os.path.basename(path) + ".gz")
        if path.endswith(".gz"):
            open(dst, "wb").write(open(path, "rb").read())
        else:
            gzip.open(dst, "wt").write(text)
        print("wrote", f"profiles/{base}.summary.txt", dst)


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
    src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
    full_summary(rnd, src)
    launch_shares(rnd, src, float(sys.argv[3]) if len(sys.argv) > 3 else 0.25)
