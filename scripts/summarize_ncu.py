"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.
usage: python scripts/summarize_ncu.py <round-tag> <launches.csv> <full.ncu-rep> [steps_in_capture]"""
import collections, csv, gzip, re, shutil, subprocess, sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name, v, unit = row.get("Kernel Name"), row.get("Metric Value"), row.get("Metric Unit")
    if not name or not v:
        continue
    t = float(v.replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    short = re.sub(r"\(.*", "", name)[:100]
    agg[short][0] += 1
    agg[short][1] += t
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launch_list_summary.md", "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 1` (gpu__time_duration.sum, --clock-control none)\n\n")
    f.write(f"{sum(v[0] for v in agg.values())} launches over {steps} steps (warm-up + timed + e2e); per-launch times are cold-cache and "
            f"serialised -> compare SHARES.  Total {tot/1e6:.1f} ms.\n\n| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        f.write(f"| `{k}` | {v[0]} | {v[1]/1e6:.2f} | {100*v[1]/tot:.1f}% | {v[1]/v[0]/1e3:.1f} |\n")
with open(launches, "rb") as fi, gzip.open(f"profiles/{tag}_launch_list.csv.gz", "wb") as fo:
    shutil.copyfileobj(fi, fo)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]
want = [w for w in want if w in idx]
with open(f"profiles/{tag}_ncu_full_summary.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(want)
    w.writerow([units[idx[c]] for c in want])
    for d in data:
        w.writerow([d[idx[c]][:110] for c in want])
print("wrote profiles/", tag)
