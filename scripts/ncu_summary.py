"""Turn the ncu exports of a gpurun call into the markdown committed under profiles/.
usage: python scripts/ncu_summary.py <launches.csv> <raw.csv (ncu -i X.ncu-rep --page raw --csv)> <out.md> <title> [kernels.csv]
kernels.csv (optional): one row per kernel -- time, DRAM bytes read / written per launch -- the table bench.py reads
`roofline.traffic` from."""
import csv, collections, sys
launches, raw, out, title = sys.argv[1:5]
kernels_csv = sys.argv[5] if len(sys.argv) > 5 else None
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split('(')[0].replace("void ", ""); v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
L = [f"# {title}", "", "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)", "",
     "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    L.append(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {t / tot:.3f} |")
L += ["", f"{len(rows) - 1} launches, {tot / 1e3:.1f} ms of kernel time.", ""]
# one or more raw exports (comma separated); every file has its own unit row, so everything is normalised here
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
want = [("gpu__time_duration.sum", "time [ms]"), ("dram__bytes_read.sum", "DRAM rd [GB]"), ("dram__bytes_write.sum", "DRAM wr [GB]"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
def num(x):
    return float(str(x).replace(',', ''))
records = collections.OrderedDict()   # kernel -> {label: value}; later files / later launches overwrite (steady state)
for path in raw.split(","):
    rr = list(csv.reader(open(path))); h, units = rr[0], rr[1]
    kn = h.index("Kernel Name")
    for r in rr[2:]:
        n = r[kn].split('(')[0].replace("void ", "").replace("fg::", "")
        rec = {}
        for w, lab in want:
            if w not in h: continue
            i = h.index(w)
            try: v = num(r[i])
            except ValueError: continue
            if lab.startswith("time"): v *= TIME.get(units[i], 1.0)
            elif "[GB]" in lab: v *= BYTES.get(units[i], 1.0) / 1e9
            rec[lab] = v
        records[n] = rec
labs = [lab for _, lab in want]
L += ["## `ncu --set full --clock-control none` (the launch of each kernel in a steady-state step; scripts/ncu_target.py)", "",
      "| kernel | " + " | ".join(labs) + " |", "|---|" + "---:|" * len(labs)]
for n, rec in records.items():
    L.append(f"| `{n}` | " + " | ".join(f"{rec[l]:.4g}" if l in rec else "" for l in labs) + " |")
open(out, "w").write("\n".join(L) + "\n")
print(out, "written")
if kernels_csv:
    with open(kernels_csv, "w") as f:
        f.write("kernel,time_ms,dram_read_bytes,dram_write_bytes\n")
        for n, rec in records.items():
            f.write(f'"{n}",{rec.get("time [ms]", 0):.6f},{rec.get("DRAM rd [GB]", 0) * 1e9:.0f},{rec.get("DRAM wr [GB]", 0) * 1e9:.0f}\n')
    print(kernels_csv, "written")
