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
rr = list(csv.reader(open(raw))); h, units = rr[0], rr[1]
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
idx = [(h.index(w), lab) for w, lab in want if w in h]
kn = h.index("Kernel Name")
L += ["## `ncu --set full --clock-control none` (one launch of each kernel of a timed step)", "",
      "| kernel | " + " | ".join(f"{lab} [{units[i]}]" if units[i] else lab for i, lab in idx) + " |", "|---|" + "---:|" * len(idx)]
seen = set()
rr = rr[:2] + rr[:1:-1]          # newest launch of every kernel first: the steady-state one
for r in rr[2:]:
    n = r[kn].split('(')[0].replace("void ", "")
    if n in seen: continue
    seen.add(n)
    def fmt(x):
        try: return f"{float(x):.4g}"
        except ValueError: return x
    L.append(f"| `{n}` | " + " | ".join(fmt(r[i]) for i, _ in idx) + " |")
open(out, "w").write("\n".join(L) + "\n")
print(out, "written")
if kernels_csv:
    def to_bytes(v, unit):
        v = float(str(v).replace(',', ''))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    def to_ms(v, unit):
        v = float(str(v).replace(',', ''))
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)
    it, ir, iw = h.index("gpu__time_duration.sum"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    seen = set()
    with open(kernels_csv, "w") as f:
        f.write("kernel,time_ms,dram_read_bytes,dram_write_bytes\n")
        for r in rr[2:]:
            n = r[kn].split('(')[0].replace("void ", "")
            if n in seen: continue
            seen.add(n)
            f.write(f'"{n}",{to_ms(r[it], units[it]):.6f},{to_bytes(r[ir], units[ir]):.0f},{to_bytes(r[iw], units[iw]):.0f}\n')
    print(kernels_csv, "written")
