"""Hot-spot table from an ncu report's source page (needs -lineinfo builds and `ncu --set full --import-source on`).
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [top=30] [chunk=100]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 100
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[2:]:
    try:
        d = dict(addr=r[col["Address"]], src=r[col["Source"]], samples=int(r[col["Warp Stall Sampling (All Samples)"]] or 0),
                 execd=int(r[col["Instructions Executed"]] or 0), thr=float(r[col["Avg. Threads Executed"]] or 0))
        d["st"] = {n: int(r[col[n]] or 0) for n in stalls}
        data.append(d)
    except Exception:
        pass
ts, ti = sum(d["samples"] for d in data), sum(d["execd"] for d in data)
print(f"{len(data)} instructions, {ts} samples, {ti} warp-instructions executed")
for lo in range(0, len(data), chunk):
    seg = data[lo:lo + chunk]
    s, i = sum(d["samples"] for d in seg), sum(d["execd"] for d in seg)
    agg = {n: sum(d["st"][n] for d in seg) for n in stalls}
    top3 = sorted(agg.items(), key=lambda kv: -kv[1])[:3]
    thr = sum(d["thr"] * d["execd"] for d in seg) / max(i, 1)
    print(f"{lo:5d}-{lo + len(seg) - 1:5d} samples {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%  thr {thr:4.1f}  " +
          " ".join(f"{k[6:]}={100 * v / max(s, 1):.0f}%" for k, v in top3))
print(f"top {top} instructions by samples")
for n, d in sorted(enumerate(data), key=lambda x: -x[1]["samples"])[:top]:
    t3 = sorted(d["st"].items(), key=lambda kv: -kv[1])[:2]
    print(f"#{n:5d} {100 * d['samples'] / ts:5.2f}% exec {d['execd']:10d} thr {d['thr']:4.1f}  {d['src'].strip()[:70]:70s} " + " ".join(f"{k[6:]}={v}" for k, v in t3))
