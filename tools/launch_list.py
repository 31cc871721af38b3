"""ncu --metrics gpu__time_duration.sum --csv launch list -> per-kernel totals + the list (profiles/*_launch_list_*.txt)."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
iK, iG, iB, iM, iV, iU = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value", "Metric Unit"))
tot, cnt, lines = defaultdict(float), defaultdict(int), []
for n, r in enumerate(rows[1:]):
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    ms = v / 1e6 if r[iU] in ("ns", "nsecond") else v / 1e3 if r[iU] in ("us", "usecond") else v
    k = r[iK][:90]
    tot[k] += ms; cnt[k] += 1
    lines.append(f"{len(lines)},{k},{r[iG]},{r[iB]},{ms:.4f}")
total = sum(tot.values())
print("# summary: total_ms launches kernel")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"# {tot[k]:10.3f} ms {cnt[k]:4d} {100 * tot[k] / total:5.1f}%  {k}")
print("id,kernel,grid,block,ms")
print("\n".join(lines))
