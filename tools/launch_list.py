"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (the share of each kernel in the step).
python tools/launch_list.py gpurun_out/x_launches.csv profiles/x_launch_list.txt "header comment" """
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].strip()
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu].strip(), 1.0)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
with open(sys.argv[2], "w") as f:
    for line in sys.argv[3].split("\\n"):
        f.write("# " + line + "\n")
    f.write("%-80s %8s %12s %8s\n" % ("kernel", "launches", "total us", "share"))
    for k in sorted(tot, key=lambda k: -tot[k]):
        f.write("%-80s %8d %12.1f %7.1f%%\n" % (k[:80], cnt[k], tot[k], 100 * tot[k] / total))
print(open(sys.argv[2]).read())
