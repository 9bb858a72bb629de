"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/launch_summary.py launches.csv [top]"""
import csv, collections, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
t = collections.Counter(); c = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v  # -> us
    name = re.sub(r"\(.*", "", r[ki])[:70]
    t[name] += v; c[name] += 1
tot = sum(t.values())
print("total_us %.1f launches %d" % (tot, sum(c.values())))
for k, v in t.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%10.1f us %5.1f%% %5d  %s" % (v, 100 * v / tot, c[k], k))
