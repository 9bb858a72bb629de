"""Aggregate an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv ; python tools/ncu_lines.py cs.csv <kernel substr> [top]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
hdr = None; cur_file = None; func = None
agg = collections.Counter(); samp = collections.Counter(); src = {}; curline = None; seen = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or func is None or want not in func: continue
    if r[0].isdigit():
        curline = (cur_file, int(r[0])); src[curline] = r[1]
    elif len(r) > ii and r[2].startswith("0x") and r[ii].isdigit():
        agg[curline] += int(r[ii]); samp[curline] += int(r[si])
tot = sum(agg.values()); ts = sum(samp.values())
print("total inst", tot, "samples", ts)
for k, v in agg.most_common(top):
    print(f"{k[0]}:{k[1]:4d} inst {v:10d} {100*v/tot:5.1f}%  samp {100*samp[k]/max(ts,1):5.1f}%  {src[k].strip()[:100]}")
