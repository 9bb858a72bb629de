"""Top stall sites of one launch in an ncu report (SASS view of `--page source`).
usage: python tools/ncu_stalls.py report.ncu-rep launch_index [top]"""
import csv, subprocess, sys
rep, li = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
h = rows[1]
idx = {k: i for i, k in enumerate(h)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(h):
        data.append(r)
S = idx["# Samples"]
tot = sum(int(r[S]) for r in data)
print("samples", tot, "sass lines", len(data), "warp-instructions", sum(int(r[idx["Instructions Executed"]]) for r in data))
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[idx[k]]) for r in data) for k in stalls}
print(", ".join("%s %d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[S]))[:top_n]:
    st = sorted([(int(r[idx[k]]), k[6:]) for k in stalls], reverse=True)[:2]
    print(r[S].rjust(6), r[idx["Instructions Executed"]].rjust(9), r[idx["Source"]].strip()[:66].ljust(66), st)
