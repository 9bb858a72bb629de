"""Per-role summary of a warp-specialised kernel in an ncu report: sync sites in SASS order with samples.
usage: python tools/dbg/ncu_roles.py report.ncu-rep launch_index"""
import csv, subprocess, sys
rep, li = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
h = rows[1]; idx = {k: i for i, k in enumerate(h)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(h):
        data.append(r)
S, I, src = idx["# Samples"], idx["Instructions Executed"], idx["Source"]
print("lines", len(data), "samples", sum(int(r[S]) for r in data), "instr", sum(int(r[I]) for r in data))
keys = ("NANOSLEEP", "UTCBAR", "BAR.SYNC", "SYNCS.ARRIVE", "LDTM", "LDGDEPBAR", "DEPBAR")
last = 0
for n, r in enumerate(data):
    s = r[src]
    if any(k in s for k in keys) and (int(r[I]) > 0 or int(r[S]) > 0):
        seg = data[last:n]
        print("   [%5d..%5d] samples %6d instr %10d" % (last, n, sum(int(x[S]) for x in seg), sum(int(x[I]) for x in seg)))
        print(n, r[S].rjust(6), r[I].rjust(9), s.strip()[:70])
        last = n + 1
