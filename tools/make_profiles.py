"""Turn the raw outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/make_profiles.py [round]"""
import csv, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True, cwd=ROOT).stdout


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


for src, dst in (("bench_r01.json", "bench_%s_sample16.json"), ("bench_r01_train.json", "bench_%s_train8.json"),
                 ("bench_n2.json", "bench_%s_sample16_n2.json"), ("bench_ref.json", "bench_%s_reference_arm.json")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        with open(os.path.join(P, dst % R), "w") as f:
            f.write(json.dumps(last_json(p)) + "\n")

for src, dst, head in (
        ("launches_sample.csv", "launches_%s_sample16.txt",
         "# ncu --metrics gpu__time_duration.sum --clock-control none : python tools/prof_sample.py 16 2\n"
         "# two Generator forwards (16 events, 256x256, bf16) incl. module construction; cold-cache serialised times: compare SHARES\n"
         "# thin::conv_thin_kernel<CPR, IS3, NB, MT>: CPR = 16-byte chunks per input pixel, IS3 = 3x3, NB = Cout/16, MT = sub-tiles per macro tile\n"),
        ("launches_train.csv", "launches_%s_train8.txt",
         "# ncu --metrics gpu__time_duration.sum --clock-control none : python tools/prof_train.py 8 2\n"
         "# two full G+D train steps (8 events, 256x256, bf16); cold-cache serialised times: compare SHARES\n")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        with open(os.path.join(P, dst % R), "w") as f:
            f.write(head + run(sys.executable, "tools/launch_summary.py", p, "60"))

for src, dst in (("layers_train_full.txt", "layers_%s_train8.txt"), ("layers_sample.txt", "layers_%s_sample16.txt")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        with open(p) as f:
            body = [l for l in f.read().splitlines() if not l.startswith(("Param count", "Adding attention"))]
        with open(os.path.join(P, dst % R), "w") as f:
            f.write("# python tools/prof_layers.py: every C-ABI call of one step timed with CUDA events on the launching stream;\n"
                    "# GB/s = algorithmic bytes (input + output of the layer, bf16) / call time\n" + "\n".join(body[:140]) + "\n")

rep = os.path.join(G, "prof_thin.ncu-rep")
if os.path.exists(rep):
    out = run("ncu", "-i", rep, "--page", "raw", "--csv")
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    best = None
    with open(os.path.join(P, "ncu_top_kernel_%s.txt" % R), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:conv_thin_kernel -s 12 -c 12 : python tools/prof_sample.py 4 2\n"
                "# the macro-tile tcgen05 conv launches of the second Generator forward (n = 160 images); the 16->16 3x3 @256x256\n"
                "# layer of bench.py's `roofline` is conv_thin_kernel<2, 1, 1, 4, 2, 1> (CPR, IS3, NB, MT, EPI, TMA) with ~336 MB read\n")
        for r in rows[2:]:
            d = dict(zip(h, r))
            f.write("\n%s\n" % d["Kernel Name"])
            for k in keys:
                if k in d:
                    f.write("  %-70s %s %s\n" % (k, d[k], rows[1][h.index(k)]))
            if "conv_thin_kernel<2, 1, 1, 4," in d["Kernel Name"].replace("(int)", "").replace("(bool)", ""):
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
                ur, uw = rows[1][h.index("dram__bytes_read.sum")], rows[1][h.index("dram__bytes_write.sum")]
                mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * mul[ur] + wr * mul[uw]
                if best is None or rd * mul[ur] > best[1]:
                    best = (tot, rd * mul[ur], float(d["gpu__time_duration.sum"]), rows.index(r) - 2, d["Kernel Name"])
    if best:
        with open(os.path.join(P, "top_kernel_traffic.json"), "w") as f:
            json.dump({"kernel": "thin::%s 16->16 3x3 @256x256, n=160 (+BN/ReLU prologue, stats epilogue)" % best[4].replace("(Params, CUtensorMap_st)", ""),
                       "dram_bytes_per_launch": best[0], "ncu_duration_us": best[2],
                       "source": "profiles/ncu_top_kernel_%s.txt (dram__bytes_read.sum + dram__bytes_write.sum)" % R}, f)
            f.write("\n")
    st = run(sys.executable, "tools/ncu_stalls.py", rep, str(best[3] if best else 9), "30")
    with open(os.path.join(P, "ncu_top_kernel_%s_stalls.txt" % R), "w") as f:
        f.write("# python tools/ncu_stalls.py gpurun_out/prof_thin.ncu-rep <launch> 30 : warp-state samples per SASS line of the roofline launch\n" + st)
print(sorted(os.listdir(P)))
