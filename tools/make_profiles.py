"""Turn the raw outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/make_profiles.py [round]"""
import csv, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True, cwd=ROOT).stdout


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


for src, dst in (("bench_%s.json" % R, "bench_%s_sample16.json"), ("bench_%s_train.json" % R, "bench_%s_train8_graph.json"),
                 ("bench_%s_train_eager.json" % R, "bench_%s_train8_eager.json"), ("bench_%s_train1.json" % R, "bench_%s_train1_graph.json"),
                 ("bench_%s_sweep.json" % R, "bench_%s_attn_sweep.json"), ("bench_%s_ref.json" % R, "bench_%s_reference_arm.json"),
                 ("bench_%s_n2.json" % R, "bench_%s_sample16_n2.json"), ("bench_%s_train_n2.json" % R, "bench_%s_train8_n2.json"),
                 ("bench_%s_strong64_n2.json" % R, "bench_%s_train_strong64_n2.json")):
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
    picks = {"dominant": ("<2, 0, 2, 4, 1,", "16->32 1x1 @256x256 + up2 residual + stats, n=640"),
             "best": ("<2, 1, 1, 4, 2,", "16->16 3x3 @256x256 (+BN/ReLU prologue, stats epilogue), n=640")}
    found = {}
    mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    with open(os.path.join(P, "ncu_top_kernel_%s.txt" % R), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:conv_thin_kernel -s 21 -c 2 : python tools/prof_sample.py 16 2\n"
                "# the macro-tile tcgen05 conv launches of the second Generator forward (16 events = 640 images);\n"
                "# template arguments <CPR, IS3, NB, MT, EPI, TMA>; bench.py's `roofline` (dominant launch by time) is the\n"
                "# <2, 0, 2, 4, 1, 1> launch with the largest DRAM traffic, `roofline_best_kernel` the <2, 1, 1, 4, 2, 1> one\n")
        for r in rows[2:]:
            d = dict(zip(h, r))
            f.write("\n%s\n" % d["Kernel Name"])
            for k in keys:
                if k in d:
                    f.write("  %-70s %s %s\n" % (k, d[k], rows[1][h.index(k)]))
            name = d["Kernel Name"].replace("(int)", "").replace("(bool)", "")
            for key, (sig, what) in picks.items():
                if "conv_thin_kernel" + sig in name:
                    rd = float(d["dram__bytes_read.sum"]) * mul[rows[1][h.index("dram__bytes_read.sum")]]
                    wr = float(d["dram__bytes_write.sum"]) * mul[rows[1][h.index("dram__bytes_write.sum")]]
                    if key not in found or rd + wr > found[key]["dram_bytes_per_launch"]:
                        found[key] = {"kernel": "thin::%s %s" % (name.replace("(Params, CUtensorMap_st)", ""), what),
                                      "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                      "ncu_duration_us": float(d["gpu__time_duration.sum"]), "launch_index": rows.index(r) - 2,
                                      "source": "profiles/ncu_top_kernel_%s.txt (dram__bytes_read.sum + dram__bytes_write.sum)" % R}
    if found:
        with open(os.path.join(P, "top_kernel_traffic.json"), "w") as f:
            json.dump(found, f, indent=1)
            f.write("\n")
    best = (0, 0, 0, found["dominant"]["launch_index"]) if "dominant" in found else None
    st = run(sys.executable, "tools/ncu_stalls.py", rep, str(best[3] if best else 9), "30")
    with open(os.path.join(P, "ncu_top_kernel_%s_stalls.txt" % R), "w") as f:
        f.write("# python tools/ncu_stalls.py gpurun_out/prof_thin.ncu-rep <launch> 30 : warp-state samples per SASS line of the DOMINANT launch (16->32 1x1 + up2 residual)\n" + st)
print(sorted(os.listdir(P)))
