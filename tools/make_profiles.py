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
                 ("bench_%s_strong64_n2.json" % R, "bench_%s_train_strong64_n2.json"),
                 ("bench_%s_n4.json" % R, "bench_%s_sample16_n4.json"), ("bench_%s_train_n4.json" % R, "bench_%s_train8_n4.json")):
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

keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
what = {"l32_64": "32->64 1x1 @128x128 + ccbn/ReLU prologue + up2 residual + stats, n=640 (tc2::conv_tc2_kernel)",
        "l32_1": "32->1 3x3 @256x256 + bn/ReLU prologue (output conv), n=640 (thin::conv_thin_kernel)",
        "l16_32": "16->32 1x1 @256x256 + ccbn/ReLU prologue + up2 residual + stats, n=640 (thin::conv_thin_kernel, TMA-store flavour)",
        "l16_16": "16->16 3x3 @256x256 + ccbn/ReLU prologue + stats, n=640 (thin::conv_thin_kernel)"}
mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
found, text = {}, []
for tag in ("l32_64", "l32_1", "l16_32", "l16_16"):
    rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        continue
    rows = list(csv.reader(run("ncu", "-i", rep, "--page", "raw", "--csv").splitlines()))
    h, r = rows[0], rows[-1]
    d = dict(zip(h, r))
    text.append("\n[%s] %s\n%s" % (tag, what[tag], d["Kernel Name"]))
    for k in keys:
        if k in d:
            text.append("  %-78s %s %s" % (k, d[k], rows[1][h.index(k)]))
    rd = float(d["dram__bytes_read.sum"]) * mul[rows[1][h.index("dram__bytes_read.sum")]]
    wr = float(d["dram__bytes_write.sum"]) * mul[rows[1][h.index("dram__bytes_write.sum")]]
    found[tag] = {"kernel": what[tag], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                  "ncu_duration_us": float(d["gpu__time_duration.sum"]),
                  "source": "profiles/ncu_top_kernel_%s.txt [%s] (dram__bytes_read.sum + dram__bytes_write.sum)" % (R, tag)}
if found:
    with open(os.path.join(P, "ncu_top_kernel_%s.txt" % R), "w") as f:
        f.write("# ncu --set full --clock-control none -k regex:conv_t -s 2 -c 1 : python tools/prof_kernel.py <tag> 3\n"
                "# the four heaviest layers of a 16-event sampling pass (bench.py ROOFLINE_LAYERS), each launched alone on 640 images;\n"
                "# the third (warm) launch is captured.  bench.py's `roofline` is the one with the longest live-measured launch.\n"
                + "\n".join(text) + "\n")
    with open(os.path.join(P, "top_kernel_traffic.json"), "w") as f:
        json.dump(found, f, indent=1)
        f.write("\n")
    for tag in ("l32_64", "l32_1", "l16_32", "l16_16"):
        rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
        if os.path.exists(rep):
            st = run(sys.executable, "tools/ncu_stalls.py", rep, "0", "30")
            ro = run(sys.executable, "tools/dbg/ncu_roles.py", rep, "0")
            with open(os.path.join(P, "ncu_top_kernel_%s_stalls_%s.txt" % (R, tag)), "w") as f:
                f.write("# python tools/ncu_stalls.py gpurun_out/prof_%s.ncu-rep 0 30 : warp-state samples per SASS line, %s\n" % (tag, what[tag]) + st
                        + "\n# python tools/dbg/ncu_roles.py: samples / instructions between the synchronisation sites, in SASS order (the warp roles)\n" + ro)
print(sorted(os.listdir(P)))
