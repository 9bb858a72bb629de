"""Write profiles/README.md from the tracked bench lines (run after tools/make_profiles.py)."""
import json, os, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.loads(f.read().splitlines()[-1])


def g(d, *ks, default=None):
    for k in ks:
        if not isinstance(d, dict) or k not in d:
            return default
        d = d[k]
    return d


b, tg, te, t1, sw, ref = (load("bench_%s_%s.json" % (R, k)) for k in (
    "sample16", "train8_graph", "train8_eager", "train1_graph", "attn_sweep", "reference_arm"))
n2s, n2t, n2strong = (load("bench_%s_%s.json" % (R, k)) for k in ("sample16_n2", "train8_n2", "train_strong64_n2"))
n4s, n4t = (load("bench_%s_%s.json" % (R, k)) for k in ("sample16_n4", "train8_n4"))
L = ["# profiles/ -- measured evidence, round %s" % R[1:], "",
     "Everything here comes from one B200 box per call through `gpurun`; raw outputs land in `gpurun_out/` (scratch) and",
     "`tools/profile_round.sh` -> `tools/make_profiles.py` -> `tools/profiles_readme.py` turn them into these tracked files.",
     "Bench numbers are never taken under a profiler; ncu launch times are cold-cache and serialised (compare shares, not",
     "absolutes).  Round-1 files (`*_r01_*`) are kept for comparison.  One kernel changed after the per-call / ncu files",
     "of this round were taken: the 32->1 output conv moved to 4-wide macro tiles (1.18 -> 1.01 ms per 16-event launch);",
     "`bench_%s_sample16.json` (`roofline_candidates_ms`) was re-measured with it, `layers_%s_sample16.txt` and the" % (R, R),
     "`[l32_1]` capture still show the 2-wide version.", "",
     "| File | What |", "|---|---|",
     "| `bench_%s_sample16.json` | default `python bench.py` line: Generator sampling, 16 events, + `train_step` (graph, 20 steps, per-step spread, CPU baseline) / `hbase3` / `roofline` (longest launch, chosen live) / `roofline_r1_dominant` / `roofline_best_kernel` / `stock_torch_gpu` / `cpu_baseline` |" % R,
     "| `bench_%s_train8_graph.json`, `_train8_eager.json`, `_train1_graph.json` | `--workload train`: the full G+D step, 8 events as a CUDA graph / launched kernel by kernel, and 1 event as a graph |" % R,
     "| `bench_%s_attn_sweep.json` | `--workload attn-sweep` (BASELINE configs[4]): both RRMs and the self-attention block over 1..256 events, fwd and fwd+bwd |" % R,
     "| `bench_%s_reference_arm.json` | `--impl reference`: the UNMODIFIED reference (baseline/_ref) on the box's host cores: `model.generate` and the `train_fns` G+D step |" % R,
     "| `bench_%s_sample16_n2.json`, `_train8_n2.json`, `_train_strong64_n2.json`, `_sample16_n4.json`, `_train8_n4.json` | the same lines under `torch.distributed.run --nproc-per-node 2 / 4` (weak scaling) and 64 events over 2 GPUs (strong scaling, 8-event micro-batches) |" % R,
     "| `layers_%s_{train8,sample16}.txt` | every C-ABI call of one step timed with CUDA events, grouped by (entry point, shape), algorithmic GB/s per conv (`tools/prof_layers.py`) |" % R,
     "| `launches_%s_{train8,sample16}.txt` | ncu `gpu__time_duration.sum` launch lists, summed per kernel (`tools/launch_summary.py`) |" % R,
     "| `ncu_top_kernel_%s.txt` | `ncu --set full` of the four heaviest layers of the sampling pass, each launched alone on 640 images (`tools/prof_kernel.py`: 32->64 1x1 @128^2 + up2 residual, 32->1 3x3 @256^2, 16->32 1x1 @256^2 + up2 residual, 16->16 3x3 @256^2) |" % R,
     "| `ncu_top_kernel_%s_stalls_l32_64.txt`, `_stalls_l16_32.txt` | warp-state samples per SASS line of the heaviest `conv_tc2` launch and of round 1's dominant launch (`tools/ncu_stalls.py`) |" % R,
     "| `ncu_top_kernel_%s_stalls_l32_1.txt`, `_stalls_l16_16.txt` | the same for the output conv (longest launch at the end of the round) and the 3x3 layer; each file ends with the per-role view of `tools/dbg/ncu_roles.py` |" % R,
     "| `ncu_top_kernel_%s_stalls_l16_32_before_tma.txt` | the 16->32 layer BEFORE its epilogue moved to TMA (1.82 ms: 45 %% of the epilogue's samples on per-pixel residual loads) |" % R,
     "| `top_kernel_traffic.json` | DRAM bytes of those four launches (read by `bench.py` for `roofline.traffic`) |",
     "| `ncu_attention_r01.txt` | `ncu --set full` of the tcgen05 / TMEM self-attention kernels (unchanged since round 1) |", ""]
if b:
    ts = tg or b.get("train_step")
    L += ["## Headline numbers (1 B200, bf16 activations, fp32 master weights, H_base = 1, synthetic data)", "",
          "| Quantity | Value |", "|---|---|",
          "| G sampling, 16 events/step, device-timed | **%.1f events/s** (%.2f ms/step, spread %s, %d launches) |" % (
              b["value"], b["ms_per_step"], g(b, "ms_spread"), b["gpu_launches"]),
          "| G sampling end to end (pinned z in, post-processed images out, %.0f MB D2H/step) | %.1f events/s |" % (b["e2e"]["d2h_bytes_per_step"] / 1e6, b["e2e"]["value"]),
          "| sampling step vs HBM roofline (2.03 GB/event algorithmic) | %.0f GB/s = %.1f %% of %.0f GB/s measured peak |" % (
              b["step_roofline"]["achieved_GBs"], 100 * b["step_roofline"]["frac_of_hbm_peak"], b["roofline"]["peak"])]
    if ts:
        L += ["| full G+D train step, 8 events/step, CUDA graph | **%.2f events/s** (%.1f ms/step, spread %s, %d kernels per replay) |" % (
            ts["value"], ts["ms_per_step"], g(ts, "ms_spread"), ts["gpu_launches"]),
              "| train step end to end (pinned real images in, 5 loss floats out) | %.2f events/s |" % ts["e2e"]["value"],
              "| train step vs HBM roofline (31.2 GB/event algorithmic) | %.0f GB/s = %.1f %%, %.1f TFLOP/s |" % (
                  ts["step_roofline"]["achieved_GBs"], 100 * ts["step_roofline"]["frac_of_hbm_peak"], ts["step_roofline"]["achieved_TFLOPs"])]
    if te:
        L += ["| same, launched kernel by kernel from Python | %.2f events/s mean (%.1f ms), median %.1f ms, max %.1f ms |" % (
            te["value"], te["ms_per_step"], g(te, "ms_spread", "median", default=0), g(te, "ms_spread", "max", default=0))]
    if t1:
        L += ["| train step at the reference's own batch (1 event), CUDA graph | %.2f events/s (%.1f ms/step) |" % (t1["value"], t1["ms_per_step"])]
    if b.get("hbase3"):
        L += ["| G sampling at the shipped geometry 256x768 (H_base = 3), 8 events | %.1f events/s (%.1f ms/step), %.1f %% of the HBM roofline |" % (
            b["hbase3"]["value"], b["hbase3"]["ms_per_step"], 100 * b["hbase3"]["step_roofline"]["frac_of_hbm_peak"])]
    for key, what in (("roofline", "LONGEST launch of the sampling step (chosen live)"),
                      ("roofline_r1_dominant", "round 1's dominant launch (0.30 then)"), ("roofline_best_kernel", "best launch")):
        r = b.get(key)
        if r and "achieved" in r:
            tr = r.get("traffic")
            if not tr:  # (a bench line written before the capture of the same round: take the committed capture)
                tj = os.path.join(P, "top_kernel_traffic.json")
                if os.path.exists(tj):
                    with open(tj) as f:
                        tt = json.load(f)
                    for sub, tag in (("32->1 3x3", "l32_1"), ("32->64 1x1", "l32_64"), ("16->32 1x1", "l16_32"), ("16->16 3x3", "l16_16")):
                        if sub in r["kernel"] and tag in tt:
                            tr = tt[tag]["dram_bytes_per_launch"]
            L += ["| %s: %s | %.3f ms, %.0f GB/s algorithmic = **%.1f %% of HBM peak**; DRAM traffic %s vs %.2f GB algorithmic%s |" % (
                what, r["kernel"].split(" (thin")[0].split(" (tc2")[0].replace("iea_conv_fprop ", ""), r["ms_per_launch"], r["achieved"], 100 * r["frac"],
                ("%.2f GB (%.0f GB/s = %.1f %% of peak)" % (tr / 1e9, tr / r["ms_per_launch"] / 1e6, tr / r["ms_per_launch"] / 1e6 / r["peak"] * 100)) if tr else "n/a",
                r["algorithmic_bytes_per_launch"] / 1e9,
                (" (+%.2f GB fused residual read, not counted)" % (r["fused_extra_bytes"] / 1e9)) if r.get("fused_extra_bytes") else "")]
    if b.get("cpu_baseline") and "value" in b["cpu_baseline"]:
        c = b["cpu_baseline"]
        L += ["| CPU baseline, sampling (%s, fp32, %d host threads, 1 event) | %.2f events/s |" % (c["kind"], c["cores"], c["value"])]
    ct = g(b, "train_step", "cpu_baseline")
    if ct and "value" in ct:
        L += ["| CPU baseline, G+D train step (%s `train_fns` step, fp32, %d host threads, 1 event) | %.3f events/s |" % (ct["kind"], ct["cores"], ct["value"])]
    st = b.get("stock_torch_gpu")
    if st and "sample_events_per_s" in st:
        L += ["| the unmodified reference through stock PyTorch / cuDNN on the same B200 (fp32, 1 event per call) | %.1f events/s sampling (%.1f ms), %.2f events/s training (%.0f ms) |" % (
            st["sample_events_per_s"], st["sample_ms_per_event"], st["train_events_per_s"], st["train_ms_per_event"])]
    if n2s:
        L += ["| 2 GPUs, weak scaling: sampling 16 events/GPU | %.1f events/s |" % n2s["value"]]
    if n2t:
        L += ["| 2 GPUs, weak scaling: train step 8 events/GPU, flat-buffer NCCL all-reduce inside the captured graph | %.2f events/s (%.1f ms/step) |" % (n2t["value"], n2t["ms_per_step"])]
    if n2strong:
        L += ["| 2 GPUs, strong scaling (BASELINE configs[3]): 64 events per step, 32 per GPU in 8-event micro-batches | %.2f events/s (%.1f ms/step) |" % (n2strong["value"], n2strong["ms_per_step"])]
    if n4s:
        L += ["| 4 GPUs, weak scaling: sampling 16 events/GPU | %.1f events/s |" % n4s["value"]]
    if n4t:
        L += ["| 4 GPUs, weak scaling: train step 8 events/GPU | %.2f events/s (%.1f ms/step) |" % (n4t["value"], n4t["ms_per_step"])]
    if sw:
        L += ["", "## RRM + self-attention sweep (BASELINE configs[4], `bench_%s_attn_sweep.json`)" % R, "",
              "| op | events | fwd ms | fwd+bwd ms | fwd TFLOP/s | fwd frac of bf16 tensor peak |", "|---|---|---|---|---|---|"]
        for r in sw["sweep"]:
            L += ["| %s | %d | %.3f | %.3f | %.2f | %.4f |" % (r["op"], r["events"], r["fwd_ms"], r["fwd_bwd_ms"], r["fwd_TFLOPs"], r["fwd_frac_tensor_peak"])]
    L += ["", "Clocks during the timed region: %s MHz of %s MHz, throttle reasons %s." % (
        b["clocks"]["sm_mhz"], b["clocks"]["sm_max_mhz"], b["clocks"]["reasons"] or "none")]
with open(os.path.join(P, "README.md"), "w") as f:
    f.write("\n".join(L) + "\n")
print("\n".join(L[-40:]))
