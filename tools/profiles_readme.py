"""Write profiles/README.md from the tracked bench lines (run after tools/make_profiles.py)."""
import json, os, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.loads(f.read().splitlines()[-1])


b, t, n2, ref = (load("bench_%s_%s.json" % (R, k)) for k in ("sample16", "train8", "sample16_n2", "reference_arm"))
L = ["# profiles/ -- measured evidence, round %s" % R[1:], "",
     "Everything here comes from one B200 box per call through `gpurun`; raw outputs land in `gpurun_out/` (scratch) and",
     "`tools/make_profiles.py` + `tools/profiles_readme.py` turn them into these tracked files.  Bench numbers are never",
     "taken under a profiler; ncu launch times are cold-cache and serialised (compare shares, not absolutes).", "",
     "| File | What | Produced by |", "|---|---|---|",
     "| `bench_%s_sample16.json` | default `python bench.py` line: Generator sampling, 16 events, + train_step / roofline / cpu_baseline | `tools/profile_round.sh` |" % R,
     "| `bench_%s_train8.json` | `python bench.py --workload train`: full G+D step, 8 events | same |" % R,
     "| `bench_%s_sample16_n2.json` | default line under `torch.distributed.run --nproc-per-node 2` (2 B200, NCCL) | `gpurun --gpus 2` |" % R,
     "| `bench_%s_reference_arm.json` | `python bench.py --impl reference` (CPU oracle on the box's host cores) | |" % R,
     "| `layers_%s_{train8,sample16}.txt` | every C-ABI call of one step timed with CUDA events, grouped by (entry point, shape), algorithmic GB/s per conv | `tools/prof_layers.py` |" % R,
     "| `launches_%s_{train8,sample16}.txt` | ncu `gpu__time_duration.sum` launch lists, summed per kernel | `tools/launch_summary.py` |" % R,
     "| `ncu_top_kernel_%s.txt` | `ncu --set full` of the macro-tile tcgen05 conv launches of one Generator forward (dominant kernel) | `tools/make_profiles.py` |" % R,
     "| `ncu_top_kernel_%s_stalls.txt` | warp-state samples per SASS line of the roofline launch | `tools/ncu_stalls.py` |" % R,
     "| `ncu_attention_%s.txt` | `ncu --set full` of the tcgen05 / TMEM self-attention kernels (forward, query- and key-parallel backward) in a train step | `ncu -k regex:attn_` on `tools/prof_train.py` |" % R,
     "| `top_kernel_traffic.json` | DRAM bytes of the roofline launch (read by `bench.py` for `roofline.traffic`) | `tools/make_profiles.py` |", ""]
if b:
    L += ["## Headline numbers (1 B200, bf16 activations, fp32 master weights, H_base = 1, synthetic data)", "",
          "| Quantity | Value |", "|---|---|",
          "| G sampling, 16 events/step, device-timed | **%.1f events/s** (%.2f ms/step, %d C-ABI launches) |" % (b["value"], b["ms_per_step"], b["gpu_launches"]),
          "| G sampling end to end (pinned z in, post-processed images out, %.0f MB D2H/step) | %.1f events/s |" % (b["e2e"]["d2h_bytes_per_step"] / 1e6, b["e2e"]["value"]),
          "| sampling step vs HBM roofline (2.03 GB/event algorithmic) | %.0f GB/s = %.1f %% of %.0f GB/s measured peak |" % (b["step_roofline"]["achieved_GBs"], 100 * b["step_roofline"]["frac_of_hbm_peak"], b["roofline"]["peak"])]
    ts = b.get("train_step") or t
    if t:
        ts = t
    if ts:
        L += ["| full G+D train step, 8 events/step, device-timed | **%.2f events/s** (%.1f ms/step, %d launches) |" % (ts["value"], ts["ms_per_step"], ts["gpu_launches"]),
              "| train step end to end (pinned real images in, 5 loss floats out) | %.2f events/s |" % ts["e2e"]["value"],
              "| train step vs HBM roofline (31.2 GB/event algorithmic) | %.0f GB/s = %.1f %%, %.1f TFLOP/s |" % (ts["step_roofline"]["achieved_GBs"], 100 * ts["step_roofline"]["frac_of_hbm_peak"], ts["step_roofline"]["achieved_TFLOPs"])]
    r = b["roofline"]
    L += ["| dominant kernel (16->16 3x3 @256^2, 160 images, fused prologue + statistics) | %.3f ms, %.0f GB/s algorithmic = **%.1f %% of HBM peak**; DRAM traffic %s MB vs %.0f MB algorithmic |" % (
        r["ms_per_launch"], r["achieved"], 100 * r["frac"], ("%.0f" % (r["traffic"] / 1e6)) if r.get("traffic") else "n/a", r["algorithmic_bytes_per_launch"] / 1e6)]
    if b.get("cpu_baseline"):
        c = b["cpu_baseline"]
        L += ["| CPU baseline (oracle port of the reference algorithm, fp32, %d host threads, 1 event) | %.2f events/s sampling |" % (c["cores"], c["value"])]
    if n2:
        L += ["| 2 GPUs (weak scaling, 16 events/GPU sampling; 8 events/GPU training, NCCL gradient all-reduce) | %.1f events/s sampling, %.2f events/s training |" % (n2["value"], n2["train_step"]["value"])]
    L += ["", "Train-step timing noise: consecutive steps measured one by one (`tools/step_times.py`) take 174 ms (46 events/s) with",
          "sporadic 220-380 ms outliers at random steps (also with the Python GC disabled and with expandable allocator",
          "segments): the step issues ~1900 launches from Python and sits near the host launch rate (81 ms per step at 1 event,",
          "`tools/cpu_bound.py`), so any host-side hiccup on the shared box starves the GPU.  The bench lines are means over",
          "their K steps and therefore scatter between 31 and 46 events/s from run to run; the e2e leg and the median agree on ~45."]
    L += ["", "Clocks during the timed region: %s MHz of %s MHz, throttle reasons %s." % (b["clocks"]["sm_mhz"], b["clocks"]["sm_max_mhz"], b["clocks"]["reasons"] or "none")]
with open(os.path.join(P, "README.md"), "w") as f:
    f.write("\n".join(L) + "\n")
print("\n".join(L[-14:]))
