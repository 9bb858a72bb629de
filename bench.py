#!/usr/bin/env python
"""bench.py -- IEA-GAN hot-path throughput on B200 (one process per GPU).

    python bench.py --gpus N --steps K --warmup W [--workload sample|train] [--impl reference]

Workload (N=1 default, BASELINE.json configs[1]): Generator sampling, 16 events per GPU
(640 images of 256x256, H_base=1), bf16 activations, train-mode batch statistics exactly as
model.generate runs it (model.py:1130-1139), random-init weights, synthetic z.  A "step" is
one Generator forward over that batch; value = events/s over all ranks with z already in
HBM; e2e = the same through the public API with pinned-host z in and the ADU-post-processed
images (model.py:1139-1147) copied back to pinned host memory every step.
`--workload train` times the full G+D step (configs[2], 8 events per GPU) instead; the
default run reports it as the extra "train_step" object.

`--impl reference` times the reference algorithm on the host cores: the CPU oracle
(oracle/iea_oracle.py, pinned to the real reference by tests/golden), one event per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes / flops per event at H_base=1 (bf16, layer granular)
G_FWD_BYTES_PER_EVENT = 2.03e9
G_FWD_FLOP_PER_EVENT = 0.134e12
TRAIN_BYTES_PER_EVENT = 31.2e9
TRAIN_FLOP_PER_EVENT = 1.73e12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason samples during the timed region: NVML at 10 Hz (nvidia-smi, 5 Hz,
    when NVML is not importable)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.stop_flag = index, [], [], set(), False
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the (possibly CUDA_VISIBLE_DEVICES-remapped) ordinal by PCI bus id
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hi).bus == bus:
                        h = hi
            self.nv, self.h = pynvml, h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _nvml(self):
        nv, h = self.nv, self.h
        self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        self.mx.append(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [t.strip() for t in line.split(",")]
            if r[1].isdigit():
                self.sm.append(int(r[1]))
            if r[2].isdigit():
                self.mx.append(int(r[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self._nvml()
                else:
                    self._smi()
            except Exception:
                if self.nv is not None:
                    self.nv = None  # fall back to nvidia-smi
            # (10 Hz: the G+D step is launched from Python and sits close to the launch-rate limit, a 50 Hz poller
            #  thread measurably slowed it through the GIL: 221 ms against 181 ms)
            time.sleep(0.1 if self.nv is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def timed(fn, steps, warmup, dist_on, finish=None):
    """W untimed + K timed calls bracketed by barrier + synchronize; CUDA-event time per step (ms).
    finish() runs before the closing event (joins side streams into the timed stream)."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if finish is not None:
        finish()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()
    b.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = a.elapsed_time(b) / steps
    if dist_on:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def build_nets(cfg, device, with_d):
    import iea_gan_b200 as P
    torch.manual_seed(0)
    G = P.Generator(**cfg).to(device)
    D = P.Discriminator(**cfg).to(device) if with_d else None
    return G, D


def top_kernel_roofline(E_, events, hbm_peak, which):
    """The dominant kernel in isolation: the 16->16 3x3 conv at 256x256 (12 GMAC/event, the
    heaviest layer of G), with its fused BN+ReLU prologue and statistics
    epilogue.  Algorithmic bytes = input + output (bf16) + weights."""
    import iea_gan_b200.sn_layers as SL
    from iea_gan_b200 import _lib as L
    n, h, w, c = 40 * events, 256, 256, 16
    m = SL.SNConv2d(c, c, 3, padding=1, eps=1e-6).cuda()
    grp = E_.SNGroup()
    l = grp.add(m, E_.act_dtype())
    grp.run(True, False)
    x = torch.randn(n, h, w, c, device="cuda").to(E_.act_dtype())
    ss = E_.ScaleShift(torch.rand(n, c, device="cuda") + 0.5, torch.randn(n, c, device="cuda"))
    tape = E_.Tape(False)
    fn = lambda: E_.conv(tape, E_.Var(x, need=False), l, n, h, w, 3, bias=m.bias, in_relu=True, ss=ss, stats=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    esz = 2 if E_.act_dtype() == torch.bfloat16 else 4
    bytes_ = 2 * n * h * w * c * esz + 9 * c * c * esz
    ach = bytes_ / (ms * 1e-3) / 1e9
    traffic, traffic_src = None, None  # DRAM bytes of this launch from the committed `ncu --set full` capture
    tp = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
    if os.path.exists(tp) and events == 4:
        with open(tp) as f:
            t = json.load(f)
        traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
    return {"bound": "hbm", "kernel": "iea_conv_fprop 16->16 3x3 @256x256 (+BN/ReLU prologue, stats epilogue), "
                                      "thin::conv_thin_kernel<2,1,1,4>, %d images" % n,
            "achieved": round(ach, 1), "peak": hbm_peak, "peak_source": which, "unit": "GB/s",
            "frac": round(ach / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
            "ms_per_launch": round(ms, 4), "algorithmic_bytes_per_launch": bytes_,
            "impl": os.environ.get("IEA_CONV_IMPL", "auto")}


def cpu_oracle_sample(cfg, reps):
    """Reference algorithm (CPU oracle) timed on the host cores: Generator sampling, 1 event/step."""
    from oracle import iea_oracle as O
    import iea_gan_b200 as P
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    c = dict(cfg, device="cpu")
    G = P.Generator(**c)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    y = torch.arange(40)
    ts = []
    for i in range(reps + 1):
        z, rd = torch.randn(40, c["dim_z"]), torch.randn(40, c["rdof_dim"])
        t0 = time.perf_counter()
        with torch.no_grad():
            O.generator_forward(sd, c, z, y, rd, training=True)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[1:])
    return ts[len(ts) // 2], torch.get_num_threads()


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    sec, cores = cpu_oracle_sample(cfg, max(1, args.steps))
    v = 1.0 / sec
    line = {"impl": "reference", "metric": "G-sample events/s (40 PXD imgs/event)", "value": round(v, 4),
            "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": 1,
            "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Generator sampling, H_base=1 (256x256), train-mode BN, 1 event per step on CPU",
                       "events_per_step": 1},
            "cpu_baseline": {"value": round(v, 4), "unit": "events/s", "cores": cores, "kind": "port",
                             "sample": "CPU oracle (oracle/iea_oracle.py, fp32 torch ops, pinned to the reference by "
                                       "tests/golden) generator forward of 1 event; median of %d" % max(1, args.steps)},
            "e2e": {"value": round(v, 4), "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sample", choices=["sample", "train"])
    ap.add_argument("--events", type=int, default=0, help="events per GPU (default 16 sampling / 8 training)")
    ap.add_argument("--hbase", type=int, default=1)
    ap.add_argument("--no-extras", action="store_true", help="skip the cpu baseline / roofline / train extras")
    args = ap.parse_args()
    from iea_gan_b200.default_config import shipped_config
    cfg = shipped_config(H_base=args.hbase, clip_norm=1e9)  # clip_norm: otherwise G never steps (train_fns.py:190)
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cfg["device"] = "cuda"
    import iea_gan_b200 as P
    from iea_gan_b200 import engine as E_
    from iea_gan_b200.train_step import make_train_step, NormalNoise, EMA
    from iea_gan_b200 import dp
    hbm_peak, tf_peak, which = peaks()
    W = max(args.warmup, 3)
    K_ = args.steps
    res_w = 256 * args.hbase
    line = {}
    sampler = ClockSampler(local)

    if args.workload == "sample":
        ev = args.events or 16
        G, _ = build_nets(cfg, dev, False)
        G.train()  # model.generate never calls eval(): batch statistics
        n = 40 * ev
        z = torch.randn(n, cfg["dim_z"], device=dev)
        y = torch.arange(40, device=dev).repeat(ev)

        def step():
            with torch.no_grad():
                return G(z, y)
        step()
        torch.cuda.synchronize()
        sampler.start()
        l0 = E_.LAUNCHES[0]
        ms = timed(step, K_, W, dist_on)
        launches = (E_.LAUNCHES[0] - l0) // (K_ + W)
        clocks = sampler.summary()
        value = ev * world / (ms * 1e-3)
        # end to end through the public API: pinned host z -> device, forward, ADU post-process, D2H
        zh = torch.randn(n, cfg["dim_z"]).pin_memory()
        yh = torch.arange(40).repeat(ev).pin_memory()
        # two pinned result buffers and a copy stream: the device->host read of step i overlaps the forward of
        # step i+1 (every step still uploads its z, y and downloads its full result; the closing event waits
        # for the last download)
        outh = [torch.empty((n, 250, res_w), dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        turn = [0]

        def step_e2e():
            with torch.no_grad():
                img = G(zh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True))
                post = E_.adu_postprocess(img)
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                outh[turn[0] & 1].copy_(post, non_blocking=True)
            post.record_stream(copy_stream)
            turn[0] += 1
        ms_e = timed(step_e2e, K_, W, dist_on, finish=lambda: torch.cuda.current_stream().wait_stream(copy_stream))
        line = {"metric": "G-sample events/s (40 PXD imgs/event)", "value": round(value, 2), "unit": "events/s",
                "n_gpus": world, "steps": K_, "warmup": W, "ms_per_step": round(ms, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if E_.act_dtype() == torch.bfloat16 else "f32", "data": "synthetic",
                "config": {"workload": "Generator sampling (RRM + ccbn + SNConv2d), %d events/GPU, 256x%d, "
                                       "train-mode BN, random-init weights" % (ev, res_w),
                           "events_per_gpu": ev, "H_base": args.hbase, "parallelism": "dp%d" % world,
                           "l2": "per-step working set (>1 GB of activations) far exceeds the 126 MB L2"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": round(ev * world / (ms_e * 1e-3), 2), "unit": "events/s",
                        "h2d_bytes_per_step": zh.numel() * 4 + yh.numel() * 8,
                        "d2h_bytes_per_step": outh[0].numel() * 4, "ms_per_step": round(ms_e, 3),
                        "overlap": "D2H of step i on a copy stream under the forward of step i+1"},
                "step_roofline": {"algorithmic_GB_per_step": round(G_FWD_BYTES_PER_EVENT * ev * args.hbase / 1e9, 2),
                                  "achieved_GBs": round(G_FWD_BYTES_PER_EVENT * ev * args.hbase / (ms * 1e-3) / 1e9, 1),
                                  "frac_of_hbm_peak": round(G_FWD_BYTES_PER_EVENT * ev * args.hbase / (ms * 1e-3) / 1e9 / hbm_peak, 4)}}
        del G, z
        torch.cuda.empty_cache()
    if args.workload == "train" or (not args.no_extras and args.workload == "sample"):
        ev = (args.events or 8) if args.workload == "train" else 8
        G, D = build_nets(cfg, dev, True)
        G.train(); D.train()
        GD = P.G_D(G, D)
        if dist_on:
            dp.broadcast_state(G)
            dp.broadcast_state(D)
        n = 40 * ev
        tcfg = dict(cfg, batch_size=n)
        z_ = NormalNoise(n, cfg["dim_z"], dev)
        G_ema = P.Generator(**dict(cfg, skip_init=True, no_optim=True)).to(dev)
        ema = EMA(G, G_ema, cfg["ema_decay"], cfg["ema_start"])
        train = make_train_step(G, D, GD, z_, tcfg, ema=ema, grad_hook=dp.allreduce_grads if dist_on else None)
        x = torch.rand(n, 1, 256, res_w, device=dev) * 2 - 1
        y = torch.arange(40, device=dev).repeat(ev)
        xh = (torch.rand(n, 1, 256, res_w) * 2 - 1).pin_memory()
        yh = torch.arange(40).repeat(ev).pin_memory()
        # (at least 10 warm-up steps for the train step: the caching allocator needs that many full G+D steps to
        #  stop growing -- 3 warm-ups measured 221 ms, 6 measured 205 ms, the settled step is 194 ms; the e2e leg
        #  that runs afterwards always saw the settled time)
        kt, wt = (K_, max(W, 10)) if args.workload == "train" else (max(2, min(K_, 4)), max(W, 10))
        l0 = E_.LAUNCHES[0]
        if args.workload == "train":
            sampler.start()
        ms_t = timed(lambda: train(x, y), kt, wt, dist_on)
        launches_t = (E_.LAUNCHES[0] - l0) // (kt + wt)
        ms_te = timed(lambda: train(xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True)), kt, wt, dist_on)
        tr = {"metric": "G+D train-step events/s (40 PXD imgs/event)", "value": round(ev * world / (ms_t * 1e-3), 3),
              "unit": "events/s", "ms_per_step": round(ms_t, 2), "events_per_gpu": ev, "steps": kt, "warmup": wt,
              "gpu_launches": launches_t,
              "e2e": {"value": round(ev * world / (ms_te * 1e-3), 3), "unit": "events/s",
                      "h2d_bytes_per_step": xh.numel() * 4 + yh.numel() * 8, "d2h_bytes_per_step": 5 * 4},
              "step_roofline": {"achieved_GBs": round(TRAIN_BYTES_PER_EVENT * ev * args.hbase / (ms_t * 1e-3) / 1e9, 1),
                                "frac_of_hbm_peak": round(TRAIN_BYTES_PER_EVENT * ev * args.hbase / (ms_t * 1e-3) / 1e9 / hbm_peak, 4),
                                "achieved_TFLOPs": round(TRAIN_FLOP_PER_EVENT * ev * args.hbase / (ms_t * 1e-3) / 1e12, 1)}}
        if args.workload == "train":
            line = dict(tr, n_gpus=world, higher_is_better=True, scaling="weak", vs_baseline=None,
                        dtype="bf16" if E_.act_dtype() == torch.bfloat16 else "f32", data="synthetic",
                        config={"workload": "full G+D hinge/contrastive train step with DiffAugment, ortho-reg, Adam, "
                                            "EMA, %d events/GPU, 256x%d" % (ev, res_w), "events_per_gpu": ev,
                                "H_base": args.hbase, "parallelism": "dp%d" % world, "clip_norm": 1e9},
                        clocks=sampler.summary())
        else:
            line["train_step"] = tr
        del G, D, GD, train, x
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_extras:
        try:
            line["roofline"] = top_kernel_roofline(E_, 4, hbm_peak, which)
        except Exception as e:  # never lose the headline because an extra failed
            line["roofline"] = {"error": repr(e)}
        if world == 1:
            sec, cores = cpu_oracle_sample(cfg, 2)
            line["cpu_baseline"] = {"value": round(1.0 / sec, 4), "unit": "events/s", "cores": cores, "kind": "port",
                                    "sample": "CPU oracle generator forward (fp32), 1 event, median of 2 after 1 warm-up"}
    if rank == 0:
        print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
