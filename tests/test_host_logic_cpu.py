"""Host-side logic that needs no GPU: configuration guard of the train step, the multi-tensor chunk table, the
data-parallel sharding rule, and the roofline bookkeeping of bench.py."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_train_step_refuses_configurations_it_does_not_implement():
    """make_train_step is the shipped configuration's path through train_fns.train (train_fns.py:20-194); any flag that
    would change the objective is refused, never silently ignored."""
    from iea_gan_b200.train_step import check_config
    ok = {"num_D_steps": 1, "num_D_accumulations": 1, "num_G_accumulations": 1, "toggle_grads": True,
          "conditional_strategy": "Contra", "split_D": True, "Con_reg": False, "batch_size": 40}
    check_config(ok)
    check_config({"batch_size": 40})  # absent keys are the defaults
    for k, v in (("num_D_steps", 2), ("num_D_accumulations", 4), ("num_G_accumulations", 2), ("toggle_grads", False),
                 ("conditional_strategy", "Proj"), ("split_D", False), ("Con_reg", True)):
        with pytest.raises(NotImplementedError) as e:
            check_config(dict(ok, **{k: v}))
        assert k in str(e.value)


def test_chunk_table_covers_every_element_once():
    """optim._chunk_table: the iea_mt_chunk records tile each tensor exactly once, in order, with 16-byte aligned
    starts (the kernels use 16-byte vector accesses) and the tensor index the clip reduction needs."""
    from iea_gan_b200 import _lib as L
    from iea_gan_b200 import optim
    sizes = [5, optim.CHUNK, optim.CHUNK + 64, 3 * optim.CHUNK + 128, 64]
    ts = [torch.zeros((n + 63) // 64 * 64) for n in sizes]
    rows = [(t, t, None, None, None, n) for t, n in zip(ts, sizes)]
    tab, cnt = optim._chunk_table(rows, "cpu")
    rec = np.frombuffer(tab.numpy().tobytes(), dtype=np.dtype([("p", "u8"), ("g", "u8"), ("m", "u8"), ("v", "u8"), ("ema", "u8"),
                                                               ("n", "i4"), ("tensor", "i4")]))
    assert rec.dtype.itemsize == C.sizeof(L.MtChunk) and len(rec) == cnt
    for ti, (t, n) in enumerate(zip(ts, sizes)):
        mine = rec[rec["tensor"] == ti]
        assert int(mine["n"].sum()) == n and int(mine["n"].max()) <= optim.CHUNK
        starts = (mine["p"].astype(np.int64) - t.data_ptr()) // 4
        assert list(starts) == list(range(0, n, optim.CHUNK))
        assert all(int(s) % 16 == 0 for s in mine["p"].astype(np.int64) - t.data_ptr())
        assert np.array_equal(mine["p"], mine["g"]) and not mine["m"].any() and not mine["ema"].any()


def test_shard_events_rejects_uneven_shards():
    """dp.shard_events: [begin, end) in events, the same number on every rank (a mean of per-rank means is only the
    global per-event mean for equal shards)."""
    from iea_gan_b200 import dp
    spans = [dp.shard_events(6, rank=r, world=3) for r in range(3)]
    assert spans == [(0, 2), (2, 4), (4, 6)]
    with pytest.raises(ValueError):
        dp.shard_events(6, rank=0, world=4)


def test_roofline_layers_and_committed_traffic_agree():
    """bench.py picks the `roofline` launch among ROOFLINE_LAYERS and reads its DRAM traffic from the committed ncu
    summary: every candidate has a traffic record, and the recorded traffic is at least the algorithmic bytes of the
    layer (input + output + weights, bf16) and below twice that plus the fused residual read."""
    import bench
    with open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json")) as f:
        traffic = json.load(f)
    assert set(bench.ROOFLINE_LAYERS) == set(traffic)
    for tag, (h, w, cin, cout, k, res_c, stats, label) in bench.ROOFLINE_LAYERS.items():
        n = 640
        alg = n * h * w * (cin + cout) * 2 + k * k * cin * cout * 2
        res = n * (h // 2) * (w // 2) * 2 * res_c * 2  # whole lines of the channel-dropped skip
        got = traffic[tag]["dram_bytes_per_launch"]
        assert 0.97 * alg <= got <= 1.1 * (alg + res), (tag, got, alg, res)
