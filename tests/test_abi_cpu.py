"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/iea_b200.h declares, the ctypes mirrors of the POD structs have the C layout, and the
product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "iea_b200.h")


@pytest.fixture(scope="module")
def lib():
    from iea_gan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iea_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    L = ctypes.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and the python binding table covers exactly the header
    assert sorted(lib.exported_symbols()) == names


def test_version_and_error_channel(lib):
    L = lib.lib()
    assert L.iea_version() >= 100
    assert isinstance(L.iea_last_error(), bytes)


def test_struct_layouts_match_the_header(lib):
    """sizeof of the four POD structs, measured by compiling the header with gcc."""
    code = '#include <stdio.h>\n#include "iea_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(iea_sn_layer), ' \
           'sizeof(iea_conv_desc), sizeof(iea_aug_draws), sizeof(iea_sn_bwd_item), sizeof(iea_mt_chunk), ' \
           'sizeof(iea_ortho_item), sizeof(iea_bwd1x1_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "s.c"), os.path.join(td, "s")
        open(src, "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(t) for t in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(lib.SnLayer), ctypes.sizeof(lib.ConvDesc), ctypes.sizeof(lib.AugDraws),
                     ctypes.sizeof(lib.SnBwdItem), ctypes.sizeof(lib.MtChunk), ctypes.sizeof(lib.OrthoItem),
                     ctypes.sizeof(lib.Bwd1x1Args)]


def test_no_cpu_fallback(small_cfg):
    """Modules construct on the CPU (for state-dict work) but every forward demands a B200."""
    import iea_gan_b200 as P
    from iea_gan_b200 import losses, augment
    torch.manual_seed(0)
    G = P.Generator(**dict(small_cfg, device="cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G(torch.randn(40, small_cfg["dim_z"]), torch.arange(40))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.unif_loss(torch.randn(40, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        augment.DiffAugment(torch.randn(4, 1, 8, 8), policy="color")


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "iea_gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle's", ""), os.path.join(dirpath, f)


def test_dropin_module_names(small_cfg):
    """The reference imports `model`, `layers`, `RRM`, `diff_aug`, `loss` by name (train.py:12-14,
    train_fns.py:6-10, model.py:8-13): the dropin directory provides exactly those, with the symbols
    SURVEY.md section 8(b) lists."""
    import importlib
    import sys
    d = os.path.join(ROOT, "iea_gan_b200", "dropin")
    sys.path.insert(0, d)
    try:
        want = {"model": ["Generator", "Discriminator", "G_D", "Model", "generate", "GBlock", "DBlock", "G_arch", "D_arch"],
                "layers": ["SN", "SNConv2d", "SNLinear", "SNEmbedding", "ccbn", "bn", "myBN", "Attention",
                           "power_iteration", "identity", "prior"],
                "RRM": ["RelationalReasoning", "EncoderBlock", "MultiheadAttention", "scaled_dot_product"],
                "diff_aug": ["DiffAugment", "AUGMENT_FNS", "rand_brightness", "rand_saturation", "rand_contrast",
                             "rand_translation", "rand_cutout"],
                "loss": ["loss_hinge_dis", "loss_hinge_gen", "Conditional_Contrastive_loss", "IEA_loss", "unif_loss",
                         "l2_loss"]}
        for mod, names in want.items():
            sys.modules.pop(mod, None)
            m = importlib.import_module(mod)
            assert os.path.dirname(m.__file__) == d
            for n in names:
                assert hasattr(m, n), (mod, n)
    finally:
        sys.path.remove(d)
        for mod in ("model", "layers", "RRM", "diff_aug", "loss"):
            sys.modules.pop(mod, None)


def test_fused_adam_is_a_torch_optimizer_with_adams_state_layout():
    """G.optim / D.optim are optim.FusedAdam: constructor, param_groups and state_dict keys of torch.optim.Adam (the
    reference's checkpoint code and LR schedulers touch exactly these); stepping on a CPU tensor raises (no fallback)."""
    import torch
    from iea_gan_b200.optim import FusedAdam
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = FusedAdam(ps, lr=2e-4, betas=(0.0, 0.999), weight_decay=0, eps=1e-6)
    ref = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=2e-4, betas=(0.0, 0.999), eps=1e-6)
    assert isinstance(opt, torch.optim.Optimizer)
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert opt.param_groups[0][k] == ref.param_groups[0][k], k
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10, eta_min=5e-5)
    sched.step()
    assert opt.param_groups[0]["lr"] < 2e-4
    opt.load_state_dict(opt.state_dict())
    for p in ps:
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(NotImplementedError):
        FusedAdam(ps, weight_decay=0.1)
