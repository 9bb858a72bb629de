"""Generator forward on the B200 against the CPU oracle and the reference-run golden vectors."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def build_G(cfg, seed=0):
    import iea_gan_b200 as P
    torch.manual_seed(seed)
    G = P.Generator(**cfg)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    return G.cuda(), sd


# tolerance: fp32 activations 2e-4 (summation order); bf16 activations 4e-2 relative L2 on the image
@pytest.mark.parametrize("adt,tol", [("fp32", 2e-4), ("bf16", 4e-2)])
def test_generator_train_mode_vs_golden(small_cfg, golden_fwd, adt, tol):
    os.environ["IEA_ACT_DTYPE"] = adt
    try:
        cfg = dict(small_cfg, device="cuda")
        G, _ = build_G(cfg)
        G.train()
        torch.manual_seed(101)
        z = torch.randn(40, cfg["dim_z"]).cuda()
        rd = torch.randn(40, cfg["rdof_dim"])  # the CPU stream's next draw, as in the golden run
        import iea_gan_b200.engine as E
        y = torch.arange(40, device="cuda")
        with torch.no_grad():
            real_randn = torch.randn
            try:
                torch.randn = lambda *a, **k: rd.cuda() if (len(a) == 2 and a[1] == cfg["rdof_dim"]) else real_randn(*a, **k)
                img = G(z, y)
            finally:
                torch.randn = real_randn
        assert img.shape == (40, 1, 64, 64)
        assert rel(img, golden_fwd["g_train_img"]) < tol
        assert rel(G.linear.u0, golden_fwd["g_u0_linear_after"]) < 1e-4
        assert rel(G.linear.sv0, golden_fwd["g_sv0_linear_after"]) < 1e-4
        btol = 1e-4 if adt == "fp32" else 2e-2
        assert rel(G.blocks[0][0].bn1.stored_mean, golden_fwd["g_bn_mean_after"]) < btol
        assert rel(G.blocks[0][0].bn1.stored_var, golden_fwd["g_bn_var_after"]) < btol
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)


def test_generator_multi_event_vs_oracle(small_cfg):
    """E = 2 events in one call == two independent single-event oracle forwards."""
    from oracle import iea_oracle as O
    os.environ["IEA_ACT_DTYPE"] = "fp32"
    try:
        cfg = dict(small_cfg, device="cuda")
        G, sd = build_G(cfg)
        G.train()
        torch.manual_seed(7)
        z = torch.randn(80, cfg["dim_z"])
        rd = torch.randn(80, cfg["rdof_dim"])
        y = torch.arange(40).repeat(2)
        with torch.no_grad():
            ref = O.generator_forward(sd, cfg, z, y, rd, training=True)
            real_randn = torch.randn
            try:
                torch.randn = lambda *a, **k: rd.cuda() if (len(a) == 2 and a[1] == cfg["rdof_dim"]) else real_randn(*a, **k)
                img = G(z.cuda(), y.cuda())
            finally:
                torch.randn = real_randn
        assert rel(img, ref) < 2e-4
        assert rel(G.blocks[3][0].bn2.stored_var, sd["blocks.3.0.bn2.stored_var"]) < 1e-4
    finally:
        os.environ.pop("IEA_ACT_DTYPE", None)
