"""Fused dgrad + GroupNorm(+ReLU)-backward kernels (``csrc/dgrad_gn.cu``): the 1x1 data-gradient GEMM runs twice and the
gradient of the normalised activation never touches HBM.  Default ON for bf16 since round 2 (``DLB_FUSED_DGRAD=0`` disables);
kernel-level tests against fp64 references and a stage-level test against the fp32 truth.  Own file / own process when
debugging: a protocol bug in a tcgen05 kernel traps the context.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def g():
    from dynamic_load_balance_distributeddnn_b200.ops import gemm_tc
    assert gemm_tc.available() and gemm_tc.dgrad_gn_available()
    return gemm_tc


DTYPES = [torch.bfloat16, torch.float32]      # fp32 storage = TF32 math (kind::tf32)


def _code(dtype):
    from dynamic_load_balance_distributeddnn_b200.ops import _native as nat
    return nat.dtype_code(dtype)


def _problem(ns, hw, cl, cm, ct, seed, dtype=torch.bfloat16):
    torch.manual_seed(seed)
    m = ns * hw
    dy = torch.randn(m, cm, device="cuda").to(dtype)
    w = (torch.randn(cm, cl, device="cuda") / cm ** 0.5).to(dtype)
    big = torch.randn(m, ct, device="cuda").to(dtype)
    off = ct - cl
    kp = (cl + 63) // 64 * 64
    ca = torch.zeros(ns, kp, device="cuda"); cb = torch.zeros(ns, kp, device="cuda")
    ca[:, :cl] = torch.rand(ns, cl, device="cuda") + 0.5
    cb[:, :cl] = torch.randn(ns, cl, device="cuda") * 0.3
    # keep the recomputed ReLU mask away from its decision boundary (fp32 fma vs fp64 reference)
    x = big[:, off:]
    for _ in range(3):
        z = ca[:, None, :cl].double() * x.double().view(ns, hw, cl) + cb[:, None, :cl].double()
        near = (z.abs() < 1e-3).view(m, cl)
        if not near.any():
            break
        x[near] = (x[near].float() + 1.0).to(dtype)
    return dy, w, big, off, ca, cb, kp


def _reference(dy, w, x, ca, cb, ns, hw, cl):
    da = dy.double() @ w.double()
    xv = x.double().view(ns, hw, cl)
    z = ca[:, None, :cl].double() * xv + cb[:, None, :cl].double()
    dz = da.view(ns, hw, cl) * (z > 0)
    return dz, torch.stack([dz.sum(1), (dz * xv).sum(1)], dim=-1)          # [ns, cl, 2]


@pytest.mark.parametrize("ns,hw,cl,cm,ct", [(4, 64, 96, 128, 160), (2, 1024, 256, 128, 256), (4, 32, 64, 128, 64), (9, 32, 72, 128, 72),
                                            (5, 256, 200, 128, 328), (2, 64, 1000, 128, 1024), (16, 256, 416, 64, 512)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_dgrad_gn_stats_pass(g, dtype, ns, hw, cl, cm, ct):
    dy, w, big, off, ca, cb, kp = _problem(ns, hw, cl, cm, ct, ns + hw + cl, dtype)
    x = big[:, off:]
    table = torch.zeros(ns, cl, 2, device="cuda")
    g.dgrad_gn_raw(1, dy.data_ptr(), cm, w.data_ptr(), cl, x.data_ptr(), ct, 0, 0, ns * hw, cl, cm, hw, ca, cb, None, None,
                   table.data_ptr(), 2 * cl, dy.device, dtype=_code(dtype))
    torch.cuda.synchronize()
    _, ref = _reference(dy, w, x, ca, cb, ns, hw, cl)
    scale = ref.abs().max().item()
    # bf16 operands are exact in the tensor core (fp32 accumulate); TF32 rounds the fp32 operands to 10 mantissa bits
    tol = 2e-3 if dtype == torch.bfloat16 else 6e-3
    assert (table.double() - ref).abs().max().item() < tol * scale, ((table.double() - ref).abs().max().item(), scale)


@pytest.mark.parametrize("ns,hw,cl,cm,ct", [(4, 64, 96, 128, 160), (2, 1024, 256, 128, 256), (4, 32, 64, 128, 64), (9, 32, 72, 128, 72),
                                            (5, 256, 200, 128, 328), (2, 64, 1000, 128, 1024)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_dgrad_gn_apply_pass(g, dtype, ns, hw, cl, cm, ct):
    dy, w, big, off, ca, cb, kp = _problem(ns, hw, cl, cm, ct, 7 + ns + hw + cl, dtype)
    x = big[:, off:]
    k2 = torch.zeros(ns, kp, device="cuda"); k3 = torch.zeros(ns, kp, device="cuda")
    k2[:, :cl] = torch.randn(ns, cl, device="cuda") * 0.1
    k3[:, :cl] = torch.randn(ns, cl, device="cuda") * 0.1
    dbig = torch.randn(ns * hw, ct, device="cuda").to(dtype)
    before = dbig.clone()
    dx = dbig[:, off:]
    g.dgrad_gn_raw(2, dy.data_ptr(), cm, w.data_ptr(), cl, x.data_ptr(), ct, dx.data_ptr(), ct, ns * hw, cl, cm, hw, ca, cb, k2, k3,
                   0, 0, dy.device, dtype=_code(dtype))
    torch.cuda.synchronize()
    dz, _ = _reference(dy, w, x, ca, cb, ns, hw, cl)
    xv = x.double().view(ns, hw, cl)
    ref = before[:, off:].double().view(ns, hw, cl) + ca[:, None, :cl].double() * dz + k2[:, None, :cl].double() * xv + k3[:, None, :cl].double()
    err = (dx.double().view(ns, hw, cl) - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
    if off:
        assert torch.equal(dbig[:, :off], before[:, :off])                   # the other channels of the buffer are untouched


def test_gn_bwd_coeff_matches_apply_kernel_math(g):
    """k2/k3 + dgamma/dbeta from the (sum dz, sum dz*x) table == the formulas inside gn_bwd_apply / nc_reduce2."""
    torch.manual_seed(3)
    ns, c, groups, hw = 6, 96, 32, 64
    table = torch.randn(ns, c, 2, device="cuda")
    gamma = torch.rand(c, device="cuda") + 0.5
    mean = torch.randn(ns * groups, device="cuda") * 0.2
    rstd = torch.rand(ns * groups, device="cuda") + 0.5
    kp = 128
    k2 = torch.empty(ns, kp, device="cuda"); k3 = torch.empty(ns, kp, device="cuda")
    dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda")
    g.gn_bwd_coeff_raw(table.data_ptr(), 2 * c, gamma, mean, rstd, k2, k3, dg.data_ptr(), db.data_ptr(), ns, c, groups, hw, table.device)
    torch.cuda.synchronize()
    cpg = c // groups
    A, B = table[..., 0].double(), table[..., 1].double()
    mu = mean.double().view(ns, groups).repeat_interleave(cpg, 1); r = rstd.double().view(ns, groups).repeat_interleave(cpg, 1)
    xh = r * (B - mu * A)
    s1 = (gamma.double() * A).view(ns, groups, cpg).sum(-1).repeat_interleave(cpg, 1)
    s2 = (gamma.double() * xh).view(ns, groups, cpg).sum(-1).repeat_interleave(cpg, 1)
    inv_m = 1.0 / (cpg * hw)
    q = r * r * s2 * inv_m
    assert torch.allclose(k2[:, :c].double(), -q, atol=1e-4, rtol=1e-4)
    assert torch.allclose(k3[:, :c].double(), -r * s1 * inv_m + q * mu, atol=1e-4, rtol=1e-4)
    assert k2[:, c:].abs().sum().item() == 0 and k3[:, c:].abs().sum().item() == 0
    assert torch.allclose(db.double(), A.sum(0), atol=1e-3, rtol=1e-4)
    assert torch.allclose(dg.double(), xh.sum(0), atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("dtype", DTYPES)
def test_dense_block_backward_with_fused_dgrad(g, dtype):
    """Whole dense stage: gradients with the fused dgrad+GN-backward path == the validated three-kernel chain."""
    from dynamic_load_balance_distributeddnn_b200.models import densenet
    torch.manual_seed(0)
    stage = densenet.DenseNet([4], growth_rate=32, num_classes=10).dense1.cuda()
    for m in stage.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = m.weight.data.to(dtype)
    x0 = torch.randn(8, 64, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last).to(dtype)
    outs = []
    saved = (g.FUSED_DGRAD, g.FUSED_DGRAD_TF32)
    g.FUSED_DGRAD_TF32 = True
    try:
        for fused in (False, True):
            g.FUSED_DGRAD = fused
            for p in stage.parameters():
                p.grad = None
            x = x0.clone().requires_grad_(True)
            y = stage(x)
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(5)).to(dtype)
            y.backward(gy)
            torch.cuda.synchronize()
            outs.append((x.grad.float(), [p.grad.float().clone() for p in stage.parameters()]))
    finally:
        g.FUSED_DGRAD, g.FUSED_DGRAD_TF32 = saved
    (dx0, g0), (dx1, g1) = outs
    assert float((dx0 - dx1).abs().max()) < 4e-2 * max(1.0, float(dx0.abs().max()))
    # parameter gradients: both paths against a plain fp32 torch evaluation of the same stage on the same (bf16-valued) weights.
    # The two paths differ by the bf16 rounding of dA (the chain writes it, the fused kernel keeps it in fp32 in TMEM), which a
    # max-norm comparison of 2 000-term sums with cancellation mistakes for an error: the fused path must simply not be
    # further from the fp32 truth than the chain is.
    import torch.nn.functional as F
    xr = x0.float().clone().requires_grad_(True)
    ps = [p.detach().float().clone().requires_grad_(True) for p in stage.parameters()]
    ref = xr
    for i in range(len(ps) // 6):
        g1w, g1b, w1, g2w, g2b, w2 = ps[6 * i:6 * i + 6]
        y = F.conv2d(F.relu(F.group_norm(ref, 32, g1w, g1b)), w1)
        z = F.conv2d(F.relu(F.group_norm(y, 32, g2w, g2b)), w2, padding=1)
        ref = torch.cat([z, ref], 1)
    torch.backends.cudnn.allow_tf32 = False
    ref.backward(gy.float())

    def rel_l2(a, b):
        return float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-12))
    for i, (a, b, r) in enumerate(zip(g0, g1, ps)):
        e_chain, e_fused = rel_l2(a, r.grad), rel_l2(b, r.grad)
        assert e_fused < max(1.5 * e_chain, 2e-2), (i, tuple(a.shape), e_chain, e_fused)
