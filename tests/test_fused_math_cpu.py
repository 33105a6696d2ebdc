"""The algebra implemented by the fused GroupNorm kernels, checked against autograd on CPU (fp64):

* forward:   relu(GN(x)) == relu(ca*x + cb)  with per-(sample, channel) coefficients  ca = gamma*rstd, cb = beta - mean*ca
             (``csrc/norm.cu:gn_coeff_kernel``; the A-operand prologue of ``csrc/gemm_tc.cu``)
* backward:  dX = ca*dz + k2*x + k3,  dz = dA * [ca*x + cb > 0],  with k2, k3 from the per-(sample, channel) sums
             A = sum_p dz, B = sum_p dz*x  (``nc_reduce2<MODE 3>`` / ``gn_bwd_apply<RELU 2>`` and the experimental
             two-pass ``csrc/dgrad_gn.cu`` + ``gn_bwd_coeff_kernel``);  dgamma = sum_n rstd*(B - mean*A), dbeta = sum_n A.
"""
import torch


def test_gn_relu_conv1x1_backward_algebra():
    torch.manual_seed(0)
    n, c, h, w, groups, cm, eps = 3, 24, 4, 4, 4, 10, 1e-5
    hw, cpg = h * w, c // groups
    x = torch.randn(n, c, h, w, dtype=torch.float64, requires_grad=True)
    gamma = (torch.rand(c, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, dtype=torch.float64) * 0.3).requires_grad_(True)
    wt = torch.randn(cm, c, 1, 1, dtype=torch.float64)
    gy = torch.randn(n, cm, h, w, dtype=torch.float64)
    y = torch.nn.functional.conv2d(torch.relu(torch.nn.functional.group_norm(x, groups, gamma, beta, eps)), wt)
    (y * gy).sum().backward()

    with torch.no_grad():
        xf = x.permute(0, 2, 3, 1).reshape(n, hw, c)                       # NHWC rows, as the kernels see it
        xg = xf.view(n, hw, groups, cpg)
        mean = xg.mean(dim=(1, 3)); var = xg.var(dim=(1, 3), unbiased=False)
        rstd = (var + eps).rsqrt()                                         # [n, G]
        mu_c, r_c = mean.repeat_interleave(cpg, 1), rstd.repeat_interleave(cpg, 1)      # [n, c]
        ca = gamma * r_c; cb = beta - mu_c * ca
        z = ca[:, None] * xf + cb[:, None]
        assert torch.allclose(torch.relu(z), torch.relu(torch.nn.functional.group_norm(x, groups, gamma, beta, eps))
                              .permute(0, 2, 3, 1).reshape(n, hw, c))
        dA = gy.permute(0, 2, 3, 1).reshape(n, hw, cm) @ wt.view(cm, c)     # the dgrad GEMM
        dz = dA * (z > 0)
        A, B = dz.sum(1), (dz * xf).sum(1)                                  # the (sum dz, sum dz*x) table
        xh = r_c * (B - mu_c * A)
        s1 = (gamma * A).view(n, groups, cpg).sum(-1).repeat_interleave(cpg, 1)
        s2 = (gamma * xh).view(n, groups, cpg).sum(-1).repeat_interleave(cpg, 1)
        inv_m = 1.0 / (cpg * hw)
        q = r_c * r_c * s2 * inv_m
        k2, k3 = -q, -r_c * s1 * inv_m + q * mu_c
        dx = ca[:, None] * dz + k2[:, None] * xf + k3[:, None]
        assert torch.allclose(dx, x.grad.permute(0, 2, 3, 1).reshape(n, hw, c), atol=1e-10)
        assert torch.allclose(xh.sum(0), gamma.grad, atol=1e-10)
        assert torch.allclose(A.sum(0), beta.grad, atol=1e-10)


def test_group_stats_from_channel_table():
    """GroupNorm statistics of a growing channel set derived from a per-(sample, channel) (sum, sumsq) table
    (``dlb_gn_finalize`` / the concat-free dense block) == statistics of the concatenated tensor."""
    torch.manual_seed(1)
    n, hw, groups, eps = 2, 16, 4, 1e-5
    parts = [torch.randn(n, hw, 8, dtype=torch.float64), torch.randn(n, hw, 4, dtype=torch.float64) * 2 + 1,
             torch.randn(n, hw, 4, dtype=torch.float64) - 0.5]
    table = torch.cat([torch.stack([p.sum(1), (p * p).sum(1)], -1) for p in parts], dim=1)      # built slice by slice
    full = torch.cat(parts, dim=2)
    c = full.shape[2]; cpg = c // groups
    s = table[..., 0].view(n, groups, cpg).sum(-1); ss = table[..., 1].view(n, groups, cpg).sum(-1)
    m = 1.0 / (cpg * hw)
    mean = s * m; var = (ss * m - mean * mean).clamp_min(0)
    ref = full.view(n, hw, groups, cpg)
    assert torch.allclose(mean, ref.mean(dim=(1, 3)))
    assert torch.allclose((var + eps).rsqrt(), (ref.var(dim=(1, 3), unbiased=False) + eps).rsqrt())
