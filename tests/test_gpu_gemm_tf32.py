"""fp32-storage / TF32-math flavour of the tcgen05 kernels (tcgen05.mma kind::tf32) vs fp32 PyTorch.  This is the precision
class of the reference's default path (fp32 tensors, cuDNN convolutions with allow_tf32=True; reference dbs.py:363 keeps the
model in fp32).  Own file / own process: a protocol bug traps the context."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 4e-3          # TF32 keeps 10 mantissa bits: ~5e-4 per product, random-sign accumulation over K


@pytest.fixture(scope="module")
def g():
    from dynamic_load_balance_distributeddnn_b200.ops import gemm_tc
    assert gemm_tc.available() and torch.float32 in gemm_tc.TC_DTYPES
    torch.backends.cuda.matmul.allow_tf32 = False          # references below are true fp32
    torch.backends.cudnn.allow_tf32 = False
    return gemm_tc


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (256, 128, 128), (4096, 128, 256), (1000, 40, 72), (65536, 128, 64),
                                   (8192, 256, 512), (300, 32, 1152), (5000, 200, 200), (128 * 148 + 77, 128, 192)])
def test_plain_gemm_tf32(g, m, n, k):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(n, k, device="cuda") / k ** 0.5
    d = g.gemm(a, b)
    torch.cuda.synchronize()
    assert d.dtype == torch.float32
    r = a @ b.t()
    err = (d - r).abs().max().item()
    assert err < TOL * max(1.0, r.abs().max().item()), err


def test_strided_operands_and_output_tf32(g):
    torch.manual_seed(0)
    big = torch.randn(4096, 320, device="cuda")
    a = big[:, 64:64 + 192]
    b = torch.randn(128, 192, device="cuda") / 14
    out_big = torch.zeros(4096, 256, device="cuda")
    g.gemm(a, b, out=out_big[:, 32:160])
    torch.cuda.synchronize()
    r = a @ b.t()
    assert (out_big[:, 32:160] - r).abs().max().item() < TOL * r.abs().max().item()
    assert out_big[:, :32].abs().sum().item() == 0 and out_big[:, 160:].abs().sum().item() == 0


@pytest.mark.parametrize("hw,k,n", [(1024, 64, 128), (256, 200, 128), (64, 512, 128), (16, 1024, 128)])
def test_gn_relu_prologue_and_stats_tf32(g, hw, k, n):
    torch.manual_seed(1)
    ns = 8
    m = ns * hw
    x = torch.randn(m, k, device="cuda")
    kp = (k + 63) // 64 * 64
    pa = torch.zeros(ns, kp, device="cuda"); pb = torch.zeros(ns, kp, device="cuda")
    pa[:, :k] = torch.rand(ns, k, device="cuda") + 0.5
    pb[:, :k] = torch.randn(ns, k, device="cuda") * 0.3
    b = torch.randn(n, k, device="cuda") / k ** 0.5
    table = torch.zeros(ns, n, 2, device="cuda")
    d = g.gemm(x, b, pro_a=pa, pro_b=pb, rows_per_sample=hw, stats=table if hw % 32 == 0 else None, stats_ns=2 * n)
    torch.cuda.synchronize()
    xa = torch.relu(x.view(ns, hw, k) * pa[:, None, :k] + pb[:, None, :k]).view(m, k)
    r = xa @ b.t()
    err = (d - r).abs().max().item()
    assert err < TOL * max(1.0, r.abs().max().item()), err
    if hw % 32 == 0:
        df = d.view(ns, hw, n)
        assert torch.allclose(table[..., 0], df.sum(1), atol=1e-2, rtol=1e-3)
        assert torch.allclose(table[..., 1], (df * df).sum(1), atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("m,co,ci", [(4096, 128, 64), (65536, 128, 256), (8192, 128, 1024), (1000, 128, 96), (3000, 32, 128),
                                     (512 * 16, 256, 512), (130, 64, 72)])
def test_wgrad_mn_major_tf32(g, m, co, ci):
    torch.manual_seed(m + co + ci)
    dy = torch.randn(m, co, device="cuda") / 8
    x = torch.randn(m, ci, device="cuda")
    dw = g.wgrad(dy, x)
    torch.cuda.synchronize()
    r = dy.t() @ x
    err = (dw - r).abs().max().item()
    assert err < TOL * max(1.0, r.abs().max().item()), (err, r.abs().max().item())


@pytest.mark.parametrize("hw,co,ci", [(1024, 128, 64), (256, 128, 200), (64, 128, 512), (16, 128, 1024)])
def test_wgrad_with_gn_relu_prologue_tf32(g, hw, co, ci):
    torch.manual_seed(7)
    ns = 6
    m = ns * hw
    big = torch.randn(m, ci + 64, device="cuda")
    x = big[:, 32:32 + ci]
    dy = torch.randn(m, co, device="cuda") / 8
    kp = (ci + 63) // 64 * 64
    pa = torch.zeros(ns, kp, device="cuda"); pb = torch.zeros(ns, kp, device="cuda")
    pa[:, :ci] = torch.rand(ns, ci, device="cuda") + 0.5
    pb[:, :ci] = torch.randn(ns, ci, device="cuda") * 0.3
    dw = g.wgrad(dy, x, pa, pb, hw)
    torch.cuda.synchronize()
    xa = torch.relu(x.reshape(ns, hw, ci) * pa[:, None, :ci] + pb[:, None, :ci]).reshape(m, ci)
    r = dy.t() @ xa
    err = (dw - r).abs().max().item()
    assert err < TOL * max(1.0, r.abs().max().item()), (err, r.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(4096, 256, 128), (65536, 64, 128), (1000, 200, 72), (8192, 1024, 128), (300, 600, 200)])
def test_gemm_b_mn_major_tf32(g, m, n, k):
    torch.manual_seed(m + n)
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(k, n, device="cuda") / k ** 0.5
    d = g.gemm_bmn(a, b)
    torch.cuda.synchronize()
    r = a @ b
    err = (d - r).abs().max().item()
    assert err < TOL * max(1.0, r.abs().max().item()), err


@pytest.mark.parametrize("n,c,o,hw", [(8, 128, 32, 32), (6, 128, 32, 16), (9, 128, 32, 8), (21, 128, 32, 4), (4, 64, 64, 32),
                                        (5, 96, 160, 16)])
def test_conv3x3_tf32_fwd_dgrad(g, n, c, o, hw):
    torch.manual_seed(n + c + o)
    x = torch.randn(n, c, hw, hw, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(o, c, 3, 3, device="cuda") / (3 * c ** 0.5)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = g._Conv3x3Fn.apply(x, w)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, padding=1)
    yr.backward(gy)
    assert (y - yr).abs().max().item() < TOL * max(1.0, yr.abs().max().item())
    assert (x.grad - xr.grad).abs().max().item() < TOL * max(1.0, xr.grad.abs().max().item())


def test_conv1x1_and_linear_autograd_tf32(g):
    torch.manual_seed(3)
    x = torch.randn(8, 96, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(128, 96, 1, 1, device="cuda") / 10).requires_grad_(True)
    assert g.conv_supported(x, w, 1, 0, 1)
    y = g.conv2d(x, w, None, 1, 0)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr)
    yr.backward(gy)
    assert (y - yr).abs().max().item() < TOL * yr.abs().max().item()
    assert (x.grad - xr.grad).abs().max().item() < TOL * xr.grad.abs().max().item()
    assert (w.grad - wr.grad).abs().max().item() < TOL * wr.grad.abs().max().item()
    xl = torch.randn(35, 16, 200, device="cuda", requires_grad=True)
    wl = (torch.randn(600, 200, device="cuda") / 14).requires_grad_(True)
    assert g.linear_supported(xl, wl)
    yl = g.linear(xl, wl, None)
    assert (yl - torch.nn.functional.linear(xl, wl)).abs().max().item() < TOL * yl.abs().max().item()


def test_dense_stage_tf32_matches_fp32_torch(g):
    """A whole DenseNet stage (concat-free buffer, GN-prologue GEMMs, stats epilogues, hand-written backward) in fp32
    storage / TF32 math against the plain torch.nn stage in true fp32."""
    from dynamic_load_balance_distributeddnn_b200.models.densenet import Bottleneck
    from dynamic_load_balance_distributeddnn_b200.ops import dense_block
    torch.manual_seed(5)
    stage = torch.nn.Sequential(*[Bottleneck(64 + 32 * i, 32) for i in range(3)]).cuda()
    x = torch.randn(8, 64, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert dense_block.supported(stage, x)
    out = dense_block.run(stage, x)
    gy = torch.randn_like(out)
    out.backward(gy)
    got = [p.grad.clone() for p in stage.parameters()]
    gx = x.grad.clone()
    for p in stage.parameters():
        p.grad = None
    x.grad = None
    import torch.nn.functional as F
    ref = x
    for blk in stage:
        y = F.conv2d(F.relu(F.group_norm(ref, 32, blk.gn1.weight, blk.gn1.bias, blk.gn1.eps)), blk.conv1.weight)
        z = F.conv2d(F.relu(F.group_norm(y, 32, blk.gn2.weight, blk.gn2.bias, blk.gn2.eps)), blk.conv2.weight, padding=1)
        ref = torch.cat([z, ref], 1)
    ref.backward(gy)
    # forward: element-wise.  backward: relative L2 -- TF32 rounding of the forward flips a few ReLU decisions near zero, which
    # moves individual gradient entries by O(10 %) of the maximum even in a bit-exact CPU emulation of TF32 truncation
    # (measured: 12.8 % max-norm vs 0.06 % forward), while the gradient as a whole stays within ~1 %
    def rel_l2(a, b):
        return ((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-12)).item()
    assert (out - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    assert rel_l2(gx, x.grad) < 5e-2, rel_l2(gx, x.grad)
    for a, p in zip(got, stage.parameters()):
        assert rel_l2(a, p.grad) < 5e-2, (p.shape, rel_l2(a, p.grad))
