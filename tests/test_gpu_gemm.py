"""tcgen05/TMA GEMM numerics vs fp32 PyTorch (own file: a protocol bug traps the context, so the driver script
runs this module in its own process under `timeout`)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    from dynamic_load_balance_distributeddnn_b200.ops import gemm_tc
    assert gemm_tc.available()
    return gemm_tc


def _ref(a, b):
    return a.float() @ b.float().t()


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (256, 128, 128), (4096, 128, 256), (1000, 40, 72), (65536, 128, 64),
                                   (8192, 256, 1024), (300, 32, 1152), (5000, 200, 200), (128 * 148 * 2 + 77, 128, 192)])
def test_plain_gemm(g, m, n, k):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    d = g.gemm(a, b)
    torch.cuda.synchronize()
    r = _ref(a, b)
    err = (d.float() - r).abs().max().item()
    assert err < 2e-2 * max(1.0, r.abs().max().item()), err


def test_strided_operands_and_output(g):
    """A is a channel slice of a wider NHWC buffer; D is written into a channel slice of another buffer."""
    torch.manual_seed(0)
    big = torch.randn(4096, 320, device="cuda").bfloat16()
    a = big[:, 64:64 + 192]
    b = (torch.randn(128, 192, device="cuda") / 14).bfloat16()
    out_big = torch.zeros(4096, 256, device="cuda", dtype=torch.bfloat16)
    g.gemm(a, b, out=out_big[:, 32:160])
    torch.cuda.synchronize()
    r = _ref(a, b)
    assert (out_big[:, 32:160].float() - r).abs().max().item() < 2e-2 * r.abs().max().item()
    assert out_big[:, :32].abs().sum().item() == 0 and out_big[:, 160:].abs().sum().item() == 0


@pytest.mark.parametrize("hw,k,n", [(1024, 64, 128), (256, 200, 128), (64, 512, 128), (16, 1024, 128)])
def test_gn_relu_prologue(g, hw, k, n):
    torch.manual_seed(1)
    ns = 8
    m = ns * hw
    x = torch.randn(m, k, device="cuda").bfloat16()
    kp = (k + 63) // 64 * 64
    pa = torch.zeros(ns, kp, device="cuda"); pb = torch.zeros(ns, kp, device="cuda")
    pa[:, :k] = torch.rand(ns, k, device="cuda") + 0.5
    pb[:, :k] = torch.randn(ns, k, device="cuda") * 0.3
    b = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    d = g.gemm(x, b, pro_a=pa, pro_b=pb, rows_per_sample=hw)
    torch.cuda.synchronize()
    xa = torch.relu(x.float().view(ns, hw, k) * pa[:, None, :k] + pb[:, None, :k]).bfloat16().view(m, k)
    r = _ref(xa, b)
    err = (d.float() - r).abs().max().item()
    assert err < 2e-2 * max(1.0, r.abs().max().item()), err


@pytest.mark.parametrize("hw,k,n", [(1024, 64, 128), (256, 128, 128), (64, 256, 128), (32, 64, 32)])
def test_stats_epilogue(g, hw, k, n):
    torch.manual_seed(2)
    ns = 6
    m = ns * hw
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = (torch.randn(n, k, device="cuda") / k ** 0.5).bfloat16()
    table = torch.zeros(ns, n, 2, device="cuda")
    d = g.gemm(a, b, rows_per_sample=hw, stats=table, stats_ns=2 * n)
    torch.cuda.synchronize()
    df = d.float().view(ns, hw, n)
    assert torch.allclose(table[..., 0], df.sum(1), atol=2e-2, rtol=2e-3), (table[..., 0] - df.sum(1)).abs().max()
    assert torch.allclose(table[..., 1], (df * df).sum(1), atol=2e-2, rtol=2e-3)


def test_conv1x1_autograd(g):
    from dynamic_load_balance_distributeddnn_b200.ops import gemm_tc
    torch.manual_seed(3)
    x = torch.randn(8, 96, 16, 16, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(128, 96, 1, 1, device="cuda") / 10).bfloat16().requires_grad_(True)
    y = gemm_tc.conv2d(x, w, None, 1, 0)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True); wr = w.detach().float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr)
    yr.backward(gy.float())
    assert (y.float() - yr).abs().max().item() < 3e-2 * yr.abs().max().item()
    assert (x.grad.float() - xr.grad).abs().max().item() < 3e-2 * xr.grad.abs().max().item()
    assert (w.grad.float() - wr.grad).abs().max().item() < 3e-2 * wr.grad.abs().max().item()


@pytest.mark.parametrize("m,co,ci", [(4096, 128, 64), (65536, 128, 256), (8192, 128, 1024), (1000, 128, 96), (3000, 32, 128),
                                     (512 * 16, 256, 512), (130, 64, 72)])
def test_wgrad_mn_major(g, m, co, ci):
    torch.manual_seed(m + co + ci)
    dy = (torch.randn(m, co, device="cuda") / 8).bfloat16()
    x = torch.randn(m, ci, device="cuda").bfloat16()
    dw = g.wgrad(dy, x)
    torch.cuda.synchronize()
    r = dy.float().t() @ x.float()
    err = (dw - r).abs().max().item()
    assert err < 5e-3 * max(1.0, r.abs().max().item()), (err, r.abs().max().item())


@pytest.mark.parametrize("hw,co,ci", [(1024, 128, 64), (256, 128, 200), (64, 128, 512), (16, 128, 1024)])
def test_wgrad_with_gn_relu_prologue(g, hw, co, ci):
    torch.manual_seed(7)
    ns = 6
    m = ns * hw
    big = torch.randn(m, ci + 64, device="cuda").bfloat16()
    x = big[:, 32:32 + ci] if ci % 8 == 0 else big[:, :ci]          # strided operand (channel slice)
    dy = (torch.randn(m, co, device="cuda") / 8).bfloat16()
    kp = (ci + 63) // 64 * 64
    pa = torch.zeros(ns, kp, device="cuda"); pb = torch.zeros(ns, kp, device="cuda")
    pa[:, :ci] = torch.rand(ns, ci, device="cuda") + 0.5
    pb[:, :ci] = torch.randn(ns, ci, device="cuda") * 0.3
    dw = g.wgrad(dy, x, pa, pb, hw)
    torch.cuda.synchronize()
    xa = torch.relu(x.float().reshape(ns, hw, ci) * pa[:, None, :ci] + pb[:, None, :ci]).bfloat16().reshape(m, ci)
    r = dy.float().t() @ xa.float()
    err = (dw - r).abs().max().item()
    assert err < 5e-3 * max(1.0, r.abs().max().item()), (err, r.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(4096, 256, 128), (65536, 64, 128), (1000, 200, 72), (8192, 1024, 128), (300, 600, 200)])
def test_gemm_b_mn_major(g, m, n, k):
    """D = A[M,K] @ B[K,N] with B row-major [K][N] (dgrad without a transposed weight copy)."""
    torch.manual_seed(m + n)
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = (torch.randn(k, n, device="cuda") / k ** 0.5).bfloat16()
    d = g.gemm_bmn(a, b)
    torch.cuda.synchronize()
    r = a.float() @ b.float()
    err = (d.float() - r).abs().max().item()
    assert err < 2e-2 * max(1.0, r.abs().max().item()), err


def test_linear_autograd_on_tcgen05(g):
    torch.manual_seed(5)
    x = torch.randn(35, 16, 200, device="cuda").bfloat16().requires_grad_(True)
    w = (torch.randn(600, 200, device="cuda") / 14).bfloat16().requires_grad_(True)
    b = torch.randn(600, device="cuda").bfloat16().requires_grad_(True)
    assert g.linear_supported(x, w)
    y = g.linear(x, w, b)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(gy.float())
    assert (y.float() - yr).abs().max().item() < 3e-2 * yr.abs().max().item()
    assert (x.grad.float() - xr.grad).abs().max().item() < 3e-2 * xr.grad.abs().max().item()
    assert (w.grad.float() - wr.grad).abs().max().item() < 3e-2 * wr.grad.abs().max().item()


@pytest.mark.parametrize("n,c,o,hw", [(8, 128, 32, 32), (6, 128, 32, 16), (9, 128, 32, 8), (21, 128, 32, 4), (4, 64, 64, 32),
                                        (5, 96, 160, 16), (3, 256, 256, 8)])
def test_conv3x3_tcgen05_fwd_dgrad(g, n, c, o, hw):
    """implicit-GEMM 3x3 (4-D TMA boxes shifted per tap, zero padding from TMA OOB fill) vs F.conv2d"""
    torch.manual_seed(n + c + o)
    x = torch.randn(n, c, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(o, c, 3, 3, device="cuda") / (3 * c ** 0.5)).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    assert g.conv3x3_geometry_ok(hw, hw)
    y = g._Conv3x3Fn.apply(x, w)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, padding=1)
    yr.backward(gy.float())
    assert (y.float() - yr).abs().max().item() < 3e-2 * max(1.0, yr.abs().max().item())
    assert (x.grad.float() - xr.grad).abs().max().item() < 3e-2 * max(1.0, xr.grad.abs().max().item())
    assert (w.grad.float() - wr.grad).abs().max().item() < 5e-2 * max(1.0, wr.grad.abs().max().item())


def test_conv3x3_strided_io_and_stats(g):
    """input read from / output written into channel slices of wider buffers, with the GN-statistics epilogue"""
    torch.manual_seed(11)
    n, c, o, hw = 6, 128, 32, 16
    xin = torch.randn(n, hw, hw, 192, device="cuda").bfloat16()
    x = xin[..., 32:160]
    w = (torch.randn(o, 3, 3, c, device="cuda") / 34).bfloat16()                  # [O][3][3][I] memory
    out = torch.zeros(n, hw, hw, 96, device="cuda", dtype=torch.bfloat16)
    table = torch.zeros(n, 96, 2, device="cuda")
    g.conv3x3_raw(False, x.data_ptr(), 192, w.data_ptr(), out[..., 16:48].data_ptr(), 96, n, hw, hw, c, o, x.device,
                  table[:, 16:48].data_ptr(), 2 * 96)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    got = out[..., 16:48].float()
    assert (got - ref).abs().max().item() < 3e-2 * ref.abs().max().item()
    assert out[..., :16].abs().sum().item() == 0 and out[..., 48:].abs().sum().item() == 0
    assert torch.allclose(table[:, 16:48, 0], got.sum((1, 2)), atol=5e-2, rtol=5e-3)
    assert torch.allclose(table[:, 16:48, 1], (got * got).sum((1, 2)), atol=5e-2, rtol=5e-3)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float32, 4e-3)])
@pytest.mark.parametrize("n,ci,co,hw", [(8, 128, 32, 32), (6, 128, 32, 16), (16, 128, 32, 8), (32, 128, 32, 4), (4, 64, 64, 32),
                                         (8, 96, 160, 16), (4, 256, 256, 8), (67, 128, 32, 4), (21, 128, 32, 8), (5, 96, 32, 4)])
def test_wgrad3x3_tcgen05(g, dtype, tol, n, ci, co, hw):
    """9-tap MN-major split-K weight gradient (csrc/conv_wgrad.cu): x and dy are channel slices of wider NHWC buffers"""
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(n + ci + co + hw)
    xin = torch.randn(n, hw, hw, ci + 64, device="cuda").to(dtype)
    dyin = (torch.randn(n, hw, hw, co + 96, device="cuda") / 8).to(dtype)
    x, dy = xin[..., 32:32 + ci], dyin[..., 64:64 + co]
    assert g.wgrad3x3_supported(n, hw, hw, ci, dtype)
    dw = torch.zeros(co, 3, 3, ci, device="cuda")
    from dynamic_load_balance_distributeddnn_b200.ops import _native as nat
    g.wgrad3x3_raw(x.data_ptr(), ci + 64, dy.data_ptr(), co + 96, dw, n, hw, hw, ci, co, x.device, dtype=nat.dtype_code(dtype))
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    dyr = dy.float().permute(0, 3, 1, 2).contiguous()
    wr = torch.zeros(co, ci, 3, 3, device="cuda", requires_grad=True)
    torch.nn.functional.conv2d(xr, wr, padding=1).backward(dyr)
    ref = wr.grad.permute(0, 2, 3, 1)                  # [co][3][3][ci]
    err = (dw - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), (err, ref.abs().max().item())


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 3e-2), (torch.float32, 5e-3)])
@pytest.mark.parametrize("t,n,k", [(512, 10, 1024), (64, 10, 1024), (200, 100, 384), (48, 10, 512)])
def test_small_classifier_head_on_tcgen05(g, dtype, tol, t, n, k):
    """N = 10 / 100 heads through the zero rows the flat parameter store keeps behind the matrix (_LinearPadFn)"""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(t + n)
    n_pad = (n + 7) // 8 * 8
    store = torch.zeros(n_pad * k + 64, device="cuda", dtype=dtype)
    w = torch.nn.Parameter(store[:n * k].view(n, k))
    w.data.copy_((torch.randn(n, k, device="cuda") / k ** 0.5).to(dtype))
    w._dlb_padded_rows = n_pad
    x = torch.randn(t, k, device="cuda").to(dtype).requires_grad_(True)
    b = torch.randn(n, device="cuda").to(dtype).requires_grad_(True)
    assert g.linear_supported(x, w)
    y = g.linear(x, w, b)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, wr, br = (v.detach().float().requires_grad_(True) for v in (x, w, b))
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(gy.float())
    assert (y.float() - yr).abs().max().item() < tol * max(1.0, yr.abs().max().item())
    assert (x.grad.float() - xr.grad).abs().max().item() < tol * max(1.0, xr.grad.abs().max().item())
    assert (w.grad.float() - wr.grad).abs().max().item() < tol * max(1.0, wr.grad.abs().max().item())
