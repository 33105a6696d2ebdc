"""Numerics of the hand-written sm_100a kernels against plain PyTorch fp32 references (run on the B200
box: ``gpurun -- python -m pytest tests -m gpu``)."""
import math

import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    from dynamic_load_balance_distributeddnn_b200.ops import _native
    assert _native.available(), "native library must be built and loadable on the GPU box"
    return _native


def _cl(x):
    return x.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,groups", [((4, 64, 32, 32), 32), ((3, 128, 16, 16), 32), ((5, 96, 8, 8), 8),
                                          ((2, 1024, 4, 4), 32), ((2, 20, 7, 5), 4), ((2, 32, 9, 9), 32)])
@pytest.mark.parametrize("relu,res", [(True, False), (False, False), (True, True)])
def test_group_norm_act_fwd_bwd(nat, dtype, shape, groups, relu, res):
    from dynamic_load_balance_distributeddnn_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    x = _cl(torch.randn(shape, device=dev) * 2 + 0.5).to(dtype).requires_grad_(True)
    w = (torch.rand(shape[1], device=dev) + 0.5).requires_grad_(True)
    b = (torch.randn(shape[1], device=dev) * 0.1).requires_grad_(True)
    r = _cl(torch.randn(shape, device=dev)).to(dtype).requires_grad_(True) if res else None
    y = ops.group_norm_act(x, groups, w, b, 1e-5, relu, r)
    gy = _cl(torch.randn(shape, device=dev)).to(dtype)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    rr = r.detach().float().requires_grad_(True) if res else None
    yr = ops.group_norm_act_reference(xr, groups, wr, br, 1e-5, relu, rr)
    yr.backward(gy.float())
    tol = 2e-4 if dtype == torch.float32 else 3e-2
    assert torch.allclose(y.float(), yr, atol=tol, rtol=tol), (y.float() - yr).abs().max()
    gtol = 1e-3 if dtype == torch.float32 else 6e-2
    assert torch.allclose(x.grad.float(), xr.grad, atol=gtol, rtol=gtol), (x.grad.float() - xr.grad).abs().max()
    scale = max(1.0, float(wr.grad.abs().max()))
    assert float((w.grad - wr.grad).abs().max()) < (2e-3 if dtype == torch.float32 else 5e-2) * scale
    assert float((b.grad - br.grad).abs().max()) < (2e-3 if dtype == torch.float32 else 5e-2) * max(1.0, float(br.grad.abs().max()))
    if res:
        assert torch.allclose(r.grad.float(), rr.grad, atol=gtol, rtol=gtol)


def test_group_norm_on_channel_slice(nat):
    """Channel slices of a wider NHWC buffer (DenseNet concat buffer) are read in place."""
    from dynamic_load_balance_distributeddnn_b200 import ops
    torch.manual_seed(1)
    buf = _cl(torch.randn(3, 256, 8, 8, device="cuda")).bfloat16()
    x = buf[:, 64:192]
    w = torch.rand(128, device="cuda") + 0.5
    b = torch.randn(128, device="cuda")
    y = ops.group_norm_act(x, 32, w, b)
    yr = ops.group_norm_act_reference(x.float(), 32, w, b)
    assert torch.allclose(y.float(), yr, atol=3e-2, rtol=3e-2)


def test_pack_sumsq_sgd(nat):
    from dynamic_load_balance_distributeddnn_b200.models import build_model
    from dynamic_load_balance_distributeddnn_b200.parallel import FlatState, SingleComm
    torch.manual_seed(0)
    m_ref = build_model("resnet18", 10).cuda()
    m = build_model("resnet18", 10)
    m.load_state_dict(m_ref.state_dict())
    flat = FlatState(m, "cuda", torch.float32, SingleComm(), lr=0.1, momentum=0.9, clip_norm=0.5)
    flat.set_weights([1.0])
    opt = torch.optim.SGD(m_ref.parameters(), lr=0.1, momentum=0.9)
    for it in range(3):
        x = torch.randn(4, 3, 32, 32, device="cuda")
        y = torch.randint(0, 10, (4,), device="cuda")
        for mod in (m, m_ref):
            mod.train()
        F.cross_entropy(m(x), y).backward()
        F.cross_entropy(m_ref(x), y).backward()
        torch.nn.utils.clip_grad_norm_(m_ref.parameters(), 0.5)
        opt.step(); opt.zero_grad()
        flat.reduce_and_step(0); flat.zero_grad()
    for p, q in zip(m.parameters(), m_ref.parameters()):
        assert torch.allclose(p, q, atol=3e-4, rtol=1e-3), (p - q).abs().max()


def test_bf16_shadow_tracks_master(nat):
    from dynamic_load_balance_distributeddnn_b200.models import build_model
    from dynamic_load_balance_distributeddnn_b200.parallel import FlatState, SingleComm
    m = build_model("densenet", 10)
    flat = FlatState(m, "cuda", torch.bfloat16, SingleComm(), lr=0.05)
    flat.set_weights([1.0])
    assert m.conv1.weight.dtype == torch.bfloat16 and m.dense1[0].gn1.weight.dtype == torch.float32
    x = torch.randn(4, 3, 32, 32, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    F.cross_entropy(m(x).float(), torch.randint(0, 10, (4,), device="cuda")).backward()
    flat.reduce_and_step(0)
    torch.cuda.synchronize()
    assert torch.equal(flat.shadow, flat.master.bfloat16())
    assert flat.mom.abs().sum() > 0


def test_augment_matches_reference_semantics(nat):
    from dynamic_load_balance_distributeddnn_b200 import ops
    u8 = torch.randint(0, 256, (16, 32, 32, 3), dtype=torch.uint8, device="cuda")
    mean, std = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
    x = ops.augment(u8, mean, std, 0, False, dtype=torch.float32)
    ref = (u8.float() / 255 - torch.tensor(mean, device="cuda")) / torch.tensor(std, device="cuda")
    assert torch.allclose(x.permute(0, 2, 3, 1), ref, atol=1e-5)
    xa = ops.augment(u8, mean, std, 4, True, seed=1, step=2, dtype=torch.bfloat16)
    assert xa.shape == (16, 3, 32, 32) and xa.dtype == torch.bfloat16 and torch.isfinite(xa.float()).all()
    # every output row is either padding (normalised zero) or a (possibly flipped/shifted) copy of the source
    z = ((0 - torch.tensor(mean)) / torch.tensor(std)).bfloat16().float()
    vals = xa.float().permute(0, 2, 3, 1)[0, 0, 0].cpu()
    src = ((u8[0].float() / 255).cpu() - torch.tensor(mean)) / torch.tensor(std)
    assert torch.allclose(vals, z, atol=2e-2) or ((src.bfloat16().float() - vals).abs().sum(-1) < 5e-2).any()


def test_single_rank_collectives(nat):
    from dynamic_load_balance_distributeddnn_b200.parallel import SymmComm
    c = SymmComm("cuda")
    gin, gout = c.alloc_grad_buffers(1 << 16, torch.float32, "cuda")
    gin.copy_(torch.randn(1 << 16, device="cuda"))
    w = torch.tensor([0.5], device="cuda")
    for algo in ("oneshot", "twoshot"):
        c.algo = algo
        gout.zero_()
        c.allreduce_buckets(gin, gout, [(0, 1 << 15), (1 << 15, 1 << 15)], w)
        torch.cuda.synchronize()
        assert torch.allclose(gout, 0.5 * gin)
    c.barrier()
    assert c.gather_times(1.25) == [1.25]
    c.check_errors()
    c.close()


def test_whole_step_cuda_graph_matches_eager(nat, tmp_path):
    """Graph replay and eager execution of the same steps give the same parameters."""
    from dynamic_load_balance_distributeddnn_b200.config import DBSConfig
    from dynamic_load_balance_distributeddnn_b200.engine import Trainer
    from dynamic_load_balance_distributeddnn_b200.utils import init_logger
    res = []
    for graphs in (False, True):
        cfg = DBSConfig(debug=False, world_size=1, batch_size=32, model="resnet18", dataset="cifar10", synthetic=True,
                        train_samples=32 * 8, test_samples=64, epoch_size=1, validate=False, cuda_graphs=graphs,
                        dtype="fp32", log_dir=str(tmp_path / f"l{int(graphs)}"), stats_dir=str(tmp_path / "s"))
        t = Trainer(cfg, 0, 1, "cuda:0", init_logger(cfg, 0, stream=False))
        t.train_set.pad, t.train_set.flip = 0, False          # deterministic inputs
        for s in range(6):
            xb, yb = t.stager.stage(list(range(s * 32, s * 32 + 32)))
            t.train_step(xb, yb)
            t.stager.release()
        torch.cuda.synchronize()
        assert graphs == bool(t._graphs)
        res.append((t.flat.master.clone(), float(t.loss_acc.item())))
        t.close()
    assert abs(res[0][1] - res[1][1]) < 2e-2 * abs(res[0][1])
    assert float((res[0][0] - res[1][0]).abs().max()) < 5e-3


def test_densenet_trains_bf16(nat, tmp_path):
    from dynamic_load_balance_distributeddnn_b200.config import DBSConfig
    from dynamic_load_balance_distributeddnn_b200.engine import Trainer
    from dynamic_load_balance_distributeddnn_b200.utils import init_logger
    cfg = DBSConfig(debug=False, world_size=1, batch_size=64, model="densenet", dataset="cifar10", synthetic=True,
                    train_samples=64 * 12, test_samples=128, epoch_size=2, validate=True, learning_rate=0.05,
                    log_dir=str(tmp_path / "l"), stats_dir=str(tmp_path / "s"))
    t = Trainer(cfg, 0, 1, "cuda:0", init_logger(cfg, 0, stream=False))
    rec = t.run()
    assert rec.data["train_loss"][-1] < rec.data["train_loss"][0]
    assert math.isfinite(rec.data["val_loss"][-1])
    t.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_avg_pool_matches_torch(nat, dtype):
    from dynamic_load_balance_distributeddnn_b200 import ops
    for shape, k in (((3, 64, 8, 8), 2), ((2, 128, 4, 4), 4), ((2, 24, 8, 8), 8), ((2, 20, 6, 6), 2)):
        x = _cl(torch.randn(shape, device="cuda")).to(dtype).requires_grad_(True)
        y = ops.avg_pool2d(x, k)
        gy = torch.randn_like(y)
        y.backward(gy)
        xr = x.detach().float().requires_grad_(True)
        yr = F.avg_pool2d(xr, k)
        yr.backward(gy.float())
        tol = 1e-5 if dtype == torch.float32 else 2e-2
        assert torch.allclose(y.float(), yr, atol=tol, rtol=tol)
        assert torch.allclose(x.grad.float(), xr.grad, atol=tol, rtol=tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 8e-2)])
def test_dense_block_fused_matches_unfused(nat, dtype, tol):
    """Concat-free dense stage (in-place buffer, stats table, hand-written backward) == cat-based autograd."""
    from dynamic_load_balance_distributeddnn_b200.models import densenet
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False          # compare the two paths, not TF32 algorithm choices
    stage = densenet.DenseNet([3], growth_rate=32, num_classes=10).dense1.cuda()
    trans_gn = densenet.GroupNormAct(32, 64 + 3 * 32).cuda()
    for m in stage.modules():
        if isinstance(m, torch.nn.Conv2d):
            m.weight.data = m.weight.data.to(dtype)
    x0 = _cl(torch.randn(4, 64, 16, 16, device="cuda")).to(dtype)
    outs = []
    for fused in (False, True):
        densenet.DenseStage.fused = fused
        for p in list(stage.parameters()) + list(trans_gn.parameters()):
            p.grad = None
        x = x0.clone().requires_grad_(True)
        y = trans_gn(stage(x))                     # the following GN consumes the stats table in the fused path
        gy = torch.randn(y.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(5)).to(dtype)
        y.backward(gy)
        outs.append((y.detach().float(), x.grad.float(), [p.grad.float().clone() for p in stage.parameters()],
                     [p.grad.float().clone() for p in trans_gn.parameters()]))
    densenet.DenseStage.fused = True
    (y0, dx0, g0, t0), (y1, dx1, g1, t1) = outs
    torch.backends.cudnn.allow_tf32 = True
    assert torch.allclose(y0, y1, atol=tol, rtol=tol), (y0 - y1).abs().max()
    # gradients: relative L2.  The two paths round differently (statistics from the GEMM epilogue vs from the stored tensor, tap
    # order of the 3x3), which flips a handful of ReLU decisions sitting at ~0: individual gradient entries then differ by a
    # few % of the maximum while the gradient as a whole agrees to 1e-3 -- a max-norm bound on them is a coin toss
    def rel_l2(a, b):
        return float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-12))
    # fp32 case: the fused path computes on TF32 tensor cores against a true-fp32 cat-based reference -- 2^-11 operand rounding
    # plus the flipped ReLU decisions give 0.5-1 % relative L2 on the deepest weight gradients (measured 0.0103 on B200)
    gtol = 2.5e-2 if dtype == torch.float32 else 5 * tol
    assert rel_l2(dx1, dx0) < gtol, rel_l2(dx1, dx0)
    for i, (a, b) in enumerate(zip(g0 + t0, g1 + t1)):
        assert rel_l2(b, a) < gtol, (i, tuple(a.shape), rel_l2(b, a))


@pytest.mark.parametrize("model,dataset,bs", [("mnistnet", "mnist", 32), ("resnet50", "cifar10", 16), ("googlenet", "cifar10", 16),
                                                ("regnet", "cifar100", 16), ("resnet18", "cifar10", 32)])
def test_every_family_trains_on_gpu(nat, tmp_path, model, dataset, bs):
    """bf16 flat-state training step through the fused GN / tcgen05 paths for every CNN family."""
    from dynamic_load_balance_distributeddnn_b200.config import DBSConfig
    from dynamic_load_balance_distributeddnn_b200.engine import Trainer
    from dynamic_load_balance_distributeddnn_b200.utils import init_logger
    cfg = DBSConfig(debug=False, world_size=1, batch_size=bs, model=model, dataset=dataset, synthetic=True,
                    train_samples=bs * 6, test_samples=64, epoch_size=1, validate=True, learning_rate=0.02,
                    log_dir=str(tmp_path / "l"), stats_dir=str(tmp_path / "s"))
    t = Trainer(cfg, 0, 1, "cuda:0", init_logger(cfg, 0, stream=False))
    before = t.flat.master.clone()
    rec = t.run()
    assert math.isfinite(rec.data["train_loss"][-1]) and math.isfinite(rec.data["val_loss"][-1])
    assert not torch.equal(before, t.flat.master)
    t.close()


def test_transformer_lm_trains_bf16(nat, tmp_path):
    from dynamic_load_balance_distributeddnn_b200.config import DBSConfig
    from dynamic_load_balance_distributeddnn_b200.engine import Trainer
    from dynamic_load_balance_distributeddnn_b200.utils import init_logger
    cfg = DBSConfig(debug=False, world_size=1, batch_size=16, model="transformer", dataset="wikitext2", synthetic=True,
                    train_samples=16 * 35 * 14, test_samples=4000, epoch_size=2, validate=True, learning_rate=0.5,
                    log_dir=str(tmp_path / "l"), stats_dir=str(tmp_path / "s"))
    t = Trainer(cfg, 0, 1, "cuda:0", init_logger(cfg, 0, stream=False))
    rec = t.run()
    assert rec.data["train_loss"][-1] < rec.data["train_loss"][0] + 0.5 and math.isfinite(rec.data["val_loss"][-1])
    t.close()


def test_linear_cross_entropy_chunked_matches_reference(nat):
    from dynamic_load_balance_distributeddnn_b200 import ops
    torch.manual_seed(0)
    feats = torch.randn(300, 200, device="cuda", requires_grad=True)
    w = (torch.randn(5000, 200, device="cuda") * 0.05).requires_grad_(True)
    b = torch.zeros(5000, device="cuda", requires_grad=True)
    tgt = torch.randint(0, 5000, (300,), device="cuda")
    l1 = ops.linear_cross_entropy(feats, w, b, tgt, chunk=128)
    l1.backward()
    g1 = (feats.grad.clone(), w.grad.clone(), b.grad.clone())
    feats.grad = w.grad = b.grad = None
    l2 = ops.linear_cross_entropy_reference(feats, w, b, tgt)
    l2.backward()
    assert torch.allclose(l1, l2, atol=1e-4)
    for a, r in zip(g1, (feats.grad, w.grad, b.grad)):
        assert torch.allclose(a, r, atol=2e-5, rtol=1e-3), (a - r).abs().max()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
def test_add_layer_norm_native(nat, dtype, tol):
    from dynamic_load_balance_distributeddnn_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(35, 16, 200, device="cuda").to(dtype).requires_grad_(True)
    r = torch.randn(35, 16, 200, device="cuda").to(dtype).requires_grad_(True)
    w = (torch.rand(200, device="cuda") + 0.5).requires_grad_(True)
    b = torch.randn(200, device="cuda").requires_grad_(True)
    y = ops.add_layer_norm(x, r, w, b)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, rr = x.detach().float().requires_grad_(True), r.detach().float().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = ops.add_layer_norm_reference(xr, rr, wr, br)
    yr.backward(gy.float())
    assert torch.allclose(y.float(), yr, atol=tol, rtol=tol)
    assert torch.allclose(x.grad.float(), xr.grad, atol=tol * 3, rtol=tol * 3) and torch.allclose(r.grad.float(), rr.grad, atol=tol * 3, rtol=tol * 3)
    assert float((w.grad - wr.grad).abs().max()) < tol * 30 and float((b.grad - br.grad).abs().max()) < tol * 30


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
def test_fused_attention_matches_reference(nat, dtype, tol):
    from dynamic_load_balance_distributeddnn_b200.ops import attention
    from dynamic_load_balance_distributeddnn_b200 import ops
    torch.manual_seed(1)
    s, b, h, hd = 35, 6, 2, 100
    qkv = (torch.randn(s, b, 3 * h * hd, device="cuda") * 0.5).to(dtype).requires_grad_(True)
    out = attention.causal_attention_packed(qkv, h, 0.0)
    go = torch.randn_like(out)
    out.backward(go)
    ref_in = qkv.detach().float().requires_grad_(True)
    q, k, v = ref_in.view(s, b, 3, h, hd).permute(2, 1, 3, 0, 4)
    o = ops.causal_attention_reference(q, k, v).permute(2, 0, 1, 3).reshape(s, b, h * hd)
    o.backward(go.float())
    assert torch.allclose(out.float(), o, atol=tol, rtol=tol), (out.float() - o).abs().max()
    assert torch.allclose(qkv.grad.float(), ref_in.grad, atol=tol * 2, rtol=tol * 2), (qkv.grad.float() - ref_in.grad).abs().max()
    # dropout: kept fraction ~ (1-p), rows renormalised by 1/(1-p), same mask regenerated in backward
    out_d = attention.causal_attention_packed(qkv.detach().requires_grad_(True), h, 0.2)
    assert torch.isfinite(out_d.float()).all() and (out_d.float() - out.float()).abs().max() > 0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_fused_linear_cross_entropy_native(nat, dtype, tol):
    from dynamic_load_balance_distributeddnn_b200.ops import lm_native
    from dynamic_load_balance_distributeddnn_b200 import ops
    assert lm_native.has_linear_ce()
    torch.manual_seed(0)
    t, d, v = 700, 200, 33278
    feats = torch.randn(t, d, device="cuda").to(dtype).requires_grad_(True)
    w = (torch.randn(v, d, device="cuda") * 0.05).to(dtype).requires_grad_(True)
    b = (torch.randn(v, device="cuda") * 0.1).to(dtype).requires_grad_(True)
    tgt = torch.randint(0, v, (t,), device="cuda")
    l1 = lm_native.linear_cross_entropy(feats, w, b, tgt)
    l1.backward()
    fr, wr, br = (x.detach().float().requires_grad_(True) for x in (feats, w, b))
    l2 = ops.linear_cross_entropy_reference(fr, wr, br, tgt)
    l2.backward()
    assert abs(float(l1) - float(l2)) < tol * max(1.0, float(l2)), (float(l1), float(l2))
    for a, r in ((feats.grad, fr.grad), (w.grad, wr.grad), (b.grad, br.grad)):
        assert float((a.float() - r).abs().max()) < tol * max(1e-3, float(r.abs().max())) * 3, (a.float() - r).abs().max()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_fused_embedding_frontend(nat, dtype, tol):
    """gather * sqrt(d) + positional table (+ dropout) in one kernel, scatter-add backward (csrc/embed.cu)"""
    from dynamic_load_balance_distributeddnn_b200.ops import embedding as E
    torch.manual_seed(3)
    v, d, s, b = 1000, 200, 35, 64
    w = (torch.randn(v, d, device="cuda") * 0.1).to(dtype).requires_grad_(True)
    pe = torch.randn(5000, 1, d, device="cuda")
    tok = torch.randint(0, v, (s, b), device="cuda")
    assert E.supported(tok, w)
    out = E.embed_pe_dropout(tok, w, pe, 0.2, training=False)
    ref = F.embedding(tok, w.float()) * math.sqrt(d) + pe[:s]
    assert (out.float() - ref).abs().max().item() < tol * ref.abs().max().item()
    go = torch.randn_like(out)
    out.backward(go)
    wr = w.detach().float().requires_grad_(True)
    (F.embedding(tok, wr) * math.sqrt(d)).backward(go.float())
    assert (w.grad.float() - wr.grad).abs().max().item() < max(tol, 1e-4) * wr.grad.abs().max().item()
    # dropout: ~p of the elements are zero, the survivors are scaled by 1/(1-p), and the backward uses the same mask
    w.grad = None
    o2 = E.embed_pe_dropout(tok, w, pe, 0.25, training=True)
    frac = (o2 == 0).float().mean().item()
    assert abs(frac - 0.25) < 0.02, frac
    keep = o2 != 0
    assert (o2.float()[keep] - (ref / 0.75)[keep]).abs().max().item() < max(tol, 1e-4) * ref.abs().max().item() / 0.75 + 1e-3
    o2.backward(torch.ones_like(o2))
    cnt = torch.zeros(v, d, device="cuda").index_put_((tok.reshape(-1),), keep.reshape(-1, d).float(), accumulate=True)
    assert (w.grad.float() - cnt * math.sqrt(d) / 0.75).abs().max().item() < 2e-2 * (cnt.max().item() * math.sqrt(d) / 0.75)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("n,c,cs,hw", [(8, 32, 8, 32), (5, 160, 40, 8), (3, 384, 96, 4)])
def test_fused_squeeze_excite(nat, dtype, tol, n, c, cs, hw):
    from dynamic_load_balance_distributeddnn_b200.ops import se
    torch.backends.cudnn.allow_tf32 = False          # the reference below must be true fp32 (cuDNN defaults to TF32 convolutions)
    torch.manual_seed(n + c)
    x = _cl(torch.randn(n, c, hw, hw, device="cuda")).to(dtype).requires_grad_(True)
    w1 = (torch.randn(cs, c, 1, 1, device="cuda") / c ** 0.5).to(dtype).requires_grad_(True)
    b1 = (torch.randn(cs, device="cuda") * 0.1).to(dtype).requires_grad_(True)
    w2 = (torch.randn(c, cs, 1, 1, device="cuda") / cs ** 0.5).to(dtype).requires_grad_(True)
    b2 = (torch.randn(c, device="cuda") * 0.1).to(dtype).requires_grad_(True)
    assert se.supported(x, w1, b1, w2, b2)
    out = se.squeeze_excite(x, w1, b1, w2, b2)
    go = torch.randn_like(out)
    out.backward(go)
    ps = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    s = F.adaptive_avg_pool2d(ps[0], (1, 1))
    ref = ps[0] * F.conv2d(F.relu(F.conv2d(s, ps[1], ps[2])), ps[3], ps[4]).sigmoid()
    ref.backward(go.float())
    assert (out.float() - ref).abs().max().item() < tol * ref.abs().max().item()
    for a, r in zip((x, w1, b1, w2, b2), ps):
        assert (a.grad.float() - r.grad).abs().max().item() < 3 * tol * max(1e-3, r.grad.abs().max().item()), a.shape


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("n,co,hw", [(16, 64, 32), (5, 32, 28), (3, 128, 8)])
def test_stem_conv_kernel(nat, dtype, tol, n, co, hw):
    """RGB stem 3x3 convolution (direct SIMT kernel) forward + weight gradient vs F.conv2d in true fp32"""
    from dynamic_load_balance_distributeddnn_b200 import ops
    from dynamic_load_balance_distributeddnn_b200.ops import stem
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(n + co)
    x = _cl(torch.randn(n, 3, hw, hw, device="cuda")).to(dtype)
    w = _cl(torch.randn(co, 3, 3, 3, device="cuda") / 5).to(dtype).requires_grad_(True)
    assert stem.supported(x, w, 1, 1, 1)
    y = ops.conv2d(x, w, None, 1, 1)
    gy = torch.randn_like(y)
    y.backward(gy)
    wr = w.detach().float().requires_grad_(True)
    yr = F.conv2d(x.float(), wr, padding=1)
    yr.backward(gy.float())
    assert (y.float() - yr).abs().max().item() < tol * max(1.0, yr.abs().max().item())
    assert (w.grad.float() - wr.grad).abs().max().item() < 2 * tol * max(1.0, wr.grad.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,c,hw,k,s,p", [(4, 64, 32, 3, 2, 1), (3, 480, 16, 3, 1, 1), (5, 20, 8, 2, 2, 0), (2, 10, 24, 2, 2, 0)])
def test_max_pool_kernel(nat, dtype, n, c, hw, k, s, p):
    from dynamic_load_balance_distributeddnn_b200 import ops
    torch.manual_seed(n + c)
    x = _cl(torch.randn(n, c, hw, hw, device="cuda")).to(dtype).requires_grad_(True)
    y = ops.max_pool2d(x, k, s, p)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    yr = F.max_pool2d(xr, k, s, p)
    yr.backward(gy.float())
    assert torch.equal(y.float(), yr)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert (x.grad.float() - xr.grad).abs().max().item() <= tol * max(1.0, xr.grad.abs().max().item())


def test_global_avg_pool_routes_to_kernel(nat):
    from dynamic_load_balance_distributeddnn_b200 import ops
    x = _cl(torch.randn(6, 384, 4, 4, device="cuda")).bfloat16()
    assert (ops.global_avg_pool2d(x).float() - F.adaptive_avg_pool2d(x.float(), 1)).abs().max().item() < 2e-2


@pytest.mark.parametrize("bulk", [1, 0])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n,hw,c,ct,groups,acc", [(8, 64, 128, 128, 32, 0), (6, 256, 96, 160, 32, 1), (3, 1024, 224, 256, 32, 1),
                                                  (5, 16, 512, 1024, 32, 1), (4, 64, 992, 1024, 32, 1), (67, 16, 128, 128, 32, 0)])
def test_fused_gn_backward_kernel_both_flavours(bulk, dtype, n, hw, c, ct, groups, acc):
    """dlb_gn_bwd_fused (reduce + per-sample barrier + apply in one launch), register flavour and bulk-copy (TMA unit) flavour,
    against the closed-form GroupNorm(+ReLU-from-coefficients) backward in fp64.  x / dX are channel slices of wider buffers."""
    from dynamic_load_balance_distributeddnn_b200.ops import _native as nat
    lib = nat.require()
    torch.manual_seed(n * hw + c)
    dev = "cuda"
    xbig = torch.randn(n, hw, ct, device=dev).to(dtype)
    dbig = torch.randn(n, hw, ct, device=dev).to(dtype)
    off = ct - c
    x, dxv = xbig[..., off:], dbig[..., off:]
    dy = torch.randn(n, hw, c, device=dev).to(dtype)
    before = dbig.clone()
    gamma = torch.rand(c, device=dev) + 0.5
    mean = torch.randn(n, groups, device=dev) * 0.2
    rstd = torch.rand(n, groups, device=dev) + 0.5
    kp = (c + 63) // 64 * 64
    ca = torch.zeros(n, kp, device=dev); cb = torch.zeros(n, kp, device=dev)
    ca[:, :c] = torch.rand(n, c, device=dev) + 0.5
    cb[:, :c] = torch.randn(n, c, device=dev) * 0.3
    z = ca[:, None, :c].double() * x.double() + cb[:, None, :c].double()
    x.masked_fill_(z.abs() < 1e-3, 2.0)                       # keep the recomputed ReLU mask off its decision boundary
    table = torch.zeros(n, 2 * c, device=dev)
    dg = torch.zeros(c, device=dev); db = torch.zeros(c, device=dev)
    done = torch.zeros(n, dtype=torch.int32, device=dev)
    lib.dlb_norm_bulk(bulk, 0)
    restore = 1 if os.environ.get("DLB_GN_BWD_BULK", "0") == "1" else 0
    try:
        rc = lib.dlb_gn_bwd_fused(nat.dtype_code(dtype), x.data_ptr(), ct, dy.data_ptr(), c, dxv.data_ptr(), ct, gamma.data_ptr(),
                                  mean.data_ptr(), rstd.data_ptr(), table.data_ptr(), 0, dg.data_ptr(), db.data_ptr(), ca.data_ptr(),
                                  cb.data_ptr(), kp, done.data_ptr(), n, hw, c, groups, acc, nat.stream_ptr(torch.device(dev)))
    finally:
        lib.dlb_norm_bulk(restore, 0)
    if rc == 1 and not bulk:
        pytest.skip("shape not covered by the register flavour (caller falls back to the two-kernel chain)")
    assert rc == 0, rc
    torch.cuda.synchronize()
    xd, gd = x.double(), dy.double()
    zz = ca[:, None, :c].double() * xd + cb[:, None, :c].double()
    dz = gd * (zz > 0)
    A, B = dz.sum(1), (dz * xd).sum(1)                                        # [n, c]
    cpg = c // groups
    mu = mean.double().repeat_interleave(cpg, 1); r = rstd.double().repeat_interleave(cpg, 1)
    s1 = (gamma.double() * A).view(n, groups, cpg).sum(-1).repeat_interleave(cpg, 1)
    s2 = (gamma.double() * r * (B - mu * A)).view(n, groups, cpg).sum(-1).repeat_interleave(cpg, 1)
    m = cpg * hw
    q = r * r * s2 / m
    ref = (gamma.double() * r)[:, None, :] * dz - q[:, None, :] * xd + (-r * s1 / m + q * mu)[:, None, :]
    if acc:
        ref = ref + before[..., off:].double()
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-4
    err = (dxv.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), err
    if off:
        assert torch.equal(dbig[..., :off], before[..., :off])
    assert torch.allclose(db.double(), A.sum(0), rtol=1e-3, atol=1e-2)
    assert torch.allclose(dg.double(), (r * (B - mu * A)).sum(0), rtol=1e-3, atol=2e-2)
