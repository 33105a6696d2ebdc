"""Model-zoo parity: parameter / tensor counts, state_dict interchange and forward
equality with the reference ``Net/*`` classes (imported read-only from /root/reference when mounted)."""
import importlib.util
import os
import sys

import pytest
import torch

from dynamic_load_balance_distributeddnn_b200.models import build_model, model_names

REF = "/root/reference/Net"
EXPECTED = {"mnistnet": (21840, 8), "resnet": (42512970, 314), "resnet50": (23520842, 161),
            "densenet": (6956298, 362), "googlenet": (6166250, 258), "regnet": (5714362, 303),
            "transformer": (13828478, 27)}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_param_counts(name):
    m = build_model(name, 10)
    ps = list(m.parameters())
    assert (sum(p.numel() for p in ps), len(ps)) == EXPECTED[name]


def test_every_registered_model_builds_and_runs():
    x = torch.randn(2, 3, 32, 32)
    for name in model_names():
        if name in ("transformer", "mnistnet") or name in ("resnet152", "densenet161", "densenet201", "resnet101", "resnet"):
            continue
        m = build_model(name, 100).eval()
        with torch.no_grad():
            assert m(x).shape == (2, 100), name


def _ref_module(fname):
    path = os.path.join(REF, fname)
    if not os.path.isfile(path):
        pytest.skip("reference not mounted")
    spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("fname,ctor,ours", [("Densenet.py", "DenseNet121", "densenet"), ("Resnet.py", "ResNet50", "resnet50"),
                                              ("Resnet.py", "ResNet18", "resnet18"), ("RegNet.py", "RegNetY_400MF", "regnet"),
                                              ("MnistNet.py", "MnistNet", "mnistnet")])
def test_state_dict_interchange_and_forward_equality(fname, ctor, ours):
    ref_mod = _ref_module(fname)
    torch.manual_seed(0)
    ref = getattr(ref_mod, ctor)(10) if ctor != "MnistNet" else ref_mod.MnistNet()
    mine = build_model(ours, 10)
    missing = mine.load_state_dict(ref.state_dict(), strict=True)
    ref.eval(); mine.eval()
    x = torch.randn(2, 1, 28, 28) if ours == "mnistnet" else torch.randn(2, 3, 32, 32)
    with torch.no_grad():
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a, b = ref(x), mine(x)
    assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), (a - b).abs().max()


def test_transformer_matches_reference_module():
    ref_mod = _ref_module("Transformer.py")
    torch.manual_seed(0)
    ref = ref_mod.TransformerModel(1000, 200, 2, 200, 2, 0.2).eval()
    mine = build_model("transformer", ntoken=1000).eval()
    sd = ref.state_dict()
    mine.load_state_dict(sd, strict=True)
    src = torch.randint(0, 1000, (35, 3))
    with torch.no_grad():
        a, b = ref(src), mine(src)
    assert a.shape == b.shape == (35, 3, 1000)
    assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), (a - b).abs().max()
    tgt = torch.randint(0, 1000, (35 * 3,))
    l1 = torch.nn.functional.nll_loss(b.view(-1, 1000), tgt)
    l2 = mine.forward_loss(src, tgt)
    assert torch.allclose(l1, l2, atol=1e-4)


def test_googlenet_fixed_order_runs_backward():
    m = build_model("googlenet", 10)
    out = m(torch.randn(2, 3, 32, 32))
    out.sum().backward()
    assert all(p.grad is not None for p in m.parameters())
