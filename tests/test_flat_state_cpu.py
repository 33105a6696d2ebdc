"""FlatState (flat parameter / gradient / momentum store) on CPU: clipping modes, checkpoint layout."""
import torch

from dynamic_load_balance_distributeddnn_b200.parallel import FlatState
from dynamic_load_balance_distributeddnn_b200.parallel.comm import SingleComm


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))


def _step(clip_mode, clip):
    m = _model()
    fs = FlatState(m, "cpu", torch.float32, SingleComm(), lr=0.1, momentum=0.0, clip_norm=clip, clip_mode=clip_mode)
    fs.set_weights([1.0])
    before = fs.master.clone()
    x = torch.randn(32, 8) * 10
    m(x).pow(2).sum().backward()
    g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    fs.reduce_and_step(0)
    return before, fs, g


def test_global_clip_clips_the_reduced_gradient():
    before, fs, g = _step("global", 0.25)
    delta = (before - fs.master)[:fs.base_numel]
    assert g.norm() > 1.0                                          # the raw gradient is far above the threshold
    assert abs(float(delta.norm()) / 0.1 - 0.25) < 1e-3            # update = lr * clipped gradient, |clipped| = 0.25


def test_local_clip_matches_reference_semantics():
    before, fs, g = _step("local", 0.25)
    delta = (before - fs.master)[:fs.base_numel]
    assert abs(float(delta.norm()) / 0.1 - 0.25) < 1e-3            # world 1: local == global


def test_no_clip_leaves_gradient_alone():
    before, fs, g = _step("local", 0.0)
    delta = (before - fs.master)[:fs.base_numel]
    assert abs(float(delta.norm()) / 0.1 - float(g.norm())) < 1e-2 * float(g.norm())


def test_state_dict_is_world_size_independent():
    _, fs, _ = _step("local", 0.0)
    sd = fs.state_dict()
    assert sd["master"].numel() == fs.base_numel <= fs.numel
    fs2 = FlatState(_model(), "cpu", torch.float32, SingleComm(), lr=0.1)
    fs2.load_state_dict(sd)
    assert torch.equal(fs2.master[:fs.base_numel], fs.master[:fs.base_numel])
