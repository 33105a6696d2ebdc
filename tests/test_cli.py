"""CLI parity with reference parser.py:40-80 and artifact naming (§2.8)."""
import pytest

from dynamic_load_balance_distributeddnn_b200.cli import config_from_args, get_parser, gpu_list, str2bool
from dynamic_load_balance_distributeddnn_b200.config import DBSConfig


def test_defaults_match_reference():
    c = config_from_args([])
    assert c.debug is True and c.world_size == 4 and c.batch_size == 64 and c.learning_rate == 0.01
    assert c.epoch_size == 10 and c.dataset == "wikitext2" and c.dynamic_batch_size is True and c.gpu == 0
    assert c.model == "transformer" and c.fault_tolerance is False and c.fault_tolerance_chance == 0.1
    assert c.one_cycle_policy is False and c.disable_enhancements is False


def test_all_13_flags_short_and_long():
    argv = "-d false -ws 8 -b 512 -lr 0.1 -e 3 -ds cifar10 -dbs false -gpu 0,0,0,1 -m densenet -ft true -ftc 0.5 -ocp true -de true".split()
    c = config_from_args(argv)
    assert (c.debug, c.world_size, c.batch_size, c.learning_rate, c.epoch_size) == (False, 8, 512, 0.1, 3)
    assert (c.dataset, c.dynamic_batch_size, c.gpu, c.model) == ("cifar10", False, [0, 0, 0, 1], "densenet")
    assert (c.fault_tolerance, c.fault_tolerance_chance, c.one_cycle_policy, c.disable_enhancements) == (True, 0.5, True, True)
    argv = ("--debug 0 --world_size 2 --batch_size 32 --learning_rate 0.5 --epoch_size 1 --dataset mnist "
            "--dynamic_batch_size yes --gpu 1 --model mnistnet --fault_tolerance n --fault_tolerance_chance 0.2 "
            "--one_cycle_policy t --disable_enhancements f").split()
    c = config_from_args(argv)
    assert c.gpu == 1 and c.model == "mnistnet" and c.debug is False and c.one_cycle_policy is True


def test_validators():
    assert str2bool("Yes") and not str2bool("0")
    with pytest.raises(Exception):
        str2bool("maybe")
    assert gpu_list("0,1,2") == [0, 1, 2] and gpu_list("3") == 3      # bare `-gpu 3` accepted (reference rejects it)
    with pytest.raises(SystemExit):
        get_parser().parse_args(["-m", "vgg"])
    with pytest.raises(SystemExit):
        get_parser().parse_args(["-ds", "imagenet"])
    assert config_from_args(["-m", "resnet50"]).model == "resnet50"   # needed by BASELINE config #3


def test_experiment_id_matches_reference_format():
    c = DBSConfig(model="transformer", dataset="wikitext2", debug=True, world_size=2, batch_size=64,
                  learning_rate=0.01, epoch_size=3)
    assert c.experiment_id(0) == "transformer-wikitext2-debug1-n2-bs64-lr0.0100-ep3-dbs1-ft0-ftc0.100000-node0-ocp0"
    assert c.replace(disable_enhancements=True).experiment_id(1).startswith("puredbs=transformer-")
    assert c.device_for_rank(1) == "cpu"
    g = c.replace(debug=False, gpu=[0, 0, 0, 1])
    assert [g.device_for_rank(r) for r in range(4)] == ["cuda:0", "cuda:0", "cuda:0", "cuda:1"]
    assert c.replace(debug=False, gpu=2).device_for_rank(3) == "cuda:2"


def test_dbs_model_and_profile_extension_flags():
    from dynamic_load_balance_distributeddnn_b200.cli import config_from_args
    cfg = config_from_args(["-ws", "2", "--dbs_model", "affine", "--profile", "true"])
    assert cfg.dbs_model == "affine" and cfg.profile is True
    base = config_from_args([])
    assert base.dbs_model == "auto"
    # auto: the reference's rule in CPU debug mode, the latency-aware model on CUDA devices
    assert config_from_args(["-d", "true"]).resolved_dbs_model() == "proportional"
    assert config_from_args(["-d", "false"]).resolved_dbs_model() == "affine"
    assert config_from_args(["-d", "false", "--dbs_model", "proportional"]).resolved_dbs_model() == "proportional"
