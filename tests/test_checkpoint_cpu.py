"""Checkpoint / resume on CPU (absent from the reference)."""
import torch

from dynamic_load_balance_distributeddnn_b200.config import DBSConfig
from dynamic_load_balance_distributeddnn_b200.engine import Trainer
from dynamic_load_balance_distributeddnn_b200.utils import init_logger
from dynamic_load_balance_distributeddnn_b200.utils.checkpoint import checkpoint_path


def test_resume_continues_from_saved_epoch(tmp_path):
    torch.set_num_threads(2)
    base = dict(debug=True, world_size=1, batch_size=32, model="mnistnet", dataset="mnist", synthetic=True, train_samples=256,
                test_samples=64, validate=False, checkpoint_dir=str(tmp_path / "ck"), log_dir=str(tmp_path / "l"),
                stats_dir=str(tmp_path / "s"))
    cfg = DBSConfig(epoch_size=2, **base)
    t = Trainer(cfg, 0, 1, "cpu", init_logger(cfg, 0, stream=False))
    t.run()
    w_end = t.flat.master.clone()
    cfg2 = DBSConfig(epoch_size=3, resume=True, **base)
    assert checkpoint_path(cfg) == checkpoint_path(cfg2)                 # independent of the epoch budget
    t2 = Trainer(cfg2, 0, 1, "cpu", init_logger(cfg2, 0, stream=False))
    assert t2.start_epoch == 2 and torch.equal(t2.flat.master, w_end)
    rec = t2.run()
    assert rec.data["epoch"] == [0, 1, 2]                               # the stats history is restored and continued
    assert t2.global_step == 3 * (256 // 32)                            # step counters continue too (augmentation stream)
