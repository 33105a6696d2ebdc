"""Artifacts and compatibility surface: log line format / file names, stats schema, Net aliases,
run.sh interface, launcher backend selection."""
import os
import re
import subprocess
import sys

import numpy as np

from dynamic_load_balance_distributeddnn_b200.config import DBSConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log_line_format_and_file_names(tmp_path):
    from dynamic_load_balance_distributeddnn_b200.utils import done_marker, init_logger, log_path
    cfg = DBSConfig(model="densenet", dataset="cifar10", debug=True, world_size=4, batch_size=512, log_dir=str(tmp_path / "logs"))
    lg = init_logger(cfg, 2, stream=False)
    lg.info("Rank 2, epoch 0: 10, train_loss 3.21")
    for h in lg.logger.handlers:
        h.flush()
    path = log_path(cfg, 2)
    assert os.path.basename(path) == "densenet-cifar10-debug1-n4-bs512-lr0.0100-ep10-dbs1-ft0-ftc0.100000-node2-ocp0.log"
    line = open(path).read().strip()
    # "<asctime> [ws:lr:dbs_x:ft_y] [file:line] LEVEL message"   (reference dbs_logging.py:21-22)
    assert re.match(r"^\d{4}-\d\d-\d\d \d\d:\d\d:\d\d,\d{3} \[4:0\.01:dbs_enabled:ft_disabled\] \[\w+\.py:\d+\] INFO Rank 2, epoch 0: 10, train_loss 3\.21$", line), line
    assert done_marker(cfg).endswith("node0-ocp0.done")
    init_logger(cfg, 2, stream=False)            # re-initialising must not fail (race-free mkdir, handler reset)


def test_stats_schema_matches_reference(tmp_path):
    from dynamic_load_balance_distributeddnn_b200.utils import StatsRecorder, load_stats
    cfg = DBSConfig(stats_dir=str(tmp_path / "statis"))
    rec = StatsRecorder(cfg)
    rec.append(epoch=0, train_loss=1.0, train_time=2.0, sync_time=0.1, val_loss=0.9, accuracy=50.0, partition=np.array([.5, .5]),
               node_time=[2.0, 2.1], wallclock_time=2.2, local_batches=[32, 32], samples_per_sec=10.0,
               straggler_wait_ms_per_step=0.3, steps=7, lr=0.01)
    path = rec.save()
    d = load_stats(path)
    for k in ("epoch", "train_loss", "train_time", "sync_time", "val_loss", "accuracy", "partition", "node_time", "wallclock_time"):
        assert k in d and len(d[k]) == 1, k        # reference dbs.py:317-326
    assert os.path.isfile(path.replace(".npy", ".json"))


def test_net_aliases_import_like_the_reference():
    from dynamic_load_balance_distributeddnn_b200 import Net
    assert not os.path.exists(os.path.join(ROOT, "Net")), "a top-level Net package would shadow the reference's in bench.py"
    assert sum(p.numel() for p in Net.Densenet.DenseNet121(10).parameters()) == 6956298
    assert sum(p.numel() for p in Net.Resnet.ResNet50(10).parameters()) == 23520842
    assert Net.MnistNet.MnistNet and Net.GoogleNet.GoogLeNet and Net.RegNet.RegNetY_400MF and Net.Transformer.TransformerModel


def test_run_sh_usage():
    r = subprocess.run(["bash", os.path.join(ROOT, "run.sh")], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage: ./run.sh [WORLD_SIZE] [BATCH_SIZE] [EPOCH_SIZE] [LEARNING_RATE] [GPUSET]" in r.stdout


def test_launcher_backend_selection():
    from dynamic_load_balance_distributeddnn_b200.launch import _backend_for
    assert _backend_for(DBSConfig(debug=True)) == "gloo"
    assert _backend_for(DBSConfig(debug=False, world_size=4, gpu=[0, 1, 2, 3])) == "nccl"
    assert _backend_for(DBSConfig(debug=False, world_size=4, gpu=[0, 0, 0, 1])) == "gloo"      # shared GPU: NCCL refuses duplicates
    assert _backend_for(DBSConfig(debug=False, world_size=2, gpu=0)) == "gloo"
    assert _backend_for(DBSConfig(debug=False, world_size=2, gpu=0), from_env=True) == "nccl"  # torchrun: LOCAL_RANK devices
    assert _backend_for(DBSConfig(debug=False, world_size=1, gpu=0)) == "nccl"


def test_print_layer_and_prepare_data(tmp_path, capsys):
    from dynamic_load_balance_distributeddnn_b200.models import build_model
    from dynamic_load_balance_distributeddnn_b200.utils import print_layer
    m = build_model("mnistnet")
    assert print_layer(m, "fc2.bias") is m.fc2.bias and print_layer(m, "nope") is None


def test_profile_flag_writes_trace_kernel_table_and_phase_log(tmp_path):
    """--profile: NVTX/record_function ranges per phase, a Chrome trace + per-op table of a few steps,
    and the host-side phase table in the rank log."""
    import json
    from dynamic_load_balance_distributeddnn_b200.config import DBSConfig
    from dynamic_load_balance_distributeddnn_b200.engine import Trainer
    from dynamic_load_balance_distributeddnn_b200.utils import init_logger
    cfg = DBSConfig(debug=True, world_size=1, batch_size=16, model="mnistnet", dataset="mnist", synthetic=True,
                    train_samples=16 * 12, test_samples=32, epoch_size=1, validate=True, profile=True,
                    log_dir=str(tmp_path / "logs"), stats_dir=str(tmp_path / "statis"))
    t = Trainer(cfg, 0, 1, "cpu", init_logger(cfg, 0, stream=False))
    t.run()
    t.close()
    stem = os.path.join(cfg.log_dir, cfg.experiment_id(0))
    trace = json.load(open(stem + ".trace.json"))
    names = {e.get("name") for e in trace["traceEvents"]}
    assert {"forward", "backward", "reduce_and_step"} <= names
    assert os.path.getsize(stem + ".kernels.txt") > 0
    log = open(stem + ".log").read()
    assert "host-side phase table" in log and "stage_h2d" in log and "validate" in log
    assert t.tracer.phase_n["forward"] == 12 and t.tracer.phase_n["rebalance"] == 1


def test_graph_nodes_tool_counts_and_critical_path(tmp_path):
    """tools/graph_nodes.py on a cudaGraphDebugDotPrint-style dump: node kinds, kernel histogram, longest chain."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("graph_nodes", os.path.join(ROOT, "tools", "graph_nodes.py"))
    gn = importlib.util.module_from_spec(spec); spec.loader.exec_module(gn)
    dot = tmp_path / "g.dot"
    dot.write_text(r'''digraph dot {
subgraph cluster_1 {
label="graph_1" graph[style="dashed"];
"n0"[style="solid" shape="rectangle" label="0\nMEMSET\nnode handle: 0x1"];
"n1"[style="bold" shape="octagon" label="1\n_Z3fooPf\nnode handle: 0x2"];
"n2"[style="bold" shape="octagon" label="2\n_Z3barPf\nnode handle: 0x3"];
"n3"[style="solid" shape="rectangle" label="3\nEVENT_RECORD\nnode handle: 0x4"];
"n4"[style="bold" shape="octagon" label="4\n_Z3fooPf\nnode handle: 0x5"];
"n0" -> "n1";
"n1" -> "n2";
"n1" -> "n3";
"n2" -> "n4";
}
}
''')
    nodes, edges = gn.parse(str(dot))
    assert len(nodes) == 5 and len(edges) == 4
    kinds = sorted(k for k, _ in nodes.values())
    assert kinds == ["EVENT_RECORD", "KERNEL", "KERNEL", "KERNEL", "MEMSET"]
    roots, leaves, (path_nodes, path_kernels), acyclic = gn.critical_path(nodes, edges)
    assert (roots, leaves, path_nodes, path_kernels, acyclic) == (1, 2, 4, 3, True)
    assert gn.main([str(dot)]) == 0


def test_trace_overlap_tool(tmp_path):
    """tools/trace_overlap.py on a synthetic Kineto trace: per-stream totals, nccl detection, overlap of the collective with
    compute on other streams."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ev = []
    # compute stream 7: two kernels [0, 100) and [150, 250);  comm stream 9: allreduce [50, 200) -> 50 + 50 of 150 us overlapped
    ev.append({"cat": "kernel", "name": "void gemm_tc_kernel<float>(int)", "ts": 0, "dur": 100, "args": {"stream": 7}})
    ev.append({"cat": "kernel", "name": "void gn_fwd_apply_kernel<float>(int)", "ts": 150, "dur": 100, "args": {"stream": 7}})
    ev.append({"cat": "kernel", "name": "void weighted_allreduce_kernel<float, 2, 8>(CommArgs)", "ts": 50, "dur": 150, "args": {"stream": 9}})
    ev.append({"cat": "cpu_op", "name": "aten::add", "ts": 0, "dur": 5})
    p = tmp_path / "t.trace.json"
    p.write_text(json.dumps({"traceEvents": ev}))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "trace_overlap.py"), str(p)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "3 kernel events on 2 streams" in out.stdout
    assert "nccl kernels in trace: 0" in out.stdout
    assert "1 launches" in out.stdout and "66.7% of it concurrent" in out.stdout
    ev.append({"cat": "kernel", "name": "ncclDevKernel_AllReduce_Sum_f32_RING_LL(x)", "ts": 300, "dur": 10, "args": {"stream": 11}})
    p.write_text(json.dumps(ev))                                     # bare-list flavour of the format
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "trace_overlap.py"), str(p)], capture_output=True, text=True)
    assert "nccl kernels in trace: 1" in out.stdout
