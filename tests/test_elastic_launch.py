"""Elastic launcher (restart-based) with a fake host-discovery script, launcher flag wiring, the
device-watchdog knobs and the single-decision dataset source (all CPU / Gloo)."""
import os
import stat
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCH = [sys.executable, "-m", "distributed_torch_horovod_gcp_b200.launch"]


def _write(path, text, exe=False):
    with open(path, "w") as f:
        f.write(textwrap.dedent(text))
    if exe:
        os.chmod(path, os.stat(path).st_mode | stat.S_IXUSR)


def test_elastic_resize_through_relaunch(tmp_path):
    """2 ranks -> the discovery script starts reporting 3 slots -> workers persist the commit and exit
    75 -> the launcher relaunches 3 ranks that resume from the persisted epoch."""
    marker = tmp_path / "grow"
    disc = tmp_path / "discover.sh"
    _write(disc, f"""\
        #!/bin/sh
        if [ -f {marker} ]; then echo localhost:3; else echo localhost:2; fi
        """, exe=True)
    worker = tmp_path / "worker.py"
    _write(worker, f"""\
        import os, sys, time
        sys.path.insert(0, {ROOT!r})
        os.environ["B200DP_FORCE_CPU"] = "1"
        os.environ["B200DP_DISCOVERY_INTERVAL_S"] = "0"
        import torch
        import distributed_torch_horovod_gcp_b200.torch as hvd
        hvd.init()
        model = torch.nn.Linear(4, 2)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        state = hvd.elastic.TorchState(model, opt, epoch=0)

        @hvd.elastic.run
        def train(state):
            while state.epoch < 6:
                print(f"EPOCH {{state.epoch}} size {{hvd.size()}} w0 {{float(model.weight.sum()):.6f}}", flush=True)
                with torch.no_grad():
                    model.weight.add_(1.0)
                state.epoch += 1
                state.commit()
                if state.epoch == 3 and hvd.rank() == 0 and hvd.size() == 2:
                    open({str(marker)!r}, "w").close()
                state.check_host_updates()
            return state.epoch

        print("DONE", train(state), "size", hvd.size(), flush=True)
        hvd.shutdown()
        """)
    env = dict(os.environ, PYTHONPATH=ROOT, B200DP_ELASTIC_STATE_DIR=str(tmp_path / "state"))
    os.makedirs(env["B200DP_ELASTIC_STATE_DIR"], exist_ok=True)
    r = subprocess.run(LAUNCH + ["--min-np", "2", "--max-np", "4", "--host-discovery-script", str(disc),
                                 sys.executable, str(worker)], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=300)
    out = r.stdout
    assert r.returncode == 0, (out[-2000:], r.stderr[-2000:])
    assert "EPOCH 0 size 2" in out and "EPOCH 2 size 2" in out
    assert "EPOCH 3 size 3" in out                      # resumed from the persisted commit on 3 ranks
    assert "EPOCH 3 size 2" not in out
    assert out.count("DONE 6 size 3") == 3
    # the weights carried over: epoch 3 starts from w0(initial) + 3 * 8 elements
    e0 = [l for l in out.splitlines() if "EPOCH 0 size 2" in l][0]
    e3 = [l for l in out.splitlines() if "EPOCH 3 size 3" in l][0]
    w0, w3 = float(e0.rsplit(" ", 1)[1]), float(e3.rsplit(" ", 1)[1])
    assert abs((w3 - w0) - 24.0) < 1e-3
    assert "hosts changed" in r.stderr


def test_elastic_needs_discovery_script():
    r = subprocess.run(LAUNCH + ["-np", "2", "--min-np", "1", sys.executable, "-c", "pass"], cwd=ROOT,
                       capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "--host-discovery-script" in r.stderr


def test_ignored_horovod_flags_are_reported():
    r = subprocess.run(LAUNCH + ["-np", "1", "--cycle-time-ms", "3", "--cache-capacity", "64",
                                 "--hierarchical-allreduce", sys.executable, "-c", "print('ok')"],
                       cwd=ROOT, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ok" in r.stdout
    for flag in ("--cycle-time-ms", "--cache-capacity", "--hierarchical-allreduce"):
        assert f"ignored: {flag}" in r.stderr


def test_stall_check_flags_reach_the_watchdog(monkeypatch):
    sys.path.insert(0, ROOT)
    from distributed_torch_horovod_gcp_b200.runtime.symm import watchdog_seconds
    for k in ("B200DP_KERNEL_TIMEOUT_S", "HOROVOD_STALL_CHECK_DISABLE", "HOROVOD_STALL_SHUTDOWN_TIME_SECONDS",
              "HOROVOD_STALL_CHECK_TIME_SECONDS"):
        monkeypatch.delenv(k, raising=False)
    assert watchdog_seconds() == 300.0
    monkeypatch.setenv("HOROVOD_STALL_CHECK_TIME_SECONDS", "45")
    assert watchdog_seconds() == 45.0
    monkeypatch.setenv("HOROVOD_STALL_SHUTDOWN_TIME_SECONDS", "90")
    assert watchdog_seconds() == 90.0
    monkeypatch.setenv("HOROVOD_STALL_CHECK_DISABLE", "1")
    assert watchdog_seconds() > 24 * 3600
    monkeypatch.setenv("B200DP_KERNEL_TIMEOUT_S", "3")
    assert watchdog_seconds() == 3.0
    # and the launcher exports them
    r = subprocess.run(LAUNCH + ["-np", "1", "--stall-check-shutdown-time-seconds", "77", "--log-level", "DEBUG",
                                 sys.executable, "-c",
                                 "import os;print(os.environ['HOROVOD_STALL_SHUTDOWN_TIME_SECONDS'], "
                                 "os.environ['HOROVOD_LOG_LEVEL'])"],
                       cwd=ROOT, capture_output=True, text=True, timeout=60)
    assert r.stdout.split() == ["77", "DEBUG"]


def test_dataset_source_is_decided_once(tmp_path, monkeypatch):
    """Rank > 0 must follow rank 0's decision even when the file appears late."""
    sys.path.insert(0, ROOT)
    monkeypatch.setenv("B200DP_OFFLINE", "1")
    monkeypatch.setenv("B200DP_SYNTH_ROWS", "300")
    from distributed_torch_horovod_gcp_b200.data import ensure_dataset
    missing = str(tmp_path / "none.csv")
    df0, s0 = ensure_dataset(missing, rank=0, decide=lambda v: v)
    df1, s1 = ensure_dataset(missing, rank=1, decide=lambda v: "synthetic")
    assert s0 == s1 == "synthetic" and len(df0) == len(df1) == 300
    # rank 1 is told the file exists -> reads the file, never the synthetic frame
    p = tmp_path / "data.csv"
    df0.to_csv(p, index=False)
    df2, s2 = ensure_dataset(str(p), rank=1, decide=lambda v: "file")
    assert s2 == "file" and len(df2) == 300
