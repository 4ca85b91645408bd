"""End-to-end: the entry script at the reference's path, zero-arg semantics, printed lines
(reference app/torch_train.py:208-312), plus the [DRIVER] CPU plumbing config
"torch_train.py ResNet-18 synthetic 32x32 on CPU/gloo world_size=1"."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "app", "torch_train.py")


def _run(args, env_extra=None, launcher=None, timeout=300, cwd=None):
    env = dict(os.environ, PYTHONPATH=ROOT, B200DP_OFFLINE="1", B200DP_SYNTH_ROWS="400",
               OMP_NUM_THREADS="2")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(env_extra or {})
    cmd = (launcher or []) + [sys.executable, SCRIPT] + args
    return subprocess.run(cmd, cwd=cwd or ROOT, env=env, capture_output=True, text=True,
                          timeout=timeout)


def test_no_gpu_message_zero_arg(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run([], cwd=str(tmp_path))
    assert r.returncode == 0 and "Needs a GPU to run!" in r.stdout


def test_lstm_cpu_single(tmp_path):
    r = _run(["--device", "cpu", "--epochs", "2"], cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    assert "horovod has distributed to the following devices: [" in out
    assert "this process is using device - cpu" in out
    assert len(re.findall(r"epoch: \d+, train_loss: [\d.e-]+", out)) == 2
    assert len(re.findall(r"epoch: \d+, test_loss: [\d.e-]+", out)) == 2
    assert re.search(r"device: 0, avg_time_per_epoch:[\d.]+", out)
    assert re.search(r"total training time in minutes: [\d.e-]+", out)


def test_lstm_cpu_two_ranks_epoch_scaling(tmp_path):
    launcher = [sys.executable, "-m", "distributed_torch_horovod_gcp_b200.launch", "-np", "2",
                "-H", "localhost:2"]
    r = _run(["--device", "cpu", "--epochs", "3", "--max-steps", "3"], launcher=launcher,
             cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    # ceil(3 / 2) = 2 epochs; only rank 0 prints losses; every rank prints its avg time
    assert len(re.findall(r"\[0\]<stdout>:epoch: \d+, train_loss", r.stdout)) == 2
    assert "[1]<stdout>:epoch" not in r.stdout
    assert "[0]<stdout>:device: 0, avg_time_per_epoch:" in r.stdout
    assert "[1]<stdout>:device: 1, avg_time_per_epoch:" in r.stdout


def test_resnet18_cpu_plumbing_config(tmp_path):
    r = _run(["--model", "resnet18", "--device", "cpu", "--epochs", "1", "--batch-size", "4",
              "--steps-per-epoch", "2"], cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert re.search(r"epoch: 0, train_loss: [\d.]+", r.stdout)
