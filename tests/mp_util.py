"""Spawn N local worker processes over Gloo (the no-cluster multi-process test tier,
SURVEY.md §4) and collect per-rank results / exceptions."""
import os
import socket
import sys
import traceback

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, modname, fname, args, q, cuda=False):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "LOCAL_WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port), "OMP_NUM_THREADS": "1"})
    if cuda:
        os.environ.pop("B200DP_FORCE_CPU", None)
        os.environ.setdefault("B200DP_KERNEL_TIMEOUT_S", "10")   # inherited from the test if it set one
    else:
        os.environ["B200DP_FORCE_CPU"] = "1"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    torch.set_num_threads(1)
    if cuda:
        torch.cuda.set_device(rank)
    try:
        import importlib
        import distributed_torch_horovod_gcp_b200.torch as hvd
        hvd.init()
        fn = getattr(importlib.import_module(modname), fname)
        res = fn(hvd, *args)
        hvd.shutdown()
        q.put((rank, "ok", res))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "err", f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))


def run_workers(world, modname, fname, args=(), timeout=180, cuda=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, modname, fname, args, q, cuda))
             for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    try:
        for _ in range(world):
            rank, status, payload = q.get(timeout=timeout)
            if status == "err":
                raise AssertionError(f"rank {rank} failed:\n{payload}")
            results[rank] = payload
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
    return [results[r] for r in range(world)]
