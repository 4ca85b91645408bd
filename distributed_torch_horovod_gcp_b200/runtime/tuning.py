"""Algorithm-selection table for the comm kernels — the replacement for Horovod's autotuner
(`HOROVOD_AUTOTUNE` / parameter_manager: Bayesian search over fusion threshold and cycle time;
SURVEY.md §2.2 N11, §5.6).  There is nothing to search at run time here: the bucket plan is
static and the only choice is WHICH kernel reduces a bucket of a given size on a given world
size.  The default table below was measured on 8xB200 and 2xB200 (profiles/allreduce_sweep_*.json);
`benchmarks/allreduce_sweep.py --write-tuning FILE` re-derives it on another machine and
`B200DP_TUNING_FILE=FILE` makes every rank load it.

Table format (JSON): {"<world>": {"oneshot_max_bytes": int, "prefer": "nvls" | "twoshot"}, ...}
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

ONESHOT, TWOSHOT, NVLS = 0, 1, 2

DEFAULT_TABLE: Dict[str, dict] = {
    "2": {"oneshot_max_bytes": 8 << 10, "prefer": "twoshot"},
    "default": {"oneshot_max_bytes": 8 << 10, "prefer": "nvls"},
}

_loaded: Optional[Dict[str, dict]] = None


def table() -> Dict[str, dict]:
    global _loaded
    if _loaded is None:
        _loaded = dict(DEFAULT_TABLE)
        path = os.environ.get("B200DP_TUNING_FILE")
        if path and os.path.exists(path):
            with open(path) as f:
                user = json.load(f)
            for k, v in user.items():
                if isinstance(v, dict) and "oneshot_max_bytes" in v:
                    _loaded[str(k)] = {"oneshot_max_bytes": int(v["oneshot_max_bytes"]),
                                       "prefer": str(v.get("prefer", "nvls"))}
    return _loaded


def reset() -> None:
    global _loaded
    _loaded = None


def choose(world: int, nbytes: int, multicast: bool) -> int:
    """Pure function: algorithm for a bucket of ``nbytes`` on ``world`` ranks."""
    t = table()
    row = t.get(str(world), t["default"])
    if nbytes <= row["oneshot_max_bytes"]:
        return ONESHOT
    if row["prefer"] == "nvls" and multicast and world > 2:
        return NVLS
    return TWOSHOT


def derive_from_sweep(rows, world: int, multicast: bool) -> dict:
    """Turn allreduce_sweep rows ({bytes, oneshot_us, twoshot_us, nvls_us}) into a table row: the
    largest size at which one-shot is still the fastest kernel, and the better sliced kernel at
    the largest measured size."""
    oneshot_max = 0
    for r in sorted(rows, key=lambda r: r["bytes"]):
        cands = {k: r[k + "_us"] for k in ("oneshot", "twoshot", "nvls") if k + "_us" in r}
        if not cands:
            continue
        if min(cands, key=cands.get) == "oneshot":
            oneshot_max = r["bytes"]
        else:
            break
    last = max(rows, key=lambda r: r["bytes"])
    prefer = "twoshot"
    if multicast and "nvls_us" in last and last["nvls_us"] <= last.get("twoshot_us", float("inf")):
        prefer = "nvls"
    return {"oneshot_max_bytes": int(oneshot_max), "prefer": prefer}
