"""Launch counters for the hand-written compute kernels (bench.py ``gpu_launches``)."""
_counts = {}


def bump(name: str, n: int = 1) -> None:
    _counts[name] = _counts.get(name, 0) + n


def total() -> int:
    return sum(_counts.values())


def snapshot() -> dict:
    return dict(_counts)
