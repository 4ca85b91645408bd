"""Time-series data layer of the reference script, rebuilt.

Reference parity (SURVEY.md §2.1 C2–C5):
  * C2 ``read_file_from_aws``    — app/torch_train.py:22-31
  * C3 ``x_cols`` / ``y_cols``   — app/torch_train.py:33-38
  * C4 ``reshape_and_scale_data_for_training`` — app/torch_train.py:40-81
  * C5 ``TimeSeriesDataSet``     — app/torch_train.py:84-103

Behaviour kept on purpose: the scaler is fit over the WHOLE frame before the chronological
80/20 split (test leakage, app/torch_train.py:61-64,77); windows are ``x[i:i+W] ->
y[i+W : i+W+y_len]`` for ``i in range(len - W)``.

Behaviour fixed on purpose (SURVEY.md §7.3 "pure defects"): the download no longer runs at
import time on every rank (N ranks racing on one file, handle never closed); there is no
network here, so ``ensure_dataset`` falls back to a deterministic synthetic frame with the
same 23-column schema.  The Python window loop (app/torch_train.py:72-74) is replaced by a
zero-copy ``sliding_window_view``.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

DATA_URL = "https://medium-post-data.s3.amazonaws.com/data_es.csv"

x_cols = ['close', 'ask', 'bid', 'md_0_ask', 'md_0_bid',
          'md_1_ask', 'md_1_bid', 'md_2_ask', 'md_2_bid', 'md_3_ask', 'md_3_bid',
          'md_4_ask', 'md_4_bid', 'md_5_ask', 'md_5_bid', 'md_6_ask', 'md_6_bid',
          'md_7_ask', 'md_7_bid', 'md_8_ask', 'md_8_bid', 'md_9_ask', 'md_9_bid']

y_cols = ['close']


class MinMaxScaler:
    """Feature-wise min-max scaling to [0, 1] (the reference's default scaler,
    app/torch_train.py:40; sklearn-compatible ``fit/transform/fit_transform/inverse_transform``)."""

    def __init__(self, feature_range=(0.0, 1.0)):
        self.feature_range = feature_range

    def fit(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.data_min_ = X.min(axis=0)
        self.data_max_ = X.max(axis=0)
        rng = self.data_max_ - self.data_min_
        rng[rng == 0.0] = 1.0
        lo, hi = self.feature_range
        self.scale_ = (hi - lo) / rng
        self.min_ = lo - self.data_min_ * self.scale_
        return self

    def transform(self, X):
        return np.asarray(X, dtype=np.float64) * self.scale_ + self.min_

    def fit_transform(self, X):
        return self.fit(X).transform(X)

    def inverse_transform(self, X):
        return (np.asarray(X, dtype=np.float64) - self.min_) / self.scale_


class StandardScaler:
    def fit(self, X):
        X = np.asarray(X, dtype=np.float64)
        self.mean_ = X.mean(axis=0)
        self.scale_ = X.std(axis=0)
        self.scale_[self.scale_ == 0.0] = 1.0
        return self

    def transform(self, X):
        return (np.asarray(X, dtype=np.float64) - self.mean_) / self.scale_

    def fit_transform(self, X):
        return self.fit(X).transform(X)

    def inverse_transform(self, X):
        return np.asarray(X, dtype=np.float64) * self.scale_ + self.mean_


def read_file_from_aws(path: str = "data_es.csv", url: str = DATA_URL, timeout: float = 10.0) -> bool:
    """Stream the ES-futures CSV to ``path``.  Returns True on success.  Writes to a
    temporary name then renames atomically, so concurrent ranks cannot observe a partial
    file (the reference's import-time race)."""
    try:
        import requests
        response = requests.get(url, stream=True, timeout=timeout)
        response.raise_for_status()
        tmp = f"{path}.{os.getpid()}.part"
        with open(tmp, "wb") as f:
            for chunk in response.iter_content(chunk_size=1 << 16):
                f.write(chunk)
        os.replace(tmp, path)
        return True
    except Exception:
        return False


def synthetic_market_frame(n_rows: int = 20000, seed: int = 0):
    """Deterministic stand-in for ``data_es.csv``: a geometric random-walk close price with a
    10-level bid/ask ladder around it — same 23-feature schema as the reference."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    close = 3000.0 * np.exp(np.cumsum(rng.normal(0.0, 2e-4, n_rows)))
    tick = 0.25
    spread = tick * (1 + rng.integers(0, 2, n_rows))
    cols = {"close": close, "ask": close + spread / 2, "bid": close - spread / 2}
    for lvl in range(10):
        cols[f"md_{lvl}_ask"] = close + spread / 2 + lvl * tick + rng.normal(0, 0.01, n_rows)
        cols[f"md_{lvl}_bid"] = close - spread / 2 - lvl * tick + rng.normal(0, 0.01, n_rows)
    df = pd.DataFrame(cols)
    return df[x_cols]


def ensure_dataset(path: str = "data_es.csv", rank: int = 0, n_rows: Optional[int] = None,
                   allow_download: bool = True, decide=None):
    """Return ``(frame, source)``: the CSV at ``path`` if present, else the downloaded CSV, else the
    synthetic frame.  Never runs at import time (the reference downloads at import on every rank,
    app/torch_train.py:22-31).

    Every rank must train on the SAME data (same length => same number of batches => matching
    collectives), so the source is decided ONCE, by rank 0, and shared: ``decide`` is a callable
    ``decide(value_or_None) -> value`` that broadcasts rank 0's decision (``hvd.broadcast_object`` in
    ``app/torch_train.py``); without it (single process) the decision is local.  Rank 0 performs the
    download; the other ranks read the file only after the decision says it exists."""
    import pandas as pd
    source = None
    if rank == 0:
        if os.path.exists(path):
            source = "file"
        elif allow_download and os.environ.get("B200DP_OFFLINE", "0") != "1" and \
                read_file_from_aws(path, timeout=float(os.environ.get("B200DP_DOWNLOAD_TIMEOUT", "3"))):
            source = "download"
        else:
            source = "synthetic"
    if decide is not None:
        source = decide(source)          # blocks until rank 0 has finished (download included)
    elif source is None:                 # rank > 0 without a broadcast: fall back to what is on disk
        source = "file" if os.path.exists(path) else "synthetic"
    if source in ("file", "download"):
        return pd.read_csv(path), source
    rows = n_rows or int(os.environ.get("B200DP_SYNTH_ROWS", "20000"))
    return synthetic_market_frame(rows), "synthetic"


def reshape_and_scale_data_for_training(data, window_size: int, x_cols: Sequence[str],
                                        y_cols: Sequence[str], y_len: int = 1, scale: bool = True,
                                        scaler=MinMaxScaler, test_size: float = 0.2,
                                        backend: str = 'keras'):
    """Given a ``pandas.DataFrame`` and a window size, reshape and scale it for training.

    Returns ``(x_train, x_test, y_train, y_test, fitted_scaler)``; numpy arrays for
    ``backend='keras'``, ``torch.Tensor`` (fp32) for ``backend='torch'``.
    Shapes: x ``[n, window, len(x_cols)]``, y ``[n, y_len, len(y_cols)]``.
    """
    columns = list(data.columns)
    x_idx = [columns.index(c) for c in x_cols]
    y_idx = [columns.index(c) for c in y_cols]
    subset = sorted(set(x_idx + y_idx))          # de-duplicated, sorted column subset
    pos = {orig: k for k, orig in enumerate(subset)}
    arr = np.asarray(data.iloc[:, subset], dtype=np.float64)
    s = None
    if scale:
        s = scaler()
        arr = np.asarray(s.fit_transform(arr))
    xs = arr[:, [pos[i] for i in x_idx]]
    ys = arr[:, [pos[i] for i in y_idx]]

    n = len(arr) - window_size
    if n <= 0:
        raise ValueError(f"need more than window_size={window_size} rows, got {len(arr)}")
    from numpy.lib.stride_tricks import sliding_window_view
    xw = sliding_window_view(xs, window_size, axis=0)[:n]          # [n, F, W]
    xw = np.ascontiguousarray(np.transpose(xw, (0, 2, 1)))         # [n, W, F]
    # y window i = ys[i+W : i+W+y_len]; the tail windows may be short in the reference (ragged
    # python lists); with y_len=1 (the only value the reference uses) every window is full.
    n_full = len(arr) - window_size - y_len + 1
    if n_full < n:
        n = n_full
        xw = xw[:n]
    yw = sliding_window_view(ys[window_size:], y_len, axis=0)[:n]   # [n, Fy, y_len]
    yw = np.ascontiguousarray(np.transpose(yw, (0, 2, 1)))         # [n, y_len, Fy]

    # chronological split, shuffle=False; sklearn semantics: n_test = ceil(test_size * n)
    n_test = int(np.ceil(test_size * n)) if isinstance(test_size, float) else int(test_size)
    n_train = n - n_test
    x_train, x_test, y_train, y_test = xw[:n_train], xw[n_train:], yw[:n_train], yw[n_train:]
    if backend == 'keras':
        return x_train, x_test, y_train, y_test, s
    if backend == 'torch':
        f = lambda a: torch.from_numpy(np.array(a, dtype=np.float32, order='C'))
        return f(x_train), f(x_test), f(y_train), f(y_test), s
    raise ValueError(f"unknown backend {backend!r} (expected 'keras' or 'torch')")


class TimeSeriesDataSet(Dataset):
    """Map-style dataset over pre-materialised tensors (app/torch_train.py:84-103)."""

    def __init__(self, X, Y):
        self.X = X
        self.Y = Y
        if len(self.X) != len(self.Y):
            raise Exception("The length of X does not match the length of Y")

    def __len__(self):
        return len(self.X)

    def __getitem__(self, index):
        return self.X[index], self.Y[index]
