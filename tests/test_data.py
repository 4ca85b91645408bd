"""C3-C5: schema, windowing/scaling golden shapes, dataset (reference app/torch_train.py:33-103)."""
import numpy as np
import pandas as pd
import pytest
import torch

from distributed_torch_horovod_gcp_b200.data import (
    x_cols, y_cols, reshape_and_scale_data_for_training, TimeSeriesDataSet, MinMaxScaler,
    StandardScaler, synthetic_market_frame, ensure_dataset, DeviceBatchLoader)


def test_schema():
    assert len(x_cols) == 23 and y_cols == ["close"]
    assert x_cols[:3] == ["close", "ask", "bid"] and x_cols[-1] == "md_9_bid"


def _naive(df, W, y_len=1, test_size=0.2):
    """Straight transcription of the reference algorithm's semantics (python loop)."""
    cols = list(df.columns)
    xi = [cols.index(c) for c in x_cols]
    yi = [cols.index(c) for c in y_cols]
    sub = sorted(set(xi + yi))
    a = df.iloc[:, sub].to_numpy(dtype=np.float64)
    mn, mx = a.min(0), a.max(0)
    a = (a - mn) / np.where(mx - mn == 0, 1, mx - mn)
    xs = a[:, [sub.index(i) for i in xi]]
    ys = a[:, [sub.index(i) for i in yi]]
    X = [xs[i:i + W] for i in range(len(a) - W)]
    Y = [ys[i + W:i + W + y_len] for i in range(len(a) - W)]
    n_test = int(np.ceil(test_size * len(X)))
    return np.array(X[:len(X) - n_test]), np.array(X[len(X) - n_test:]), \
        np.array(Y[:len(X) - n_test]), np.array(Y[len(X) - n_test:])


def test_windowing_matches_reference_semantics():
    df = synthetic_market_frame(257, seed=3)
    xt, xv, yt, yv, s = reshape_and_scale_data_for_training(df, 10, x_cols, y_cols, backend="keras")
    rxt, rxv, ryt, ryv = _naive(df, 10)
    assert xt.shape == (197, 10, 23) and xv.shape == (50, 10, 23)
    assert yt.shape == (197, 1, 1) and yv.shape == (50, 1, 1)
    np.testing.assert_allclose(xt, rxt, rtol=0, atol=1e-12)
    np.testing.assert_allclose(xv, rxv, rtol=0, atol=1e-12)
    np.testing.assert_allclose(yt, ryt, rtol=0, atol=1e-12)
    np.testing.assert_allclose(yv, ryv, rtol=0, atol=1e-12)
    assert isinstance(s, MinMaxScaler)


def test_torch_backend_and_no_scale():
    df = synthetic_market_frame(100)
    xt, xv, yt, yv, s = reshape_and_scale_data_for_training(
        df, 10, x_cols, y_cols, scale=False, backend="torch")
    assert s is None and xt.dtype == torch.float32 and xt.shape[1:] == (10, 23)
    assert float(xt[0, 0, 0]) == pytest.approx(float(df["close"].iloc[0]), rel=1e-6)
    with pytest.raises(ValueError):
        reshape_and_scale_data_for_training(df, 10, x_cols, y_cols, backend="jax")


def test_scalers_roundtrip():
    a = np.random.default_rng(0).normal(size=(50, 4)) * 7 + 3
    for S in (MinMaxScaler, StandardScaler):
        s = S()
        t = s.fit_transform(a)
        np.testing.assert_allclose(s.inverse_transform(t), a, atol=1e-9)
    t = MinMaxScaler().fit_transform(a)
    assert t.min() == pytest.approx(0) and t.max() == pytest.approx(1)


def test_dataset():
    ds = TimeSeriesDataSet(torch.zeros(5, 10, 23), torch.ones(5, 1, 1))
    assert len(ds) == 5 and ds[2][0].shape == (10, 23) and ds[2][1].item() == 1
    with pytest.raises(Exception, match="does not match"):
        TimeSeriesDataSet(torch.zeros(5, 1), torch.zeros(4, 1))


def test_ensure_dataset_synthetic(tmp_path):
    df, src = ensure_dataset(str(tmp_path / "nope.csv"), rank=0, n_rows=64, allow_download=False)
    assert src == "synthetic" and list(df.columns) == x_cols and len(df) == 64
    df.to_csv(tmp_path / "d.csv", index=False)
    df2, src2 = ensure_dataset(str(tmp_path / "d.csv"))
    assert src2 == "file" and len(df2) == 64


def test_device_loader_matches_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    X = torch.arange(103).float().view(-1, 1)
    Y = X.clone()
    for world in (1, 2, 4):
        for rank in range(world):
            samp = DistributedSampler(TimeSeriesDataSet(X, Y), num_replicas=world, rank=rank)
            want = list(iter(samp))
            dl = DeviceBatchLoader(X, Y, 32, num_replicas=world, rank=rank)
            got = torch.cat([b[0] for b in dl]).view(-1).long().tolist()
            assert got == want
            assert len(dl) == -(-len(want) // 32)
