"""C6: reference LSTM model parity; model zoo shapes / parameter counts."""
import pytest
import torch

from distributed_torch_horovod_gcp_b200.models import LSTM, resnet18, resnet50, resnet152, vit_b_16, vit_tiny, build


def test_lstm_shapes_and_message_table():
    m = LSTM(n_features=23, window_size=10, output_size=1, h_size=256)
    out = m(torch.randn(32, 10, 23))
    assert out.shape == (32, 1, 1)
    sd = m.state_dict()
    assert list(sd.keys()) == [
        "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
        "linear.weight", "linear.bias", "linear2.weight", "linear2.bias",
        "linear3.weight", "linear3.bias"]
    assert sum(v.numel() for v in sd.values()) == 370049
    assert sum(v.numel() * v.element_size() for v in sd.values()) == 1480196
    assert sd["lstm.weight_hh_l0"].shape == (1024, 256) and sd["linear3.bias"].shape == (1,)


def test_lstm_random_hidden_each_forward_and_backward():
    torch.manual_seed(0)
    m = LSTM(23, 10, 1, 256)
    x = torch.randn(4, 10, 23)
    a, b = m(x), m(x)
    assert not torch.equal(a, b)          # fresh random (h0, c0) every call (reference quirk)
    a.sum().backward()
    assert all(p.grad is not None for p in m.parameters())


def test_lstm_matches_manual_composition():
    m = LSTM(23, 10, 1, 256)
    x = torch.randn(3, 10, 23)
    torch.manual_seed(1)
    out = m(x)
    torch.manual_seed(1)
    h = m.init_hidden(3)
    y, _ = m.lstm(x, h)
    ref = m.linear3(m.linear2(m.linear(y[:, 9:10, :])))
    torch.testing.assert_close(out, ref)


def test_lstm_multilayer_bidirectional():
    m = LSTM(23, 10, 1, 32, n_layers=2, bidirectional=True)
    assert m(torch.randn(5, 10, 23)).shape == (5, 1, 1)


def test_lstm_initializers():
    with pytest.warns(UserWarning, match="only one initializer"):
        m = LSTM(23, 10, 1, 16, initializers=[torch.nn.init.zeros_])
    assert float(m.linear.weight.detach().abs().sum()) == 0.0
    with pytest.raises(Exception, match="initializers were provided"):
        LSTM(23, 10, 1, 16, initializers=[torch.nn.init.zeros_] * 2)
    LSTM(23, 10, 1, 16, initializers=[torch.nn.init.xavier_uniform_] * 4)


def test_zoo_param_counts():
    def count(m):
        ps = list(m.parameters())
        return sum(p.numel() for p in ps), len(ps)
    assert count(resnet18()) == (11689512, 62)
    assert count(resnet50()) == (25557032, 161)
    assert count(resnet152()) == (60192808, 467)
    assert count(vit_b_16()) == (86567656, 152)
    r = resnet50()
    nbuf = sum(b.numel() for b in r.buffers())
    assert nbuf == 53173          # BN running stats + counters travel in broadcast_parameters
    assert "layer1.0.downsample.1.running_var" in r.state_dict()


def test_zoo_forward_backward_small():
    r = build("resnet18", small_input=True, num_classes=10)
    y = r(torch.randn(2, 3, 32, 32))
    assert y.shape == (2, 10)
    y.sum().backward()
    v = vit_tiny()
    z = v(torch.randn(2, 3, 32, 32))
    assert z.shape == (2, 10)
    z.sum().backward()
    with pytest.raises(ValueError):
        build("alexnet")
