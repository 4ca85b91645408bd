"""``hvd.SyncBatchNorm`` — batch norm with statistics reduced over all ranks.

Horovod surface parity ("not used by the reference").  Statistics
(sum, sum of squares, count) are packed into one small vector and all-reduced with the
framework's own ``allreduce`` (one-shot sm_100a kernel on CUDA: 2C+1 floats is a pure
latency message), forward and backward.
"""
from __future__ import annotations

import torch
from torch.nn.modules.batchnorm import _BatchNorm

from .. import _state
from . import mpi_ops


class _SyncBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum):
        C = x.shape[1]
        dims = [0] + list(range(2, x.dim()))
        xf = x.float()
        n_local = x.numel() // C
        stats = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
        stats[:C] = xf.sum(dims)
        stats[C:2 * C] = (xf * xf).sum(dims)
        stats[2 * C] = float(n_local)
        stats = mpi_ops.allreduce(stats, op=mpi_ops.Sum, name=None)
        n = stats[2 * C]
        mean = stats[:C] / n
        var = (stats[C:2 * C] / n - mean * mean).clamp_min_(0.0)
        invstd = torch.rsqrt(var + eps)
        if running_mean is not None:
            with torch.no_grad():
                unbiased = var * (n / (n - 1).clamp_min(1.0))
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype) * momentum)
                running_var.mul_(1 - momentum).add_(unbiased.to(running_var.dtype) * momentum)
        shape = [1, C] + [1] * (x.dim() - 2)
        xhat = (xf - mean.view(shape)) * invstd.view(shape)
        y = xhat
        if weight is not None:
            y = y * weight.float().view(shape)
        if bias is not None:
            y = y + bias.float().view(shape)
        ctx.save_for_backward(xhat, invstd, weight)
        ctx.n = n
        ctx.has_bias = bias is not None
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xhat, invstd, weight = ctx.saved_tensors
        C = xhat.shape[1]
        dims = [0] + list(range(2, xhat.dim()))
        shape = [1, C] + [1] * (xhat.dim() - 2)
        dyf = dy.float()
        sum_dy = dyf.sum(dims)
        sum_dy_xhat = (dyf * xhat).sum(dims)
        dweight = sum_dy_xhat.to(weight.dtype) if weight is not None else None
        dbias = sum_dy.to(weight.dtype if weight is not None else dy.dtype) \
            if ctx.has_bias else None
        packed = torch.cat([sum_dy, sum_dy_xhat])
        packed = mpi_ops.allreduce(packed, op=mpi_ops.Sum, name=None)
        g_sum_dy, g_sum_dy_xhat = packed[:C], packed[C:]
        w = weight.float().view(shape) if weight is not None else 1.0
        dx = (dyf - (g_sum_dy / ctx.n).view(shape)
              - xhat * (g_sum_dy_xhat / ctx.n).view(shape)) * invstd.view(shape) * w
        return dx.to(dy.dtype), dweight, dbias, None, None, None, None


class SyncBatchNorm(_BatchNorm):
    """Applies synchronous BatchNorm: statistics are computed over the global batch."""

    def _check_input_dim(self, input):
        if input.dim() < 2:
            raise ValueError("expected at least 2D input (got {}D input)".format(input.dim()))

    def forward(self, input):
        self._check_input_dim(input)
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        momentum = self.momentum
        if momentum is None:
            momentum = 1.0 / float(self.num_batches_tracked) if self.track_running_stats else 0.0
        use_batch = self.training or not self.track_running_stats
        if not use_batch or not _state.is_initialized() or _state.size() == 1:
            return torch.nn.functional.batch_norm(
                input, self.running_mean, self.running_var, self.weight, self.bias,
                use_batch, momentum if momentum is not None else 0.0, self.eps)
        return _SyncBNFn.apply(input, self.weight, self.bias, self.running_mean,
                               self.running_var, self.eps, momentum)
