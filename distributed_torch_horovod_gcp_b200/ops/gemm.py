"""tcgen05 GEMM binding (csrc/gemm_sm100.cu) and the autograd ``linear`` built on it.

    y = act(x @ W^T + b) (+ residual)          forward : A = x (K-major),  B = W (K-major)
    dx = dy' @ W                                dgrad   : A = dy' (K-major), B = W (MN-major)
    dW = dy'^T @ x                              wgrad   : A = dy' (MN-major), B = x (MN-major),
                                                          split-K, fp32 atomics
No transposes are materialised: the kernel takes MN-major operands through its UMMA
descriptors.  ``dy' = dy * act'(z)`` is fused into the dgrad of the *next* layer's epilogue
where possible; here it is computed by the epilogue modes 3/4 of the kernel when the layer
has an activation.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import counters

_lib = None
ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}


def register(lib, have):
    global _lib
    if not hasattr(lib, "b200dp_gemm_bf16"):
        return
    _lib = lib
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.b200dp_gemm_bf16.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, i, vp, vp, vp, vp, i, i, f, i,
                                     i, i, ctypes.c_uint64]
    lib.b200dp_gemm_bf16.restype = i
    lib.b200dp_gemm_last_error.restype = ctypes.c_char_p
    have["gemm"] = True
    have["linear"] = True


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, M: int, N: int, K: int, *,
         a_mn: bool = False, b_mn: bool = False, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, preact: Optional[torch.Tensor] = None,
         act: int = 0, out_mode: int = 0, alpha: float = 1.0, splits: int = 1, block_n: int = 0,
         max_ctas: int = 0) -> torch.Tensor:
    """Raw kernel call.  ``a``: [M,K] (K-major) or [K,M] (MN-major) bf16 with contiguous rows;
    ``b``: [N,K] or [K,N]; ``out``: [M,N] bf16 (out_mode 0) or fp32 (1: atomic add, 2: store)."""
    assert _lib is not None, "libb200dp_kernels.so not loaded"
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1 and out.stride(-1) == 1
    bias_bf = bias.data_ptr() if bias is not None and bias.dtype == torch.bfloat16 else None
    bias_f32 = bias.data_ptr() if bias is not None and bias.dtype == torch.float32 else None
    rc = _lib.b200dp_gemm_bf16(
        a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
        int(a_mn), int(b_mn), bias_bf, bias_f32,
        residual.data_ptr() if residual is not None else None,
        preact.data_ptr() if preact is not None else None,
        act, out_mode, float(alpha), splits, block_n, max_ctas,
        torch.cuda.current_stream(a.device).cuda_stream)
    if rc != 0:
        raise RuntimeError("b200dp_gemm_bf16: " + (_lib.b200dp_gemm_last_error() or b"").decode())
    counters.bump("gemm_sm100")
    return out


def supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    if _lib is None or x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        return False
    N, K = weight.shape
    return N % 8 == 0 and K % 8 == 0 and x.shape[-1] == K and x.numel() // K >= 1


def _splits_for(M_out: int, N_out: int, K_red: int, sms: int = 148) -> int:
    tiles = ((M_out + 127) // 128) * ((N_out + 127) // 128)
    kb = (K_red + 63) // 64
    want = max(1, (2 * sms) // max(tiles, 1))
    return max(1, min(want, kb // 4 if kb >= 8 else 1))


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, residual):
        K = weight.shape[1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1 or (x2.stride(0) % 8) or (x2.data_ptr() % 16):
            x2 = x2.contiguous()
        M = x2.shape[0]
        y = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
        need_z = act != 0 and (x.requires_grad or weight.requires_grad)
        z = torch.empty_like(y) if need_z else None
        r2 = None
        if residual is not None:
            r2 = residual.reshape(-1, N)
            if not r2.is_contiguous():
                r2 = r2.contiguous()
        gemm(x2, weight, y, M, N, K, bias=bias, residual=r2, preact=z, act=act)
        ctx.save_for_backward(x2, weight, z)
        ctx.act, ctx.has_bias, ctx.has_res = act, bias is not None, residual is not None
        ctx.x_shape = x.shape
        ctx.bias_dtype = bias.dtype if bias is not None else None
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, z = ctx.saved_tensors
        N, K = weight.shape
        M = x2.shape[0]
        dy2 = dy.reshape(M, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dres = dy2.view(*ctx.x_shape[:-1], N) if ctx.has_res else None
        if ctx.act != 0:
            # dz = dy * act'(z): elementwise epilogue of an identity-free pass is not worth a GEMM;
            # use the fused torch op (memory-bound) — fused variants live in the dgrad epilogue.
            if ctx.act == 1:
                dz = dy2 * (z > 0).to(dy2.dtype)
            else:
                zf = z.float()
                cdf = 0.5 * (1.0 + torch.erf(zf * 0.7071067811865476))
                pdf = 0.3989422804014327 * torch.exp(-0.5 * zf * zf)
                dz = (dy2.float() * (cdf + zf * pdf)).to(torch.bfloat16)
        else:
            dz = dy2
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.bfloat16, device=dy.device)
            gemm(dz, weight, dx, M, K, N, b_mn=True)                       # dx = dz @ W
            dx = dx.view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            acc = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
            gemm(dz, x2, acc, N, K, M, a_mn=True, b_mn=True, out_mode=1,
                 splits=_splits_for(N, K, M))                              # dW = dz^T @ x
            dw = acc.to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dz.float().sum(0).to(ctx.bias_dtype)
        return dx, dw, db, None, dres


def linear(x, weight, bias=None, act: Optional[str] = None, residual=None):
    return _LinearFn.apply(x, weight, bias, ACT[act], residual)
