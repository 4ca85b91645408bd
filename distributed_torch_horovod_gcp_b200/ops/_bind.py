"""ctypes signatures + autograd wrappers for the compute kernels.  ``register`` marks an op
available only if its C symbol exists in the built library."""
from __future__ import annotations

from typing import Dict

from . import gemm as _gemm
from . import bn as _bn
from . import lstm_fused as _lstm
from . import ln as _ln
from . import conv as _conv
from . import lstm_rec as _lstm_rec
from . import attention as _attn
from .attention import attention_fused  # noqa: F401
from .lstm_rec import lstm_recurrent  # noqa: F401
from .conv import conv3x3, conv2d as conv2d_implicit  # noqa: F401
from .ln import layer_norm  # noqa: F401
from .gemm import linear, mlp, qkv_proj  # noqa: F401  (re-exported as kernels.linear / .mlp / .qkv_proj)
from .bn import conv_bn_act, bn_act, max_pool_3x3_s2, global_avg_pool  # noqa: F401


def register(lib, have: Dict[str, bool]) -> None:
    _gemm.register(lib, have)
    _bn.register(lib, have)
    _lstm.register(lib, have)
    _ln.register(lib, have)
    _conv.register(lib, have)
    _lstm_rec.register(lib, have)
    _attn.register(lib, have)


def linear_supported(x, weight) -> bool:
    return _gemm.supported(x, weight)


def layer_norm_supported(x, weight) -> bool:
    return _ln.supported(x, weight)
