"""Training path of the conditioner's dense layers on the tensor cores.

`linear(x, weight, bias, mask)` is F.linear(x, weight * mask, bias) (flowcon/nn/nets/resnet.py:26-28,
flowcon/transforms/made.py:71-72) as an autograd function whose forward and input-gradient GEMMs run on the
3xTF32 tcgen05 kernel (`fc_linear_apply`, fp32-faithful):

    forward   y  = x  @ (W*mask)^T + b          tensor cores
    backward  gx = gy @ (W*mask)                tensor cores (the packed operand is the transposed weight)
              gW = (gy^T @ x) * mask            tensor cores, split-K: the reduction runs over the batch, so it is cut
              gb = gy.sum(0)                    into ranges that run as independent work units (fc_linear_splitk_apply)

The weights change every optimizer step, so they are re-packed on every call (two tiny kernels per layer).
Anything the kernel does not cover (CPU tensors, other dtypes, misaligned operands) takes F.linear.
"""
import torch
from torch.nn import functional as F

from .. import linear as fl

ENABLED = True
# Weight gradients through fc_linear_splitk_apply: grad_y and x are transposed by fc_linear_transpose /
# fc_linear_pack_transposed (the batch has to be the contiguous reduction axis; x is split into its hi / lo planes on the
# way), then one split-K product.  Measured over 262144 rows: 256 x 256: 0.60 ms (0.135 + 0.19 + 0.28) vs 0.77 ms for the
# cuBLAS fp32 torch.mm; 752 x 256: 1.40 vs 2.03 ms.
WGRAD_TC = True
WGRAD_T = True  # grad_y enters the weight-gradient product untransposed (fc_linear_splitk_t_apply)
# Shortest reduction the tensor-core GEMM is used for.  Every 3xTF32 product carries ~2^-22 relative error (dropped
# lo*lo term, truncation inside the MMA); an fp32 FMA chain rounds at 2^-24 per step, so its error grows with the
# chain length.  Measured (scripts/check_linear.py, Gaussian operands): at K = 32 the tensor-core result is 1.3x
# noisier than cuBLAS fp32, at K = 64 equal, at K = 256 15% better.  Below 64 there is no time to gain either.
MIN_K = 64


def _eligible(x, weight):
    return (ENABLED and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 2
            and x.shape[0] > 0 and x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0
            and weight.shape[1] >= MIN_K        # forward reduction length
            and weight.shape[0] >= MIN_K)       # input-gradient reduction length


def _gemm(a, packed, n):
    """a @ packed^T, first n columns.  The store epilogue writes 16-byte vectors: n is rounded up into the packed
    layer's zero-weight padding and sliced off again."""
    n4 = (n + 3) // 4 * 4
    out = fl.linear(a, packed, n_out=n4)
    return out if n4 == n else out[:, :n]


class _TCLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mask):
        y = _gemm(x, fl.pack(weight, bias, mask=mask), weight.shape[0])
        ctx.save_for_backward(x, weight, mask)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, mask = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wm = weight if mask is None else weight * mask
            gx = _gemm(gy, fl.pack(wm.t().contiguous(), None), weight.shape[1])
            if gx.stride(0) != weight.shape[1]:
                gx = gx.contiguous()
        if ctx.needs_input_grad[1]:
            if WGRAD_TC and weight.shape[1] >= MIN_K and x.shape[0] >= 4096 and x.shape[0] % 4 == 0:
                # grad_W[N, K] = grad_y^T[N, B] @ x[B, K]: a reduction over the batch -> split-K on the tensor cores
                # (both operands transposed once so that the batch is the contiguous reduction axis)
                # (grad_y is read as it lies: the kernel transposes it on its way into tensor memory)
                if WGRAD_T and ctx.has_bias and ctx.needs_input_grad[2]:
                    gw, gb = fl.linear_splitk_t(gy, fl.pack_transposed(x), column_sums=True)  # bias gradient for free
                elif WGRAD_T:
                    gw = fl.linear_splitk_t(gy, fl.pack_transposed(x))
                else:
                    gw = fl.linear_splitk(fl.transpose(gy), fl.pack_transposed(x))
                if gw.stride(0) != weight.shape[1]:
                    gw = gw.contiguous()
            else:
                gw = gy.t().mm(x)
            if mask is not None:
                gw = gw * mask
        if gb is None and ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None


def linear(x, weight, bias=None, mask=None):
    if not _eligible(x, weight) or gy_misaligned(weight):
        return F.linear(x, weight if mask is None else weight * mask, bias)
    return _TCLinear.apply(x, weight, bias, mask)


def gy_misaligned(weight):
    # the input-gradient GEMM reads gy [M, N] through TMA: N * 4 bytes must be a multiple of 16
    return weight.shape[0] % 4 != 0


def module_linear(module, x):
    """`module(x)` for an nn.Linear / MaskedLinear, through `linear`."""
    return linear(x, module.weight, module.bias, getattr(module, "mask", None))
