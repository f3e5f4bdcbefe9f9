"""Host side of the tensor-core conditioner path (csrc/fc_linear.cu, C ABI `fc_linear_*`).

`PackedLinear` is the packed ("3xTF32" hi/lo planes, padded) form of one nn.Linear / MaskedLinear
(flowcon/nn/nets/resnet.py:26-28,69-91, flowcon/transforms/made.py:15-72); `linear` and `linear_rqs` launch the
kernels on torch's current stream.  No fallback: CPU tensors or a missing library raise.
"""
import ctypes

import torch

from . import _cabi

N_TILE_STORE = 256  # fc_linear_apply
N_TILE_RQS = 192    # fc_linear_rqs_apply
RQS_PPAD = {8: 24, 16: 48}  # accumulator columns per feature for the supported bin counts (P = 3K-1 -> P_pad)


def _ceil_to(v, m):
    return (v + m - 1) // m * m


class PackedLinear:
    """Packed weights of one dense layer.  `n_out` = logical output width (before padding)."""

    def __init__(self, w, b, n_out, k_in):
        self.w = w          # [2, n_pad, k_pad] fp32 (tf32 hi plane, tf32 lo plane)
        self.b = b          # [n_pad]
        self.n_out = n_out
        self.k_in = k_in    # width of the activation matrix the layer multiplies (after any column scatter)
        self.struct = _cabi.LinearWeights(w.data_ptr(), b.data_ptr(), w.shape[1], w.shape[2])

    @property
    def n_pad(self):
        return self.w.shape[1]

    @property
    def k_pad(self):
        return self.w.shape[2]


def pack(weight, bias, mask=None, row_map=None, col_map=None, k_in=None, n_pad=None, n_tile=N_TILE_STORE):
    """fc_linear_pack.  weight [N, K] (+ optional {0,1} mask [N, K], made.py:72), bias [N] or None.
    row_map int32[N]: packed row of each weight row; col_map int32[K]: activation column each weight column
    multiplies (k_in = width of that activation matrix)."""
    _cabi.require_cuda_f32(weight, "weight")
    L = _cabi.lib()
    weight = weight.detach()
    N, K = weight.shape
    if weight.stride(1) != 1:
        weight = weight.contiguous()
    k_in = K if k_in is None else k_in
    k_pad = _ceil_to(k_in, 32)
    rows = N if row_map is None else int(row_map.max().item()) + 1
    n_pad = _ceil_to(max(rows, n_pad or 0), n_tile)
    dev = weight.device
    w = torch.empty((2, n_pad, k_pad), dtype=torch.float32, device=dev)
    b = torch.empty((n_pad,), dtype=torch.float32, device=dev)
    if bias is not None:
        bias = _cabi.require_cuda_f32(bias.detach(), "bias").contiguous()
    if mask is not None:
        mask = _cabi.require_cuda_f32(mask.detach(), "mask")
        if mask.stride(1) != 1:
            mask = mask.contiguous()
    for m in (row_map, col_map):
        if m is not None:
            assert m.dtype == torch.int32 and m.is_cuda and m.is_contiguous()
    with torch.cuda.device(dev), _cabi.launch("fc_linear_pack", dev):
        rc = L.fc_linear_pack(weight.data_ptr(), weight.stride(0), mask.data_ptr() if mask is not None else None,
                              mask.stride(0) if mask is not None else 0,
                              bias.data_ptr() if bias is not None else None, N, K,
                              row_map.data_ptr() if row_map is not None else None,
                              col_map.data_ptr() if col_map is not None else None, n_pad, k_pad, w.data_ptr(),
                              b.data_ptr(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "fc_linear_pack")
    return PackedLinear(w, b, rows, k_in)


def grouped_row_map(n_groups, group, group_pad, device):
    """row j*group + i  ->  j*group_pad + i  (per-feature parameter groups padded to the epilogue's stride)."""
    j = torch.arange(n_groups, device=device).repeat_interleave(group)
    i = torch.arange(group, device=device).repeat(n_groups)
    return (j * group_pad + i).to(torch.int32)


def linear(a, packed, relu_in=False, relu_out=False, residual=None, out=None):
    """out = act_out(act_in(a) @ W.T + b (+ residual)); a [M, K] fp32 (row stride a multiple of 4 floats)."""
    _cabi.require_cuda_f32(a, "activations")
    L = _cabi.lib()
    a, ap, lda = _cabi.rows(a)
    M = a.shape[0]
    if a.shape[1] != packed.k_in:
        raise ValueError("activations have {} columns, the packed layer expects {}".format(a.shape[1], packed.k_in))
    if out is None:
        out = torch.empty((M, packed.n_out), dtype=torch.float32, device=a.device)
    rp, ldr = None, 0
    if residual is not None:
        residual, rp, ldr = _cabi.rows(_cabi.require_cuda_f32(residual, "residual"))
    with torch.cuda.device(a.device), _cabi.launch("fc_linear_apply", a.device):
        rc = L.fc_linear_apply(ap, lda, M, a.shape[1], ctypes.byref(packed.struct), int(relu_in), out.data_ptr(),
                               out.stride(0), packed.n_out, int(relu_out), rp, ldr, _cabi.stream_ptr(a.device))
    _cabi.check(rc, "fc_linear_apply")
    return out


def linear_rqs(hidden, packed, x, y, logabsdet, accumulate, d_t, tcols, ccols, cfg, status, relu_in=False):
    """Final conditioner layer + rational-quadratic spline in one kernel (fc_linear_rqs_apply).
    Writes y[:, tcols] (and y[:, ccols] = x[:, ccols] unless y is x) and logabsdet in place."""
    _cabi.require_cuda_f32(hidden, "hidden activations")
    _cabi.require_cuda_f32(x, "inputs")
    L = _cabi.lib()
    hidden, hp, ldh = _cabi.rows(hidden)
    assert x.stride(1) == 1 and y.stride(1) == 1 and logabsdet.is_contiguous()
    B = hidden.shape[0]
    with torch.cuda.device(x.device), _cabi.launch("fc_linear_rqs_apply", x.device):
        rc = L.fc_linear_rqs_apply(hp, ldh, B, hidden.shape[1], ctypes.byref(packed.struct), int(relu_in),
                                   x.data_ptr(), x.stride(0), y.data_ptr(), y.stride(0), logabsdet.data_ptr(),
                                   int(accumulate), d_t, _cabi.cols(tcols), _cabi.cols(ccols), ctypes.byref(cfg),
                                   status.data_ptr() if status is not None else None, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_linear_rqs_apply")
    return y, logabsdet
