"""Host side of the incremental autoregressive inverse (csrc/fc_made_inverse.cu, C ABI `fc_made_inverse_*`).

`AutoregressiveTransform.inverse` of the reference (flowcon/transforms/autoregressive/autoregressive.py:44-53) runs the whole
MADE D times.  `compile_made` turns a residual MADE (flowcon/transforms/made.py:205-283) into the straight-line program the
kernel interprets: the hidden units are ordered by degree, so that every masked weight row (made.py:28-51,72) reads a PREFIX
of the layer below; pass f then computes only the units that become valid with feature f - 1 and the P parameters of
feature f, after which the kernel inverts that feature.  The schedule is derived from the MASKS, not from the degree
formula: any mask pattern with the prefix property compiles, anything else returns None (the D-pass inverse runs).

No fallback inside: `apply_*` needs CUDA fp32 tensors and the library.
"""
import ctypes

import numpy as np
import torch

from . import _cabi

MAX_NJ = 24      # FC_MADE_MAX_NJ
ROWS = 32        # rows per CTA of the kernel
SMEM_LIMIT = 232448  # opt-in shared memory per CTA on sm_100

STEP_FIELDS = ("in_array", "out_array", "k_count", "j0", "nj", "nj4", "relu_in", "res_array", "feature", "w_off4", "b_off",
               "reserved")


class MadeProgram:
    """Compiled program + its device buffers (`steps`, `weights`, `bias` keep the memory alive)."""

    def __init__(self, steps_np, weights, bias, features, params_per_feature, n_arrays, hidden, device):
        self.steps_np = steps_np  # [n_steps, 12] int32 (host copy: tests, debugging)
        self.steps = torch.from_numpy(steps_np.copy()).to(device)
        self.weights, self.bias = weights, bias
        self.features, self.params_per_feature, self.n_arrays, self.hidden = features, params_per_feature, n_arrays, hidden
        self.struct = None
        if weights.is_cuda:
            st = _cabi.MadeProgramStruct()
            st.steps, st.weights, st.bias = self.steps.data_ptr(), weights.data_ptr(), bias.data_ptr()
            st.n_steps, st.features, st.params_per_feature = steps_np.shape[0], features, params_per_feature
            st.n_arrays, st.hidden = n_arrays, hidden
            self.struct = st

    @property
    def n_steps(self):
        return self.steps_np.shape[0]


def smem_bytes(features, params_per_feature, n_arrays, hidden):
    """Mirror of fc_made_inverse_smem_bytes (kept in sync by tests/test_host_api.py)."""
    ps = params_per_feature if params_per_feature % 2 else params_per_feature + 1
    ring = 4 * 64 * MAX_NJ * 4
    return ring + 4 * (features * ROWS + n_arrays * hidden * ROWS + ROWS * ps + 8 * MAX_NJ * ROWS) + 8 * 2 * 4 + 128


def _prefix_counts(mask):
    """mask [N, K] of 0/1 -> per-row count of leading ones, or None if some row is not of the form 1..1 0..0."""
    cnt = mask.sum(dim=1).round().long()
    k = torch.arange(mask.shape[1], device=mask.device)
    if not bool(((k[None, :] < cnt[:, None]).to(mask.dtype) == mask).all()):
        return None
    return cnt


def supported_made(net):
    """Structure the compiler handles: residual blocks, ReLU, no context / batch norm / active dropout."""
    from .nn import tensorcore

    if not getattr(net, "use_residual_blocks", False) or hasattr(net, "context_layer"):
        return False
    if not tensorcore._is_relu(net.activation):
        return False
    for blk in net.blocks:
        if getattr(blk, "use_batch_norm", False) or not tensorcore._is_relu(blk.activation) or hasattr(blk, "context_layer"):
            return False
        if blk.dropout.p > 0 and blk.training:
            return False
    return True


def compile_made(net, params_per_feature):
    """-> MadeProgram, or None when the network does not have the structure the kernel needs."""
    init, fin = net.initial_layer, net.final_layer
    dev = init.weight.device
    H, D = init.weight.shape
    P = params_per_feature
    if fin.weight.shape[0] != D * P:
        return None
    nb = len(net.blocks)
    n_arrays = 1 + 2 * nb
    if smem_bytes(D, P, n_arrays, H) > SMEM_LIMIT:
        return None
    with torch.no_grad():
        deg = init.degrees.detach().cpu()
        for blk in net.blocks:
            for lin in blk.linear_layers:
                if not torch.equal(lin.degrees.detach().cpu(), deg):
                    return None
        perm = torch.argsort(deg, stable=True).to(dev)
        # layers: (module, input array, output array (0 = parameter tile), relu on the input, skip-connection array)
        layers = [(init, 0, 1, 0, 0)]
        for b, blk in enumerate(net.blocks):
            h_in = 1 if b == 0 else 2 * b + 1
            layers.append((blk.linear_layers[0], h_in, 2 * b + 2, 1, 0))
            layers.append((blk.linear_layers[1], 2 * b + 2, 2 * b + 3, 1, h_in))
        layers.append((fin, n_arrays, 0, 0, 0))
        sorted_w, counts, bias_parts, bias_base = [], [], [], []
        off = 0
        for idx, (lin, in_a, out_a, _, _) in enumerate(layers):
            w = (lin.weight * lin.mask).detach().float()
            m = lin.mask.detach().float()
            b = lin.bias.detach().float() if lin.bias is not None else torch.zeros(w.shape[0], device=dev)
            if in_a != 0:
                w, m = w[:, perm], m[:, perm]
            if out_a != 0:
                w, m, b = w[perm], m[perm], b[perm]
            cnt = _prefix_counts(m)
            if cnt is None:
                return None
            sorted_w.append(w.contiguous())
            counts.append(cnt.cpu().numpy())
            bias_parts.append(b)
            bias_base.append(off)
            off += b.numel()
        producer = {layers[i][2]: i for i in range(len(layers) - 1)}  # hidden array -> index of the layer that writes it
        ready = [0] * (n_arrays + 1)
        steps, blocks = [], []
        w_floats = [0]

        def emit(li, j_lo, j_hi, k_count, feature, j_base):
            lin, in_a, out_a, relu_in, res_a = layers[li]
            j = j_lo
            while j < j_hi:
                nj = min(MAX_NJ, j_hi - j)
                nj4 = (nj + 3) // 4
                last = j + nj >= j_hi
                assert w_floats[0] % 4 == 0
                steps.append([in_a, out_a, k_count, j - j_base, nj, nj4, relu_in, res_a,
                              feature if (last and feature is not None) else -1, w_floats[0] // 4, bias_base[li] + j, 0])
                if k_count > 0:
                    blk = torch.zeros((k_count, 4 * nj4), dtype=torch.float32, device=dev)
                    blk[:, :nj] = sorted_w[li][j:j + nj, :k_count].t()
                    blocks.append(blk.reshape(-1))
                    w_floats[0] += k_count * 4 * nj4
                j += nj

        def ensure(arr, upto, n_inverted):
            """Make units [0, upto) of hidden array `arr` available (recursively what they read)."""
            if arr == 0:
                return upto <= n_inverted
            if ready[arr] >= upto:
                return True
            li = producer[arr]
            _, in_a, _, _, res_a = layers[li]
            lo = ready[arr]
            need = int(counts[li][lo:upto].max())
            if not ensure(in_a, need, n_inverted):
                return False
            if res_a and not ensure(res_a, upto, n_inverted):
                return False
            emit(li, lo, upto, need, None, 0)
            ready[arr] = upto
            return True

        fl = len(layers) - 1
        for f in range(D):
            c = counts[fl][f * P:(f + 1) * P]
            need = int(c.max())
            if not ensure(n_arrays, need, f):  # feature f may read features < f only
                return None
            emit(fl, f * P, (f + 1) * P, need, f, f * P)
        steps_np = np.asarray(steps, dtype=np.int32).reshape(-1, len(STEP_FIELDS))
        weights = torch.cat(blocks) if blocks else torch.zeros((4,), dtype=torch.float32, device=dev)
        bias = torch.cat(bias_parts).contiguous()
    return MadeProgram(steps_np, weights.contiguous(), bias, D, P, n_arrays, H, dev)


def _io(prog, z, inplace_ok=False):
    _cabi.require_cuda_f32(z, "inputs")
    if z.dim() != 2 or z.shape[1] != prog.features:
        raise ValueError("expected inputs of shape [B, {}]".format(prog.features))
    z, zp, ldz = _cabi.rows(z)
    x = torch.empty((z.shape[0], prog.features), dtype=torch.float32, device=z.device)
    lad = torch.empty((z.shape[0],), dtype=torch.float32, device=z.device)
    return z, zp, ldz, x, lad


def apply_rqs(prog, z, cfg, status=None):
    """Inverse of a MAF layer with rational-quadratic splines: (x, logabsdet) = layer.inverse(z)."""
    L = _cabi.lib()
    z, zp, ldz, x, lad = _io(prog, z)
    with torch.cuda.device(z.device), _cabi.launch("fc_made_inverse_rqs", z.device):
        rc = L.fc_made_inverse_rqs(ctypes.byref(prog.struct), zp, ldz, x.data_ptr(), x.stride(0), lad.data_ptr(), 0,
                                   z.shape[0], ctypes.byref(cfg), status.data_ptr() if status is not None else None,
                                   _cabi.stream_ptr(z.device))
    _cabi.check(rc, "fc_made_inverse_rqs")
    return x, lad


def apply_affine(prog, z, activation):
    """Inverse of a MaskedAffineAutoregressiveTransform layer."""
    L = _cabi.lib()
    z, zp, ldz, x, lad = _io(prog, z)
    with torch.cuda.device(z.device), _cabi.launch("fc_made_inverse_affine", z.device):
        rc = L.fc_made_inverse_affine(ctypes.byref(prog.struct), zp, ldz, x.data_ptr(), x.stride(0), lad.data_ptr(), 0,
                                      z.shape[0], int(activation), _cabi.stream_ptr(z.device))
    _cabi.check(rc, "fc_made_inverse_affine")
    return x, lad


ENABLED = True  # False: every autoregressive inverse takes the D-pass path


def program_for(net, params_per_feature):
    """Compiled program of `net`, cached on the module with the same key rules as the packed tensor-core weights
    (nn/tensorcore.py: data pointer + in-place version of every parameter).  None if the net does not compile."""
    from .nn import tensorcore

    key = (tensorcore._param_key(net), params_per_feature, bool(net.training))
    cached = getattr(net, "_fc_made_plan", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    prog = compile_made(net, params_per_feature) if supported_made(net) else None
    object.__setattr__(net, "_fc_made_plan", (key, prog))
    tensorcore._generation[0] += 1
    return prog


def usable(net, inputs, context):
    """Should this inverse call take the incremental kernel?"""
    from .nn import tensorcore
    from .transforms import made as made_module

    if not ENABLED or context is not None or not isinstance(net, made_module.MADE):
        return False
    if tensorcore.wants_grad(net, inputs):
        return False
    return inputs.is_cuda and inputs.dtype == torch.float32 and inputs.dim() == 2 and inputs.shape[0] > 0
