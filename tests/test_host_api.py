"""CPU tests of the host side: API mirror, state_dict compatibility with the reference, C-ABI exports,
and the 'no CPU fallback' contract."""
import ctypes
import os
import re

import pytest
import torch

import flowconductor_b200 as fcb
from flowconductor_b200 import _cabi, transforms, workloads
from flowconductor_b200.nn import nets
from flowconductor_b200.utils import torchutils
from tests.helpers import ROOT, golden_state, load_golden

MODELS = ["cfg1", "cfg2_small", "cfg3_small", "cfg4_small", "affine_coupling_small", "cond_prq_small",
          "maf_sos_small", "prq_coupling_notails_small", "prq_coupling_uncond_small", "plin_coupling_small",
          "maf_plin_small", "pquad_coupling_small", "maf_pquad_small", "pcubic_coupling_small", "maf_pcubic_small", "actnorm_maf_small"]


@pytest.mark.parametrize("name", MODELS)
def test_reference_state_dict_loads_strictly(name):
    """Golden state_dicts were saved from the unmodified reference: same keys, same shapes."""
    gold = load_golden(name)
    flow = workloads.build_flow(workloads.get_workload(name))
    state = golden_state(gold)
    assert set(flow.state_dict().keys()) == set(state.keys())
    flow.load_state_dict(state, strict=True)
    for k, v in flow.state_dict().items():
        assert v.shape == state[k].shape, k


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "flowcon_b200.h")).read()
    declared = set(re.findall(r"\b(fc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert os.path.exists(_cabi.LIB_PATH), "build the library first: python -m flowconductor_b200.build"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), "libflowcon_b200.so does not export " + sym
    assert set(_cabi.EXPORTS) <= declared
    lib.fc_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.fc_version()
    assert lib.fc_built_for_sm() == 100


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _cabi.lib()
    cfg = _cabi.RqsConfig(8, 1, 0, 0, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 1.0)
    none = _cabi.Cols(None, 0)
    # null pointers -> FC_ERR_INVALID_ARGUMENT before any CUDA call
    assert lib.fc_rqs_apply(None, 0, None, 0, None, 0, None, 0, 4, 2, none, none, ctypes.byref(cfg), None, None) == -1
    bad = _cabi.RqsConfig(2000, 1, 0, 0, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 1.0)
    assert lib.fc_rqs_apply(None, 0, None, 0, None, 0, None, 0, 4, 2, none, none, ctypes.byref(bad), None, None) < 0
    assert lib.fc_sos_apply(None, 0, None, 0, None, 0, None, 0, 4, 2, 10, 0.0, 0, 50, 120.0, None) == -1
    assert lib.fc_affine_apply(None, 0, None, 0, None, 0, None, 0, 4, 2, none, none, 0, 0, 0, None) == -1


def test_cpu_tensors_fail_loudly():
    """north_star: no CPU fallback.  A CPU tensor must raise, not silently take another path."""
    flow = workloads.build_flow(workloads.get_workload("cfg2_small"))
    with pytest.raises(RuntimeError, match="CUDA"):
        flow.log_prob(torch.randn(4, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        transforms.unconstrained_rational_quadratic_spline(torch.zeros(3), torch.zeros(3, 4), torch.zeros(3, 4),
                                                           torch.zeros(3, 3))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libflowcon_b200.so")
    with pytest.raises(_cabi.LibraryMissing):
        _cabi.lib()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flowconductor_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
                assert "hostmath" not in src, os.path.join(dirpath, f)


def test_masks_and_helpers():
    assert torchutils.create_alternating_binary_mask(5, even=True).tolist() == [1, 0, 1, 0, 1]
    assert torchutils.create_alternating_binary_mask(5, even=False).tolist() == [0, 1, 0, 1, 0]
    assert torchutils.create_mid_split_binary_mask(5).tolist() == [1, 1, 1, 0, 0]
    assert int(torchutils.create_random_binary_mask(7).sum()) == 4
    assert torchutils.tile(torch.tensor([1, 2, 3]), 2).tolist() == [1, 1, 2, 2, 3, 3]
    x = torch.arange(24.0).reshape(2, 3, 4)
    assert torchutils.sum_except_batch(x).tolist() == x.reshape(2, -1).sum(1).tolist()  # torchutils_test.py:106-114
    assert torchutils.repeat_rows(torch.tensor([[1], [2]]), 2).flatten().tolist() == [1, 1, 2, 2]
    assert torchutils.merge_leading_dims(x, 2).shape == (6, 4)
    assert torchutils.split_leading_dim(torch.zeros(6, 4), [2, 3]).shape == (2, 3, 4)


def test_made_masks_are_autoregressive():
    """tests/transforms/made_test.py:109-137: the product of all masks is strictly lower triangular
    (output of feature d sees only features < d)."""
    d, h, mult = 6, 16, 3
    made = transforms.MADE(features=d, hidden_features=h, num_blocks=2, output_multiplier=mult)
    total = made.initial_layer.mask
    for block in made.blocks:
        total = block.linear_layers[1].mask @ (block.linear_layers[0].mask @ total) + total
    total = made.final_layer.mask @ total
    total = (total > 0).reshape(d, mult, d)
    for out_f in range(d):
        for in_f in range(d):
            if in_f >= out_f:
                assert not total[out_f, :, in_f].any()
    assert not hasattr(made, "hidden_features")  # autoregressive.py:589 relies on this


def test_coupling_constructor_contract():
    net = lambda i, o: nets.ResidualNet(i, o, hidden_features=8)  # noqa: E731
    with pytest.raises(ValueError):
        transforms.PiecewiseRationalQuadraticCouplingTransform(torch.zeros(2, 2), net)
    with pytest.raises(ValueError):
        transforms.PiecewiseRationalQuadraticCouplingTransform(torch.zeros(0), net)
    t = transforms.PiecewiseRationalQuadraticCouplingTransform([1, 0, 1, 0], net, num_bins=8, tails="linear")
    assert t.transform_features.tolist() == [0, 2] and t.identity_features.tolist() == [1, 3]
    assert t.transform_net.final_layer.out_features == 2 * 23
    t = transforms.PiecewiseRationalQuadraticCouplingTransform([1, 0, 1, 0], net, num_bins=8, tails=None)
    assert t.transform_net.final_layer.out_features == 2 * 25
    with pytest.raises(ValueError):
        t.forward(torch.zeros(3, 5))
    with pytest.raises(TypeError):
        transforms.ConditionalSumOfSigmoidsTransform(4, 8, context_features=2).forward(torch.zeros(3, 4))
    with pytest.raises(transforms.InverseNotAvailable):
        transforms.Transform().inverse(torch.zeros(1, 1))


def test_package_metadata():
    assert fcb.__version__


def test_inplace_ownership_protocol():
    """nn/tensorcore.py: a layer may overwrite its input only if the cascade declared it private AND the previous
    layer call allocated it; user tensors and stale pointers never qualify."""
    import torch
    from flowconductor_b200.nn import tensorcore as tcm

    user = torch.zeros(4, 4)
    with torch.no_grad():
        tcm.end_cascade()
        tcm.begin_layer(None)                      # first layer: input belongs to the caller
        assert not tcm.may_overwrite(user)
        y0 = torch.empty(4, 4)
        tcm.mark_fresh(y0)                         # layer 0 allocated its output
        tcm.begin_layer(y0)                        # cascade: y0 is my private intermediate
        assert tcm.may_overwrite(y0)
        assert not tcm.may_overwrite(user)
        view = user                                # layer 1 returns the user's tensor itself (identity-like layer)
        tcm.begin_layer(view)
        assert not tcm.may_overwrite(view)         # nobody marked it fresh in the call that produced it
        tcm.mark_fresh(y0)
        tcm.end_cascade()
        tcm.begin_layer(y0)                        # stale pointer from a finished cascade
        assert not tcm.may_overwrite(y0)
        tcm.end_cascade()
    tcm.mark_fresh(y0)
    tcm.begin_layer(y0)
    assert not tcm.may_overwrite(y0)               # autograd enabled: never in place
    tcm.end_cascade()


def test_t128_layout_round_trip_and_row_maps():
    """Host side of the tensor-core path (linear.py): the T128 activation layout of include/flowcon_b200.h and the
    packed-row maps of the fused final layers, checked on CPU tensors."""
    import torch
    from flowconductor_b200 import _cabi, linear as fl

    t = torch.arange(300 * 32, dtype=torch.float32).reshape(300, 32)
    tiled = fl.T128.from_rows(t)
    assert tiled.buf.numel() == 4 * 128 * 32  # 3 tiles of data, rounded up to an even count: the kernels work on tile pairs
    assert torch.equal(tiled.to_rows(), t)
    # element (r, c) -> (r // 128) * 128 * W + ((c // 4) * 128 + r % 128) * 4 + c % 4
    for r, c in ((0, 0), (5, 7), (129, 31), (299, 16)):
        off = (r // 128) * 128 * 32 + ((c // 4) * 128 + r % 128) * 4 + c % 4
        assert tiled.buf[off].item() == t[r, c].item()
    try:
        fl.T128(10, 20, torch.device("cpu"))
        raise AssertionError("width 20 must be rejected")
    except ValueError:
        pass
    rm = fl.grouped_row_map(3, 23, 24, torch.device("cpu"))
    assert rm.tolist()[:3] == [0, 1, 2] and rm.tolist()[23] == 24 and rm.tolist()[-1] == 2 * 24 + 22
    blocked = fl.affine_row_map(3, _cabi.AFFINE_BLOCKED, torch.device("cpu")).tolist()
    assert blocked == [1, 3, 5, 0, 2, 4]          # [shift_0..2 | raw_0..2] -> (raw_j, shift_j) pairs
    assert fl.affine_row_map(3, _cabi.AFFINE_INTERLEAVED, torch.device("cpu")).tolist() == [0, 1, 2, 3, 4, 5]


def test_tensorcore_path_selection():
    """nn/tensorcore.py + nn/tc_autograd.py: which conditioners / layers are allowed on the tensor cores."""
    import torch
    from flowconductor_b200 import transforms as T
    from flowconductor_b200.nn import tc_autograd, tensorcore
    from flowconductor_b200.nn.nets import ResidualNet

    assert tensorcore.supported_net(ResidualNet(32, 100, 64), None)
    assert not tensorcore.supported_net(ResidualNet(32, 100, 32), None)            # narrower than MIN_K
    assert not tensorcore.supported_net(ResidualNet(32, 100, 64, context_features=4), None)
    assert not tensorcore.supported_net(ResidualNet(32, 100, 64, use_batch_norm=True), None)
    assert not tensorcore.supported_net(ResidualNet(32, 100, 64), torch.zeros(1, 4))  # context passed at call time
    made = T.MaskedAffineAutoregressiveTransform(6, 64).autoregressive_net
    assert tensorcore.supported_net(made, None)
    x = torch.zeros(8, 64)
    assert not tensorcore.usable(ResidualNet(64, 100, 64), x, None)                # CPU tensors never
    w = torch.zeros(128, 128)
    assert not tc_autograd._eligible(x, w)                                          # CPU tensors never
    # CPU tensors take F.linear (torch), results identical to the reference formula
    lin = torch.nn.Linear(16, 8)
    xi = torch.randn(4, 16)
    assert torch.equal(tc_autograd.module_linear(lin, xi), torch.nn.functional.linear(xi, lin.weight, lin.bias))


def test_graph_capture_refuses_cpu_tensors():
    """graphs.capture has no CPU fallback: it raises before touching CUDA when given host tensors."""
    from flowconductor_b200 import graphs

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        graphs.capture(lambda x: x * 2, torch.zeros(4, 2))
