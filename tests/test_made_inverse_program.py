"""Host logic of the incremental autoregressive inverse: the program `made_inverse.compile_made` emits,
interpreted on the CPU exactly as csrc/fc_made_inverse.cu does, must reproduce the reference's D-pass inverse
(flowcon/transforms/autoregressive/autoregressive.py:44-53, restated in oracle/restated.py `_autoregressive`)."""
import math

import numpy as np
import pytest
import torch

from flowconductor_b200 import made_inverse, workloads
from flowconductor_b200.transforms import made as made_module
from oracle import restated

from .helpers import emulate_made_program, golden_state, load_golden


def _maf_layers(name):
    gold = load_golden(name)
    wl = workloads.get_workload(name)
    flow = workloads.build_flow(wl)
    flow.load_state_dict(golden_state(gold), strict=True)
    state = golden_state(gold, dtype=torch.float64)
    specs = workloads.oracle_specs(wl)
    layers = [(spec, t) for spec, t in zip(specs, flow._transform._transforms) if spec["kind"].startswith("maf_")]
    return state, layers, wl


@pytest.mark.parametrize("name", ["cfg3_small", "cfg1"])
def test_program_reproduces_d_pass_inverse(name):
    state, layers, wl = _maf_layers(name)
    torch.manual_seed(3)
    z = torch.randn(37, wl["features"], dtype=torch.float64) * 1.3
    for spec, layer in layers:
        P = layer._output_dim_multiplier()
        prog = made_inverse.compile_made(layer.autoregressive_net, P)
        assert prog is not None
        assert prog.features == wl["features"] and prog.params_per_feature == P
        assert made_inverse.smem_bytes(prog.features, P, prog.n_arrays, prog.hidden) <= made_inverse.SMEM_LIMIT

        if spec["kind"] == "maf_prq":
            def invert(zf, params):
                y, lad = restated.rq_elementwise(zf[:, None], params, spec["num_bins"], spec.get("tails"),
                                                 spec.get("tail_bound", 1.0), True, None, True, constrained_bound=1.2)
                return y[:, 0], lad
        else:
            def invert(zf, params):
                y, lad = restated.affine_elementwise(zf[:, None], params, "interleaved", "softplus_eps", True)
                return y[:, 0], lad

        x, lad = emulate_made_program(prog, z, invert)
        x_ref, lad_ref = restated.apply_layer(state, spec, z, inverse=True)
        assert (x - x_ref).abs().max() < 1e-9
        assert (lad - lad_ref).abs().max() < 1e-9


@pytest.mark.parametrize("features,hidden,blocks", [(2, 256, 2), (1, 32, 1), (3, 100, 2), (5, 7, 1), (48, 64, 3), (7, 200, 1)])
def test_program_shapes(features, hidden, blocks):
    """Shapes with very large degree groups (2 features: every hidden unit has degree 1), a single feature, more features
    than hidden units: the compiled program obeys the kernel's limits (checked by the emulator) and reproduces the oracle."""
    torch.manual_seed(features * 1000 + hidden)
    net = made_module.MADE(features=features, hidden_features=hidden, num_blocks=blocks, output_multiplier=2)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    prog = made_inverse.compile_made(net, 2)
    assert prog is not None
    for _, hdr, tasks in prog.tasks():
        assert 1 <= len(tasks) <= made_inverse.TASKS and all(1 <= t["nj"] <= made_inverse.MAX_NJ for t in tasks)
    state = {"autoregressive_net." + k: v.double() for k, v in net.state_dict().items()}
    spec = {"kind": "maf_affine", "prefix": "", "num_blocks": blocks, "hidden_features": hidden}
    z = torch.randn(9, features, dtype=torch.float64)

    def invert(zf, params):
        y, lad = restated.affine_elementwise(zf[:, None], params, "interleaved", "softplus_eps", True)
        return y[:, 0], lad

    x, lad = emulate_made_program(prog, z, invert)
    x_ref, lad_ref = restated.apply_layer(state, spec, z, inverse=True)
    assert (x - x_ref).abs().max() < 1e-9 and (lad - lad_ref).abs().max() < 1e-9


def test_program_structure_cfg3():
    """cfg 3 (D=16, H=256, 2 blocks, P=47): per pass one wide phase + 4 chain phases + the feature's chain phase."""
    wl = workloads.get_workload("cfg3")
    flow = workloads.build_flow(wl)
    layer = [t for t in flow._transform._transforms if hasattr(t, "autoregressive_net")][0]
    prog = made_inverse.compile_made(layer.autoregressive_net, 47)
    assert prog is not None
    phases = prog.tasks()
    feats = [h["feature"] for _, h, _ in phases if h["feature"] >= 0]
    assert feats == list(range(16))
    assert prog.n_phases == 1 + 15 * 6  # feature 0: bias only; then (wide, 4 layers, parameters) per pass
    macs = 0
    for _, hdr, tasks in phases:
        assert hdr["width"] == sum(4 * ((t["nj"] + 3) // 4) for t in tasks) and hdr["width"] <= 2048 - 128
        assert hdr["rows"] == max(t["kn"] for t in tasks)
        macs += sum(t["kn"] * t["nj"] for t in tasks)
    # multiply-adds of the whole inverse vs 16 full conditioner passes (the point of n2)
    full = 16 * 256 + 4 * 256 * 256 + 256 * 752
    assert macs < 0.6 * full and 16 * full / macs > 25
    assert made_inverse.smem_bytes(16, 47, prog.n_arrays, 256) <= made_inverse.SMEM_LIMIT


def test_unsupported_structures_do_not_compile():
    net = made_module.MADE(features=6, hidden_features=16, num_blocks=1, output_multiplier=2, use_residual_blocks=False)
    assert not made_inverse.supported_made(net)
    # a mask without the prefix property (rows permuted against the degree order of the layer below)
    net = made_module.MADE(features=6, hidden_features=16, num_blocks=1, output_multiplier=2)
    with torch.no_grad():
        net.blocks[0].linear_layers[0].mask.copy_(net.blocks[0].linear_layers[0].mask.flip(1))
    assert made_inverse.compile_made(net, 2) is None
    # too large for one CTA's shared memory
    big = made_module.MADE(features=8, hidden_features=1024, num_blocks=2, output_multiplier=2)
    assert made_inverse.compile_made(big, 2) is None
