"""The CTA-level tile ring (csrc/fc_pipeline.cuh: tiled_apply_kernel / tiled_backward_kernel) against the general-strides
kernels (fc_staged.cuh) on the same inputs: same element arithmetic, so the results must agree bit for bit.  Shapes cover
rows that are not multiples of 16 bytes (D = 21 autoregressive, 3 + 3 coupling: full tiles travel as aligned blocks, a ragged
last tile with plain loads / stores), ragged batches (1 row, one short tile, many tiles + remainder) and every family."""
import os

import pytest
import torch

from flowconductor_b200 import _cabi, ops

pytestmark = pytest.mark.gpu
STAGED, WARP_RING, TILE_RING = 0, 1, 2


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def _env(**kw):
    for k in list(os.environ):
        if k.startswith("FC_TILE") or k.startswith("FC_PIPE"):
            del os.environ[k]
    for k, v in kw.items():
        os.environ[k] = str(v)


def _columns(D, coupling, dev):
    if not coupling:
        return D, None, None
    tc = torch.arange(0, D, 2, dtype=torch.int32, device=dev)
    cc = torch.arange(1, D, 2, dtype=torch.int32, device=dev)
    return tc.numel(), tc, cc


LIN = (_cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0)


def _family(name, d_t, tc, cc):
    """(params per feature, forward(x, p, inverse), backward(x, p, gy, gl))"""
    if name == "rqs":
        rest = (8, _cabi.TAILS_LINEAR, False, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 0.25)
        return 23, (lambda x, p, inv: ops.rqs_layer(x, p, tc, cc, 8, _cabi.TAILS_LINEAR, inv, *rest[3:])[:2]), \
            (lambda x, p, gy, gl: ops.rqs_layer_backward(x, p, gy, gl, tc, cc, *rest))
    if name == "affine":
        return 2, (lambda x, p, inv: ops.affine_layer(x, p, tc, cc, _cabi.AFFINE_BLOCKED, _cabi.SCALE_SIGMOID2, inv)), \
            (lambda x, p, gy, gl: ops.affine_layer_backward(x, p, gy, gl, tc, cc, _cabi.AFFINE_BLOCKED, _cabi.SCALE_SIGMOID2, False))
    if name == "linspline":
        return 8, (lambda x, p, inv: ops.linspline_layer(x, p, tc, cc, 8, _cabi.TAILS_LINEAR, inv, -3.0, 3.0, -3.0, 3.0)[:2]), \
            (lambda x, p, gy, gl: ops.linspline_layer_backward(x, p, gy, gl, tc, cc, 8, *LIN))
    if name == "quadspline":
        tail = (1e-3, 1e-3, 0.25)
        return 15, (lambda x, p, inv: ops.quadspline_layer(x, p, tc, cc, 8, _cabi.TAILS_LINEAR, inv, -3.0, 3.0, -3.0, 3.0, *tail)[:2]), \
            (lambda x, p, gy, gl: ops.quadspline_layer_backward(x, p, gy, gl, tc, cc, 8, *LIN, *tail))
    if name == "cubicspline":
        tail = (1e-3, 1e-3, 0.25)
        return 18, (lambda x, p, inv: ops.cubicspline_layer(x, p, tc, cc, 8, _cabi.TAILS_LINEAR, inv, -3.0, 3.0, -3.0, 3.0, *tail)[:2]), \
            (lambda x, p, gy, gl: ops.cubicspline_layer_backward(x, p, gy, gl, tc, cc, 8, *LIN, *tail))
    assert name == "sos" and tc is None
    return 31, (lambda x, p, inv: ops.sos_layer(x, p, 10, -0.5, inv, 50, 120.0)), \
        (lambda x, p, gy, gl: ops.sos_layer_backward(x, p, gy, gl, 10))


@pytest.mark.parametrize("family", ["rqs", "affine", "linspline", "quadspline", "cubicspline", "sos"])
@pytest.mark.parametrize("D,coupling", [(21, False), (6, True), (64, True), (16, False)])
@pytest.mark.parametrize("B", [1, 37, 4099])
def test_tile_ring_matches_the_general_kernel_bit_for_bit(dev, family, D, coupling, B):
    if family == "sos" and coupling:
        pytest.skip("the sum-of-sigmoids layers transform every column")
    lib = _cabi.lib()
    g = torch.Generator(device=dev).manual_seed(B * 131 + D)
    d_t, tc, cc = _columns(D, coupling, dev)
    P, fwd, bwd = _family(family, d_t, tc, cc)
    x = torch.randn(B, D, generator=g, device=dev)
    p = torch.randn(B, d_t * P, generator=g, device=dev)
    gy, gl = torch.randn(B, D, generator=g, device=dev), torch.randn(B, generator=g, device=dev)
    try:
        for inverse in (False, True):
            _env()
            got = fwd(x, p, inverse)
            assert lib.fc_elementwise_last_path() == TILE_RING
            _env(FC_TILE=0, FC_PIPE=0)
            want = fwd(x, p, inverse)
            assert lib.fc_elementwise_last_path() == STAGED
            for u, v in zip(got, want):
                assert torch.equal(u, v), (family, inverse)
        _env()
        got = bwd(x, p, gy, gl)
        assert lib.fc_elementwise_last_path() == TILE_RING
        _env(FC_TILE=0, FC_PIPE=0)
        want = bwd(x, p, gy, gl)
        assert lib.fc_elementwise_last_path() == STAGED
        for u, v in zip(got, want):
            assert torch.equal(u, v), family
    finally:
        _env()


def test_per_warp_ring_is_still_selectable(dev):
    """FC_TILE=0 keeps the round-1 per-warp ring (A/B switch of scripts/bench_tile_ring.py)."""
    lib = _cabi.lib()
    x = torch.randn(1000, 64, device=dev)
    p = torch.randn(1000, 32 * 23, device=dev)
    tc = torch.arange(0, 64, 2, dtype=torch.int32, device=dev)
    cc = torch.arange(1, 64, 2, dtype=torch.int32, device=dev)
    args = (x, p, tc, cc, 8, _cabi.TAILS_LINEAR, False, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 0.25)
    try:
        _env(FC_TILE=0)
        a = ops.rqs_layer(*args)
        assert lib.fc_elementwise_last_path() == WARP_RING
        _env()
        b = ops.rqs_layer(*args)
        assert lib.fc_elementwise_last_path() == TILE_RING
    finally:
        _env()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("B,D", [(4099, 64), (1000, 21)])
def test_c_abi_accumulate_in_place_and_status(dev, B, D):
    """Straight through the C ABI (fc_rqs_apply): accumulate_logabsdet = 1 adds to the caller's vector (the producer warp's
    read-modify-write), y == x works in place (the tile is complete in shared memory before its rows are stored), and a
    spline without tails reports inputs outside the domain through the status word — tile ring and general kernel alike."""
    import ctypes

    from flowconductor_b200 import ops as _ops

    lib = _cabi.lib()
    g = torch.Generator(device=dev).manual_seed(B + D)
    x = torch.rand(B, D, generator=g, device=dev) * 0.98 + 0.01
    x[B // 2, D // 3] = 1.7  # outside [0, 1]
    p = torch.randn(B, D * 25, generator=g, device=dev)
    cfg = _ops._cfg(8, _cabi.TAILS_NONE, False, False, 0.0, 1.0, 0.0, 1.0, 1e-3, 1e-3, 1e-3, 1.0)
    res = {}
    try:
        for name, env in (("tile", {}), ("staged", {"FC_TILE": 0, "FC_PIPE": 0})):
            _env(**env)
            y = x.clone()
            lad = torch.full((B,), 3.0, device=dev)
            status = torch.zeros((1,), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                rc = lib.fc_rqs_apply(y.data_ptr(), D, p.data_ptr(), p.shape[1], y.data_ptr(), D, lad.data_ptr(), 1, B, D,
                                      _cabi.cols(None), _cabi.cols(None), ctypes.byref(cfg), status.data_ptr(),
                                      _cabi.stream_ptr(dev))
            assert rc == 0
            assert lib.fc_elementwise_last_path() == (TILE_RING if name == "tile" else STAGED)
            res[name] = (y, lad, int(status.item()))
    finally:
        _env()
    fresh = ops.rqs_layer(x, p, None, None, 8, _cabi.TAILS_NONE, False, False, 0.0, 1.0, 0.0, 1.0, 1e-3, 1e-3, 1e-3, 1.0)
    assert torch.equal(res["tile"][0], res["staged"][0]) and torch.equal(res["tile"][1], res["staged"][1])
    assert torch.equal(res["tile"][0], fresh[0]) and torch.allclose(res["tile"][1], fresh[1] + 3.0, rtol=0, atol=1e-5)
    assert res["tile"][2] == res["staged"][2] != 0


def test_very_long_rows_shrink_the_tile_instead_of_leaving_the_ring(dev):
    """256 transformed features x 23 parameters = 23.5 KB per row: a tile of 8 rows no longer fits twice in shared memory, so
    the launcher halves the consumer warps (rows per tile) rather than falling back."""
    lib = _cabi.lib()
    g = torch.Generator(device=dev).manual_seed(9)
    B, D = 300, 256
    P, fwd, bwd = _family("rqs", D, None, None)
    x = torch.randn(B, D, generator=g, device=dev)
    p = torch.randn(B, D * P, generator=g, device=dev)
    gy, gl = torch.randn(B, D, generator=g, device=dev), torch.randn(B, generator=g, device=dev)
    try:
        _env()
        got = fwd(x, p, False) + bwd(x, p, gy, gl)
        assert lib.fc_elementwise_last_path() == TILE_RING
        _env(FC_TILE=0, FC_PIPE=0)
        want = fwd(x, p, False) + bwd(x, p, gy, gl)
    finally:
        _env()
    for u, v in zip(got, want):
        assert torch.equal(u, v)
