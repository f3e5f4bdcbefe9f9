"""torch custom ops (`torch.ops.flowcon_b200.*`) over the C-ABI kernels, with autograd.

Each op is one kernel launch on torch's current CUDA stream.  Layer ops take the FULL-width input
[B, D] plus int32 column lists, so the coupling split / scatter of the reference
(flowcon/transforms/coupling.py:82-83,96-98) happens inside the kernel; autoregressive / conditional
layers pass no column lists.  Backward ops recompute the forward intermediates from (x, params)
 — nothing else is saved.
"""
import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _cabi


def _cfg(num_bins, tails, inverse, identity_init, left, right, bottom, top, min_w, min_h, min_d, wh_scale):
    return _cabi.RqsConfig(int(num_bins), int(tails), int(identity_init), int(inverse), left, right, bottom, top,
                           min_w, min_h, min_d, wh_scale)


# ------------------------------------------------------------------------------------------------
# RQ-spline layer
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::rqs_layer", mutates_args=())
def rqs_layer(x: Tensor, params: Tensor, tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int,
              tails: int, inverse: bool, identity_init: bool, left: float, right: float, bottom: float, top: float,
              min_bin_width: float, min_bin_height: float, min_derivative: float,
              wh_scale: float) -> Tuple[Tensor, Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params, broadcast_ok=True)
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    y = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    status = torch.zeros((1,), dtype=torch.int32, device=x.device)
    cfg = _cfg(num_bins, tails, inverse, identity_init, left, right, bottom, top, min_bin_width, min_bin_height,
               min_derivative, wh_scale)
    p_per = 3 * num_bins - 1 if tails == _cabi.TAILS_LINEAR else 3 * num_bins + 1
    if params.shape[1] != d_t * p_per:
        raise ValueError("transform params have {} columns, expected {} x {}".format(params.shape[1], d_t, p_per))
    with torch.cuda.device(x.device), _cabi.launch("fc_rqs_apply", x.device):
        rc = L.fc_rqs_apply(xp, xs, pp, ps, y.data_ptr(), y.shape[1], lad.data_ptr(), 0, B, d_t, _cabi.cols(tcols),
                            _cabi.cols(ccols), ctypes.byref(cfg), status.data_ptr(), _cabi.stream_ptr(x.device))
    if rc == -1 and (min_bin_width * num_bins > 1.0 or min_bin_height * num_bins > 1.0):
        # flowcon/transforms/splines/rational_quadratic.py:86-89
        raise ValueError("Minimal bin width/height too large for the number of bins")
    _cabi.check(rc, "fc_rqs_apply")
    return y, lad, status


def rqs_bins(x, params, tcols, num_bins, tails, inverse, identity_init, left, right, bottom, top, min_bin_width,
             min_bin_height, min_derivative, wh_scale):
    """Parity aid (fc_rqs_bins): (bins int32 [B, D_t], knot distance [B, D_t]) as the kernels' own arithmetic sees them;
    bin -1 = outside the linear tails.  Same argument meaning as `rqs_layer`."""
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    bins = torch.empty((B, d_t), dtype=torch.int32, device=x.device)
    dist = torch.empty((B, d_t), dtype=torch.float32, device=x.device)
    cfg = _cfg(num_bins, tails, inverse, identity_init, left, right, bottom, top, min_bin_width, min_bin_height,
               min_derivative, wh_scale)
    with torch.cuda.device(x.device), _cabi.launch("fc_rqs_bins", x.device):
        rc = L.fc_rqs_bins(xp, xs, pp, ps, B, d_t, _cabi.cols(tcols), ctypes.byref(cfg), bins.data_ptr(), dist.data_ptr(),
                           _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_rqs_bins")
    return bins, dist


@rqs_layer.register_fake
def _(x, params, tcols, ccols, num_bins, tails, inverse, identity_init, left, right, bottom, top, min_bin_width,
      min_bin_height, min_derivative, wh_scale):
    return torch.empty_like(x), x.new_empty((x.shape[0],)), x.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("flowcon_b200::rqs_layer_backward", mutates_args=())
def rqs_layer_backward(x: Tensor, params: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                       tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int, tails: int, inverse: bool,
                       identity_init: bool, left: float, right: float, bottom: float, top: float,
                       min_bin_width: float, min_bin_height: float, min_derivative: float,
                       wh_scale: float) -> Tuple[Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    gx = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    gp = torch.empty((B, params.shape[1]), dtype=x.dtype, device=x.device)
    if tcols is not None and (tcols.numel() + (ccols.numel() if ccols is not None else 0)) < x.shape[1]:
        gx.zero_()
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    cfg = _cfg(num_bins, tails, inverse, identity_init, left, right, bottom, top, min_bin_width, min_bin_height,
               min_derivative, wh_scale)
    with torch.cuda.device(x.device), _cabi.launch("fc_rqs_backward", x.device):
        rc = L.fc_rqs_backward(xp, xs, pp, ps, gyp, gys, glp, gx.data_ptr(), gx.shape[1], gp.data_ptr(), gp.shape[1],
                               B, d_t, _cabi.cols(tcols), _cabi.cols(ccols), ctypes.byref(cfg),
                               _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_rqs_backward")
    return gx, gp


@rqs_layer_backward.register_fake
def _(x, params, grad_y, grad_lad, tcols, ccols, *args):
    return torch.empty_like(x), torch.empty_like(params)


def _rqs_setup(ctx, inputs, output):
    x, params, tcols, ccols = inputs[:4]
    ctx.save_for_backward(x, params, tcols, ccols)
    ctx.hyper = inputs[4:]


def _rqs_backward(ctx, gy, gl, gstatus):
    x, params, tcols, ccols = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gp = rqs_layer_backward(x, params, gy, gl, tcols, ccols, *ctx.hyper)
    return (gx, gp) + (None,) * (2 + len(ctx.hyper))


rqs_layer.register_autograd(_rqs_backward, setup_context=_rqs_setup)


# ------------------------------------------------------------------------------------------------
# piecewise-linear spline layer
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::linspline_layer", mutates_args=())
def linspline_layer(x: Tensor, params: Tensor, tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int,
                    tails: int, inverse: bool, left: float, right: float, bottom: float,
                    top: float) -> Tuple[Tensor, Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params, broadcast_ok=True)
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    if params.shape[1] != d_t * num_bins:
        raise ValueError("transform params have {} columns, expected {} x {}".format(params.shape[1], d_t, num_bins))
    y = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    status = torch.zeros((1,), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device), _cabi.launch("fc_linspline_apply", x.device):
        rc = L.fc_linspline_apply(xp, xs, pp, ps, y.data_ptr(), y.shape[1], lad.data_ptr(), 0, B, d_t,
                                  _cabi.cols(tcols), _cabi.cols(ccols), int(num_bins), int(tails), left, right, bottom,
                                  top, int(inverse), status.data_ptr(), _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_linspline_apply")
    return y, lad, status


@linspline_layer.register_fake
def _(x, params, tcols, ccols, num_bins, tails, inverse, left, right, bottom, top):
    return torch.empty_like(x), x.new_empty((x.shape[0],)), x.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("flowcon_b200::linspline_layer_backward", mutates_args=())
def linspline_layer_backward(x: Tensor, params: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                             tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int, tails: int, inverse: bool,
                             left: float, right: float, bottom: float, top: float) -> Tuple[Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    gx = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    gp = torch.empty((B, params.shape[1]), dtype=x.dtype, device=x.device)
    if tcols is not None and (tcols.numel() + (ccols.numel() if ccols is not None else 0)) < x.shape[1]:
        gx.zero_()
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    with torch.cuda.device(x.device), _cabi.launch("fc_linspline_backward", x.device):
        rc = L.fc_linspline_backward(xp, xs, pp, ps, gyp, gys, glp, gx.data_ptr(), gx.shape[1], gp.data_ptr(),
                                     gp.shape[1], B, d_t, _cabi.cols(tcols), _cabi.cols(ccols), int(num_bins),
                                     int(tails), left, right, bottom, top, int(inverse), _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_linspline_backward")
    return gx, gp


@linspline_layer_backward.register_fake
def _(x, params, grad_y, grad_lad, tcols, ccols, *args):
    return torch.empty_like(x), torch.empty_like(params)


def _linspline_setup(ctx, inputs, output):
    x, params, tcols, ccols = inputs[:4]
    ctx.save_for_backward(x, params, tcols, ccols)
    ctx.hyper = inputs[4:]


def _linspline_backward(ctx, gy, gl, gstatus):
    x, params, tcols, ccols = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gp = linspline_layer_backward(x, params, gy, gl, tcols, ccols, *ctx.hyper)
    return (gx, gp) + (None,) * (2 + len(ctx.hyper))


linspline_layer.register_autograd(_linspline_backward, setup_context=_linspline_setup)


# ------------------------------------------------------------------------------------------------
# piecewise-quadratic spline layer
# ------------------------------------------------------------------------------------------------
def _quad_cfg(num_bins, tails, inverse, left, right, bottom, top, min_w, min_h, wh_scale):
    return _cabi.QuadSplineConfig(int(num_bins), int(tails), int(inverse), left, right, bottom, top, min_w, min_h,
                                  wh_scale)


@torch.library.custom_op("flowcon_b200::quadspline_layer", mutates_args=())
def quadspline_layer(x: Tensor, params: Tensor, tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int,
                     tails: int, inverse: bool, left: float, right: float, bottom: float, top: float,
                     min_bin_width: float, min_bin_height: float, wh_scale: float) -> Tuple[Tensor, Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params, broadcast_ok=True)
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    p_per = 2 * num_bins - 1 if tails == _cabi.TAILS_LINEAR else 2 * num_bins + 1
    if params.shape[1] != d_t * p_per:
        raise ValueError("transform params have {} columns, expected {} x {}".format(params.shape[1], d_t, p_per))
    if min_bin_width * num_bins > 1.0:  # quadratic.py:77-80
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    y = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    status = torch.zeros((1,), dtype=torch.int32, device=x.device)
    cfg = _quad_cfg(num_bins, tails, inverse, left, right, bottom, top, min_bin_width, min_bin_height, wh_scale)
    with torch.cuda.device(x.device), _cabi.launch("fc_quadspline_apply", x.device):
        rc = L.fc_quadspline_apply(xp, xs, pp, ps, y.data_ptr(), y.shape[1], lad.data_ptr(), 0, B, d_t,
                                   _cabi.cols(tcols), _cabi.cols(ccols), ctypes.byref(cfg), status.data_ptr(),
                                   _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_quadspline_apply")
    return y, lad, status


@quadspline_layer.register_fake
def _(x, params, tcols, ccols, num_bins, tails, inverse, left, right, bottom, top, min_bin_width, min_bin_height,
      wh_scale):
    return torch.empty_like(x), x.new_empty((x.shape[0],)), x.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("flowcon_b200::quadspline_layer_backward", mutates_args=())
def quadspline_layer_backward(x: Tensor, params: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                              tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int, tails: int,
                              inverse: bool, left: float, right: float, bottom: float, top: float,
                              min_bin_width: float, min_bin_height: float, wh_scale: float) -> Tuple[Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    gx = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    gp = torch.empty((B, params.shape[1]), dtype=x.dtype, device=x.device)
    if tcols is not None and (tcols.numel() + (ccols.numel() if ccols is not None else 0)) < x.shape[1]:
        gx.zero_()
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    cfg = _quad_cfg(num_bins, tails, inverse, left, right, bottom, top, min_bin_width, min_bin_height, wh_scale)
    with torch.cuda.device(x.device), _cabi.launch("fc_quadspline_backward", x.device):
        rc = L.fc_quadspline_backward(xp, xs, pp, ps, gyp, gys, glp, gx.data_ptr(), gx.shape[1], gp.data_ptr(),
                                      gp.shape[1], B, d_t, _cabi.cols(tcols), _cabi.cols(ccols), ctypes.byref(cfg),
                                      _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_quadspline_backward")
    return gx, gp


@quadspline_layer_backward.register_fake
def _(x, params, grad_y, grad_lad, tcols, ccols, *args):
    return torch.empty_like(x), torch.empty_like(params)


def _quadspline_setup(ctx, inputs, output):
    x, params, tcols, ccols = inputs[:4]
    ctx.save_for_backward(x, params, tcols, ccols)
    ctx.hyper = inputs[4:]


def _quadspline_backward(ctx, gy, gl, gstatus):
    x, params, tcols, ccols = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gp = quadspline_layer_backward(x, params, gy, gl, tcols, ccols, *ctx.hyper)
    return (gx, gp) + (None,) * (2 + len(ctx.hyper))


quadspline_layer.register_autograd(_quadspline_backward, setup_context=_quadspline_setup)


# ------------------------------------------------------------------------------------------------
# cubic spline layer (same configuration struct as the quadratic layer)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::cubicspline_layer", mutates_args=())
def cubicspline_layer(x: Tensor, params: Tensor, tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int,
                     tails: int, inverse: bool, left: float, right: float, bottom: float, top: float,
                     min_bin_width: float, min_bin_height: float, wh_scale: float) -> Tuple[Tensor, Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params, broadcast_ok=True)
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    p_per = 2 * num_bins + 2
    if params.shape[1] != d_t * p_per:
        raise ValueError("transform params have {} columns, expected {} x {}".format(params.shape[1], d_t, p_per))
    if min_bin_width * num_bins > 1.0:  # cubic.py:89-92
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    y = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    status = torch.zeros((1,), dtype=torch.int32, device=x.device)
    cfg = _quad_cfg(num_bins, tails, inverse, left, right, bottom, top, min_bin_width, min_bin_height, wh_scale)
    with torch.cuda.device(x.device), _cabi.launch("fc_cubicspline_apply", x.device):
        rc = L.fc_cubicspline_apply(xp, xs, pp, ps, y.data_ptr(), y.shape[1], lad.data_ptr(), 0, B, d_t,
                                   _cabi.cols(tcols), _cabi.cols(ccols), ctypes.byref(cfg), status.data_ptr(),
                                   _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_cubicspline_apply")
    return y, lad, status


@cubicspline_layer.register_fake
def _(x, params, tcols, ccols, num_bins, tails, inverse, left, right, bottom, top, min_bin_width, min_bin_height,
      wh_scale):
    return torch.empty_like(x), x.new_empty((x.shape[0],)), x.new_empty((1,), dtype=torch.int32)


@torch.library.custom_op("flowcon_b200::cubicspline_layer_backward", mutates_args=())
def cubicspline_layer_backward(x: Tensor, params: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                              tcols: Optional[Tensor], ccols: Optional[Tensor], num_bins: int, tails: int,
                              inverse: bool, left: float, right: float, bottom: float, top: float,
                              min_bin_width: float, min_bin_height: float, wh_scale: float) -> Tuple[Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    gx = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    gp = torch.empty((B, params.shape[1]), dtype=x.dtype, device=x.device)
    if tcols is not None and (tcols.numel() + (ccols.numel() if ccols is not None else 0)) < x.shape[1]:
        gx.zero_()
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    cfg = _quad_cfg(num_bins, tails, inverse, left, right, bottom, top, min_bin_width, min_bin_height, wh_scale)
    with torch.cuda.device(x.device), _cabi.launch("fc_cubicspline_backward", x.device):
        rc = L.fc_cubicspline_backward(xp, xs, pp, ps, gyp, gys, glp, gx.data_ptr(), gx.shape[1], gp.data_ptr(),
                                      gp.shape[1], B, d_t, _cabi.cols(tcols), _cabi.cols(ccols), ctypes.byref(cfg),
                                      _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_cubicspline_backward")
    return gx, gp


@cubicspline_layer_backward.register_fake
def _(x, params, grad_y, grad_lad, tcols, ccols, *args):
    return torch.empty_like(x), torch.empty_like(params)


def _cubicspline_setup(ctx, inputs, output):
    x, params, tcols, ccols = inputs[:4]
    ctx.save_for_backward(x, params, tcols, ccols)
    ctx.hyper = inputs[4:]


def _cubicspline_backward(ctx, gy, gl, gstatus):
    x, params, tcols, ccols = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gp = cubicspline_layer_backward(x, params, gy, gl, tcols, ccols, *ctx.hyper)
    return (gx, gp) + (None,) * (2 + len(ctx.hyper))


cubicspline_layer.register_autograd(_cubicspline_backward, setup_context=_cubicspline_setup)


# ------------------------------------------------------------------------------------------------
# affine layer
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::affine_layer", mutates_args=())
def affine_layer(x: Tensor, params: Tensor, tcols: Optional[Tensor], ccols: Optional[Tensor], layout: int,
                 activation: int, inverse: bool) -> Tuple[Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params, broadcast_ok=True)
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    if params.shape[1] != 2 * d_t:
        raise ValueError("affine params have {} columns, expected {}".format(params.shape[1], 2 * d_t))
    y = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device), _cabi.launch("fc_affine_apply", x.device):
        rc = L.fc_affine_apply(xp, xs, pp, ps, y.data_ptr(), y.shape[1], lad.data_ptr(), 0, B, d_t,
                               _cabi.cols(tcols), _cabi.cols(ccols), layout, activation, int(inverse),
                               _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_affine_apply")
    return y, lad


@affine_layer.register_fake
def _(x, params, tcols, ccols, layout, activation, inverse):
    return torch.empty_like(x), x.new_empty((x.shape[0],))


@torch.library.custom_op("flowcon_b200::affine_layer_backward", mutates_args=())
def affine_layer_backward(x: Tensor, params: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                          tcols: Optional[Tensor], ccols: Optional[Tensor], layout: int, activation: int,
                          inverse: bool) -> Tuple[Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B = x.shape[0]
    d_t = tcols.numel() if tcols is not None else x.shape[1]
    gx = torch.empty((B, x.shape[1]), dtype=x.dtype, device=x.device)
    gp = torch.empty((B, params.shape[1]), dtype=x.dtype, device=x.device)
    if tcols is not None and (tcols.numel() + (ccols.numel() if ccols is not None else 0)) < x.shape[1]:
        gx.zero_()
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    with torch.cuda.device(x.device), _cabi.launch("fc_affine_backward", x.device):
        rc = L.fc_affine_backward(xp, xs, pp, ps, gyp, gys, glp, gx.data_ptr(), gx.shape[1], gp.data_ptr(),
                                  gp.shape[1], B, d_t, _cabi.cols(tcols), _cabi.cols(ccols), layout, activation,
                                  int(inverse), _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_affine_backward")
    return gx, gp


@affine_layer_backward.register_fake
def _(x, params, grad_y, grad_lad, tcols, ccols, layout, activation, inverse):
    return torch.empty_like(x), torch.empty_like(params)


def _affine_setup(ctx, inputs, output):
    x, params, tcols, ccols = inputs[:4]
    ctx.save_for_backward(x, params, tcols, ccols)
    ctx.hyper = inputs[4:]


def _affine_backward(ctx, gy, gl):
    x, params, tcols, ccols = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gp = affine_layer_backward(x, params, gy, gl, tcols, ccols, *ctx.hyper)
    return (gx, gp) + (None,) * (2 + len(ctx.hyper))


affine_layer.register_autograd(_affine_backward, setup_context=_affine_setup)


# ------------------------------------------------------------------------------------------------
# activation normalisation (per-feature affine, parameters shared by the batch)
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::actnorm_layer", mutates_args=())
def actnorm_layer(x: Tensor, log_scale: Tensor, shift: Tensor, inverse: bool) -> Tuple[Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(log_scale, "log_scale")
    _cabi.require_cuda_f32(shift, "shift")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    B, D = x.shape
    if log_scale.numel() != D or shift.numel() != D:
        raise ValueError("log_scale / shift must have {} entries".format(D))
    log_scale, shift = log_scale.contiguous(), shift.contiguous()
    y = torch.empty((B, D), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device), _cabi.launch("fc_actnorm_apply", x.device):
        rc = L.fc_actnorm_apply(xp, xs, log_scale.data_ptr(), shift.data_ptr(), y.data_ptr(), D, lad.data_ptr(), 0, B, D,
                                int(inverse), _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_actnorm_apply")
    return y, lad


@actnorm_layer.register_fake
def _(x, log_scale, shift, inverse):
    return torch.empty_like(x), x.new_empty((x.shape[0],))


@torch.library.custom_op("flowcon_b200::actnorm_layer_backward", mutates_args=())
def actnorm_layer_backward(x: Tensor, log_scale: Tensor, shift: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                           inverse: bool) -> Tuple[Tensor, Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B, D = x.shape
    log_scale, shift = log_scale.contiguous(), shift.contiguous()
    gx = torch.empty((B, D), dtype=x.dtype, device=x.device)
    gls = torch.empty((D,), dtype=x.dtype, device=x.device)
    gsh = torch.empty((D,), dtype=x.dtype, device=x.device)
    ws = torch.empty((max(int(L.fc_actnorm_workspace_floats(B, D)), 1),), dtype=x.dtype, device=x.device)
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    with torch.cuda.device(x.device), _cabi.launch("fc_actnorm_backward", x.device):
        rc = L.fc_actnorm_backward(xp, xs, log_scale.data_ptr(), shift.data_ptr(), gyp, gys, glp, gx.data_ptr(), D,
                                   gls.data_ptr(), gsh.data_ptr(), ws.data_ptr(), B, D, int(inverse),
                                   _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_actnorm_backward")
    return gx, gls, gsh


@actnorm_layer_backward.register_fake
def _(x, log_scale, shift, grad_y, grad_lad, inverse):
    return torch.empty_like(x), torch.empty_like(log_scale), torch.empty_like(shift)


def _actnorm_setup(ctx, inputs, output):
    x, log_scale, shift, inverse = inputs
    ctx.save_for_backward(x, log_scale, shift)
    ctx.inverse = inverse


def _actnorm_backward(ctx, gy, gl):
    x, log_scale, shift = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gls, gsh = actnorm_layer_backward(x, log_scale, shift, gy, gl, ctx.inverse)
    return gx, gls.reshape(log_scale.shape), gsh.reshape(shift.shape), None


actnorm_layer.register_autograd(_actnorm_backward, setup_context=_actnorm_setup)


# ------------------------------------------------------------------------------------------------
# sum-of-sigmoids layer
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::sos_layer", mutates_args=())
def sos_layer(x: Tensor, params: Tensor, n_sigmoids: int, offset: float, inverse: bool, bisection_iterations: int,
              lim: float) -> Tuple[Tensor, Tensor]:
    _cabi.require_cuda_f32(x, "inputs")
    _cabi.require_cuda_f32(params, "transform params")
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params, broadcast_ok=True)
    B, D = x.shape
    if params.shape[1] != D * (3 * n_sigmoids + 1):
        raise ValueError("sum-of-sigmoids params have {} columns, expected {}".format(
            params.shape[1], D * (3 * n_sigmoids + 1)))
    y = torch.empty((B, D), dtype=x.dtype, device=x.device)
    lad = torch.empty((B,), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device), _cabi.launch("fc_sos_apply", x.device):
        rc = L.fc_sos_apply(xp, xs, pp, ps, y.data_ptr(), D, lad.data_ptr(), 0, B, D, n_sigmoids, offset,
                            int(inverse), bisection_iterations, lim, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_sos_apply")
    return y, lad


@sos_layer.register_fake
def _(x, params, n_sigmoids, offset, inverse, bisection_iterations, lim):
    return torch.empty_like(x), x.new_empty((x.shape[0],))


@torch.library.custom_op("flowcon_b200::sos_layer_backward", mutates_args=())
def sos_layer_backward(x: Tensor, params: Tensor, grad_y: Tensor, grad_lad: Optional[Tensor],
                       n_sigmoids: int) -> Tuple[Tensor, Tensor]:
    L = _cabi.lib()
    x, xp, xs = _cabi.rows(x)
    params, pp, ps = _cabi.rows(params)
    grad_y, gyp, gys = _cabi.rows(_cabi.require_cuda_f32(grad_y, "grad outputs"))
    B, D = x.shape
    gx = torch.empty((B, D), dtype=x.dtype, device=x.device)
    gp = torch.empty((B, params.shape[1]), dtype=x.dtype, device=x.device)
    glp = None
    if grad_lad is not None:
        grad_lad = grad_lad.contiguous()
        glp = grad_lad.data_ptr()
    with torch.cuda.device(x.device), _cabi.launch("fc_sos_backward", x.device):
        rc = L.fc_sos_backward(xp, xs, pp, ps, gyp, gys, glp, gx.data_ptr(), D, gp.data_ptr(), gp.shape[1], B, D,
                               n_sigmoids, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "fc_sos_backward")
    return gx, gp


@sos_layer_backward.register_fake
def _(x, params, grad_y, grad_lad, n_sigmoids):
    return torch.empty_like(x), torch.empty_like(params)


def _sos_setup(ctx, inputs, output):
    x, params, n_sigmoids, offset, inverse = inputs[:5]
    ctx.save_for_backward(x, params)
    ctx.n_sigmoids = n_sigmoids
    ctx.inverse = inverse


def _sos_backward(ctx, gy, gl):
    if ctx.inverse:
        raise NotImplementedError("gradients through the numerical sum-of-sigmoids inverse are not implemented")
    x, params = ctx.saved_tensors
    if gy is None:
        gy = torch.zeros_like(x)
    gx, gp = sos_layer_backward(x, params, gy, gl, ctx.n_sigmoids)
    return gx, gp, None, None, None, None, None


sos_layer.register_autograd(_sos_backward, setup_context=_sos_setup)


# ------------------------------------------------------------------------------------------------
# standard-normal tail
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("flowcon_b200::stdnormal_log_prob", mutates_args=())
def stdnormal_log_prob(z: Tensor, logabsdet: Optional[Tensor]) -> Tensor:
    _cabi.require_cuda_f32(z, "noise")
    L = _cabi.lib()
    z, zp, zs = _cabi.rows(z)
    B, D = z.shape
    out = torch.empty((B,), dtype=z.dtype, device=z.device)
    ladp = None
    if logabsdet is not None:
        logabsdet = _cabi.require_cuda_f32(logabsdet, "logabsdet").contiguous()
        ladp = logabsdet.data_ptr()
    with torch.cuda.device(z.device), _cabi.launch("fc_stdnormal_log_prob", z.device):
        rc = L.fc_stdnormal_log_prob(zp, zs, ladp, out.data_ptr(), B, D, _cabi.stream_ptr(z.device))
    _cabi.check(rc, "fc_stdnormal_log_prob")
    return out


@stdnormal_log_prob.register_fake
def _(z, logabsdet):
    return z.new_empty((z.shape[0],))


def _sn_setup(ctx, inputs, output):
    z, lad = inputs
    ctx.save_for_backward(z)
    ctx.has_lad = lad is not None


def _sn_backward(ctx, g):
    (z,) = ctx.saved_tensors
    return -z * g[:, None], (g if ctx.has_lad else None)


stdnormal_log_prob.register_autograd(_sn_backward, setup_context=_sn_setup)
