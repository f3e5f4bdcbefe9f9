"""Function-level seam: `rational_quadratic_spline` / `unconstrained_rational_quadratic_spline` with the
signatures of flowcon/transforms/splines/rational_quadratic.py:13-25,66-80, backed by the CUDA kernels.

Inputs of any shape [...]; raw parameters [..., K], [..., K], [..., K-1 | K+1]; returns
(outputs [...], per-element logabsdet [...]).  These are what `flowconductor_b200.patch_reference()`
installs into the reference's modules.  The layer classes do NOT go through here: they hand the
conditioner's [B, D_t*P] output to the kernel directly (no concatenation, column split fused).
"""
import math

import torch

from .. import _cabi, ops
from .base import InputOutsideDomain

DEFAULT_MIN_BIN_WIDTH = 1e-3
DEFAULT_MIN_BIN_HEIGHT = 1e-3
DEFAULT_MIN_DERIVATIVE = 1e-3

# when True, every spline call checks the device status word (one host sync, like the reference's own
# checks at rational_quadratic.py:81-82,142) even with linear tails
STRICT = False


def check_status(status, tails):
    """Turn the device status word into the reference's exceptions."""
    if tails == _cabi.TAILS_NONE or STRICT:
        if status.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("a spline without linear tails checks its inputs' domain on the host (the reference "
                               "raises InputOutsideDomain, rational_quadratic.py:81-82 / linear.py:45-46): that "
                               "synchronisation cannot be recorded into a CUDA graph — use tails='linear'")
        code = int(status.item())
        if code & _cabi.STATUS_INPUT_OUTSIDE_DOMAIN:
            raise InputOutsideDomain()
        if code & _cabi.STATUS_NEGATIVE_DISCRIMINANT:
            raise AssertionError("negative discriminant in the rational-quadratic inverse")


def _elementwise(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives, inverse, tails, left,
                 right, bottom, top, min_bin_width, min_bin_height, min_derivative, enable_identity_init):
    num_bins = unnormalized_widths.shape[-1]
    if min_bin_width * num_bins > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * num_bins > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")
    shape = inputs.shape
    params = torch.cat((unnormalized_widths, unnormalized_heights, unnormalized_derivatives), dim=-1)
    params = params.reshape(-1, params.shape[-1])
    # every element is its own row (D_t = 1), so the kernel's per-row log-det IS the per-element one
    y, lad, status = ops.rqs_layer(inputs.reshape(-1, 1), params, None, None, num_bins, tails, bool(inverse),
                                   bool(enable_identity_init), float(left), float(right), float(bottom), float(top),
                                   float(min_bin_width), float(min_bin_height), float(min_derivative), 1.0)
    check_status(status, tails)
    return y.reshape(shape), lad.reshape(shape)


def rational_quadratic_spline(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives,
                              inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0,
                              min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
                              min_derivative=DEFAULT_MIN_DERIVATIVE, enable_identity_init=False):
    return _elementwise(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives, inverse,
                        _cabi.TAILS_NONE, left, right, bottom, top, min_bin_width, min_bin_height, min_derivative,
                        enable_identity_init)


def unconstrained_rational_quadratic_spline(inputs, unnormalized_widths, unnormalized_heights,
                                            unnormalized_derivatives, inverse=False, tails="linear", tail_bound=1.0,
                                            min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                                            min_bin_height=DEFAULT_MIN_BIN_HEIGHT,
                                            min_derivative=DEFAULT_MIN_DERIVATIVE, enable_identity_init=False):
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    num_bins = unnormalized_widths.shape[-1]
    if unnormalized_derivatives.shape[-1] != num_bins - 1:
        # The reference pads WHATEVER it is given with the boundary constant on both sides and then reads knot
        # derivatives 0..K of the padded vector (rational_quadratic.py:33-36,120-130) — its own unit tests pass K + 1 raw
        # derivatives (tests/transforms/splines/rational_quadratic_test.py:71), in which case the right boundary knot takes
        # a raw value, not the constant.  Same result here: the constrained kernel on [-tail_bound, tail_bound] with the
        # first K + 1 entries of the padded vector, identity outside.
        constant = math.log(math.exp(1 - min_derivative) - 1)
        padded = torch.nn.functional.pad(unnormalized_derivatives, (1, 1), value=constant)
        if padded.shape[-1] < num_bins + 1:
            raise RuntimeError("unnormalized_derivatives needs at least num_bins - 1 entries")
        inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
        y, lad = _elementwise(inputs.clamp(-tail_bound, tail_bound), unnormalized_widths, unnormalized_heights,
                              padded[..., : num_bins + 1], inverse, _cabi.TAILS_NONE, -tail_bound, tail_bound, -tail_bound,
                              tail_bound, min_bin_width, min_bin_height, min_derivative, enable_identity_init)
        return torch.where(inside, y, inputs), torch.where(inside, lad, torch.zeros_like(lad))
    return _elementwise(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives, inverse,
                        _cabi.TAILS_LINEAR, -tail_bound, tail_bound, -tail_bound, tail_bound, min_bin_width,
                        min_bin_height, min_derivative, enable_identity_init)


class RationalQuadraticSettings:
    """The spline hyper-parameters a layer class carries, and the one call that applies them to a
    conditioner output.  `constrained_bound` is the box used when tails is None: 1.0 -> [0,1] for couplings
    (coupling.py:566-567), 1.2 -> [-1.2,1.2] for autoregressive / conditional layers (autoregressive.py:595)."""

    def __init__(self, num_bins, tails, tail_bound, min_bin_width, min_bin_height, min_derivative,
                 identity_init, constrained_box):
        if tails not in (None, "linear"):
            raise ValueError(tails)
        self.num_bins = num_bins
        self.tails = tails
        self.tail_bound = tail_bound
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.min_derivative = min_derivative
        self.identity_init = identity_init
        self.constrained_box = constrained_box

    def params_per_feature(self):
        return self.num_bins * 3 - 1 if self.tails == "linear" else self.num_bins * 3 + 1

    def config(self, inverse, hidden_for_scaling):
        """(struct fc_rqs_config, tails constant) of these settings."""
        if self.tails == "linear":
            tails, lo, hi = _cabi.TAILS_LINEAR, -self.tail_bound, self.tail_bound
        else:
            tails, (lo, hi) = _cabi.TAILS_NONE, self.constrained_box
        wh_scale = 1.0 / math.sqrt(hidden_for_scaling) if hidden_for_scaling else 1.0
        cfg = _cabi.RqsConfig(int(self.num_bins), tails, int(bool(self.identity_init)), int(bool(inverse)), float(lo),
                              float(hi), float(lo), float(hi), float(self.min_bin_width), float(self.min_bin_height),
                              float(self.min_derivative), float(wh_scale))
        return cfg, tails

    def apply(self, inputs, params, tcols, ccols, inverse, hidden_for_scaling):
        if self.tails == "linear":
            tails, lo, hi = _cabi.TAILS_LINEAR, -self.tail_bound, self.tail_bound
        else:
            tails, (lo, hi) = _cabi.TAILS_NONE, self.constrained_box
        wh_scale = 1.0 / math.sqrt(hidden_for_scaling) if hidden_for_scaling else 1.0
        y, lad, status = ops.rqs_layer(inputs, params, tcols, ccols, self.num_bins, tails, bool(inverse),
                                       bool(self.identity_init), float(lo), float(hi), float(lo), float(hi),
                                       float(self.min_bin_width), float(self.min_bin_height),
                                       float(self.min_derivative), float(wh_scale))
        check_status(status, tails)
        return y, lad


# ------------------------------------------------------------------------------------------------
# piecewise-linear spline (flowcon/transforms/splines/linear.py)
# ------------------------------------------------------------------------------------------------
class LinearSplineSettings:
    """Hyper-parameters of a piecewise-linear spline layer and the one kernel call that applies them to a
    [B, D_t * num_bins] parameter tensor (PiecewiseLinearCouplingTransform._piecewise_cdf, coupling.py:340-352)."""

    def __init__(self, num_bins, tails, tail_bound):
        if tails not in (None, "linear"):
            raise RuntimeError("{} tails are not implemented.".format(tails))  # linear.py:22
        self.num_bins = num_bins
        self.tails = tails
        self.tail_bound = tail_bound

    def params_per_feature(self):
        return self.num_bins

    def domain(self):
        """(tails constant, lower, upper bound) of these settings."""
        if self.tails == "linear":
            return _cabi.TAILS_LINEAR, -float(self.tail_bound), float(self.tail_bound)
        return _cabi.TAILS_NONE, 0.0, 1.0

    def apply(self, inputs, params, tcols, ccols, inverse):
        if self.tails == "linear":
            tails, lo, hi = _cabi.TAILS_LINEAR, -float(self.tail_bound), float(self.tail_bound)
        else:
            tails, lo, hi = _cabi.TAILS_NONE, 0.0, 1.0
        y, lad, status = ops.linspline_layer(inputs, params, tcols, ccols, int(self.num_bins), tails, bool(inverse),
                                             lo, hi, lo, hi)
        check_status(status, tails)
        return y, lad


def _linear_elementwise(inputs, unnormalized_pdf, inverse, tails, left, right, bottom, top):
    shape = inputs.shape
    params = unnormalized_pdf.reshape(-1, unnormalized_pdf.shape[-1])
    # every element is its own row (D_t = 1), so the kernel's per-row log-det IS the per-element one
    y, lad, status = ops.linspline_layer(inputs.reshape(-1, 1), params, None, None, params.shape[-1], tails,
                                         bool(inverse), float(left), float(right), float(bottom), float(top))
    check_status(status, tails)
    return y.reshape(shape), lad.reshape(shape)


def linear_spline(inputs, unnormalized_pdf, inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0):
    """flowcon/transforms/splines/linear.py:38-105 (per-element outputs and log-dets; raises InputOutsideDomain)."""
    return _linear_elementwise(inputs, unnormalized_pdf, inverse, _cabi.TAILS_NONE, left, right, bottom, top)


def unconstrained_linear_spline(inputs, unnormalized_pdf, inverse=False, tail_bound=1.0, tails="linear"):
    """linear.py:9-35: identity outside [-tail_bound, tail_bound]."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    return _linear_elementwise(inputs, unnormalized_pdf, inverse, _cabi.TAILS_LINEAR, -tail_bound, tail_bound,
                               -tail_bound, tail_bound)


def _quad_config(settings, inverse, hidden_for_scaling):
    if settings.tails == "linear":
        tails, lo, hi = _cabi.TAILS_LINEAR, -float(settings.tail_bound), float(settings.tail_bound)
    else:
        tails, lo, hi = _cabi.TAILS_NONE, 0.0, 1.0
    wh_scale = 1.0 / math.sqrt(hidden_for_scaling) if hidden_for_scaling else 1.0
    cfg = _cabi.QuadSplineConfig(int(settings.num_bins), tails, int(bool(inverse)), lo, hi, lo, hi,
                                 float(settings.min_bin_width), float(settings.min_bin_height), float(wh_scale))
    return cfg, tails


# ------------------------------------------------------------------------------------------------
# piecewise-quadratic spline (flowcon/transforms/splines/quadratic.py)
# ------------------------------------------------------------------------------------------------
class QuadraticSplineSettings:
    """Hyper-parameters of a piecewise-quadratic spline layer and the kernel call that applies them to a
    [B, D_t * P] parameter tensor, P = 2K+1 (no tails) / 2K-1 (linear tails): per feature [K raw widths ; raw heights]
    (PiecewiseQuadraticCouplingTransform._piecewise_cdf, coupling.py:403-427)."""

    def __init__(self, num_bins, tails, tail_bound, min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=DEFAULT_MIN_BIN_HEIGHT):
        if tails not in (None, "linear"):
            raise RuntimeError("{} tails are not implemented.".format(tails))  # quadratic.py:36
        self.num_bins = num_bins
        self.tails = tails
        self.tail_bound = tail_bound
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height

    def params_per_feature(self):
        return self.num_bins * 2 - 1 if self.tails == "linear" else self.num_bins * 2 + 1

    def config(self, inverse, hidden_for_scaling):
        """(struct fc_quadspline_config, tails constant) of these settings."""
        return _quad_config(self, inverse, hidden_for_scaling)

    def apply(self, inputs, params, tcols, ccols, inverse, hidden_for_scaling):
        if self.tails == "linear":
            tails, lo, hi = _cabi.TAILS_LINEAR, -float(self.tail_bound), float(self.tail_bound)
        else:
            tails, lo, hi = _cabi.TAILS_NONE, 0.0, 1.0
        wh_scale = 1.0 / math.sqrt(hidden_for_scaling) if hidden_for_scaling else 1.0
        y, lad, status = ops.quadspline_layer(inputs, params, tcols, ccols, int(self.num_bins), tails, bool(inverse),
                                              lo, hi, lo, hi, float(self.min_bin_width), float(self.min_bin_height),
                                              float(wh_scale))
        check_status(status, tails)
        return y, lad


def _quadratic_elementwise(inputs, unnormalized_widths, unnormalized_heights, inverse, tails, left, right, bottom, top,
                           min_bin_width, min_bin_height):
    shape = inputs.shape
    num_bins = unnormalized_widths.shape[-1]
    params = torch.cat((unnormalized_widths, unnormalized_heights), dim=-1)
    params = params.reshape(-1, params.shape[-1])
    y, lad, status = ops.quadspline_layer(inputs.reshape(-1, 1), params, None, None, num_bins, tails, bool(inverse),
                                          float(left), float(right), float(bottom), float(top), float(min_bin_width),
                                          float(min_bin_height), 1.0)
    check_status(status, tails)
    return y.reshape(shape), lad.reshape(shape)


def quadratic_spline(inputs, unnormalized_widths, unnormalized_heights, inverse=False, left=0.0, right=1.0, bottom=0.0,
                     top=1.0, min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT):
    """flowcon/transforms/splines/quadratic.py:55-159 (per-element outputs and log-dets)."""
    if unnormalized_heights.shape[-1] != unnormalized_widths.shape[-1] + 1:
        raise ValueError("quadratic_spline expects num_bins + 1 raw heights (use unconstrained_quadratic_spline for the "
                         "num_bins - 1 form)")
    return _quadratic_elementwise(inputs, unnormalized_widths, unnormalized_heights, inverse, _cabi.TAILS_NONE, left,
                                  right, bottom, top, min_bin_width, min_bin_height)


def unconstrained_quadratic_spline(inputs, unnormalized_widths, unnormalized_heights, inverse=False, tail_bound=1.0,
                                   tails="linear", min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                                   min_bin_height=DEFAULT_MIN_BIN_HEIGHT):
    """quadratic.py:11-52: identity outside [-tail_bound, tail_bound]; num_bins - 1 raw heights."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    assert unnormalized_heights.shape[-1] == unnormalized_widths.shape[-1] - 1  # quadratic.py:34
    return _quadratic_elementwise(inputs, unnormalized_widths, unnormalized_heights, inverse, _cabi.TAILS_LINEAR,
                                  -tail_bound, tail_bound, -tail_bound, tail_bound, min_bin_width, min_bin_height)


# ------------------------------------------------------------------------------------------------
# cubic spline (flowcon/transforms/splines/cubic.py)
# ------------------------------------------------------------------------------------------------
class CubicSplineSettings:
    """Hyper-parameters of a cubic-spline layer and the kernel call that applies them to a [B, D_t * (2K+2)] parameter
    tensor: per feature [K raw widths ; K raw heights ; raw left derivative ; raw right derivative]
    (PiecewiseCubicCouplingTransform._piecewise_cdf, coupling.py:468-500)."""

    def __init__(self, num_bins, tails, tail_bound, min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=DEFAULT_MIN_BIN_HEIGHT):
        if tails not in (None, "linear"):
            raise RuntimeError("{} tails are not implemented.".format(tails))  # cubic.py:40
        self.num_bins = num_bins
        self.tails = tails
        self.tail_bound = tail_bound
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height

    def params_per_feature(self):
        return self.num_bins * 2 + 2

    def config(self, inverse, hidden_for_scaling):
        """(struct fc_quadspline_config, tails constant) of these settings (the cubic family shares the struct)."""
        return _quad_config(self, inverse, hidden_for_scaling)

    def apply(self, inputs, params, tcols, ccols, inverse, hidden_for_scaling):
        if self.tails == "linear":
            tails, lo, hi = _cabi.TAILS_LINEAR, -float(self.tail_bound), float(self.tail_bound)
        else:
            tails, lo, hi = _cabi.TAILS_NONE, 0.0, 1.0
        wh_scale = 1.0 / math.sqrt(hidden_for_scaling) if hidden_for_scaling else 1.0
        y, lad, status = ops.cubicspline_layer(inputs, params, tcols, ccols, int(self.num_bins), tails, bool(inverse),
                                               lo, hi, lo, hi, float(self.min_bin_width), float(self.min_bin_height),
                                               float(wh_scale))
        check_status(status, tails)
        return y, lad


def _cubic_elementwise(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left,
                       unnorm_derivatives_right, inverse, tails, left, right, bottom, top, min_bin_width, min_bin_height):
    shape = inputs.shape
    num_bins = unnormalized_widths.shape[-1]
    params = torch.cat((unnormalized_widths, unnormalized_heights, unnorm_derivatives_left, unnorm_derivatives_right),
                       dim=-1)
    params = params.reshape(-1, params.shape[-1])
    y, lad, status = ops.cubicspline_layer(inputs.reshape(-1, 1), params, None, None, num_bins, tails, bool(inverse),
                                           float(left), float(right), float(bottom), float(top), float(min_bin_width),
                                           float(min_bin_height), 1.0)
    check_status(status, tails)
    return y.reshape(shape), lad.reshape(shape)


def cubic_spline(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left, unnorm_derivatives_right,
                 inverse=False, left=0.0, right=1.0, bottom=0.0, top=1.0, min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=DEFAULT_MIN_BIN_HEIGHT, eps=1e-5, quadratic_threshold=1e-3):
    """flowcon/transforms/splines/cubic.py:63-267 (per-element outputs and log-dets).  `eps` and `quadratic_threshold`
    steer the reference's closed-form root selection; the kernel's Newton inverse does not need them."""
    return _cubic_elementwise(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left,
                              unnorm_derivatives_right, inverse, _cabi.TAILS_NONE, left, right, bottom, top,
                              min_bin_width, min_bin_height)


def unconstrained_cubic_spline(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left,
                               unnorm_derivatives_right, inverse=False, tail_bound=1.0, tails="linear",
                               min_bin_width=DEFAULT_MIN_BIN_WIDTH, min_bin_height=DEFAULT_MIN_BIN_HEIGHT, eps=1e-5,
                               quadratic_threshold=1e-3):
    """cubic.py:15-60: identity outside [-tail_bound, tail_bound]."""
    if tails != "linear":
        raise RuntimeError("{} tails are not implemented.".format(tails))
    return _cubic_elementwise(inputs, unnormalized_widths, unnormalized_heights, unnorm_derivatives_left,
                              unnorm_derivatives_right, inverse, _cabi.TAILS_LINEAR, -tail_bound, tail_bound,
                              -tail_bound, tail_bound, min_bin_width, min_bin_height)
