"""Unconditional element-wise CDF layers on the hot path: `PiecewiseRationalQuadraticCDF`
(flowcon/transforms/nonlinearities.py:406-487) — a rational-quadratic spline whose parameters are learnable tensors
shared across the batch (`_share_across_batch` :246-247) instead of a conditioner output.  Same constructor, same
parameter names and shapes (`unnormalized_widths` [*shape, K], `unnormalized_heights` [*shape, K],
`unnormalized_derivatives` [*shape, K-1 | K+1]), so reference state_dicts load.  It is what
`PiecewiseRationalQuadraticCouplingTransform(apply_unconditional_transform=True)` puts on the identity features
(coupling.py:524-535).

The evaluation is the same kernel as every other RQ layer (`fc_rqs_apply` / `fc_rqs_backward` through
`splines.RationalQuadraticSettings`): no 1/sqrt(H) scaling, identity-init off, [0, 1] box when tails is None."""
import numpy as np
import torch
from torch import nn

from . import splines
from .base import Transform

__all__ = ["PiecewiseCubicCDF", "PiecewiseLinearCDF", "PiecewiseQuadraticCDF", "PiecewiseRationalQuadraticCDF"]


class PiecewiseRationalQuadraticCDF(Transform):
    def __init__(self, shape, num_bins=10, tails=None, tail_bound=1.0, identity_init=False,
                 min_bin_width=splines.DEFAULT_MIN_BIN_WIDTH, min_bin_height=splines.DEFAULT_MIN_BIN_HEIGHT,
                 min_derivative=splines.DEFAULT_MIN_DERIVATIVE):
        super().__init__()
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.min_derivative = min_derivative
        self.tail_bound = tail_bound
        self.tails = tails
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        if len(shape) != 1:
            raise NotImplementedError("image-shaped CDF layers are outside the B200 hot path")
        num_derivatives = (num_bins - 1) if tails == "linear" else (num_bins + 1)
        if identity_init:  # nonlinearities.py:430-440
            constant = float(np.log(np.exp(1 - min_derivative) - 1))
            self.unnormalized_widths = nn.Parameter(torch.zeros(*shape, num_bins))
            self.unnormalized_heights = nn.Parameter(torch.zeros(*shape, num_bins))
            self.unnormalized_derivatives = nn.Parameter(constant * torch.ones(*shape, num_derivatives))
        else:  # :441-450 (same draw order as the reference: widths, heights, derivatives)
            self.unnormalized_widths = nn.Parameter(torch.rand(*shape, num_bins))
            self.unnormalized_heights = nn.Parameter(torch.rand(*shape, num_bins))
            self.unnormalized_derivatives = nn.Parameter(torch.rand(*shape, num_derivatives))
        # the reference never enables the identity-init boundary rule for this layer (it calls the spline functions
        # with their default enable_identity_init=False, :465-477)
        self._spline = splines.RationalQuadraticSettings(num_bins, tails, tail_bound, min_bin_width, min_bin_height,
                                                         min_derivative, identity_init=False,
                                                         constrained_box=(0.0, 1.0))

    def _shared_params(self, batch_size):
        """[D, P] per-feature blocks [w ; h ; d] -> one row, shared by every sample (expanded view)."""
        p = torch.cat((self.unnormalized_widths, self.unnormalized_heights, self.unnormalized_derivatives), dim=-1)
        return p.reshape(1, -1).expand(batch_size, -1)

    def apply_on_columns(self, inputs, tcols, ccols, inverse):
        """The layer applied to columns `tcols` of a wider tensor (the other columns, `ccols`, are copied): what a
        coupling layer needs for its identity features, without gathering them first."""
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), tcols, ccols, inverse, None)

    def _run(self, inputs, inverse):
        if inputs.dim() != 2 or inputs.shape[1] != self.unnormalized_widths.shape[0]:
            raise ValueError("expected inputs of shape [batch, {}]".format(self.unnormalized_widths.shape[0]))
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), None, None, inverse, None)

    def forward(self, inputs, context=None):
        return self._run(inputs, inverse=False)

    def inverse(self, inputs, context=None):
        return self._run(inputs, inverse=True)


class PiecewiseLinearCDF(Transform):
    """flowcon/transforms/nonlinearities.py:250-283: a piecewise-linear spline whose `unnormalized_pdf` [*shape, K] is a
    learnable tensor shared across the batch."""

    def __init__(self, shape, num_bins=10, tails=None, tail_bound=1.0):
        super().__init__()
        self.tail_bound = tail_bound
        self.tails = tails
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        if len(shape) != 1:
            raise NotImplementedError("image-shaped CDF layers are outside the B200 hot path")
        self.unnormalized_pdf = nn.Parameter(torch.randn(*shape, num_bins))
        self._spline = splines.LinearSplineSettings(num_bins, tails, tail_bound)

    def _shared_params(self, batch_size):
        return self.unnormalized_pdf.reshape(1, -1).expand(batch_size, -1)

    def apply_on_columns(self, inputs, tcols, ccols, inverse):
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), tcols, ccols, inverse)

    def _run(self, inputs, inverse):
        if inputs.dim() != 2 or inputs.shape[1] != self.unnormalized_pdf.shape[0]:
            raise ValueError("expected inputs of shape [batch, {}]".format(self.unnormalized_pdf.shape[0]))
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), None, None, inverse)

    def forward(self, inputs, context=None):
        return self._run(inputs, inverse=False)

    def inverse(self, inputs, context=None):
        return self._run(inputs, inverse=True)


class PiecewiseQuadraticCDF(Transform):
    """flowcon/transforms/nonlinearities.py:286-340: learnable `unnormalized_widths` [*shape, K] and
    `unnormalized_heights` [*shape, K+1] (K-1 with linear tails), shared across the batch."""

    def __init__(self, shape, num_bins=10, tails=None, tail_bound=1.0, min_bin_width=splines.DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=splines.DEFAULT_MIN_BIN_HEIGHT):
        super().__init__()
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.tail_bound = tail_bound
        self.tails = tails
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        if len(shape) != 1:
            raise NotImplementedError("image-shaped CDF layers are outside the B200 hot path")
        self.unnormalized_widths = nn.Parameter(torch.randn(*shape, num_bins))
        self.unnormalized_heights = nn.Parameter(torch.randn(*shape, num_bins + 1 if tails is None else num_bins - 1))
        self._spline = splines.QuadraticSplineSettings(num_bins, tails, tail_bound, min_bin_width, min_bin_height)

    def _shared_params(self, batch_size):
        p = torch.cat((self.unnormalized_widths, self.unnormalized_heights), dim=-1)
        return p.reshape(1, -1).expand(batch_size, -1)

    def apply_on_columns(self, inputs, tcols, ccols, inverse):
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), tcols, ccols, inverse, None)

    def _run(self, inputs, inverse):
        if inputs.dim() != 2 or inputs.shape[1] != self.unnormalized_widths.shape[0]:
            raise ValueError("expected inputs of shape [batch, {}]".format(self.unnormalized_widths.shape[0]))
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), None, None, inverse, None)

    def forward(self, inputs, context=None):
        return self._run(inputs, inverse=False)

    def inverse(self, inputs, context=None):
        return self._run(inputs, inverse=True)


class PiecewiseCubicCDF(Transform):
    """flowcon/transforms/nonlinearities.py:342-404: learnable `unnormalized_widths` / `unnormalized_heights` [*shape, K]
    and `unnorm_derivatives_left` / `unnorm_derivatives_right` [*shape, 1], shared across the batch."""

    def __init__(self, shape, num_bins=10, tails=None, tail_bound=1.0, min_bin_width=splines.DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=splines.DEFAULT_MIN_BIN_HEIGHT):
        super().__init__()
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.tail_bound = tail_bound
        self.tails = tails
        if isinstance(shape, int):
            shape = (shape,)
        shape = tuple(shape)
        if len(shape) != 1:
            raise NotImplementedError("image-shaped CDF layers are outside the B200 hot path")
        self.unnormalized_widths = nn.Parameter(torch.randn(*shape, num_bins))
        self.unnormalized_heights = nn.Parameter(torch.randn(*shape, num_bins))
        self.unnorm_derivatives_left = nn.Parameter(torch.randn(*shape, 1))
        self.unnorm_derivatives_right = nn.Parameter(torch.randn(*shape, 1))
        self._spline = splines.CubicSplineSettings(num_bins, tails, tail_bound, min_bin_width, min_bin_height)

    def _shared_params(self, batch_size):
        p = torch.cat((self.unnormalized_widths, self.unnormalized_heights, self.unnorm_derivatives_left,
                       self.unnorm_derivatives_right), dim=-1)
        return p.reshape(1, -1).expand(batch_size, -1)

    def apply_on_columns(self, inputs, tcols, ccols, inverse):
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), tcols, ccols, inverse, None)

    def _run(self, inputs, inverse):
        if inputs.dim() != 2 or inputs.shape[1] != self.unnormalized_widths.shape[0]:
            raise ValueError("expected inputs of shape [batch, {}]".format(self.unnormalized_widths.shape[0]))
        return self._spline.apply(inputs, self._shared_params(inputs.shape[0]), None, None, inverse, None)

    def forward(self, inputs, context=None):
        return self._run(inputs, inverse=False)

    def inverse(self, inputs, context=None):
        return self._run(inputs, inverse=True)
