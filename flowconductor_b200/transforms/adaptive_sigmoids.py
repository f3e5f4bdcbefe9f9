"""`SumOfSigmoids` as a stand-alone element-wise transform (flowcon/transforms/adaptive_sigmoids.py:13-142):
either with its own learnable raw parameters (shared across the batch) or wrapping per-sample `raw_params`
[B, features, 3n+1].  Forward and numerical inverse both run in the sum-of-sigmoids kernel."""
import torch
from torch import nn

from .. import ops
from .base import Transform


class _ExtendedSoftplusParams(nn.Module):
    """Holder of the `shift` parameter under the reference's sub-module name (`extended_softplus.shift`,
    flowcon/transforms/nonlinearities.py:490-512), so reference checkpoints load with strict=True.  The arithmetic lives in
    the kernel."""

    def __init__(self, features):
        super().__init__()
        self.shift = nn.Parameter(torch.ones(1, features) * 3)  # nonlinearities.py:503


class SumOfSigmoids(Transform):
    PREACT_SCALE_MIN = 0.1
    PREACT_SCALE_MAX = 10.0
    PREACT_SHIFT_MAX = 10

    def __init__(self, features, n_sigmoids=10, iterations_bisection_inverse=50, lim_bisection_inverse=120,
                 raw_params=None):
        super().__init__()
        self.features = features
        self.n_sigmoids = n_sigmoids
        self.num_iterations = iterations_bisection_inverse
        self.lim = lim_bisection_inverse
        self._raw = None
        if raw_params is None:
            self.shift_preact = nn.Parameter(torch.randn(1, features, n_sigmoids))
            self.log_scale_preact = nn.Parameter(torch.zeros(1, features, n_sigmoids))
            self.raw_softmax = nn.Parameter(torch.ones(1, features, n_sigmoids))
            self.extended_softplus = _ExtendedSoftplusParams(features)
        else:
            assert raw_params.shape[1:] == (features, 3 * n_sigmoids + 1)
            self._raw = raw_params
        # adaptive_sigmoids.py:67: a frozen zero the reference keeps in its state_dict
        self.log_scale_postact = nn.Parameter(torch.zeros(1), requires_grad=False)

    def get_raw_params(self):
        if self._raw is not None:
            return self._raw
        return torch.cat((self.shift_preact, self.log_scale_preact, self.raw_softmax,
                          self.extended_softplus.shift.reshape(1, self.features, 1)), dim=-1)

    def _params_for(self, inputs):
        raw = self.get_raw_params()
        if raw.shape[0] != inputs.shape[0]:
            raw = raw.expand(inputs.shape[0], -1, -1)
        return raw.reshape(inputs.shape[0], -1)

    def forward(self, inputs, context=None):
        return ops.sos_layer(inputs, self._params_for(inputs), self.n_sigmoids, 0.0, False, self.num_iterations,
                             float(self.lim))

    def inverse(self, inputs, context=None):
        return ops.sos_layer(inputs, self._params_for(inputs), self.n_sigmoids, 0.0, True, self.num_iterations,
                             float(self.lim))
