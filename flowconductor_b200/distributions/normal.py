"""StandardNormal base density (flowcon/distributions/normal.py:11-50).  `_log_prob` is one kernel:
-0.5 * sum z^2 - 0.5 D log(2 pi), optionally fused with the flow's `+ logabsdet` (flows/base.py:48)."""
import numpy as np
import torch

from .. import ops
from ..utils import torchutils
from .base import Distribution


class StandardNormal(Distribution):
    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)
        self.register_buffer("_log_z", torch.tensor(0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64),
                             persistent=False)

    def _check(self, inputs):
        if inputs.shape[1:] != self._shape:
            raise ValueError("Expected input of shape {}, got {}".format(self._shape, inputs.shape[1:]))

    def _log_prob(self, inputs, context):
        self._check(inputs)
        return ops.stdnormal_log_prob(inputs.reshape(inputs.shape[0], -1), None)

    def log_prob_plus(self, inputs, logabsdet):
        """log N(inputs) + logabsdet in one pass (the tail of Flow._log_prob)."""
        self._check(inputs)
        return ops.stdnormal_log_prob(inputs.reshape(inputs.shape[0], -1), logabsdet)

    def _sample(self, num_samples, context):
        if context is None:
            return torch.randn(num_samples, *self._shape, device=self._log_z.device)
        samples = torch.randn(context.shape[0] * num_samples, *self._shape, device=context.device)
        return torchutils.split_leading_dim(samples, [context.shape[0], num_samples])

    def _mean(self, context):
        if context is None:
            return self._log_z.new_zeros(self._shape)
        return context.new_zeros(context.shape[0], *self._shape)
