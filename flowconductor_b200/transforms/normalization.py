"""`ActNorm` with the API and state_dict of flowcon/transforms/normalization.py:144-218 (2-D inputs): a per-feature affine
map whose scale and shift are learnable and, in training mode, initialised from the first batch so that the outputs have
zero mean and unit variance.  Forward / inverse / gradients run in csrc/fc_actnorm.cu."""
import torch
from torch import nn

from .. import ops
from ..utils import typechecks as check
from .base import Transform


class ActNorm(Transform):
    def __init__(self, features):
        if not check.is_positive_int(features):
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.register_buffer("initialized", torch.tensor(False, dtype=torch.bool))
        self.log_scale = nn.Parameter(torch.zeros(features))
        self.shift = nn.Parameter(torch.zeros(features))

    @property
    def scale(self):
        return torch.exp(self.log_scale)

    @staticmethod
    def _check(inputs):
        if inputs.dim() not in [2, 4]:
            raise ValueError("Expecting inputs to be a 2D or a 4D tensor.")
        if inputs.dim() == 4:
            raise NotImplementedError("image-shaped (4-D) inputs are outside the B200 hot path")

    def forward(self, inputs, context=None):
        self._check(inputs)
        if self.training and not self.initialized:
            self._initialize(inputs)
        return ops.actnorm_layer(inputs, self.log_scale, self.shift, False)

    def inverse(self, inputs, context=None):
        self._check(inputs)
        return ops.actnorm_layer(inputs, self.log_scale, self.shift, True)

    def _initialize(self, inputs):
        """Data-dependent initialisation (normalization.py:204-218): one-off statistics of the first training batch."""
        with torch.no_grad():
            std = inputs.std(dim=0)
            mu = (inputs / std).mean(dim=0)
            self.log_scale.data = -torch.log(std)
            self.shift.data = -mu
            self.initialized.data = torch.tensor(True, dtype=torch.bool)
