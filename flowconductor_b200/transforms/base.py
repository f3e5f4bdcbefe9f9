"""Transform protocol and composition — the API contract of flowcon/transforms/base.py kept as is:
`forward(inputs, context=None) -> (outputs, logabsdet[B])`, `inverse(...)` the same."""
import torch
from torch import nn


class InverseNotAvailable(Exception):
    """Raised by transforms without an inverse (flowcon/transforms/base.py:10-13)."""


class InputOutsideDomain(Exception):
    """Raised when an input is outside a transform's domain (flowcon/transforms/base.py:16-19)."""


class Transform(nn.Module):
    def forward(self, inputs, context=None):
        raise NotImplementedError()

    def inverse(self, inputs, context=None):
        raise InverseNotAvailable()


class CompositeTransform(Transform):
    """Applies transforms in order; inverse walks them backwards (flowcon/transforms/base.py:32-60)."""

    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    @staticmethod
    def _cascade(inputs, funcs, context):
        from .. import _cabi
        from ..nn import tensorcore

        outputs = inputs
        total_logabsdet = inputs.new_zeros(inputs.shape[0])
        try:
            for i, func in enumerate(funcs):
                # the previous layer's output is referenced by this loop only: a layer may overwrite it in place
                # (honoured by the tensor-core inference path when that layer allocated the tensor itself)
                tensorcore.begin_layer(outputs if i > 0 else None)
                if _cabi.NVTX:
                    owner = getattr(func, "__self__", func)
                    name = "layer {}: {}.{}".format(i, type(owner).__name__, getattr(func, "__name__", "forward"))
                    with _cabi.nvtx_range(name):
                        outputs, logabsdet = func(outputs, context)
                else:
                    outputs, logabsdet = func(outputs, context)
                total_logabsdet = total_logabsdet + logabsdet
        finally:
            tensorcore.end_cascade()
        return outputs, total_logabsdet

    def forward(self, inputs, context=None):
        return self._cascade(inputs, self._transforms, context)

    def inverse(self, inputs, context=None):
        return self._cascade(inputs, (t.inverse for t in self._transforms[::-1]), context)


class InverseTransform(Transform):
    """Swaps forward and inverse of a transform (flowcon/transforms/base.py:215-231)."""

    def __init__(self, transform):
        super().__init__()
        self._transform = transform

    def forward(self, inputs, context=None):
        return self._transform.inverse(inputs, context)

    def inverse(self, inputs, context=None):
        return self._transform(inputs, context)
