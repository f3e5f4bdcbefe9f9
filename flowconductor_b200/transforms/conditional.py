"""Conditional ("hypernetwork") element-wise layers on the hot path, API of
flowcon/transforms/conditional.py (ConditionalTransform :23-95, ConditionalPiecewiseRationalQuadraticTransform
:656-743, ConditionalSumOfSigmoidsTransform :746-787): a ResidualNet maps the context to ALL D*P parameters and
one kernel applies the bijection.  Sub-module name `conditional_net` is kept.
"""
import torch
from torch.nn import functional as F

from .. import ops
from ..nn import tensorcore
from ..nn.nets import ResidualNet
from . import splines
from .base import Transform


class ConditionalTransform(Transform):
    def __init__(self, features, hidden_features=64, context_features=1, num_blocks=2, use_residual_blocks=True,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False, conditional_net=None):
        super().__init__()
        self.features = features
        if conditional_net is not None:
            assert isinstance(conditional_net, torch.nn.Module)
            self.conditional_net = conditional_net
        else:
            if not use_residual_blocks:
                raise NotImplementedError("the plain-MLP conditioner (use_residual_blocks=False) is outside the "
                                          "B200 hot path; pass conditional_net= instead")
            self.conditional_net = ResidualNet(in_features=context_features, out_features=self._num_parameters(),
                                               hidden_features=hidden_features, activation=activation,
                                               num_blocks=num_blocks, dropout_probability=dropout_probability,
                                               use_batch_norm=use_batch_norm)

    def _num_parameters(self):
        return self.features * self._output_dim_multiplier()

    def forward(self, inputs, context=None):
        if context is None:
            raise TypeError("Conditional transforms require a context.")
        if tensorcore.usable(self.conditional_net, context, None, inputs):
            return self._tensorcore_layer(inputs, context, inverse=False)
        return self._forward_given_params(inputs, self.conditional_net(context))

    def inverse(self, inputs, context=None):
        if context is None:
            raise TypeError("Conditional transforms require a context.")
        if tensorcore.usable(self.conditional_net, context, None, inputs):
            return self._tensorcore_layer(inputs, context, inverse=True)
        return self._inverse_given_params(inputs, self.conditional_net(context))

    def _tensorcore_layer(self, inputs, context, inverse):
        """Hypernetwork on the tensor cores (inference): the context is the conditioner's only input."""
        params = tensorcore.params(self.conditional_net, context)
        if inverse:
            return self._inverse_given_params(inputs, params)
        return self._forward_given_params(inputs, params)

    def _output_dim_multiplier(self):
        raise NotImplementedError()

    def _forward_given_params(self, inputs, autoregressive_params):
        raise NotImplementedError()

    def _inverse_given_params(self, inputs, autoregressive_params):
        raise NotImplementedError()


class ConditionalPiecewiseRationalQuadraticTransform(ConditionalTransform):
    """Identity-init on (conditional.py:733), box [-1.2,1.2]^2 without tails (:717), and the 1/sqrt(H)
    pre-scale IS applied because the conditioner is a ResidualNet (:711-713)."""

    def __init__(self, features, hidden_features, context_features=None, num_bins=10, tails=None, tail_bound=1.0,
                 num_blocks=2, use_residual_blocks=True, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False, min_bin_width=splines.DEFAULT_MIN_BIN_WIDTH,
                 min_bin_height=splines.DEFAULT_MIN_BIN_HEIGHT, min_derivative=splines.DEFAULT_MIN_DERIVATIVE):
        self.num_bins = num_bins
        self.min_bin_width = min_bin_width
        self.min_bin_height = min_bin_height
        self.min_derivative = min_derivative
        self.tails = tails
        self.tail_bound = tail_bound
        self._spline = splines.RationalQuadraticSettings(num_bins, tails, tail_bound, min_bin_width, min_bin_height,
                                                         min_derivative, identity_init=True,
                                                         constrained_box=(-1.2, 1.2))
        super().__init__(features=features, hidden_features=hidden_features, context_features=context_features,
                         num_blocks=num_blocks, use_residual_blocks=use_residual_blocks, activation=activation,
                         dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)

    def _output_dim_multiplier(self):
        return self._spline.params_per_feature()

    def _elementwise(self, inputs, autoregressive_params, inverse=False):
        hidden = getattr(self.conditional_net, "hidden_features", None)
        return self._spline.apply(inputs, autoregressive_params, None, None, inverse, hidden)

    def _forward_given_params(self, inputs, autoregressive_params):
        return self._elementwise(inputs, autoregressive_params)

    def _inverse_given_params(self, inputs, autoregressive_params):
        return self._elementwise(inputs, autoregressive_params, inverse=True)

    def _tensorcore_layer(self, inputs, context, inverse):
        net = self.conditional_net
        if not tensorcore.rqs_fusable(self._spline, net.final_layer.weight.shape[0], inputs.shape[1], net, context.shape[1]):
            return super()._tensorcore_layer(inputs, context, inverse)
        return tensorcore.rqs_layer(net, context, inputs, self._spline, None, None, inverse,
                                    getattr(net, "hidden_features", None))


class ConditionalSumOfSigmoidsTransform(ConditionalTransform):
    def __init__(self, features, hidden_features, context_features=None, n_sigmoids=10, num_blocks=2,
                 use_residual_blocks=True, activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        self.n_sigmoids = n_sigmoids
        super().__init__(features=features, hidden_features=hidden_features, context_features=context_features,
                         num_blocks=num_blocks, use_residual_blocks=use_residual_blocks, activation=activation,
                         dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)

    def _output_dim_multiplier(self):
        return 3 * self.n_sigmoids + 1

    def _forward_given_params(self, inputs, autoregressive_params):
        return ops.sos_layer(inputs, autoregressive_params, self.n_sigmoids, 0.0, False, 50, 120.0)

    def _inverse_given_params(self, inputs, autoregressive_params):
        return ops.sos_layer(inputs, autoregressive_params, self.n_sigmoids, 0.0, True, 50, 120.0)

    def _tensorcore_layer(self, inputs, context, inverse):
        net = self.conditional_net
        if (not inverse and net.final_layer.weight.shape[0] == inputs.shape[1] * self._output_dim_multiplier()
                and tensorcore.sos_fusable(net, context.shape[1], self.n_sigmoids, inputs.shape[1])):
            # ResidualNet on the context + sum of sigmoids in one kernel: the [B, D * 31] parameters never exist
            return tensorcore.sos_layer(net, context, inputs, self.n_sigmoids, 0.0)
        return super()._tensorcore_layer(inputs, context, inverse)
