"""Residual MLP conditioner with the parameter names of flowcon/nn/nets/resnet.py (`initial_layer`,
`blocks.N.linear_layers.{0,1}`, `blocks.N.context_layer`, `final_layer`), so reference state_dicts load.

Inference runs the whole net on the tensor cores (nn/tensorcore.py, final layer fused with the bijection); with
autograd every dense layer goes through nn/tc_autograd.py (forward and input-gradient GEMMs on the tensor cores).
`hidden_features` is public on purpose: couplings use it for the 1/sqrt(H) width/height pre-scale
(coupling.py:554-556).
"""
import torch
from torch import nn
from torch.nn import functional as F

from .. import tc_autograd


class ResidualBlock(nn.Module):
    def __init__(self, features, context_features, activation=torch.nn.ReLU(), dropout_probability=0.0,
                 use_batch_norm=False, zero_initialization=True):
        super().__init__()
        self.activation = activation
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList(nn.BatchNorm1d(features, eps=1e-3) for _ in range(2))
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.linear_layers = nn.ModuleList(nn.Linear(features, features) for _ in range(2))
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:  # resnet.py:35-37: the block starts as (almost) the identity
            for tensor in (self.linear_layers[1].weight, self.linear_layers[1].bias):
                nn.init.uniform_(tensor, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        if (context is None and not self.use_batch_norm and tc_autograd.is_relu(self.activation)
                and not (self.training and self.dropout.p > 0)
                and tc_autograd.residual_block_eligible(inputs, self.linear_layers[0], self.linear_layers[1])):
            return tc_autograd.residual_block(inputs, self.linear_layers[0], self.linear_layers[1])
        h = inputs
        for i in range(2):
            if self.use_batch_norm:
                h = self.batch_norm_layers[i](h)
            h = self.activation(h)
            if i == 1:
                h = self.dropout(h)
            h = tc_autograd.module_linear(self.linear_layers[i], h)
        if context is not None:
            h = F.glu(torch.cat((h, self.context_layer(context)), dim=1), dim=1)
        return inputs + h


class ResidualNet(nn.Module):
    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2,
                 activation=torch.nn.ReLU(), dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.hidden_features = hidden_features
        self.context_features = context_features
        self.initial_layer = nn.Linear(in_features + (context_features or 0), hidden_features)
        self.blocks = nn.ModuleList(
            ResidualBlock(hidden_features, context_features, activation=activation,
                          dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
            for _ in range(num_blocks))
        self.final_layer = nn.Linear(hidden_features, out_features)

    def hidden(self, inputs, context=None):
        """Everything up to (not including) `final_layer`."""
        h = tc_autograd.module_linear(self.initial_layer,
                                      inputs if context is None else torch.cat((inputs, context), dim=1))
        for block in self.blocks:
            h = block(h, context=context)
        return h

    def forward(self, inputs, context=None):
        return tc_autograd.module_linear(self.final_layer, self.hidden(inputs, context))
