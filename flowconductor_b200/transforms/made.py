"""MADE conditioner with the module / buffer names of flowcon/transforms/made.py (`initial_layer`,
`blocks.N.linear_layers.{0,1}`, `final_layer`, each with `weight`, `bias`, `mask`, `degrees`), so reference
state_dicts load.  Hidden layers stay torch; the final masked layer is the GEMM the fused kernel replaces.
"""
import torch
from torch import nn
from torch.nn import functional as F

from ..nn import tc_autograd
from ..utils import torchutils


def _input_degrees(features):
    return torch.arange(1, features + 1)


class MaskedLinear(nn.Linear):
    """nn.Linear whose weight is multiplied by a fixed 0/1 mask on every call (made.py:17-72)."""

    def __init__(self, in_degrees, out_features, autoregressive_features, random_mask, is_output, bias=True):
        super().__init__(in_features=len(in_degrees), out_features=out_features, bias=bias)
        mask, degrees = self._get_mask_and_degrees(in_degrees, out_features, autoregressive_features, random_mask,
                                                   is_output)
        self.register_buffer("mask", mask)
        self.register_buffer("degrees", degrees)

    @classmethod
    def _get_mask_and_degrees(cls, in_degrees, out_features, autoregressive_features, random_mask, is_output):
        d = autoregressive_features
        if is_output:
            # every feature owns `out_features // d` consecutive outputs, all with that feature's degree;
            # an output may see only strictly smaller degrees (made.py:46-51)
            out_degrees = torchutils.tile(_input_degrees(d), out_features // d)
            return (out_degrees[:, None] > in_degrees).float(), out_degrees
        if random_mask:
            low = min(torch.min(in_degrees).item(), d - 1)
            out_degrees = torch.randint(low=low, high=d, size=[out_features], dtype=torch.long)
        else:
            out_degrees = torch.arange(out_features) % max(1, d - 1) + min(1, d - 1)
        return (out_degrees[:, None] >= in_degrees).float(), out_degrees

    def masked_weight(self):
        return self.weight * self.mask

    def forward(self, x):
        return tc_autograd.linear(x, self.weight, self.bias, self.mask)


class MaskedFeedforwardBlock(nn.Module):
    """made.py:75-123 (output width == input width)."""

    def __init__(self, in_degrees, autoregressive_features, context_features=None, random_mask=False,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        features = len(in_degrees)
        self.batch_norm = nn.BatchNorm1d(features, eps=1e-3) if use_batch_norm else None
        self.linear = MaskedLinear(in_degrees, features, autoregressive_features, random_mask, is_output=False)
        self.degrees = self.linear.degrees
        self.activation = activation
        self.dropout = nn.Dropout(p=dropout_probability)

    def forward(self, inputs, context=None):
        h = self.batch_norm(inputs) if self.batch_norm else inputs
        return self.dropout(self.activation(self.linear(h)))


class MaskedResidualBlock(nn.Module):
    """made.py:126-202."""

    def __init__(self, in_degrees, autoregressive_features, context_features=None, random_mask=False,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False, zero_initialization=True):
        if random_mask:
            raise ValueError("Masked residual block can't be used with random masks.")
        super().__init__()
        features = len(in_degrees)
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList(nn.BatchNorm1d(features, eps=1e-3) for _ in range(2))
        first = MaskedLinear(in_degrees, features, autoregressive_features, False, is_output=False)
        second = MaskedLinear(first.degrees, features, autoregressive_features, False, is_output=False)
        self.linear_layers = nn.ModuleList([first, second])
        self.degrees = second.degrees
        if not bool(torch.all(self.degrees >= in_degrees)):
            raise RuntimeError("In a masked residual block, the output degrees can't be less than the "
                               "corresponding input degrees.")
        self.activation = activation
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            for tensor in (second.weight, second.bias):
                nn.init.uniform_(tensor, a=-1e-3, b=1e-3)

    def forward(self, inputs, context=None):
        if (context is None and not self.use_batch_norm and tc_autograd.is_relu(self.activation)
                and not (self.training and self.dropout.p > 0)
                and tc_autograd.residual_block_eligible(inputs, self.linear_layers[0], self.linear_layers[1])):
            return tc_autograd.residual_block(inputs, self.linear_layers[0], self.linear_layers[1])
        h = self.batch_norm_layers[0](inputs) if self.use_batch_norm else inputs
        h = self.linear_layers[0](self.activation(h))
        if context is not None:
            h = h + self.context_layer(context)
        if self.use_batch_norm:
            h = self.batch_norm_layers[1](h)
        h = self.linear_layers[1](self.dropout(self.activation(h)))
        return inputs + h


class MADE(nn.Module):
    """made.py:205-283.  Deliberately has NO `hidden_features` attribute: the reference's autoregressive RQ
    layer therefore skips the 1/sqrt(H) pre-scale (autoregressive.py:589), and so does ours."""

    def __init__(self, features, hidden_features, context_features=None, num_blocks=2, output_multiplier=1,
                 use_residual_blocks=True, random_mask=False, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False):
        if use_residual_blocks and random_mask:
            raise ValueError("Residual blocks can't be used with random masks.")
        super().__init__()
        self.initial_layer = MaskedLinear(_input_degrees(features), hidden_features, features, random_mask,
                                          is_output=False)
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, hidden_features)
        self.use_residual_blocks = use_residual_blocks
        self.activation = activation
        block_cls = MaskedResidualBlock if use_residual_blocks else MaskedFeedforwardBlock
        blocks, degrees = [], self.initial_layer.degrees
        for _ in range(num_blocks):
            blocks.append(block_cls(in_degrees=degrees, autoregressive_features=features,
                                    context_features=context_features, random_mask=random_mask,
                                    activation=activation, dropout_probability=dropout_probability,
                                    use_batch_norm=use_batch_norm))
            degrees = blocks[-1].degrees
        self.blocks = nn.ModuleList(blocks)
        self.final_layer = MaskedLinear(degrees, features * output_multiplier, features, random_mask, is_output=True)

    def hidden(self, inputs, context=None):
        """Everything up to (not including) `final_layer`."""
        h = self.initial_layer(inputs)
        if context is not None:
            h = h + self.activation(self.context_layer(context))
        if not self.use_residual_blocks:
            h = self.activation(h)
        for block in self.blocks:
            h = block(h, context)
        return h

    def forward(self, inputs, context=None):
        return self.final_layer(self.hidden(inputs, context))
