"""Host-side tensor helpers with the names and semantics of flowcon/utils/torchutils.py.

Only shape / mask plumbing lives here.  The two helpers on the hot path — `searchsorted`
(torchutils.py:147-149) and `sum_except_batch` (:25-30) as used on per-element log-dets — are fused into the
kernels; `sum_except_batch` is kept as a plain torch reduction for callers outside the path.
"""
import torch

from . import typechecks as check


def tile(x, n):
    """[a, b, c] -> [a]*n + [b]*n + [c]*n (torchutils.py:14-22; MADE output degrees)."""
    if not check.is_positive_int(n):
        raise TypeError("Argument 'n' must be a positive integer.")
    return x.reshape(-1).repeat_interleave(n)


def sum_except_batch(x, num_batch_dims=1):
    if not check.is_nonnegative_int(num_batch_dims):
        raise TypeError("Number of batch dimensions must be a non-negative integer.")
    dims = list(range(num_batch_dims, x.ndimension()))
    return torch.sum(x, dim=dims) if dims else x


def split_leading_dim(x, shape):
    return torch.reshape(x, torch.Size(shape) + x.shape[1:])


def merge_leading_dims(x, num_dims):
    if not check.is_positive_int(num_dims):
        raise TypeError("Number of leading dims must be a positive integer.")
    if num_dims > x.dim():
        raise ValueError("Number of leading dims can't be greater than total number of dims.")
    return torch.reshape(x, torch.Size([-1]) + x.shape[num_dims:])


def repeat_rows(x, num_reps):
    if not check.is_positive_int(num_reps):
        raise TypeError("Number of repetitions must be a positive integer.")
    return x.repeat_interleave(num_reps, dim=0)


def create_alternating_binary_mask(features, even=True):
    mask = torch.zeros(features, dtype=torch.uint8)
    mask[(0 if even else 1)::2] = 1
    return mask


def create_mid_split_binary_mask(features):
    mask = torch.zeros(features, dtype=torch.uint8)
    mask[: (features + 1) // 2] = 1
    return mask


def create_random_binary_mask(features):
    mask = torch.zeros(features, dtype=torch.uint8)
    picks = torch.multinomial(torch.ones(features), num_samples=(features + 1) // 2, replacement=False)
    mask[picks] = 1
    return mask
