"""Minimal stand-in for the `torchtestcase` package used by the reference's unit tests.

Test infrastructure only. Provides a tensor-aware unittest.TestCase with an `eps` tolerance.
"""
import unittest

import torch


class TorchTestCase(unittest.TestCase):
    _eps = 0.0

    @property
    def eps(self):
        return self._eps

    @eps.setter
    def eps(self, value):
        self._eps = float(value)

    def _fail_with_message(self, msg, standard_msg):
        self.fail(self._formatMessage(msg, standard_msg))

    def _tensors_close(self, first, second):
        if first.shape != second.shape:
            return False, "shapes differ: {} vs {}".format(tuple(first.shape), tuple(second.shape))
        if first.numel() == 0:
            return True, ""
        if self._eps and first.is_floating_point():
            diff = (first - second).abs().max().item()
            return diff <= self._eps, "max abs diff {} > eps {}".format(diff, self._eps)
        return bool(torch.equal(first, second)), "tensors are not equal"

    def assertEqual(self, first, second, msg=None):
        if torch.is_tensor(first) and torch.is_tensor(second):
            ok, why = self._tensors_close(first, second)
            if not ok:
                self._fail_with_message(msg, why)
        else:
            super().assertEqual(first, second, msg)

    def assert_tensor_equal(self, first, second, msg=None):
        self.assertEqual(first, second, msg)

    def assert_tensor_less(self, first, second, msg=None):
        if not bool((first < second).all()):
            self._fail_with_message(msg, "tensor is not element-wise less")
