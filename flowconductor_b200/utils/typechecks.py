"""Argument checks (mirror of flowcon/utils/typechecks.py)."""


def is_bool(x):
    return isinstance(x, bool)


def is_int(x):
    return isinstance(x, int) and not isinstance(x, bool)


def is_positive_int(x):
    return is_int(x) and x > 0


def is_nonnegative_int(x):
    return is_int(x) and x >= 0
