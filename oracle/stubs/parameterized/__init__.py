"""Minimal stand-in for the `parameterized` package used by a few reference tests (test infrastructure only)."""
import functools
import sys


def parameterized_class(attrs, input_values=None):
    """Create one subclass per value tuple; the undecorated base is kept out of collection."""
    if isinstance(attrs, str):
        attrs = [attrs]

    def decorator(base):
        module = sys.modules[base.__module__]
        for idx, values in enumerate(input_values):
            if not isinstance(values, (tuple, list)):
                values = (values,)
            name = "{}_{}".format(base.__name__, idx)
            sub = type(name, (base,), dict(zip(attrs, values)))
            sub.__test__ = True
            setattr(module, name, sub)
        base.__test__ = False
        return base

    return decorator


class parameterized:
    @staticmethod
    def expand(cases):
        def decorator(fn):
            @functools.wraps(fn)
            def runner(self):
                for case in cases:
                    if not isinstance(case, (tuple, list)):
                        case = (case,)
                    with self.subTest(case=case):
                        fn(self, *case)

            return runner

        return decorator
