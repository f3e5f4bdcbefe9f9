#!/usr/bin/env python
"""Accuracy of the RQ-spline element arithmetic on large fresh samples: error of OURS vs the fp64 oracle next
to the error of the fp32 ORACLE (== reference op chain) vs the same fp64 oracle, as quantiles.  'ratio' < ~1.3 at
every quantile means we are statistically as accurate as the reference's own fp32 evaluation.
Runs the CUDA kernels when a GPU is present, else the host-compiled copy of the same header (tests/hostmath)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import restated  # noqa: E402

N, D = 16384, 16
Q = torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999], dtype=torch.float64)


def ours_gpu(x, p, k, tb, inverse, ident):
    from flowconductor_b200 import transforms
    dev = torch.device("cuda:0")
    xx, pp = x.to(dev), p.to(dev)
    y, lad = transforms.unconstrained_rational_quadratic_spline(
        xx, pp[..., :k], pp[..., k:2 * k], pp[..., 2 * k:], inverse=inverse, tails="linear", tail_bound=tb,
        enable_identity_init=ident)
    return y.cpu(), lad.cpu()


def ours_host(x, p, k, tb, inverse, ident):
    from tests.test_kernel_math_host import LIB, fptr, make_cfg
    hm = ctypes.CDLL(LIB)
    y, lad = torch.empty_like(x), torch.empty_like(x)
    status = ctypes.c_uint(0)
    cfg = make_cfg(k, 1, tb, int(inverse), int(ident))
    hm.hm_rqs_apply(fptr(x.contiguous()), fptr(p.contiguous()), fptr(y), fptr(lad), ctypes.c_long(x.numel()),
                    ctypes.byref(cfg), 0, ctypes.byref(status))
    return y, lad


def main():
    gpu = torch.cuda.is_available()
    print("backend:", "CUDA kernels" if gpu else "host shim (same header, libm primitives)")
    fn = ours_gpu if gpu else ours_host
    print("%-28s %-4s %s" % ("case", "out", "quantiles 50/90/99/99.9/99.99 of |err vs fp64|: ours | ref32 | ratio"))
    for k, tb, ident, scale in ((8, 3.0, False, 2.0), (8, 3.0, False, 1.0), (16, 3.0, True, 1.5)):
        for inverse in (False, True):
            g = torch.Generator().manual_seed(k * 10 + inverse)
            x = torch.randn(N, D, generator=g) * tb * 0.5
            p = torch.randn(N, D, 3 * k - 1, generator=g) * scale
            kw = dict(inverse=inverse, tails="linear", tail_bound=tb, enable_identity_init=ident)
            sl = lambda t: (t[..., :k], t[..., k:2 * k], t[..., 2 * k:])  # noqa: E731
            r32y, r32l = restated.unconstrained_rational_quadratic_spline(x, *sl(p), **kw)
            r64y, r64l = restated.unconstrained_rational_quadratic_spline(x.double(), *sl(p.double()), **kw)
            oy, ol = fn(x, p, k, tb, inverse, ident)
            for name, o, r32, r64 in (("y", oy, r32y, r64y), ("lad", ol, r32l, r64l)):
                eo = torch.quantile((o.double() - r64).abs().flatten(), Q)
                er = torch.quantile((r32.double() - r64).abs().flatten(), Q)
                ratio = eo / er.clamp_min(1e-12)
                print("K=%-2d scale=%.1f %s %-4s %s | %s | %s" % (
                    k, scale, "inv" if inverse else "fwd", name, " ".join("%.1e" % v for v in eo.tolist()),
                    " ".join("%.1e" % v for v in er.tolist()), " ".join("%.2f" % v for v in ratio.tolist())))


if __name__ == "__main__":
    main()
