"""Guard-band check of fc_linear_apply: every output buffer is a window inside a larger sentinel-filled buffer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import linear as fl
dev = torch.device("cuda:0")
SENT = 12345.0
GUARD = 1 << 18  # floats on each side


def guarded(n):
    big = torch.full((n + 2 * GUARD,), SENT, device=dev)
    return big, big[GUARD:GUARD + n]


def check(big, n, tag):
    torch.cuda.synchronize()
    lo = big[:GUARD] != SENT
    hi = big[GUARD + n:] != SENT
    if lo.any() or hi.any():
        hidx = torch.nonzero(hi)
        print("  OOB WRITE", tag, "below:", int(lo.sum()), "above:", int(hi.sum()),
              "first/last offset past the end:", (hidx[0].item(), hidx[-1].item()) if hi.any() else None)
        return True
    return False


torch.manual_seed(0)
bad = False
for M in (1, 31, 32, 100, 128, 129, 500):
    for K, N in ((256, 256), (256, 384), (64, 64), (256, 752), (128, 100)):
        W = torch.randn(N, K, device=dev) * 0.1
        b = torch.randn(N, device=dev)
        packed = fl.pack(W, b)
        a_rows = torch.randn(M, K, device=dev)
        ref = a_rows @ W.t() + b
        t128_ok = K % 16 == 0 and N % 16 == 0
        for a_t, o_t, res in ((False, False, False), (False, True, False), (True, True, False), (True, True, True), (True, False, False),
                              (False, False, True)):
            if (a_t or o_t) and not t128_ok:
                continue
            if res and N != K:
                continue
            a = fl.T128.from_rows(a_rows) if a_t else a_rows
            n4 = (N + 3) // 4 * 4
            if o_t:
                out = fl.T128(M, N, dev)
                big, win = guarded(out.buf.numel())
                out.buf = win
                n = win.numel()
                r = (fl.T128.from_rows(a_rows) if res else None)
            else:
                big, flat = guarded(M * n4)
                out = flat.view(M, n4)
                n = M * n4
                r = a_rows if res else None
            got = fl.linear(a, packed, residual=r, out=out, out_t128=o_t, n_out=(N if o_t else n4))
            tag = "M=%d K=%d N=%d a_t128=%s out_t128=%s residual=%s" % (M, K, N, a_t, o_t, res)
            if check(big, n, tag):
                bad = True
            g = got.to_rows() if o_t else got[:, :N]
            want = ref + (a_rows if res else 0)
            err = (g - want).abs().max().item()
            if err > 1e-3:
                print("  WRONG RESULT", tag, err)
                bad = True
print("guard-band check of fc_linear_apply:", "FAILED" if bad else "clean")
