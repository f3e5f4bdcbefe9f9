import os, sys, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import linear as fl, _cabi
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for (N, K, B) in [(256, 256, 8192), (64, 100, 4096), (752, 256, 20000)]:
    gy = torch.randn(B, N, generator=g, device=dev); x = torch.randn(B, K, generator=g, device=dev)
    want = gy.double().t() @ x.double()
    pk = fl.pack_transposed(x)
    for sl in (1, 2, 7, None):
        L = _cabi.lib()
        M = N; n4 = (K + 3)//4*4
        ks = sl or 8
        slice_rows = (M + 255)//256*256
        partials = torch.full((ks, slice_rows, n4), float('nan'), device=dev)
        rc = L.fc_linear_splitk_t_apply(gy.data_ptr(), N, M, B, ctypes.byref(pk.struct), ks, partials.data_ptr(), slice_rows, n4, n4, _cabi.stream_ptr(dev))
        torch.cuda.synchronize()
        p = partials[:, :M]
        nan_frac = torch.isnan(p).float().mean().item()
        got = torch.nan_to_num(p).sum(0)
        err = (got.double() - want[:, :n4]).abs().max().item() / want.abs().max().item()
        print(N, K, B, "slices", ks, "rc", rc, "nan_frac", round(nan_frac, 4), "relerr", err, "per-slice absmax", [round(torch.nan_to_num(p[i]).abs().max().item(), 2) for i in range(min(ks, 8))], flush=True)
