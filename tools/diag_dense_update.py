import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import littlemcmc_b200 as lmc
from littlemcmc_b200 import _lib as L
import ctypes as C
dev = torch.device("cuda", 0)
D, Cn = 1000, 256
pot = lmc.QuadPotentialFullAdapt(D, np.zeros(D), np.eye(D), 10)
pot._to_device(dev, Cn)
q = torch.randn(Cn, D, dtype=torch.float64, device=dev)
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for k in (256, 64, 8):
    idx = torch.arange(k, device=dev)
    print(k, "update_rows total %.1f ms" % t(lambda: pot._update_rows(idx, q)))
    sel = idx
    print("   gather cov %.1f ms" % t(lambda: pot._cov_all[sel][:, :, :D]))
    A = pot._cov_all[sel][:, :, :D]
    print("   cholesky_ex %.1f ms" % t(lambda: torch.linalg.cholesky_ex(A)))
    chol, info = torch.linalg.cholesky_ex(A)
    print("   isfinite %.1f ms" % t(lambda: bool(((info == 0) & torch.isfinite(chol).all(dim=2).all(dim=1)).all())))
    ok = (info == 0)
    print("   scatter %.1f ms" % t(lambda: pot._chol_all.__setitem__(sel[ok], chol[ok])))
    lib = L.load()
    sel32 = sel.to(torch.int32)
    p = lambda x: C.c_void_p(x.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    print("   cov_update kernel %.1f ms" % t(lambda: lib.lmc_dense_cov_update(p(sel32), k, D, q.shape[1], pot._lda, p(q), p(pot._mean_fg), p(pot._raw_fg), p(pot._mean_bg), p(pot._raw_bg), p(pot._nsamp), p(pot._cov_all), st)))
