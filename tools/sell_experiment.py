"""GPU experiment: would a SELL-32-sigma SpMV (lane per row, sorted windows, renumbered system) beat the
CSR-stream kernel on the velocity block of the W matrix?  usage: python tools/sell_experiment.py [n] [sigma]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402

from fenics_b200 import _abi as abi  # noqa: E402
from fenics_b200 import dolfin_api as dapi, mesh as pmesh  # noqa: E402
from fenics_b200.drivers import ldc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
sigma = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
drv = ldc.CompLDC(pmesh.Skew_Mesh(n, 1, 1))
drv.t = 0.45
drv.step()
A = dapi.assemble(drv.pr.form(abi.FX_FORM_C2_A1)).to_scipy().tocsr()
A.eliminate_zeros()
N = A.shape[0]
elim = np.zeros(N, bool)
elim[np.unique(drv.S["W"].cell_dofs[:, 12:15])] = True
A = sp.diags((~elim).astype(float)) @ A  # eliminated rows empty
A.eliminate_zeros()
A = A.tocsr()
lens = np.diff(A.indptr)
print("rows", N, "nnz", A.nnz, "mean row", A.nnz / (~elim).sum())
# window-local sort by descending length -> new numbering
perm = np.concatenate([w0 + np.argsort(-lens[w0:w0 + sigma], kind="stable") for w0 in range(0, N, sigma)])  # new -> old
iperm = np.empty(N, np.int64)
iperm[perm] = np.arange(N)
Npad = (N + 31) // 32 * 32
nsl = Npad // 32
plen = np.zeros(Npad, np.int64)
plen[:N] = lens[perm]
slen = plen.reshape(nsl, 32).max(1)
sptr = np.concatenate([[0], np.cumsum(slen * 32)])
total = int(sptr[-1])
print("slices", nsl, "padded entries", total, "padding overhead %.1f%%" % (100.0 * (total - A.nnz) / A.nnz))
val = np.zeros(total)
col = np.zeros(total, np.int32)
indptr, indices, data = A.indptr, A.indices, A.data
for s in range(nsl):  # vectorised per slice
    rows = perm[s * 32:min(N, s * 32 + 32)]
    for lane, r in enumerate(rows):
        b, e = indptr[r], indptr[r + 1]
        idx = sptr[s] + np.arange(e - b) * 32 + lane
        val[idx] = data[b:e]
        col[idx] = iperm[indices[b:e]]
dev = "cuda"
t = lambda a: torch.as_tensor(a, device=dev)  # noqa: E731
d_sptr, d_slen, d_val, d_col = t(sptr.astype(np.int64)), t(slen.astype(np.int32)), t(val), t(col)
x_old = np.random.default_rng(0).standard_normal(N)
x_new = np.zeros(Npad)
x_new[:N] = x_old[perm]
d_x, d_y = t(x_new), torch.zeros(Npad, dtype=torch.float64, device=dev)
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "libsell.so"))
ref = (A @ x_old)[perm]
for unroll in (4, 8, 16):
    for grid in (148 * 4, 148 * 8, 148 * 16):
        run = lambda: lib.sell_spmv(nsl, C.c_void_p(d_sptr.data_ptr()), C.c_void_p(d_slen.data_ptr()),  # noqa: E731
                                    C.c_void_p(d_val.data_ptr()), C.c_void_p(d_col.data_ptr()), C.c_void_p(d_x.data_ptr()),
                                    C.c_void_p(d_y.data_ptr()), unroll, grid, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        for _ in range(5):
            assert run() == 0
        torch.cuda.synchronize()
        err = np.abs(d_y.cpu().numpy()[:N] - ref).max() / np.abs(ref).max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print("unroll %2d grid %4d: %.1f us  (%.0f GB/s of 12 B/entry streamed)  err %.1e" % (unroll, grid, us, 12 * total / us / 1e3, err))
