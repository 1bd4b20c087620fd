"""CPU: depth of the dependency graph natural-ordering ILU(0) walks (the level schedule of the strictly lower triangle of
DOLFIN's cell-block pattern, in the library's dof numbering) for the W, Zc and Q spaces of an n x n cavity mesh -- the
lower bound on flag hand-overs per triangular solve of csrc/fx_ilu.cu, quoted in DESIGN.md section 4.

    python tools/ilu_levels.py 48 96
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fenics_b200 import mesh as pmesh, spaces  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [48]:
    m = pmesh.Skew_Mesh(n, 1, 1)
    seen = set()
    for s in spaces.function_spaces(m):
        if s is None or s.name not in ("W", "Zc", "Q") or s.name in seen:
            continue
        seen.add(s.name)
        cd, N = s.cell_dofs, s.dim()
        ld = cd.shape[1]
        A = sp.csr_matrix((np.ones(cd.size * ld), (np.repeat(cd, ld, axis=1).ravel(), np.tile(cd, (1, ld)).ravel())), shape=(N, N))
        A.sum_duplicates()
        A.sort_indices()
        rp, ci = A.indptr, A.indices
        lev = np.zeros(N, dtype=np.int64)
        for i in range(N):
            lo = ci[rp[i]:rp[i + 1]]
            lo = lo[lo < i]
            lev[i] = 1 + (lev[lo].max() if len(lo) else 0)
        print("n=%d %-2s rows=%d nnz=%d levels=%d rows/level=%.1f" % (n, s.name, N, A.nnz, lev.max(), N / lev.max()))
