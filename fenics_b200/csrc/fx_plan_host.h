// Host-side plan construction: CSR pattern (DOLFIN's full cell-block pattern, SparsityPatternBuilder:
// every (test local dof, trial local dof) pair of every cell, columns ascending -- SURVEY.md 8a-notes)
// and the per-cell scatter map.  Pure C++ (no CUDA) so the CPU tests exercise the same code.
#pragma once
#include <algorithm>
#include <functional>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

namespace fx {

struct HostPattern {
  int n = 0, nnz = 0, ld = 0, max_row = 0;
  std::vector<int> rowptr, col;
  std::vector<uint8_t> pos;  // [(a*ld+b)*nc + cell] = offset of column dof(b) within row dof(a)
  std::vector<int> dm_t;     // transposed dofmap [a*nc + cell]
  std::vector<int> rowblk;   // SpMV row blocks: rows [rowblk[k], rowblk[k+1]) hold <= SPMV_NNZB nnz, <= SPMV_ROWB rows
};

constexpr int SPMV_NNZB = 512;  // nonzeros staged in shared memory per row block (one warp each)
constexpr int SPMV_ROWB = 32;   // rows per row block (one lane each in the reduce phase)

inline void build_rowblocks(HostPattern& P) {
  P.rowblk.clear();
  P.rowblk.push_back(0);
  int start = 0;
  for (int r = 0; r < P.n; ++r) {
    const int nnz_if = P.rowptr[r + 1] - P.rowptr[start];
    if (nnz_if > SPMV_NNZB || r + 1 - start > SPMV_ROWB) {
      P.rowblk.push_back(r);  // close the block before row r (every row has <= 256 <= NNZB entries)
      start = r;
    }
  }
  P.rowblk.push_back(P.n);
}

inline void parallel_for(long n, const std::function<void(long, long)>& fn) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 16) nt = 16;
  if (n < 4096 || nt == 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  const long chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const long b = t * chunk, e = std::min(n, b + chunk);
    if (b < e) th.emplace_back(fn, b, e);
  }
  for (auto& t : th) t.join();
}

// cell_dofs: nc x ld row-major.  want_pattern=false builds only dm_t (spaces never used as a
// matrix space: Zd, Qt).
inline bool build_pattern(const int* cell_dofs, int nc, int ld, bool want_pattern, HostPattern& P,
                          std::string& err) {
  P.ld = ld;
  int n = 0;
  for (long i = 0; i < (long)nc * ld; ++i) {
    if (cell_dofs[i] < 0) { err = "negative dof index"; return false; }
    n = std::max(n, cell_dofs[i] + 1);
  }
  P.n = n;
  P.dm_t.resize((size_t)nc * ld);
  for (int c = 0; c < nc; ++c)
    for (int a = 0; a < ld; ++a) P.dm_t[(size_t)a * nc + c] = cell_dofs[(size_t)c * ld + a];
  if (!want_pattern) return true;
  // dof -> incident cells
  std::vector<int> iptr(n + 1, 0);
  for (long i = 0; i < (long)nc * ld; ++i) iptr[cell_dofs[i] + 1]++;
  for (int i = 0; i < n; ++i) iptr[i + 1] += iptr[i];
  std::vector<int> icell(iptr[n]), fill(iptr.begin(), iptr.end() - 1);
  for (int c = 0; c < nc; ++c)
    for (int a = 0; a < ld; ++a) icell[fill[cell_dofs[(size_t)c * ld + a]]++] = c;
  // pass 1: row lengths
  std::vector<int> len(n, 0);
  parallel_for(n, [&](long b, long e) {
    std::vector<int> tmp;
    for (long r = b; r < e; ++r) {
      tmp.clear();
      for (int k = iptr[r]; k < iptr[r + 1]; ++k) {
        const int* cd = cell_dofs + (size_t)icell[k] * ld;
        tmp.insert(tmp.end(), cd, cd + ld);
      }
      std::sort(tmp.begin(), tmp.end());
      len[r] = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  });
  P.rowptr.assign(n + 1, 0);
  long total = 0;
  int maxrow = 0;
  for (int r = 0; r < n; ++r) {
    total += len[r];
    maxrow = std::max(maxrow, len[r]);
    if (total > 2147483647L) { err = "nnz exceeds int32"; return false; }
    P.rowptr[r + 1] = (int)total;
  }
  P.nnz = (int)total;
  P.max_row = maxrow;
  if (maxrow > 256) { err = "row longer than 256 entries: uint8 scatter map overflow"; return false; }
  P.col.resize(total);
  parallel_for(n, [&](long b, long e) {
    std::vector<int> tmp;
    for (long r = b; r < e; ++r) {
      tmp.clear();
      for (int k = iptr[r]; k < iptr[r + 1]; ++k) {
        const int* cd = cell_dofs + (size_t)icell[k] * ld;
        tmp.insert(tmp.end(), cd, cd + ld);
      }
      std::sort(tmp.begin(), tmp.end());
      auto last = std::unique(tmp.begin(), tmp.end());
      std::copy(tmp.begin(), last, P.col.begin() + P.rowptr[r]);
    }
  });
  // scatter map
  P.pos.resize((size_t)nc * ld * ld);
  parallel_for(nc, [&](long cb, long ce) {
    for (long c = cb; c < ce; ++c) {
      const int* cd = cell_dofs + (size_t)c * ld;
      for (int a = 0; a < ld; ++a) {
        const int* rb = P.col.data() + P.rowptr[cd[a]];
        const int* re = P.col.data() + P.rowptr[cd[a] + 1];
        for (int b = 0; b < ld; ++b) {
          const int* it = std::lower_bound(rb, re, cd[b]);
          P.pos[((size_t)a * ld + b) * nc + c] = (uint8_t)(it - rb);
        }
      }
    }
  });
  build_rowblocks(P);
  return true;
}

}  // namespace fx
