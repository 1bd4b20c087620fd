// Host-side construction of the NODE-BLOCK solver layout of a space (pure C++, no CUDA: the CPU tests build it too).
//
// The DOLFIN-pattern CSR matrix (bit-exact sparsity, arbitrary dof numbering) is the product's external format; it is
// a poor streaming format for the Krylov kernels: 12 bytes and one scattered 8-byte x-gather per scalar entry, 20-60 %
// explicit zeros, rows of 7-57 entries.  All P2/P1 systems of the time loops are node-blocked, though: the bs
// components living on one mesh node (ux, uy of the velocity; txx, txy, tyy of the stress) have IDENTICAL column
// sets.  The solver layout stores one column index per bs x bs node block and keeps every Krylov work vector in an
// internal node-major numbering (bs components of a node adjacent, padded to bp = 2 / 4 doubles), so the x-gather
// is ONE aligned 16/32-byte load per block instead of bs^2 scattered 8-byte loads:
//     bytes per scalar entry 12 -> 9 (bs = 2) / 8.4 (bs = 3); L1 wavefronts of the gather / bs^2.
// The structure (block columns, the position of every block value inside the DOLFIN CSR value array, chunking and
// the per-warp work split) depends only on the plan, so it is built ONCE; a solve only gathers the current values
// through `src` (bval[k] = val[src[k]]) -- no scans, no grid barriers in the set-up.
//
// Value layout ("group planes"): block entries are numbered e = 0 .. ne-1 in block-row order; group g = e / 32 holds
// 32 entries.  The bs^2 values f = ci*bs + cj of an entry are stored as double2 planes (f = 2q, 2q+1) and, when bs^2
// is odd, one trailing double plane, so that a warp's loads are contiguous 512-byte (double2) / 256-byte runs:
//     pos(e, f) = g*32*bs^2 + (j < 2*NP2 ? (j/2)*64 + 2*l + (j&1) : NP2*64 + l),   l = e % 32, NP2 = bs^2 / 2,
// j = nb_slot(bs, f) the plane slot of value f (see NB_ORDER below).
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "fx_plan_host.h"

namespace fx {

constexpr int nb_bp(int bs) { return bs == 3 ? 4 : bs; }                         // padded components per node
constexpr int nb_chunk(int bs) { return bs == 1 ? 512 : bs == 2 ? 256 : 128; }   // block entries per chunk (<=)
constexpr int nb_rows(int bs) { return 32 / nb_bp(bs); }                         // block rows per chunk (<=)
// The layout splits the work of a CTA into NB_SLOTS contiguous chunk ranges; a kernel with W warps per CTA gives each warp
// NB_SLOTS / W consecutive slots, so every kernel chooses its own CTA size on the SAME layout:
//   BiCGStab: 16 warps x 128 registers -- the kernel keeps its whole working set in registers (at 32 warps x 64
//     registers it spilled 0.6-1.1 KB per thread inside the product loops).  Per iteration at n = 192, 32 / 24 / 20 / 16
//     / 12 warps: W 49.1 / 40.0 / 38.1 / 33.7 / 40.5 us, Zc 86.6 / 72.4 / 71.1 / 72.5 / 86.0 us (tools/build_variants.sh);
//   CG and the stand-alone product: 32 warps (no spills at 64 registers; more loads in flight).
constexpr int NB_SLOTS = 32;
#ifndef FX_NB_WARPS
#define FX_NB_WARPS 16
#endif
#ifndef FX_NB_WARPS_CG
#define FX_NB_WARPS_CG 32
#endif
constexpr int NB_WARPS_CG = FX_NB_WARPS_CG;
static_assert(NB_SLOTS % FX_NB_WARPS == 0 && NB_SLOTS % FX_NB_WARPS_CG == 0, "warps per CTA must divide the slot count");
constexpr int NB_WARPS = FX_NB_WARPS;                                                     // warps per CTA of k_bicgstab_nb

// Storage order of the bs^2 values of a block: value f = ci*bs + cj lives in plane slot nb_slot(bs, f).  The order pairs
// up the values that vanish TOGETHER in the scripts' operators, so a whole double2 plane can be skipped by the product
// when a matrix has none of them (the per-solve plane mask of nb_gather_values):
//   bs = 3 (stress):  (xx,xx)(xy,xy) | (yy,yy)(xx,xy) | (xy,xx)(xy,yy) | (xx,yy)(yy,xx) | (yy,xy)
//                     the upper-convected operators never couple xx with yy (2 of 9 values), a mass matrix has the
//                     diagonal only (planes 3-5 skipped: 4 of 9 values streamed)
//   bs = 2 (velocity): (x,x)(y,y) | (x,y)(y,x)
// plane slot -> f: {0,4,8,1,3,5,2,6,7} (bs = 3), {0,3,1,2} (bs = 2); written without a table so device code can call it
constexpr int nb_f_of_slot(int bs, int j) {
  return bs == 3 ? (j == 0 ? 0 : j == 1 ? 4 : j == 2 ? 8 : j == 3 ? 1 : j == 4 ? 3 : j == 5 ? 5 : j == 6 ? 2 : j == 7 ? 6 : 7)
       : bs == 2 ? (j == 0 ? 0 : j == 1 ? 3 : j == 2 ? 1 : 2)
                 : j;
}
constexpr int nb_slot(int bs, int f) {
  for (int j = 0; j < bs * bs; ++j)
    if (nb_f_of_slot(bs, j) == f) return j;
  return -1;
}

inline long long nb_pos(long long e, int f, int bs) {
  const int b2 = bs * bs, np2 = b2 / 2, j = nb_slot(bs, f);
  const long long g = e >> 5;
  const int l = (int)(e & 31);
  return g * 32 * b2 + (j < 2 * np2 ? (j >> 1) * 64 + 2 * l + (j & 1) : np2 * 64 + l);
}

struct NbHost {
  int bs = 0, bp = 0, nn = 0, N = 0, nchunks = 0, grid = 0;
  int nn_chunked = 0;            // block rows [0, nn_chunked) are covered by chunks (trailing ghost rows are not)
  long long ne = 0, nslots = 0;  // block entries; value slots = groups * 32 * bs^2
  std::vector<int> ext_of_int;   // [N]   external dof of an internal index (-1: padding)
  std::vector<int> int_of_ext;   // [n]   internal index of an external dof (-1: not in a node block)
  std::vector<int> bcol;         // [ne]  block column = internal node index
  std::vector<int> src;          // [nslots] index into the DOLFIN CSR value array (-1: structural zero / padding)
  std::vector<int> bre;          // [nn]  one past the last block entry of the block row
  std::vector<int> diag;         // [nn]  block entry of the diagonal block (-1: empty row)
  std::vector<int> desc;         // [4*nchunks] {first block row, block rows, first entry, one past the last entry}
  std::vector<int> wrange;       // [grid*NB_SLOTS + 1] chunk range of every warp slot
  std::vector<int> zchk;         // CSR value positions A[active, eliminated] that must be exactly zero
  std::vector<uint8_t> imask;    // [N] row mask in internal numbering (MASK_GHOST / MASK_IFACE bits; 0 for padding)
};

// local layout of the spaces: local dof (node a, component c) = c * nloc + a (UFC component-blocked order);
// trailing local dofs (W's three DG0 components) belong to no node block
inline bool nb_space_layout(int space, int& bs, int& nloc) {
  switch (space) {
    case 0: bs = 2; nloc = 6; return true;  // S_W : P2 velocity (+ DG0 DEVSS tensor, must be eliminated)
    case 1: bs = 3; nloc = 6; return true;  // S_ZC: P2 symmetric stress
    case 2: bs = 1; nloc = 3; return true;  // S_Q : P1
    case 5: bs = 2; nloc = 6; return true;  // S_V : P2 velocity
    case 6: bs = 1; nloc = 6; return true;  // S_VS: scalar P2
    default: return false;
  }
}

// dm_t: transposed dofmap [ld][nc]; mask: null or [n] with bit 0 = eliminated, bit 1 = ghost, bit 2 = interface.
// Returns false with `why` when the space has no node-block structure the solver can use (callers fall back to the
// scalar CSR path).
// ghost_owner: null or [n] = 1 + owner rank of a ghost dof (0: owned).  Ghost nodes are numbered LAST, grouped by owner
// and in ascending external order inside a group -- the order in which the owner walks its interface nodes -- so the
// owner's P2P stores of interface values land on consecutive addresses (coalesced NVLink writes).
inline bool nb_build(const int* dm_t, int nc, int ld, int bs, int nloc, int n, const int* rowptr, const int* col,
                     const uint8_t* mask, int num_sms, NbHost& H, std::string& why,
                     const uint8_t* ghost_owner = nullptr) {
  const int bp = nb_bp(bs), b2 = bs * bs;
  H = NbHost();
  H.bs = bs; H.bp = bp;
  if (bs * nloc > ld) { why = "local layout does not fit the dofmap"; return false; }
  std::vector<int> head(n, -1), comp(n, 0);
  for (int k = 0; k < bs; ++k)
    for (int a = 0; a < nloc; ++a) {
      const int* h0 = dm_t + (size_t)a * nc;
      const int* dk = dm_t + (size_t)(k * nloc + a) * nc;
      for (int c = 0; c < nc; ++c) {
        const int d = dk[c], h = h0[c];
        if (head[d] >= 0 && (head[d] != h || comp[d] != k)) { why = "dofs are not node-blocked"; return false; }
        head[d] = h; comp[d] = k;
      }
    }
  // nodes in the order of their first component's external index (inherits the locality of the external numbering)
  std::vector<int> rank(n, -1);
  int nn = 0;
  {
    std::vector<int> heads;
    for (int d = 0; d < n; ++d)
      if (head[d] == d) heads.push_back(d);
    if (ghost_owner != nullptr)
      std::stable_sort(heads.begin(), heads.end(), [&](int a, int b) { return ghost_owner[a] < ghost_owner[b]; });
    for (int h : heads) rank[h] = nn++;
  }
  for (int d = 0; d < n; ++d) {
    const uint8_t m = mask ? mask[d] : 0;
    if (head[d] < 0 && !(m & 1)) { why = "a dof outside the node blocks is not eliminated"; return false; }
    if (head[d] >= 0 && (m & 1)) { why = "a node-block dof is eliminated"; return false; }
  }
  if (nn == 0) { why = "no nodes"; return false; }
  H.nn = nn; H.N = nn * bp;
  H.ext_of_int.assign((size_t)H.N, -1);
  H.int_of_ext.assign((size_t)n, -1);
  H.imask.assign((size_t)H.N, 0);
  for (int d = 0; d < n; ++d)
    if (head[d] >= 0) {
      const int i = rank[head[d]] * bp + comp[d];
      H.int_of_ext[d] = i;
      H.ext_of_int[i] = d;
      H.imask[i] = mask ? (mask[d] & 6) : 0;
    }
  std::vector<int> ext0(nn);  // external dof of component 0 of every node
  for (int d = 0; d < n; ++d)
    if (head[d] == d) ext0[rank[d]] = d;
  for (int i = 0; i < nn; ++i)
    for (int k = 0; k < bs; ++k) {
      if (H.ext_of_int[(size_t)i * bp + k] < 0) { why = "a node lacks a component"; return false; }
      if ((H.imask[(size_t)i * bp + k] & 2) != (H.imask[(size_t)i * bp] & 2)) { why = "a node is partly ghost"; return false; }
    }
  // block rows: neighbours of node i = nodes whose component-0 dof is a column of row ext0[i]; ghost nodes: empty rows
  std::vector<int> len(nn, 0);
  parallel_for(nn, [&](long b, long e) {
    for (long i = b; i < e; ++i) {
      if (H.imask[(size_t)i * bp] & 2) continue;
      const int r = ext0[i];
      int cnt = 0;
      for (int j = rowptr[r]; j < rowptr[r + 1]; ++j) cnt += (head[col[j]] == col[j]);
      len[i] = cnt;
    }
  });
  H.bre.resize(nn);
  long long ne = 0;
  int maxlen = 0;
  for (int i = 0; i < nn; ++i) {
    ne += len[i];
    maxlen = std::max(maxlen, len[i]);
    if (ne > 2000000000LL / b2) { why = "too many block entries"; return false; }
    H.bre[i] = (int)ne;
  }
  if (maxlen > nb_chunk(bs)) { why = "a block row exceeds the chunk size"; return false; }
  H.ne = ne;
  const long long ngroups = (ne + 31) / 32;
  H.nslots = ngroups * 32 * b2;
  H.bcol.assign((size_t)std::max<long long>(ne, 1), 0);
  H.src.assign((size_t)std::max<long long>(H.nslots, 1), -1);
  H.diag.assign(nn, -1);
  std::vector<std::vector<int>> zparts(64);
  const long chunk = (nn + 63) / 64;
  parallel_for(64, [&](long pb, long pe) {
    std::vector<std::pair<int, int>> nb;  // (rank, external comp-0 dof)
    for (long part = pb; part < pe; ++part) {
      std::vector<int>& zl = zparts[part];
      for (long i = part * chunk; i < std::min<long>(nn, (part + 1) * chunk); ++i) {
        if (len[i] == 0) continue;
        const int r = ext0[i];
        nb.clear();
        for (int j = rowptr[r]; j < rowptr[r + 1]; ++j)
          if (head[col[j]] == col[j]) nb.emplace_back(rank[col[j]], col[j]);
        std::sort(nb.begin(), nb.end());
        long long e = (long long)H.bre[i] - len[i];
        for (const auto& pr : nb) {
          H.bcol[(size_t)e] = pr.first;
          if (pr.first == (int)i) H.diag[i] = (int)e;
          for (int ci = 0; ci < bs; ++ci) {
            const int row = H.ext_of_int[(size_t)i * bp + ci];
            const int* rb = col + rowptr[row];
            const int* re = col + rowptr[row + 1];
            for (int cj = 0; cj < bs; ++cj) {
              const int cdof = H.ext_of_int[(size_t)pr.first * bp + cj];
              const int* it = std::lower_bound(rb, re, cdof);
              H.src[(size_t)nb_pos(e, ci * bs + cj, bs)] = (it != re && *it == cdof) ? (int)(it - col) : -1;
            }
          }
          ++e;
        }
        // couplings of the active rows to eliminated dofs must vanish (block-triangular elimination)
        for (int ci = 0; ci < bs; ++ci) {
          const int row = H.ext_of_int[(size_t)i * bp + ci];
          for (int j = rowptr[row]; j < rowptr[row + 1]; ++j)
            if (head[col[j]] < 0) zl.push_back(j);
        }
      }
    }
  });
  for (auto& z : zparts) H.zchk.insert(H.zchk.end(), z.begin(), z.end());
  // work split: every warp of the grid gets a contiguous range of block rows holding ~ne/warps entries, cut into
  // chunks of <= nb_rows block rows and <= nb_chunk entries of (nearly) equal size
  const int CH = nb_chunk(bs), RM = nb_rows(bs);
  const long long per_warp_min = CH / 2;
  int grid = (int)std::min<long long>(num_sms, std::max<long long>(1, (ne + NB_SLOTS * per_warp_min - 1) / (NB_SLOTS * per_warp_min)));
  grid = std::max(grid, std::min(num_sms, (nn + NB_SLOTS * RM - 1) / (NB_SLOTS * RM)));  // rows without entries still need lanes
  H.grid = grid;
  const int nw = grid * NB_SLOTS;
  H.wrange.assign((size_t)nw + 1, 0);
  // Rows that take part: with ghost_owner the ghost nodes sit at the END of the numbering and get NO chunk at all --
  // their block rows are empty, nothing of a product or a vector pass ever touches them (the kernels zero their
  // entries of the work vectors once); a chunk of 16 empty rows would still cost a warp its ~1.5 us latency chain,
  // and a few thousand trailing ghost rows made the last warps of a partitioned run 2x late (measured).  Without
  // ghost_owner (ghost rows interleaved, host replay tests) every row is chunked, an empty one at cost 1.
  int nrows = nn;
  if (ghost_owner != nullptr)
    while (nrows > 0 && (H.imask[(size_t)(nrows - 1) * bp] & 2)) --nrows;
  H.nn_chunked = nrows;
  std::vector<long long> cpre((size_t)nrows + 1, 0);
  for (int i = 0; i < nrows; ++i) cpre[i + 1] = cpre[i] + std::max(len[i], 1);
  const long long ctot = cpre[nrows];
  int row = 0;
  for (int w = 0; w < nw; ++w) {
    const long long target = (ctot * (w + 1)) / nw;
    int rend = (int)(std::lower_bound(cpre.begin() + row, cpre.end(), target) - cpre.begin());
    if (w == nw - 1) rend = nrows;
    rend = std::max(rend, row);
    // chunks of rows [row, rend)
    const long long cost = cpre[rend] - cpre[row];
    if (rend > row) {
      const int nch = (int)std::max<long long>(std::max<long long>((cost + CH - 1) / CH, (rend - row + RM - 1) / RM), 1);
      int r0 = row;
      for (int k = 0; k < nch && r0 < rend; ++k) {
        const long long goal = cpre[row] + (cost * (k + 1)) / nch;
        int r1 = r0;
        long long ent = 0;
        while (r1 < rend && r1 - r0 < RM && ent + len[r1] <= CH && (cpre[r1] < goal || r1 == r0)) { ent += len[r1]; ++r1; }
        if (k == nch - 1 && r1 < rend) { --k; }  // hard limits cut the last chunk short: keep going
        H.desc.push_back(r0); H.desc.push_back(r1 - r0);
        H.desc.push_back(H.bre[r0] - len[r0]); H.desc.push_back(r1 > r0 ? H.bre[r1 - 1] : H.bre[r0] - len[r0]);
        r0 = r1;
      }
    }
    H.wrange[w + 1] = (int)(H.desc.size() / 4);
    row = rend;
  }
  H.nchunks = (int)(H.desc.size() / 4);
  return true;
}

}  // namespace fx
