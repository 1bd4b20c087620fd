// Device helpers shared by fx_la.cu and fx_krylov.cu: warp/block reductions and the CSR "stream"
// SpMV building block.
#pragma once
#include "fx_plan.h"
#include "fx_plan_host.h"  // SPMV_NNZB, SPMV_ROWB

namespace fx {

constexpr int LA_THREADS = 256;
constexpr int LA_WARPS = LA_THREADS / 32;
constexpr int SPMV_SMEM_DOUBLES = LA_WARPS * SPMV_NNZB;  // shared staging per CTA (32 KB)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum broadcast to every thread (fixed reduction tree => deterministic); sh: 33 doubles
__device__ __forceinline__ double block_sum_bcast(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < nw) ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
// NV sums at once (same barriers as one): sh must hold NV*33 doubles
template <int NV>
__device__ __forceinline__ void block_sum_bcast_n(double (&v)[NV], double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[i * 33 + w] = v[i];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = (lane < nw) ? sh[i * 33 + lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) sh[i * 33 + 32] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = sh[i * 33 + 32];
}

// streaming (read-once) loads: do not allocate in L1 so the x-gather lines stay resident
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// CSR matrix on the plan's pattern, optionally with a zero-compacted copy (see compact_zeros).
struct CsrView {
  int n;
  const int* rowptr;
  const int* col;
  const double* val;
  const int* rowblk;
  int nrowblk;
  // compacted copy (structural zeros of the DOLFIN cell-block pattern dropped); desc == null: not compacted
  int4* desc;   // [blocks] {first row, one past the last row, first entry, one past the last entry} of the row block
  int* re;      // [n] one past the last entry of row (rows of a block are stored back to back)
  int* col2;
  double* val2;
  // per-row mask (null = none): MASK_ELIM rows are left out of the Krylov iteration and recovered by
  // back_substitute() (requires A[active, eliminated] == 0); MASK_GHOST rows belong to another rank of a
  // partitioned run (empty in the compacted copy, their vector entries arrive by halo pushes);
  // MASK_IFACE marks owned rows some neighbour holds as ghosts
  const unsigned char* elim;
  const int* elim_rows;
  int nelim;
  // cost-balanced contiguous partition of the row blocks over the warps (balance_partition)
  int* blkprefix;  // [nrowblk + 1] exclusive prefix of the per-block cost
  int* scanpart;   // [2 * gridDim] scratch of the scans
  // re-blocking of the compacted matrix (compact_reblock): blk_target > 0 switches it on
  int blk_target;  // entries-per-block target T = SPMV_NNZB - (longest pattern row): a block never exceeds SPMV_NNZB
  int2* rowpre;    // [n] CTA-chunk-local exclusive prefixes (entries, weights) of the rows
  int* nblk;       // [1] number of blocks built
};
constexpr unsigned char MASK_ELIM = 1, MASK_GHOST = 2, MASK_IFACE = 4, MASK_INACTIVE = MASK_ELIM | MASK_GHOST;

// ---- mesh-partitioned runs: peer-mapped symmetric regions (one per GPU, same layout) -------------------
// Every rank's Krylov work vectors, cross-GPU reduction slots and barrier flags live in a region that all
// peers map over NVLink (fx_plan_set_peer_regions; torch symmetric memory on the host side).  The
// persistent Krylov kernels then do their own halo exchange (P2P stores of interface values straight into
// the neighbour's copy of the same work vector, issued by the thread that computed the value), their own
// cross-GPU reductions and barriers: no NCCL call and no host round trip inside a solve.
// value `val` of interface dof i of work vector j -> the ghost slot of every neighbour that reads it
__device__ __forceinline__ void push_iface(const DistDev& dd, int j, int i, double val) {
  int k = dd.s_first[i];
  for (;;) {
    const int pe = dd.s_peer[k];
    dd.region[pe & (DIST_MAXW - 1)][(long long)j * dd.nmax + dd.s_remote[k]] = val;
    if (pe & DIST_LAST) break;
    ++k;
  }
}

// Grid barrier for cooperative (co-resident) launches.  Unlike cooperative_groups' grid.sync(), whose
// waiting thread polls with ld.acquire -- every poll invalidates the SM's whole L1 (CCTL.IVALL), which
// evicts the x-gather lines of the CTAs still working on that SM -- the wait here polls with relaxed
// loads and acquires once.  Monotonic counter, zeroed by the host before every launch.
//
// DIST (mesh-partitioned run): the barrier spans the GPUs of the partition and carries up to RED_SLOTS
// reduction values with it.  Every CTA arrives at the local counter; when all have, thread 0 of CTA 0
// stores one LL line per value {lo, epoch, hi, epoch} into every peer's region -- data and flag travel
// together, so no fence separates them -- and meanwhile warp 1 of EVERY CTA polls the lines the peers
// store into this GPU's region (one lane per line, in parallel), acquires, and leaves the peers' values
// in shared memory.  A CTA passes when the local counter is full and every peer's line of this epoch is
// in: that implies the peer's CTAs all arrived, and (their system-scope fences precede their arrival) that
// the halo values they pushed are visible here.  Lines are double-buffered by epoch parity: a peer can
// be at most one barrier ahead.  Epochs continue across solves (persisted in the region).  A peer that
// never arrives trips a 10 s timeout: the barrier goes dead (all later waits fall through, together on
// all CTAs) and the solve reports converged = -4 instead of hanging the GPU.
template <bool DIST>
struct GridBarT {
  unsigned* ctr;
  unsigned target;
  const DistDev* dd;
  unsigned xepoch;
  bool dead;
  double* xs;  // shared memory [DIST_MAXW * RED_SLOTS]: the peers' values of the latest barrier
  __device__ __forceinline__ constexpr bool dist() const { return DIST; }
  __device__ __forceinline__ unsigned long long* state() const {
    return reinterpret_cast<unsigned long long*>(dd->region[dd->rank] + dist_state_off(dd->nmax));
  }
  __device__ __forceinline__ uint4* line(int dst, unsigned parity, int src, int i) const {
    return reinterpret_cast<uint4*>(dd->region[dst] + dist_xred_off(dd->nmax)) + ((parity * DIST_MAXW + src) * RED_SLOTS + i);
  }
  __device__ __forceinline__ void init(unsigned* c, const DistDev* d = nullptr, double* xs_shared = nullptr) {
    ctr = c; target = 0; dd = d; xepoch = 0; dead = false; xs = xs_shared;
    if (DIST) {
      xepoch = (unsigned)__ldcg(state());
      dead = __ldcg(state() + 1) != 0;
    }
  }
  // persist the epoch for the next solve (every exit path, block 0 / thread 0)
  __device__ __forceinline__ void save() const {
    if (DIST && blockIdx.x == 0 && threadIdx.x == 0) *state() = xepoch;
  }
  __device__ __forceinline__ void wait_local() {
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while ((int)(v - target) < 0);
  }
  // NV > 0: cur[0..NV) are this GPU's totals (device accumulators, complete once the local counter is full).
  // PUSHED: this CTA may have stored to peer memory since the previous barrier (system-scope release).
  template <int NV, bool PUSHED>
  __device__ __forceinline__ void sync_x(const double* cur) {
    target += gridDim.x;
    if (!DIST) {
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();  // release: everything this CTA wrote (the barrier above orders the other threads' stores)
        atomicAdd(ctr, 1u);
        wait_local();
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      }
      __syncthreads();
      return;
    }
    const unsigned e = ++xepoch, par = e & 1u;
    constexpr int NVS = NV > 0 ? NV : 1;
    const int L = (dd->world - 1) * NVS;
    __syncthreads();
    if (threadIdx.x == 0) {
      if (PUSHED) __threadfence_system(); else __threadfence();
      atomicAdd(ctr, 1u);
      if (blockIdx.x == 0) {
        wait_local();
        __threadfence_system();  // the local CTAs' releases are ordered before the lines below (cumulativity)
        double t[NVS];
#pragma unroll
        for (int i = 0; i < NVS; ++i) t[i] = (i < NV) ? __ldcg(cur + i) : 0.0;
        for (int q = 0; q < dd->world; ++q) {
          if (q == dd->rank) continue;
#pragma unroll
          for (int i = 0; i < NVS; ++i) {
            const unsigned lo = (unsigned)__double2loint(t[i]), hi = (unsigned)__double2hiint(t[i]);
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line(q, par, dd->rank, i)), "r"(lo), "r"(e),
                         "r"(hi), "r"(e) : "memory");
          }
        }
      }
      wait_local();
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } else if (threadIdx.x >= 32 && (int)threadIdx.x < 32 + L && !dead) {
      const int k = threadIdx.x - 32, qq = k / NVS, i = k - qq * NVS;
      const int q = qq + (qq >= dd->rank ? 1 : 0);
      const uint4* lp = line(dd->rank, par, q, i);
      unsigned long long* deadflag = state() + 1;
      unsigned long long t0 = 0;
      unsigned spins = 0, w0, w1, w2, w3;
      for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "l"(lp) : "memory");
        if (w1 == e && w3 == e) break;
        if ((++spins & 0x3ffu) == 0) {
          if (__ldcg(deadflag) != 0) { dead = true; break; }  // somebody gave up: all give up together
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > 10000000000ull) { dead = true; *deadflag = 1ull; break; }
        }
      }
      if (!dead) {
        unsigned f;  // acquire on the line's flag word: pairs with the peer's fence + line store
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(reinterpret_cast<const unsigned*>(lp) + 1) : "memory");
        xs[q * RED_SLOTS + i] = __hiloint2double((int)w2, (int)w0);
      }
    }
    dead = __syncthreads_or(dead);
  }
  __device__ __forceinline__ void sync() { sync_x<0, true>(nullptr); }
};
typedef GridBarT<false> GridBar;

// The warp processes row blocks k0, k0+kstep, ... < k1.
struct BlockRange {
  int k0, k1, kstep;
};
__device__ __forceinline__ BlockRange round_robin_range(const CsrView& A) {
  const int wpb = blockDim.x >> 5;
  return BlockRange{(int)(blockIdx.x * wpb + (threadIdx.x >> 5)), A.nrowblk, (int)(gridDim.x * wpb)};
}

// Drop exactly-zero entries, row block by row block, into (col2, val2) at the block's original offset.
// DOLFIN's pattern stores every (test dof, trial dof) pair of every cell whether or not the form
// couples them: 20-60 % of the stored values of the W and Zc matrices are exact
// zeros, which a Krylov solve would otherwise stream from HBM twice per iteration.
// With elimination on, eliminated rows become empty in the copy (never streamed); the return value
// counts non-zero couplings A[active, eliminated] seen by this thread (must total 0).
__device__ __forceinline__ int compact_zeros(const CsrView& A) {
  const int lane = threadIdx.x & 31;
  const BlockRange rg = round_robin_range(A);
  int bad = 0;
  for (int k = rg.k0; k < rg.k1; k += rg.kstep) {
    const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
    const int j0 = A.rowptr[r0];
    const int row = r0 + lane;
    const bool act = row < r1;
    int b = 0, e = 0, cnt = 0;
    if (act) {
      b = A.rowptr[row]; e = A.rowptr[row + 1];
      if (A.elim != nullptr && (A.elim[row] & MASK_INACTIVE)) {
        e = b;  // eliminated or ghost row: empty
      } else if (A.elim != nullptr) {
        for (int j = b; j < e; ++j) {
          const bool nz = A.val[j] != 0.0;
          cnt += nz;
          bad += (nz && (A.elim[A.col[j]] & MASK_ELIM));
        }
      } else {
        for (int j = b; j < e; ++j) cnt += (A.val[j] != 0.0);
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (act) {
      int dst = j0 + incl - cnt;
      A.re[row] = dst + cnt;
      for (int j = b; j < e; ++j) {
        const double v = A.val[j];
        if (v != 0.0) { A.val2[dst] = v; A.col2[dst] = A.col[j]; ++dst; }
      }
    }
    if (lane == 0) A.desc[k] = make_int4(r0, r1, j0, j0 + total);
    __syncwarp();
  }
  return bad;
}

// Compaction WITH re-blocking.  The plan's row blocks are cut on the PATTERN (<= SPMV_NNZB stored entries), but 36-58 %
// of the stored values of the W systems are structural zeros (and eliminated / ghost rows are empty), so a pattern block
// of a velocity system holds ~16 rows and ~300 real entries: the per-block latency chain of the SpMV is amortised over
// little data and half the lanes idle in the epilogue.  Here the blocks are rebuilt from the ACTUAL counts, inside the
// solve, with a grid-wide scan: row i gets the weight w_i = max(cnt_i, MINW) and block floor(W_i / T), W = exclusive
// prefix of the weights, T = SPMV_NNZB - (longest pattern row).  A block therefore holds < T + longest row <= SPMV_NNZB
// entries and at most T / MINW + 1 < 32 rows, and the compacted entries are stored densely at the exclusive prefix of
// the counts.  Two passes over the values (count, copy) with one grid barrier between them and one after;
// CTA b owns the contiguous rows [b L, (b+1) L).  Returns this thread's count of non-zero couplings A[active, eliminated];
// nb = number of blocks (same value in every thread).
constexpr int REBLOCK_MINW = 16;
template <class Bar>
__device__ __forceinline__ int compact_reblock(const CsrView& A, Bar& bar, int* sh_i, int& nb) {
  const int n = A.n, G = gridDim.x, T = A.blk_target;
  const int L = (n + G - 1) / G, c0 = min(n, (int)blockIdx.x * L), c1 = min(n, c0 + L);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int bad = 0, runE = 0, runW = 0;
  for (int base = c0; base < c1; base += blockDim.x) {  // CTA-uniform trip count
    const int row = base + threadIdx.x;
    int cnt = 0;
    if (row < c1 && !(A.elim != nullptr && (A.elim[row] & MASK_INACTIVE))) {
      const int b = A.rowptr[row], e = A.rowptr[row + 1];
      if (A.elim != nullptr) {
        for (int j = b; j < e; ++j) {
          const bool nz = A.val[j] != 0.0;
          cnt += nz;
          bad += (nz && (A.elim[A.col[j]] & MASK_ELIM));
        }
      } else {
        for (int j = b; j < e; ++j) cnt += (A.val[j] != 0.0);
      }
    }
    const int wt = row < c1 ? max(cnt, REBLOCK_MINW) : 0;
    // block-wide exclusive scan of (cnt, wt) in row order
    int ie = cnt, iw = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int te = __shfl_up_sync(0xffffffffu, ie, o), tw = __shfl_up_sync(0xffffffffu, iw, o);
      if (lane >= o) { ie += te; iw += tw; }
    }
    __syncthreads();
    if (lane == 31) { sh_i[w] = ie; sh_i[32 + w] = iw; }
    __syncthreads();
    int offe = 0, offw = 0, tote = 0, totw = 0;
    for (int i = 0; i < nw; ++i) {
      const int a = sh_i[i], c = sh_i[32 + i];
      if (i < w) { offe += a; offw += c; }
      tote += a; totw += c;
    }
    if (row < c1) {
      A.rowpre[row] = make_int2(runE + offe + ie - cnt, runW + offw + iw - wt);
      A.re[row] = cnt;
    }
    runE += tote; runW += totw;
  }
  if (threadIdx.x == 0) { A.scanpart[blockIdx.x] = runE; A.scanpart[G + blockIdx.x] = runW; }
  bar.sync();
  int offE = 0, offW = 0;
  for (int j = threadIdx.x; j < (int)blockIdx.x; j += blockDim.x) {
    offE += __ldcg(A.scanpart + j);
    offW += __ldcg(A.scanpart + G + j);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    offE += __shfl_down_sync(0xffffffffu, offE, o);
    offW += __shfl_down_sync(0xffffffffu, offW, o);
  }
  __syncthreads();
  if (lane == 0) { sh_i[w] = offE; sh_i[32 + w] = offW; }
  __syncthreads();
  offE = 0; offW = 0;
  for (int i = 0; i < nw; ++i) { offE += sh_i[i]; offW += sh_i[32 + i]; }
  __syncthreads();
  for (int base = c0; base < c1; base += blockDim.x) {
    const int row = base + threadIdx.x;
    if (row >= c1) continue;
    const int2 pr = A.rowpre[row];  // written by this same thread above
    const int cnt = A.re[row];
    const int E = offE + pr.x, W = offW + pr.y;
    const int blk = W / T, blk_next = (W + max(cnt, REBLOCK_MINW)) / T;
    int4* d = A.desc;
    if (row == 0) { d[0].x = 0; d[0].z = 0; }
    if (blk_next != blk || row == n - 1) {  // last row of its block: closes it and opens the next one
      d[blk].y = row + 1; d[blk].w = E + cnt;
      if (row + 1 < n) { d[blk_next].x = row + 1; d[blk_next].z = E + cnt; }
      else *A.nblk = blk + 1;
    }
    int dst = E;
    if (cnt > 0) {
      const int b = A.rowptr[row], e = A.rowptr[row + 1];
      for (int j = b; j < e; ++j) {
        const double v = A.val[j];
        if (v != 0.0) { A.val2[dst] = v; A.col2[dst] = A.col[j]; ++dst; }
      }
    }
    A.re[row] = E + cnt;
  }
  bar.sync();
  nb = __ldcg(A.nblk);
  return bad;
}

// A row block costs a warp its entries plus a fixed latency chain (descriptor -> values/columns -> x
// gather -> row sums -> epilogue), worth about this many entries of streaming.
constexpr int BLOCK_OVERHEAD_NNZ = 192;

// After compact_zeros + a grid barrier: split the row blocks into one contiguous range per warp of
// (nearly) equal cost and return this warp's range.  Round-robin dealing leaves some warps one block
// ahead of others (5 vs 6 blocks at n=192: every barrier waits ~20 % for the stragglers).
// Distributed scan: each CTA sums its chunk of blocks, a barrier, offsets from the CTAs before it, a
// barrier, then two binary searches per warp.  sh_i: 64 ints of shared memory.
template <class Bar>
__device__ __forceinline__ BlockRange balance_partition(const CsrView& A, Bar& bar, int* sh_i, int nb) {
  const int G = gridDim.x, wpb = blockDim.x >> 5;
  const int L = (nb + G - 1) / G;
  const int c0 = min(nb, (int)blockIdx.x * L), c1 = min(nb, c0 + L);
  // chunk-local inclusive scan by one warp (chunks are a few dozen blocks)
  if (threadIdx.x < 32) {
    int run = 0;
    for (int base = c0; base < c1; base += 32) {
      const int k = base + threadIdx.x;
      int v = 0;
      if (k < c1) {
        const int4 d = __ldcg(A.desc + k);
        v = d.w - d.z + BLOCK_OVERHEAD_NNZ;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)threadIdx.x >= o) v += t;
      }
      if (k < c1) A.blkprefix[k + 1] = run + v;  // inclusive, chunk-local
      run += __shfl_sync(0xffffffffu, v, 31);
    }
    if (threadIdx.x == 0) A.scanpart[blockIdx.x] = run;
  }
  bar.sync();
  // offset of my chunk and the grand total
  int off = 0, tot = 0;
  for (int j = threadIdx.x; j < G; j += blockDim.x) {
    const int v = __ldcg(A.scanpart + j);
    tot += v;
    if (j < (int)blockIdx.x) off += v;
  }
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_down_sync(0xffffffffu, off, o);
      tot += __shfl_down_sync(0xffffffffu, tot, o);
    }
    __syncthreads();
    if (lane == 0) { sh_i[w] = off; sh_i[32 + w] = tot; }
    __syncthreads();
    off = 0; tot = 0;
    for (int i = 0; i < wpb; ++i) { off += sh_i[i]; tot += sh_i[32 + i]; }
  }
  for (int k = c0 + threadIdx.x; k < c1; k += blockDim.x) A.blkprefix[k + 1] = __ldcg(A.blkprefix + k + 1) + off;
  if (blockIdx.x == 0 && threadIdx.x == 0) A.blkprefix[0] = 0;
  bar.sync();
  // warp w owns the blocks whose prefix lies in [w*tot/nw, (w+1)*tot/nw)
  const long long nw = (long long)G * wpb, w = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
  auto lower = [&](long long target) {
    int lo = 0, hi = nb;  // smallest k in [0, nb] with prefix[k] >= target
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((long long)__ldcg(A.blkprefix + mid) >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  const int k0 = (w == 0) ? 0 : lower((w * tot) / nw);
  const int k1 = (w == nw - 1) ? nb : lower(((w + 1) * tot) / nw);
  return BlockRange{k0, k1, 1};
}

// x[row] = (b[row] - sum_{j != row} A[row,j] x[j]) / A[row,row] for the eliminated rows, 8 lanes per row
// (exact for a trailing block with diagonal self-coupling, e.g. the DG0 DEVSS rows of the W systems).
// Returns the number of non-zero couplings between DIFFERENT eliminated rows seen by this thread.
__device__ __forceinline__ int back_substitute(const CsrView& A, double* x, const double* b) {
  int bad = 0;
  // (the DG0 rows hold 13 entries: 8 lanes keep twice as many rows in flight as 16 -- the loop is a chain of dependent
  // loads, and the node-block kernels run it with 512 threads per CTA)
  const int sub = threadIdx.x & 7;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, ng = (gridDim.x * blockDim.x) >> 3;
  const int rounds = (A.nelim + ng - 1) / ng;  // uniform trip count: the shuffles need whole warps
  for (int it = 0; it < rounds; ++it) {
    const int k = it * ng + g;
    const int row = k < A.nelim ? A.elim_rows[k] : -1;
    double s = 0.0, dg = 0.0;
    if (row >= 0)
      for (int j = A.rowptr[row] + sub; j < A.rowptr[row + 1]; j += 8) {
        const int c = A.col[j];
        const double v = A.val[j];
        if (c == row) dg = v;
        else if (v != 0.0) { s += v * x[c]; bad += (A.elim[c] & MASK_ELIM); }
      }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o, 8);
      dg += __shfl_xor_sync(0xffffffffu, dg, o, 8);
    }
    if (row >= 0 && sub == 0) x[row] = (b[row] - s) / dg;
  }
  return bad;
}

// CSR-stream SpMV at warp granularity (rows of 7-57 nonzeros are too short for a warp per row and a
// thread per row cannot coalesce): a warp takes a block of consecutive rows holding <= SPMV_NNZB
// nonzeros (<= 32 rows), streams val[j]*x[col[j]] into its private shared-memory slice with fully
// coalesced val/col loads, 8 independent loads per lane in flight, then one lane per row sums its
// slice and runs the fused epilogue epi(row, (A x)[row]).  Only __syncwarp() -- warps of a CTA drift
// apart, so streaming and reducing overlap on the SM.
// `x` may have been written earlier in the same (cooperative) kernel: plain coherent loads.
// This is the plain-CSR variant (stand-alone SpMV); the Krylov kernels use spmv_stream_c below.
template <class Epi>
__device__ __forceinline__ void spmv_stream(const CsrView& A, const BlockRange rg, const double* x, double* smem,
                                            Epi&& epi) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* prod = smem + wib * SPMV_NNZB;
  for (int k = rg.k0; k < rg.k1; k += rg.kstep) {
    const int r0 = A.rowblk[k], r1 = A.rowblk[k + 1];
    const int j0 = A.rowptr[r0], j1 = A.rowptr[r1];
    for (int jb = j0; jb < j1; jb += 256) {
      int c[8];
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = jb + u * 32 + lane;
        c[u] = (j < j1) ? ld_stream(A.col + j) : -1;
        v[u] = (j < j1) ? ld_stream(A.val + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c[u] >= 0) prod[jb - j0 + u * 32 + lane] = v[u] * x[c[u]];
    }
    __syncwarp();
    if (lane < r1 - r0) {
      const int row = r0 + lane;
      const int s0 = A.rowptr[row] - j0, s1 = A.rowptr[row + 1] - j0;
      double s = 0.0;
      for (int j = s0; j < s1; ++j) s += prod[j];
      epi(row, s);
    }
    __syncwarp();
  }
}

// Compacted variant inside the persistent Krylov kernels.  A row block costs a chain of dependent memory
// round trips, and with ~5 blocks per warp and SpMV that chain -- not bandwidth -- sets the pace, so it is
// kept short: ONE 16-byte descriptor per block (prefetched a block ahead), and everything that depends only
// on the row -- its end offset and the epilogue's operands, fetched by pre(row) -- is requested BEFORE the
// value/column stream so all of it is in flight together; row starts come from the neighbour lane.
//   pre(row)            -> Ops (plain struct of the vector entries the epilogue needs)
//   post(row, Ax, ops)  epilogue
#ifndef SPMV_INFLIGHT
#define SPMV_INFLIGHT 8  // value/column loads per lane in flight (x 32 lanes = entries per pass)
#endif
template <class Pre, class Post>
__device__ __forceinline__ void spmv_stream_c(const CsrView& A, const BlockRange rg, const double* x, double* smem,
                                              Pre&& pre, Post&& post) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* prod = smem + wib * SPMV_NNZB;
  const int* col = A.col2;
  const double* val = A.val2;
  int4 d = (rg.k0 < rg.k1) ? __ldcg(A.desc + rg.k0) : make_int4(0, 0, 0, 0);
  for (int k = rg.k0; k < rg.k1; k += rg.kstep) {
    const int kn = k + rg.kstep;
    const int4 dn = (kn < rg.k1) ? __ldcg(A.desc + kn) : make_int4(0, 0, 0, 0);
    const int r0 = d.x, nr = d.y - d.x, j0 = d.z, j1 = d.w;
    const int row = r0 + lane;
    const bool act = lane < nr;
    const int e1 = act ? __ldcg(A.re + row) : j0;
    decltype(pre(0)) ops{};
    if (act) ops = pre(row);
    for (int jb = j0; jb < j1; jb += 32 * SPMV_INFLIGHT) {
      int c[SPMV_INFLIGHT];
      double v[SPMV_INFLIGHT];
#pragma unroll
      for (int u = 0; u < SPMV_INFLIGHT; ++u) {
        const int j = jb + u * 32 + lane;
        c[u] = (j < j1) ? ld_stream(col + j) : -1;
        v[u] = (j < j1) ? ld_stream(val + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < SPMV_INFLIGHT; ++u)
        if (c[u] >= 0) prod[jb - j0 + u * 32 + lane] = v[u] * x[c[u]];
    }
    __syncwarp();
    const int s1 = e1 - j0;
    int s0 = __shfl_up_sync(0xffffffffu, s1, 1);
    if (lane == 0) s0 = 0;
    if (act) {
      double s = 0.0;
      for (int j = s0; j < s1; ++j) s += prod[j];
      post(row, s, ops);
    }
    __syncwarp();
    d = dn;
  }
}

}  // namespace fx
