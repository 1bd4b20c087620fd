// Device side of the node-block solver layout (fx_nb_host.h): value gather and the block CSR-stream SpMV.
#pragma once
#include "fx_la_common.cuh"
#include "fx_nb_host.h"

// tuning knob (tools/build_variants.sh): independent index -> value -> store chains per thread of the value gather
#ifndef FX_NB_GU
#define FX_NB_GU 4
#endif

namespace fx {

struct NbView {
  int nn, N, nchunks, nzchk;
  long long ne, nslots;
  const int4* desc;
  const int* wrange;
  const int* bre;
  const int* bcol;
  const int* src;
  const int* diag;
  const int* ext_of_int;
  const int* zchk;
  const unsigned char* imask;
  double* bval;  // [nslots] gathered per solve
  int l2policy;  // L2 eviction priority of the value / column stream: 0 normal, 1 evict_first, 2 evict_last
  int dual;      // BiCGStab: 1 = second product gathers r - alpha v on the fly (no s vector, 3 barriers per iteration)
  // partitioned run: halo push lists in internal numbering (null: not available)
  const int* s_first;   // [N] first (dof, destination) pair of an interface entry (pairs shared with DistDev::s_peer)
  const int* s_remote;  // [nsend] the destination's INTERNAL index of the ghost entry
  const int* s_local;   // [nsend] my INTERNAL index of the pair's dof (-1: outside the node blocks)
  unsigned* pmask;      // [1] plane slots with a non-zero value in the current matrix (zeroed before the launch)
  int planemask;        // 0: products load every plane (A/B switch FX_NB_PLANEMASK)
};

// The block values and columns are read once per product; the work vectors (a few MB) are re-read every pass.  With
// evict_first on the stream the vectors stay L2-resident while 130 MB of matrix (Zc) pass through the 126 MB L2.
// A matrix that FITS beside the vectors (W at n = 192: 61 MB of blocks) is streamed with evict_last instead and then
// served from L2 on the second product of every iteration (stand-alone product 12.7 -> 10.2 us, iteration 48.6 -> 46.9).
__device__ __forceinline__ unsigned long long nb_make_policy(int mode) {
  unsigned long long pol;
  if (mode == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (mode == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld_stream2(const double* p, unsigned long long pol) {
  double2 t;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(t.x), "=d"(t.y) : "l"(p), "l"(pol));
  return t;
}
__device__ __forceinline__ double ld_stream1(const double* p, unsigned long long pol) {
  double t;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(t) : "l"(p), "l"(pol));
  return t;
}
__device__ __forceinline__ int ld_stream1(const int* p, unsigned long long pol) {
  int t;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(t) : "l"(p), "l"(pol));
  return t;
}

template <int BS>
__device__ __forceinline__ long long nb_pos_d(int e, int f) {
  constexpr int B2 = BS * BS, NP2 = B2 / 2;
  const int j = nb_slot(BS, f);
  const long long g = e >> 5;
  const int l = e & 31;
  return g * 32 * B2 + (j < 2 * NP2 ? (j >> 1) * 64 + 2 * l + (j & 1) : NP2 * 64 + l);
}

// bval[k] = val[src[k]] for the whole grid (coalesced index reads and value writes, the DOLFIN CSR values are
// gathered); returns this thread's count of non-zero couplings A[active, eliminated] (must total 0)
// Also ORs into *B.pmask the plane slots that hold at least one non-zero value in THIS matrix (bit j = plane slot j):
// the products skip the double2 planes whose two slots are both clear.
template <int BS>
__device__ __forceinline__ int nb_gather_values(const NbView& B, const double* __restrict__ val) {
  constexpr int B2 = BS * BS, NP2 = B2 / 2;
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, gs = (long long)gridDim.x * blockDim.x;
  unsigned seen = 0;
  // GU independent (index -> value -> store) chains per thread: the gather is latency bound, not bandwidth bound
  constexpr int GU = FX_NB_GU;
  for (long long k0 = gt; k0 < B.nslots; k0 += GU * gs) {
    int s[GU];
    double x[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const long long k = k0 + u * gs;
      s[u] = k < B.nslots ? ld_stream(B.src + k) : -1;
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) x[u] = s[u] >= 0 ? __ldg(val + s[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const long long k = k0 + u * gs;
      if (k < B.nslots) {
        B.bval[k] = x[u];
        if (x[u] != 0.0) {
          const int off = (int)(k % (32 * B2));
          seen |= 1u << (off < 2 * NP2 * 32 ? 2 * (off / 64) + (off & 1) : B2 - 1);
        }
      }
    }
  }
  seen = __reduce_or_sync(0xffffffffu, seen);
  if ((threadIdx.x & 31) == 0 && seen) atomicOr(B.pmask, seen);
  int bad = 0;
  for (long long k = gt; k < B.nzchk; k += gs) bad += (__ldg(val + ld_stream(B.zchk + k)) != 0.0);
  return bad;
}

// 1 / diagonal of internal row i (Jacobi); rows of empty block rows and padding slots: 0
template <int BS>
__device__ __forceinline__ double nb_diag_inv(const NbView& B, int i, int pc) {
  constexpr int BP = nb_bp(BS);
  const int node = i / BP, c = i - node * BP;
  if (c >= BS) return 0.0;
  const int e = B.diag[node];
  if (e < 0) return 0.0;
  if (pc != FX_PC_JACOBI && pc != FX_PC_JACOBI_COARSE) return 1.0;
  const double dg = B.bval[nb_pos_d<BS>(e, c * BS + c)];
  return dg != 0.0 ? 1.0 / dg : 1.0;
}

// Block CSR-stream SpMV over the chunks [k0, k1) of this warp.
//   stream phase: lane <-> block entry: bs^2 values (coalesced double2 planes, L1-bypassing), one block column, ONE
//                 aligned gather of the bs components of x, bs partial sums to the warp's shared-memory slice;
//   reduce phase: lane <-> (block row, component) = internal row index: sums its slice entries and runs the fused
//                 epilogue post(i, (A x)_i, pre(i)); pre(i) is issued before the stream so its loads overlap it.
// Padding components (bs = 3 -> 4) run the epilogue with (A x)_i = 0 so every work vector stays zero there.
// smem: this warp's slice, nb_chunk(BS) * BS doubles.
// tuning knobs (tools/build_variants.sh): rounds of 32 entries a lane keeps in flight per pass
#ifndef FX_NB_U2
#define FX_NB_U2 2
#endif
#ifndef FX_NB_U3
#define FX_NB_U3 1
#endif
template <int BS>
struct NbU { static constexpr int v = BS == 1 ? 8 : BS == 2 ? FX_NB_U2 : FX_NB_U3; };  // groups of 32 entries in flight per pass

// DUALG: the product's input is x + coef * x2, formed on the fly from two gathers.
// pm: plane mask of the matrix (nb_gather_values); double2 planes without a non-zero value are not loaded.
template <int BS, bool DUALG = false, class Pre, class Post>
__device__ __forceinline__ void spmv_nb(const NbView& B, int k0, int k1, const double* x, double* prod,
                                        unsigned long long pol, unsigned pm, Pre&& pre, Post&& post,
                                        const double* x2 = nullptr, double coef = 0.0) {
  constexpr int BP = nb_bp(BS), B2 = BS * BS, NP2 = B2 / 2, U = NbU<BS>::v;
  const int lane = threadIdx.x & 31;
  const int rl = lane / BP, cl = lane - rl * BP;
  int4 d = (k0 < k1) ? __ldcg(B.desc + k0) : make_int4(0, 0, 0, 0);
  for (int k = k0; k < k1; ++k) {
    const int4 dn = (k + 1 < k1) ? __ldcg(B.desc + k + 1) : make_int4(0, 0, 0, 0);
    const int r0 = d.x, nr = d.y, e0 = d.z, e1 = d.w;
    const bool act = rl < nr;
    const int irow = (r0 + rl) * BP + cl;
    const int myend = act ? __ldcg(B.bre + r0 + rl) : e0;
    decltype(pre(0)) ops{};
    if (act) ops = pre(irow);
    for (int eb = e0; eb < e1; eb += 32 * U) {
      double v[U][B2];
      int c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = eb + u * 32 + lane;
        c[u] = -1;
        if (e < e1) {
          c[u] = ld_stream1(B.bcol + e, pol);
          const double* base = B.bval + (long long)(e >> 5) * (32 * B2) + 2 * (e & 31);
#pragma unroll
          for (int q = 0; q < NP2; ++q) {
            double2 t = make_double2(0.0, 0.0);
            if ((pm >> (2 * q)) & 3u) t = ld_stream2(base + q * 64, pol);  // warp-uniform
            v[u][nb_f_of_slot(BS, 2 * q)] = t.x; v[u][nb_f_of_slot(BS, 2 * q + 1)] = t.y;
          }
          if (B2 & 1) {
            double t = 0.0;
            if ((pm >> (B2 - 1)) & 1u) t = ld_stream1(B.bval + (long long)(e >> 5) * (32 * B2) + NP2 * 64 + (e & 31), pol);
            v[u][nb_f_of_slot(BS, B2 - 1)] = t;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (c[u] < 0) continue;
        double xv[BP];
        if (BS == 1) {
          xv[0] = x[c[u]];
          if (DUALG) xv[0] += coef * x2[c[u]];
        } else if (BP == 4) {  // one 256-bit load per node (sm_100: LDG.256) instead of two 128-bit ones: the gather's
                               // L1 wavefronts, not HBM, bound the 3 x 3-block product
          asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(xv[0]), "=d"(xv[1]), "=d"(xv[2]), "=d"(xv[3])
                       : "l"(x + (long long)c[u] * 4) : "memory");
          if (DUALG) {
            double y0, y1, y2, y3;
            asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(y0), "=d"(y1), "=d"(y2), "=d"(y3)
                         : "l"(x2 + (long long)c[u] * 4) : "memory");
            xv[0] += coef * y0; xv[1] += coef * y1; xv[2] += coef * y2; xv[3] += coef * y3;
          }
        } else {
#pragma unroll
          for (int h = 0; h < BP / 2; ++h) {
            const double2 t = *reinterpret_cast<const double2*>(x + (long long)c[u] * BP + 2 * h);
            xv[2 * h] = t.x; xv[2 * h + 1] = t.y;
            if (DUALG) {
              const double2 t2 = *reinterpret_cast<const double2*>(x2 + (long long)c[u] * BP + 2 * h);
              xv[2 * h] += coef * t2.x; xv[2 * h + 1] += coef * t2.y;
            }
          }
        }
        double* dst = prod + (eb - e0 + u * 32 + lane) * BS;
        double y[BS];
#pragma unroll
        for (int ci = 0; ci < BS; ++ci) {
          y[ci] = 0.0;
#pragma unroll
          for (int cj = 0; cj < BS; ++cj) y[ci] += v[u][ci * BS + cj] * xv[cj];
        }
        if (BS == 2) {
          *reinterpret_cast<double2*>(dst) = make_double2(y[0], y[BS - 1]);
        } else {
#pragma unroll
          for (int ci = 0; ci < BS; ++ci) dst[ci] = y[ci];
        }
      }
    }
    __syncwarp();
    const int s1 = myend - e0;
    int s0 = __shfl_up_sync(0xffffffffu, s1, BP);
    if (rl == 0) s0 = 0;
    if (act) {
      double s = 0.0;
      if (cl < BS)
        for (int j = s0; j < s1; ++j) s += prod[j * BS + cl];
      post(irow, s, ops);
    }
    __syncwarp();
    d = dn;
  }
}

}  // namespace fx
