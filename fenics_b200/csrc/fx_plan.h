// Internal: the plan object behind the C ABI (device-resident mesh, dofmaps, patterns, scratch).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/fenics_b200.h"
#include "fx_assemble.h"

namespace fx {

std::string& last_error();
void count_launch(int n);
int set_error(int code, const char* fmt, ...);

#define FX_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fx::set_error(FX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                           __FILE__, __LINE__);                                                \
  } while (0)

struct DevPattern {
  int n = 0, nnz = 0, ld = 0;
  bool present = false, has_csr = false;
  int* rowptr = nullptr;
  int* col = nullptr;
  uint8_t* pos = nullptr;
  int* dm_t = nullptr;
  int* rowblk = nullptr;
  int nrowblk = 0;
  // block-triangular elimination (fx_plan_set_eliminated_dofs): trailing rows solved by back-substitution
  std::vector<int> rowblk_h;  // host copy of rowblk
  int maxblknnz = 0;          // largest pattern entry count of a row block
  int maxrow = 0;             // longest pattern row
  int* agg = nullptr;         // [n] aggregate of each dof (fx_plan_set_aggregates), null = none
  int nagg = 0;
  int agg_nl = 0;             // most aggregates the rows of one CTA of the resident CG kernel touch
  uint8_t* elim = nullptr;    // [n] row mask: bit 0 eliminated, bit 1 ghost, bit 2 interface (fx_la_common.cuh MASK_*)
  int* elim_rows = nullptr;   // [nelim]
  int nelim = 0;
  std::vector<uint8_t> elim_h, halo_h;  // host copies of the two mask sources (rebuild_mask)
  // halo push lists of a partitioned run (fx_plan_set_halo), sorted by local dof
  bool has_halo = false;
  int nsend = 0;
  int *s_first = nullptr, *s_peer = nullptr, *s_local = nullptr, *s_remote = nullptr;
  // deterministic (atomic-free) assembly, built lazily by det_ensure(): contributions of every CSR entry / dof as
  // indices into the element-tensor buffer, ascending in the cell index
  int *det_cptr = nullptr, *det_contrib = nullptr;  // [nnz + 1], [nc * ld * ld]   (matrix spaces)
  int *det_aptr = nullptr, *det_adj = nullptr;      // [n + 1],   [nc * ld]
  bool det_ready = false;
  // node-block solver layout (fx_nb_host.h), built lazily at the first solve that can use it and dropped whenever
  // the row masks change; state 0 = not built yet, 1 = ready, -1 = the space has no usable node-block structure
  struct NbDev {
    int state = 0, bs = 0, bp = 0, nn = 0, N = 0, nchunks = 0, grid = 0, nzchk = 0;
    long long ne = 0, nslots = 0;
    std::string why;
    int4* desc = nullptr;
    int *wrange = nullptr, *bre = nullptr, *bcol = nullptr, *src = nullptr, *diag = nullptr, *ext_of_int = nullptr,
        *int_of_ext = nullptr, *zchk = nullptr;
    uint8_t* imask = nullptr;
    int *s_first = nullptr, *s_remote = nullptr, *s_local = nullptr;  // halo push lists in internal numbering (partitioned run)
    std::vector<int> int_of_ext_h;                // host copy (fx_space_solver_index)
    void release() {
      cudaFree(desc); cudaFree(wrange); cudaFree(bre); cudaFree(bcol); cudaFree(src); cudaFree(diag);
      cudaFree(ext_of_int); cudaFree(int_of_ext); cudaFree(zchk); cudaFree(imask); cudaFree(s_first); cudaFree(s_remote); cudaFree(s_local);
      *this = NbDev();
    }
  } nb;
  std::vector<uint8_t> ghost_owner_h; // fx_plan_set_halo: 1 + owner rank of a ghost dof (or just non-zero), 0 = owned
  std::vector<int> halo_order;       // fx_plan_set_halo: sorted position k <- caller's pair order[k]
  std::vector<int> halo_first_h;     // host copy of s_first
  std::vector<int> halo_remote_int;  // fx_plan_set_halo_solver_index: destination internal indices, caller's pair order
};

// peer-mapped regions of a partitioned run, as the Krylov kernels see them (layout: fx_la_common.cuh)
constexpr int DIST_MAXW = 8;
struct DistDev {
  int rank = 0, world = 0;       // world <= 1: single GPU
  double* region[DIST_MAXW] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // region[rank] is mine
  long long nmax = 0;            // stride of the work vectors inside a region (>= local n of every space)
  // halo push lists of the space being solved (fx_plan_set_halo), sorted by local dof
  int nsend = 0;                 // (dof, destination) pairs
  const int* s_first = nullptr;  // [n] first pair of an interface dof (MASK_IFACE rows only)
  const int* s_peer = nullptr;   // [nsend] destination rank | DIST_LAST
  const int* s_local = nullptr;  // [nsend] my local dof index
  const int* s_remote = nullptr; // [nsend] the destination's local (ghost) dof index
};

constexpr int KRYLOV_NWORK = 12;
constexpr int COARSE_KMAX = 256;   // aggregates of the two-level preconditioner in the resident kernel (E^-1 rows in shared memory)
constexpr int COARSE_KMAX_CG = 1024;  // ... inside k_cg (systems beyond the resident kernel: a finer coarse space pays)
constexpr int COARSE_NLMAX = 64;   // aggregates one CTA of the resident CG kernel may touch
constexpr int RED_SLOTS = 4;       // values reduced at once
constexpr int MAX_GRID = 148 * 16; // upper bound on cooperative grid size

constexpr int DIST_LAST = 0x100;   // s_peer flag: last destination of this dof
constexpr int DIST_XH = 8;         // work-vector slot the solution's interface values are exchanged through
// region layout in doubles: [KRYLOV_NWORK * nmax work vectors][2 x DIST_MAXW x RED_SLOTS 16-byte LL lines][2 state]
// An LL line {lo32, epoch, hi32, epoch} carries one double of a cross-GPU reduction TOGETHER with the barrier
// epoch it belongs to (two 8-byte halves, each written atomically: the protocol of NCCL's LL primitives), so
// the data needs no fence + separate flag; line [parity of the epoch][source rank][value].
__host__ __device__ inline long long dist_xred_off(long long nmax) { return (long long)KRYLOV_NWORK * nmax; }
__host__ __device__ inline long long dist_state_off(long long nmax) { return dist_xred_off(nmax) + 2 * DIST_MAXW * RED_SLOTS * 2; }
__host__ __device__ inline long long dist_region_doubles(long long nmax) { return dist_state_off(nmax) + 2; }


}  // namespace fx

struct fx_plan {
  int device = 0, num_sms = 0;
  int nv = 0, nc = 0, nbf = 0;
  int nc_owned = 0;  // functionals integrate over cells [0, nc_owned) (partitioned mode; default nc)
  double* xy = nullptr;
  int* cells = nullptr;
  double* geom = nullptr;
  int *bf_cell = nullptr, *bf_local = nullptr, *bf_marker = nullptr;
  double* bf_geo = nullptr;
  fx::DevPattern sp[fx::S_N];
  // scratch
  double* part = nullptr;      // per-block partial sums of functionals / dots
  int part_cap = 0;
  double* scal = nullptr;      // device scalars (results of fx_assemble_scalar etc.)
  double* work = nullptr;      // Krylov work vectors, KRYLOV_NWORK x work_n
  size_t work_n = 0;
  double* red = nullptr;       // grid-reduction scratch 2 x RED_SLOTS x MAX_GRID
  fx_solve_info* info = nullptr;  // device, FX_NTICKETS result slots
  // zero-compacted matrix copy used inside the Krylov kernels (sized for the largest pattern)
  double* cval = nullptr;
  int *ccol = nullptr, *cre = nullptr;
  int4* cdesc = nullptr;  // per-row-block descriptors of the compacted copy
  int *cprefix = nullptr, *scanpart = nullptr;  // nnz-balanced partition of the row blocks (balance_partition)
  int2* rowpre = nullptr;                       // compact_reblock: per-row prefixes
  int* nblk = nullptr;                          // compact_reblock: number of blocks built
  double* coarse = nullptr;                     // two-level preconditioner scratch: E0, E1 (256^2 each), cvec [3][256]
  unsigned* bar = nullptr;                      // grid-barrier counter of the persistent Krylov kernels
  double* gmres_basis = nullptr;                // GMRES Krylov basis, (restart + 1) x n, allocated on first use
  size_t gmres_cap = 0;
  double* red_big = nullptr;                    // GMRES: block partials of the Gram-Schmidt coefficients
  int asm_mode = 0;                             // 0: atomic scatter-add (default), 1: deterministic row gather
  double* elt = nullptr;                        // element-tensor buffer of the deterministic mode, elt_cap doubles
  size_t elt_cap = 0;
  unsigned char* bf_first = nullptr;            // boundary facets chained per cell (deterministic facet integrals)
  int* bf_next = nullptr;
  double* nbval = nullptr;                      // node-block value buffer (gathered inside every solve), nbval_cap doubles
  size_t nbval_cap = 0;
  void* ilu = nullptr;                          // FX_PC_ILU0 workspace (fx_ilu.cu: factors, epoch flags, work vectors), built on first use
  fx::DistDev dist;                             // world > 1: mesh-partitioned run (fx_plan_set_peer_regions)

  fx::MeshView view() const {
    fx::MeshView m;
    m.nc = nc; m.geom = geom;
    for (int s = 0; s < fx::S_N; ++s) m.dm[s] = sp[s].dm_t;
    m.nbf = nbf; m.bf_cell = bf_cell; m.bf_local = bf_local; m.bf_marker = bf_marker; m.bf_geo = bf_geo;
    m.bf_first = bf_first; m.bf_next = bf_next;
    return m;
  }
};

namespace fx {
int ensure_work(fx_plan* plan, size_t n);
// launchers implemented in fx_asm.cu
int launch_geometry(fx_plan* plan, cudaStream_t st);
int launch_matrix(fx_plan* plan, int form, const AsmArgs& base, cudaStream_t st);
int launch_vector(fx_plan* plan, int form, const AsmArgs& base, cudaStream_t st);
int launch_scalar(fx_plan* plan, int form, const AsmArgs& base, double* out_dev, cudaStream_t st);
int launch_dg0(fx_plan* plan, int expr, const AsmArgs& base, cudaStream_t st);
int form_info(int form, int* rank, int* space, int* qdeg);
// fx_la.cu
int la_spmv(fx_plan* plan, int space, const double* val, const double* x, double* y, cudaStream_t st);
int la_spmv_variant(fx_plan* plan, int space, const double* val, const double* x, double* y, int variant,
                    cudaStream_t st);
int la_apply_bc(fx_plan* plan, int space, double* val, double* b, const int* rows, const double* g, int n,
                cudaStream_t st);
int la_axpy(int n, double alpha, const double* x, double* y, cudaStream_t st);
int la_dot(fx_plan* plan, int n, const double* x, const double* y, double* out_dev, cudaStream_t st);
int la_norm(fx_plan* plan, int n, const double* x, int kind, double* out_dev, cudaStream_t st);
int la_lincomb3(int n, double* out, const double* coef, const double* x, const double* y, const double* z, cudaStream_t st);
int la_dotw(fx_plan* plan, int n, const double* x, const double* y, const double* w, double* out_dev, cudaStream_t st);
struct PcgConfig { int R, slice, grid, nlmax; };  // R == 0: the system does not fit the resident CG kernel
PcgConfig pcg_resident_config(const fx_plan* plan, int nrowblk, int maxblknnz);
int la_spmv_nb(fx_plan* plan, int space, const double* val, const double* x, double* y, int reps, cudaStream_t st);
int nb_ensure(fx_plan* plan, int space);
int nb_attach_halo(fx_plan* plan, int space);  // builds the node-block layout of a space if possible; FX_OK either way unless CUDA fails
int la_krylov(fx_plan* plan, int space, const double* val, double* x, const double* b, int method, int pc,
              double rtol, double atol, int maxit, int ticket, cudaStream_t st);
// fx_ilu.cu: BiCGStab / CG with ILU(0) in the natural ordering (PETSc's serial "default")
int la_krylov_ilu0(fx_plan* plan, int space, const double* val, double* x, const double* b, int method, double rtol,
                   double atol, int maxit, int ticket, cudaStream_t st);
void ilu_release(fx_plan* plan);
}  // namespace fx
