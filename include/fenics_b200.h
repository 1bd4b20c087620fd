/* fenics_b200 -- C ABI of the B200-native time-step hot path (assembly + Krylov solves).
 *
 * The reference (ATMackay/fenics) has no FFI of its own: its hot path is the dolfin 2019.1.0 Python
 * calls made inside each script's `while t < Tf` loop.  Every entry point below states the
 * reference call (file:line) it replaces.  All `double*`/`int*` arguments documented "device" are
 * CUDA device pointers on the plan's device (torch CUDA tensors on the Python side); "host" are
 * ordinary host pointers.  Every call returns 0 on success or a negative fx_status; the message is
 * available from fx_last_error().  All device work is enqueued on `stream` (a cudaStream_t passed as
 * void*) and is asynchronous unless stated.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef FENICS_B200_H
#define FENICS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fx_plan fx_plan;

typedef enum {
  FX_OK = 0,
  FX_ERR_ARG = -1,      /* bad argument (null pointer, unknown id, size mismatch)            */
  FX_ERR_CUDA = -2,     /* CUDA runtime error (message holds cudaGetErrorString)             */
  FX_ERR_NOCONV = -3,   /* Krylov solver hit maxit or diverged (dolfin raises RuntimeError)  */
  FX_ERR_UNSUPPORTED = -4
} fx_status;

/* Function spaces of function_spaces(mesh, 2) (lid-driven-cavity/fenics_fem.py:262-287,
 * buoyancy-driven-flow/fenics_fem.py:448-468, "flow between eccentrically rotating
 * cylinders"/fenics_base.py:470-502).  Local dofs are component-blocked in UFC order.       */
typedef enum {
  FX_SPACE_W = 0,  /* (P2)^2 x (DG0)^3, 15 local dofs: ux 0-5 | uy 6-11 | D0 12 | D1 13 | D2 14 */
  FX_SPACE_ZC = 1, /* (P2)^3, 18 local dofs                                                     */
  FX_SPACE_Q = 2,  /* P1, 3 local dofs                                                          */
  FX_SPACE_ZD = 3, /* (DG0)^3                                                                   */
  FX_SPACE_QT = 4, /* DG0                                                                       */
  FX_SPACE_V = 5,  /* (P2)^2, 12 local dofs (post-loop diagnostics)                             */
  FX_SPACE_VS = 6, /* P2 scalar, 6 local dofs: V.sub(0).collapse() of stream_function()           */
  FX_NSPACES = 7
} fx_space;

/* Slots of the coefficient ("state") table handed to the assembly calls: state[f] is a device
 * vector in the dof order of the named space, or NULL when the form does not read it.  The names
 * are the reference's (solution_functions, lid-driven-cavity/fenics_fem.py:227-247).          */
typedef enum {
  FX_F_W0 = 0, FX_F_W12, FX_F_WS, FX_F_W1,                 /* W  */
  FX_F_P0, FX_F_P1, FX_F_RHO0, FX_F_RHO12, FX_F_RHO1,      /* Q  */
  FX_F_T0, FX_F_T1, FX_F_THETA0, FX_F_THERM_INV,           /* Q  */
  FX_F_TAU0, FX_F_TAU12, FX_F_TAU1,                        /* Zc */
  FX_F_KAPP,                                               /* Qt */
  FX_F_RES_TEST,                                           /* Zd: project(restau0, Zd)          */
  FX_F_RES_ORTH,                                           /* Zc: project(restau0-res_test, Zc) */
  FX_F_QT_A,                                               /* Qt scratch (res_orth_norm_sq)     */
  FX_F_Q_A,                                                /* Q  scratch                        */
  FX_NFIELDS = 24
} fx_field;

/* Slots of the parameter table (host doubles). */
typedef enum {
  FX_P_DT = 0, FX_P_RE, FX_P_WE, FX_P_BETAV, FX_P_C1, FX_P_C2, FX_P_TH, FX_P_CONV, FX_P_MA,
  FX_P_PR, FX_P_RA, FX_P_DI, FX_P_VH, FX_P_BI, FX_P_T0, FX_P_TH_HOT, FX_P_AL1, FX_P_AL2,
  FX_P_LAMBDA_D, FX_P_B_FENE, FX_P_A0, FX_P_EPS, FX_P_T,
  FX_P_K_EWM, FX_P_B_EWM,  /* extended White-Metzner exponents k_ewm, B (comp_ewm_jpb.py:89-90)             */
  FX_P_TH_T,               /* weight of the temperature DEVSS terms: th (incompressible_benchmarkEWM.py:650) or 0 */
  FX_NPARAMS = 32
} fx_param;

/* Forms.  Bilinear forms produce CSR values on the pattern of their space, linear forms a vector,
 * functionals one double.  "cfg1" = lid-driven-cavity/incomp_LDC.py, "cfg2" =
 * lid-driven-cavity/comp_LDC_mesh_convergence.py (line numbers of each form in DESIGN.md).     */
typedef enum {
  /* generic */
  FX_FORM_MASS_ZC = 1, /* inner(w,Pv)*dx on Zc  (dolfin project(), mass matrix)   */
  FX_FORM_MASS_Q,      /* on Q                                                     */
  /* post-loop diagnostics: stream_function / comp_stream_function (lid-driven-cavity/fenics_fem.py:340-378) */
  FX_FORM_STREAM_A = 10,      /* inner(grad(psi), grad(phi))*dx on VS                              */
  FX_FORM_STREAM_L,           /* inner(u[1].dx(0) - u[0].dx(1), phi)*dx, u = velocity of FX_F_W1   */
  FX_FORM_COMP_STREAM_L,      /* inner(rho*u[1].dx(0) - rho*u[0].dx(1), phi)*dx, rho = FX_F_RHO1   */
  FX_FORM_SHEAR_L,            /* shear_rate(u): inner(inner(0.5*gamma, gamma), phi)*dx, gamma = grad u + grad u^T (:380-398) */
  /* cfg1: bilinear */
  FX_FORM_C1_A1 = 100, FX_FORM_C1_A2, FX_FORM_C1_A5, FX_FORM_C1_A7, FX_FORM_C1_A4,
  /* cfg1: linear */
  FX_FORM_C1_L1 = 120, FX_FORM_C1_L2, FX_FORM_C1_L5, FX_FORM_C1_L7, FX_FORM_C1_L4,
  FX_FORM_LDC_L_RES_ORTH, /* inner(w, restau0 - res_test)*dx on Zc (incomp_LDC.py:305)        */
  /* cfg1: functionals */
  FX_FORM_C1_EK = 140, FX_FORM_C1_EE,
  /* cfg2: bilinear */
  FX_FORM_C2_A0 = 200, FX_FORM_C2_A1, FX_FORM_C2_A2, FX_FORM_C2_A5, FX_FORM_C2_A7, FX_FORM_C2_A4,
  /* cfg2: linear */
  FX_FORM_C2_L0 = 220, FX_FORM_C2_L1, FX_FORM_C2_L2, FX_FORM_C2_L5, FX_FORM_C2_L7, FX_FORM_C2_L4,
  FX_FORM_C2_L_RHO1,      /* inner(q, rho0 + Ma^2/Re (p1-p0))*dx (comp_LDC_mesh_convergence.py:501-502) */
  /* cfg2: functionals */
  FX_FORM_C2_EK = 240, FX_FORM_C2_EE,
  /* cfg3 = "flow between eccentrically rotating cylinders"/incompressible_benchmarkEWM.py at its own
   * A_0 = k_ewm = B = 0 (time loop :465-782; parameter FX_P_AL1 carries `al`): bilinear */
  FX_FORM_C3_A0 = 300, FX_FORM_C3_A1, FX_FORM_C3_A2, FX_FORM_C3_A5, FX_FORM_C3_A6, FX_FORM_C3_A7, FX_FORM_C3_A4,
  FX_FORM_C3_A9,
  /* true extended White-Metzner variants (comp_ewm_jpb.py:421-732: A_0, k_ewm, B != 0 -> non-constant phi_s, phi_ewm);
   * the other forms of that loop are the FX_FORM_C3_* ones (A2 already carries phi_s) */
  FX_FORM_C3E_A1, FX_FORM_C3E_A4,
  /* cfg3: linear */
  FX_FORM_C3_L0 = 320, FX_FORM_C3_L1, FX_FORM_C3_L2, FX_FORM_C3_L5, FX_FORM_C3_L6, FX_FORM_C3_L7, FX_FORM_C3_L4,
  FX_FORM_C3_L9,
  FX_FORM_C3E_L1, FX_FORM_C3E_L2, FX_FORM_C3E_L4, FX_FORM_C3E_L9,
  /* cfg3: functionals; torque / forces over ds(0) with sigmacon (:690-707) */
  FX_FORM_C3_EK = 340, FX_FORM_C3_EE, FX_FORM_C3_TORQUE, FX_FORM_C3_FORCE_X, FX_FORM_C3_FORCE_Y,
  /* cfg4 = "flow between eccentrically rotating cylinders"/fenepmp_jbp.py (lambda_D = 0): bilinear */
  FX_FORM_C4_A0 = 400, FX_FORM_C4_A2, FX_FORM_C4_A5, FX_FORM_C4_A6, FX_FORM_C4_A7, FX_FORM_C4_A4, FX_FORM_C4_A9,
  /* cfg4: linear */
  FX_FORM_C4_L0 = 420, FX_FORM_C4_L2, FX_FORM_C4_L5, FX_FORM_C4_L6, FX_FORM_C4_L7, FX_FORM_C4_L4, FX_FORM_C4_L9,
  FX_FORM_C4_L_RES_ORTH,
  /* cfg4: functionals; torque / force_x / force_y are exterior-facet integrals over ds(0) (fenepmp_jbp.py:639-649) */
  FX_FORM_C4_EK = 440, FX_FORM_C4_EE, FX_FORM_C4_TORQUE, FX_FORM_C4_FORCE_X, FX_FORM_C4_FORCE_Y,
  /* cfg4 with lambda_D != 0 (FENE-P-MP proper: phi_def(u) = 1 + (3 lambda_D I_3(D)/(I_2(D) + 1e-7))^2,
   * fenics_base.py:287-291, used at fenepmp_jbp.py:464,519,597,615,639): the forms that carry phi_def -- UFL degrees
   * 13 (A4, L2, L_RES_ORTH), 14 (L9), 12 (ds functionals), 11 (restau0 -> Zd); every other cfg4 form is unchanged.  FX_P_LAMBDA_D holds lambda_D. */
  FX_FORM_C4M_A4 = 460, FX_FORM_C4M_L2, FX_FORM_C4M_L9, FX_FORM_C4M_L_RES_ORTH,
  FX_FORM_C4M_TORQUE = 470, FX_FORM_C4M_FORCE_X, FX_FORM_C4M_FORCE_Y,
  /* cfg5 = buoyancy-driven-flow/compressible_dgp.py: bilinear */
  FX_FORM_C5_A0 = 500, FX_FORM_C5_A1, FX_FORM_C5_A2, FX_FORM_C5_A5, FX_FORM_C5_A6, FX_FORM_C5_A7, FX_FORM_C5_A4,
  FX_FORM_C5_A8,
  /* cfg5: linear (incl. the right-hand sides of the in-loop project(.., Q) calls :405,:408,:530,:581) */
  FX_FORM_C5_L0 = 520, FX_FORM_C5_L1, FX_FORM_C5_L2, FX_FORM_C5_L5, FX_FORM_C5_L6, FX_FORM_C5_L7, FX_FORM_C5_L4,
  FX_FORM_C5_L8, FX_FORM_C5_L_RES_ORTH, FX_FORM_C5_L_THETA0, FX_FORM_C5_L_THETA1, FX_FORM_C5_L_THERM_INV,
  FX_FORM_C5_L_RHO1,
  /* cfg5: functionals; NUS = assemble(-inner(grad(theta1), n)*ds(2)) with theta1 in field FX_F_Q_A (:581-583) */
  FX_FORM_C5_EK = 540, FX_FORM_C5_EE, FX_FORM_C5_NUS,
  /* "cfg6" = buoyancy-driven-flow/incompressible_dgp.py (time loop :357-535).  Its a5/L5 (pure Neumann
   * pressure Poisson problem; FX_P_RE must hold 1), a4, L4, E_k, E_e and theta1 projection are shared functors */
  FX_FORM_C6_A1 = 600, FX_FORM_C6_A2, FX_FORM_C6_A5, FX_FORM_C6_A7, FX_FORM_C6_A4, FX_FORM_C6_A8,
  FX_FORM_C6_L1 = 620, FX_FORM_C6_L2, FX_FORM_C6_L5, FX_FORM_C6_L7, FX_FORM_C6_L4, FX_FORM_C6_L8, FX_FORM_C6_L_RES_ORTH,
  FX_FORM_C6_EK = 640, FX_FORM_C6_EE, FX_FORM_C6_NUS
} fx_form;

/* DG0 projections (diagonal mass => cell average), dolfin project(expr, Zd|Qt).               */
typedef enum {
  FX_EXPR_LDC_RESTAU = 1,   /* restau0 -> Zd   (incomp_LDC.py:300-304)                     */
  FX_EXPR_RES_ORTH_SQ,      /* inner(res_orth,res_orth) -> Qt (:306)                       */
  FX_EXPR_SQRT_QT_A,        /* (QT_A)**0.5 -> Qt (:307-308)                                */
  FX_EXPR_H_KAPP,           /* h*kapp -> Qt (:434)                                         */
  FX_EXPR_C4_RESTAU,        /* cfg4 restau0 -> Zd (fenepmp_jbp.py:462-467)                 */
  FX_EXPR_C5_RESTAU,        /* cfg5 restau0 -> Zd (compressible_dgp.py:382-386)            */
  FX_EXPR_C6_RESTAU,        /* incompressible_dgp.py restau0 -> Zd (:369-373)             */
  /* post-loop adaptive-refinement indicator of the journal-bearing scripts (incompressible_benchmarkEWM.py:903-917,
   * fenepmp_jbp.py:691-705, comp_ewm_jpb.py:735-747); refine() itself stays in dolfin                              */
  FX_EXPR_ADAPT_RES_SQ,     /* inner(restau0,restau0) -> Qt, restau0 = We/dt (tau1-tau0) + We Fdef(u1,tau1) + tau1 - I */
  FX_EXPR_TAU_AVG,          /* (tau1_vec[0]+tau1_vec[1]+tau1_vec[2])/3 -> Qt                */
  FX_EXPR_ABS_KAPP_RATIO,   /* |FX_F_KAPP / (FX_F_QT_A + FX_P_EPS)| -> Qt                   */
  FX_EXPR_C4M_RESTAU        /* cfg4 restau0 with the lambda_D != 0 dissipation term -> Zd (fenepmp_jbp.py:462-467) */
} fx_expr;

typedef enum { FX_KSP_BICGSTAB = 0, FX_KSP_CG = 1, FX_KSP_GMRES = 2 } fx_ksp;
typedef enum {
  FX_PC_NONE = 0, FX_PC_JACOBI = 1, FX_PC_ILU0 = 2,
  FX_PC_JACOBI_COARSE = 3 /* Jacobi + aggregate coarse-grid correction, see fx_plan_set_aggregates */
} fx_pc;

typedef struct {
  int iterations;
  int converged;      /* 1 converged, 0 maxit, -1 breakdown/diverged, -2 ILU(0): missing or zero pivot, -3 singular eliminated block */
  double residual;    /* final (preconditioned) residual 2-norm               */
  double residual0;   /* reference norm the relative tolerance is applied to  */
} fx_solve_info;

const char* fx_last_error(void);
int fx_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
long long fx_launch_count(void);
/* synchronous device -> host copy of plan-owned data (patterns, geometry) for inspection/tests */
int fx_copy_to_host(void* dst_host, const void* src_dev, long long nbytes);

/* ---- plan: mesh + dofmaps -> device-resident geometry, CSR patterns, scatter maps ----------
 * Replaces what dolfin builds on every assemble() call (SparsityPatternBuilder, DofMap tabulation;
 * e.g. lid-driven-cavity/incomp_LDC.py:328) -- built once here.  All inputs are HOST arrays:
 * xy[nv*2], cells[nc*3] (vertices ascending per cell, UFC), cell_dofs[s] = nc*ld(s) ints (row-major,
 * dolfin's V.dofmap().cell_dofs(c)) or NULL if the space is unused; boundary facets as (cell, local
 * facet opposite local vertex i, marker).                                                     */
int fx_plan_create(const double* xy, int nv, const int* cells, int nc,
                   const int* const* cell_dofs, const int* bfacet_cell, const int* bfacet_local,
                   const int* bfacet_marker, int nbf, int device, fx_plan** out);
void fx_plan_destroy(fx_plan* plan);
int fx_plan_set_markers(fx_plan* plan, const int* bfacet_marker, int nbf);

int fx_space_dim(const fx_plan* plan, int space);
int fx_space_local_dim(int space);
/* CSR pattern of space x space (DOLFIN's full cell-block pattern, columns ascending):
 * device pointers owned by the plan.                                                          */
int fx_pattern(const fx_plan* plan, int space, int* n, int* nnz, const int** rowptr_dev,
               const int** col_dev);
/* geometry produced by the K1 kernel, device, SoA [6][nc]: K00,K01,K10,K11 (=J^-1), |detJ|, h   */
int fx_geometry(const fx_plan* plan, const double** geom_dev);

/* ---- assembly (assemble(a), assemble(L), assemble(functional)) ----------------------------- */
int fx_assemble_matrix(fx_plan* plan, int form, const double* const* state, const double* params,
                       double* val_dev, void* stream);
int fx_assemble_vector(fx_plan* plan, int form, const double* const* state, const double* params,
                       double* b_dev, void* stream);
/* synchronous: waits for `stream`, writes one host double */
int fx_assemble_scalar(fx_plan* plan, int form, const double* const* state, const double* params,
                       double* out_host, void* stream);
/* asynchronous variant: result lands in out_dev[0] */
int fx_assemble_scalar_dev(fx_plan* plan, int form, const double* const* state,
                           const double* params, double* out_dev, void* stream);
int fx_project_dg0(fx_plan* plan, int expr, const double* const* state, const double* params,
                   double* out_dev, void* stream);
/* which quadrature degree the kernels of `form` use: q[0]=dx degree, q[1]=ds degree (-1: none) */
int fx_form_info(int form, int* rank, int* space, int* qdeg);

/* ---- DirichletBC.apply(A, b) (incomp_LDC.py:330): identity rows, b[row]=g ------------------ */
int fx_apply_bc(fx_plan* plan, int space, double* val_dev, double* b_dev, const int* rows_dev,
                const double* g_dev, int n, void* stream);

/* ---- linear algebra ------------------------------------------------------------------------ */
int fx_spmv(fx_plan* plan, int space, const double* val_dev, const double* x_dev, double* y_dev,
            void* stream);
/* tuning aid (tools/spmv_bench.py): 0 = the production CSR-stream kernel, 1.. = CSR-vector variants */
int fx_spmv_variant(fx_plan* plan, int space, const double* val_dev, const double* x_dev, double* y_dev,
                    int variant, void* stream);
int fx_axpy(int n, double alpha, const double* x_dev, double* y_dev, void* stream);
int fx_dot(fx_plan* plan, int n, const double* x_dev, const double* y_dev, double* out_host,
           void* stream);                      /* synchronous */
int fx_norm(fx_plan* plan, int n, const double* x_dev, int kind /*0 l2, 1 linf*/, double* out_host,
            void* stream);                     /* synchronous; norm(vec,'l2'|'linf')          */
/* asynchronous variants: the result lands in out_dev[0] */
int fx_dot_dev(fx_plan* plan, int n, const double* x_dev, const double* y_dev, double* out_dev,
               void* stream);
int fx_norm_dev(fx_plan* plan, int n, const double* x_dev, int kind, double* out_dev, void* stream);

/* ---- building blocks of the mesh-partitioned (multi-GPU) mode: SURVEY.md 8e.2 ----------------
 * Each rank owns a plan on its sub-mesh (owned cells + the ghost cells touching an owned dof).
 * out = c[0] x + c[1] y + c[2] z with the three coefficients read from DEVICE memory (so a Krylov
 * iteration needs no host round trip for its scalars); out may alias x, y or z.                */
int fx_lincomb3_dev(int n, double* out_dev, const double* coef_dev, const double* x_dev,
                    const double* y_dev, const double* z_dev, void* stream);
/* out_dev[0] = sum_i w_i x_i y_i  (w = 1 on owned dofs, 0 on ghosts: rank-local part of a dot)   */
int fx_dotw_dev(fx_plan* plan, int n, const double* x_dev, const double* y_dev, const double* w_dev,
                double* out_dev, void* stream);
/* functionals integrate over cells [0, nc_owned) only (owned cells are numbered first) and over
 * the boundary facets of those cells; pass nc (the default) to integrate over every local cell. */
int fx_plan_set_owned_cells(fx_plan* plan, int nc_owned);

/* Block-triangular elimination for the Krylov solves on `space`.  The W systems of every loop
 * (a1/a2/a7, e.g. incomp_LDC.py:317-325) are block lower triangular: the DEVSS rows
 * (D - grad u, R) read the velocity but the momentum rows never read D, and the D-D block is the
 * diagonal DG0 mass.  With rows[0..n) registered (host array of dof indices), fx_krylov_solve*
 * iterates on the remaining rows only and recovers the registered rows afterwards by exact
 * back-substitution x_i = (b_i - sum_{j!=i} A_ij x_j) / A_ii: same linear system, same solution
 * to the requested tolerance, fewer and cheaper iterations.  The structure is verified on the device
 * at every solve: a non-zero A[active, eliminated] or A[eliminated, other eliminated] entry makes
 * the solve fail with converged = -3.  n = 0 switches elimination off.  The convergence test then
 * monitors the residual of the active rows.                                                     */
int fx_plan_set_eliminated_dofs(fx_plan* plan, int space, const int* rows, int n);

/* Aggregates for the two-level preconditioner FX_PC_JACOBI_COARSE of the CG solves on `space`:
 *   M^-1 = D^-1 + P (P^T A P)^-1 P^T,   P = indicator vectors of k aggregates of dofs (agg[i] in [0,k)).
 * Stands where the reference asks PETSc for "amg"/"default" on the pressure system
 * (incompressible_benchmarkEWM.py:441-451): Jacobi-CG needs O(1/h) iterations of microseconds each; the
 * coarse space removes the smooth error modes that cost most of them.  The coarse matrix is rebuilt
 * from the current matrix values and inverted on the device inside every solve.  Requirements:
 * k a multiple of 32, every aggregate non-empty, CG, no eliminated rows.  A may be singular or nearly singular along
 * the constant vector (pure-Neumann pressure systems; compressible ones at Ma -> 0): the kernels detect it on the
 * device and invert the coarse matrix through a rank-one shift (exact Sherman-Morrison correction when the matrix is
 * only nearly singular).  Two kernels apply it: the shared-memory-resident pipelined CG (k <= 256, the system small
 * enough to stay resident -- P1 spaces up to ~148 k rows -- with aggregates local to its row blocks, rtol >= 1e-9) and,
 * for everything else (k <= 1024), the general CG kernel with two extra grid barriers per iteration.  Returns
 * FX_ERR_UNSUPPORTED (and leaves the space without aggregates) when k is out of range.  k = 0 removes them. */
int fx_plan_set_aggregates(fx_plan* plan, int space, const int* agg, int k);

/* How fx_assemble_matrix / fx_assemble_vector add the element tensors into the global tensor
 * (north_star (1): "mesh colouring, or atomics where colouring loses"):
 *   FX_ASM_ATOMIC         fp64 atomicAdd scatter through the uint8 position map (default; the sum order, and with it
 *                         the last bits of the result, vary from run to run)
 *   FX_ASM_DETERMINISTIC  atomic-free: the element kernels store every element-tensor row into a buffer, a gather
 *                         kernel then sums the contributions of each CSR entry / dof in ascending cell order (facet
 *                         integrals are folded into their cell's row first) -- bitwise reproducible, every value
 *                         written exactly once (no memset).  The environment variable FX_ASM_MODE=1 makes it the
 *                         default of new plans (bench.py A/B).                                                  */
typedef enum { FX_ASM_ATOMIC = 0, FX_ASM_DETERMINISTIC = 1 } fx_asm_mode;
int fx_plan_set_assembly_mode(fx_plan* plan, int mode);

/* ---- mesh-partitioned runs: one process per GPU, one plan per sub-mesh ---------------
 * The partitioned analogue of what PETSc's MPI Mat/Vec + VecScatter do under dolfin's solve() when the
 * reference is run with mpirun, re-designed for NVLink peer memory: every rank owns a
 * "region" of fx_peer_region_bytes(nmax) bytes that all ranks of the box have mapped into their address
 * space (torch symmetric memory, cudaIpc or cuMem fabric handles -- the library only sees pointers).
 * It holds the rank's Krylov work vectors (stride nmax doubles, nmax >= the local dimension of every
 * space on every rank), cross-GPU reduction slots and barrier flags.  Regions must be zero-filled, and
 * that zero-fill complete on EVERY rank, before any rank's first solve.
 * With regions and halos registered, fx_krylov_solve* runs the same persistent kernels, which then
 *   - treat ghost rows as absent (owned rows must be complete: assemble over owned + ghost cells),
 *   - store every freshly computed interface value of an SpMV input vector straight into the ghost
 *     slot of each neighbour's copy of that vector (P2P stores over NVLink, no pack/send/unpack),
 *   - close every grid barrier with a cross-GPU epoch-flag exchange, and sum dot products across ranks in
 *     rank order (bitwise identical on all ranks, so all ranks take the same branches),
 *   - refresh the ghost entries of the solution x before returning.
 * No NCCL call and no host synchronisation happens inside a solve.  All ranks must issue the same
 * sequence of solves.  A peer that stops responding trips a 10 s device-side timeout: converged = -4.
 * world <= 1 removes the regions (single-GPU behaviour).                                            */
long long fx_peer_region_bytes(long long nmax);
int fx_plan_set_peer_regions(fx_plan* plan, int rank, int world, void* const* regions, long long nmax);
/* Halo of one space: ghost[i] != 0 marks local dofs owned by another rank (host array, length dim);
 * (peer[k], local[k], remote[k]), k < nsend: the owned local dof local[k] is ghost dof remote[k] (in
 * rank peer[k]'s local numbering) on rank peer[k].  Host arrays, copied.  ghost == NULL removes it.   */
int fx_plan_set_halo(fx_plan* plan, int space, int nsend, const int* peer, const int* local, const int* remote,
                     const unsigned char* ghost);

/* The Krylov kernels keep their work vectors in an internal node-major numbering (all components of a mesh node
 * adjacent; fenics_b200/csrc/fx_nb_host.h).  On a partition the owner of an interface value stores it straight into the
 * neighbour's copy of the work vector, so it needs the neighbour's INTERNAL index of the ghost entry:
 *   fx_space_solver_index          int_of_ext[n] of this rank's plan (-1: dof outside the node blocks, e.g. the eliminated
 *                                  DEVSS rows); *n_internal = length of the internal vectors (peer regions must hold
 *                                  vectors that long: pass nmax >= n_internal, a multiple of 4, to fx_plan_set_peer_regions).  Call after
 *                                  fx_plan_set_eliminated_dofs / fx_plan_set_halo.  FX_ERR_UNSUPPORTED: no such layout.
 *   fx_plan_set_halo_solver_index  remote_internal[k] = the destination rank's int_of_ext[remote[k]] for every pair k of
 *                                  the preceding fx_plan_set_halo call (same order; -1 where the dof lies outside the
 *                                  node blocks on both sides).  Without it the partitioned BiCGStab solves run on the
 *                                  scalar compacted-CSR kernel.                                                        */
int fx_space_solver_index(fx_plan* plan, int space, int* int_of_ext, int* n_internal);
int fx_plan_set_halo_solver_index(fx_plan* plan, int space, int nsend, const int* remote_internal);

/* solve(A, x, b, "bicgstab"|"cg"|"gmres", pc) (incomp_LDC.py:331): one persistent cooperative kernel per
 * solve (gmres: restart 30, classical Gram-Schmidt, PETSc KSPGMRES defaults; single-GPU plans); pc none,
 * jacobi or jacobi_coarse.  FX_PC_ILU0 (bicgstab, cg; single-GPU plans): PETSc's serial "default", ILU(0) in the
 * natural ordering, dependency-driven factorisation and triangular solves (csrc/fx_ilu.cu) -- PETSc's iteration
 * counts at the cost of its serial chain; this variant waits for the solve even through the _async entry.  x holds the initial guess (parameters['krylov_solver']['nonzero_initial_guess']=True,
 * incomp_LDC.py:267).  Convergence: ||M^-1 r||_2 <= max(rtol*||M^-1 b||_2, atol) (PETSc default
 * for left preconditioning).  Synchronous: waits for the solve, fills *info (may be NULL) and
 * returns FX_ERR_NOCONV on failure, like dolfin's RuntimeError.                                 */
int fx_krylov_solve(fx_plan* plan, int space, const double* val_dev, double* x_dev,
                    const double* b_dev, int method, int pc, double rtol, double atol, int maxit,
                    fx_solve_info* info, void* stream);
/* Asynchronous variant: returns after enqueueing; the outcome is written to the plan's result slot
 * `ticket` (0 <= ticket < FX_NTICKETS) and fetched later with fx_solve_results.                */
#define FX_NTICKETS 64
int fx_krylov_solve_async(fx_plan* plan, int space, const double* val_dev, double* x_dev,
                          const double* b_dev, int method, int pc, double rtol, double atol,
                          int maxit, int ticket, void* stream);
/* waits for `stream`, copies result slots [first, first+count) to info_host */
int fx_solve_results(fx_plan* plan, int first, int count, fx_solve_info* info_host, void* stream);
/* same copy, enqueued on `stream` without waiting (info_host should be page-locked); the caller
 * synchronises with an event before reading -- lets a time loop enqueue step k+1 while step k runs */
int fx_solve_results_async(fx_plan* plan, int first, int count, fx_solve_info* info_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif
