"""Level-1 drop-in surface: modules with the NAMES of the reference's helper modules.

The reference scripts start with `from fenics_fem import *` (lid-driven-cavity/incomp_LDC.py:36,
buoyancy-driven-flow/compressible_dgp.py:34) or `from fenics_base import *`
("flow between eccentrically rotating cylinders"/fenepmp_jbp.py:35).  Putting one of

    fenics_b200/shims/lid_driven_cavity      -> fenics_fem   (lid-driven-cavity/fenics_fem.py)
    fenics_b200/shims/buoyancy_driven_flow   -> fenics_fem   (buoyancy-driven-flow/fenics_fem.py)
    fenics_b200/shims/eccentric_cylinders    -> fenics_base  ("flow between ..."/fenics_base.py)

on sys.path in place of the reference directory gives the same helper names and call signatures.  What
sits behind them:
  * hot-path names (`assemble`, `solve`, `project`, `KrylovSolver`, `DirichletBC`, `Function`, `norm`,
    `parameters`, `function_spaces`, `solution_functions`, the diagnostics) are the GPU-backed objects of
    fenics_b200.dolfin_api / .spaces / .drivers -- no CPU fallback;
  * host utilities (progress bar, skew maps, ramp functions, .npy series, `chunks`) are plain numpy;
  * the symbolic tensor helpers (`Dincomp`, `sigma`, `Fdef`, `phi_ewm`, ...) are written against an algebra
    namespace (`ufl` when dolfin is installed; see _common.set_algebra): the GPU path does not evaluate
    them -- every form they build has a hand-written functor (include/fenics_b200.h fx_form) -- they exist so
    script code that composes expressions for its own post-processing keeps working under dolfin;
  * mesh refinement (`refine_*`, `adaptive_refinement`'s refine() call) and mshr mesh generation stay in
    dolfin (north_star): without dolfin those names raise with that message, and the structured
    synthetic meshes of fenics_b200.mesh stand in.
"""
