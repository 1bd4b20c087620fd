"""The time loops on a mesh partition: one process per GPU, owned/ghost dofs, NCCL halo exchange
(SURVEY.md 8e.2).  Same loop bodies as the single-GPU drivers -- only the linear solves (distributed
Krylov + ghost refresh) and the functionals (owned cells + all-reduce) differ, so every driver gets its
partitioned variant from one mixin:

    DistCompLDC (config 2), DistEwmJBP (config 3), DistFenepmpJBP (config 4), DistCompDGP (config 5),
    DistIncompDGP (incompressible_dgp.py)
"""
import torch  # noqa: F401
import torch.distributed as dist

from .. import dist_la, la
from .._lib import check, lib
from ..partition import LocalPartition, rcb_partition
from ..spaces import function_spaces
from .dgp import CompDGP, IncompDGP
from .ewm import EwmJBP
from .jbp import FenepmpJBP
from .ldc import CompLDC

_RESULT_SPACE = dict(w1="W", p1="Q", tau1_vec="Zc", rho1="Q", T1="Q")


def distributed(base):
    class Dist(base):
        def __init__(self, mesh, rank=None, world=None, group=None, fused=None, replicate_q=None, **kw):
            on = dist.is_available() and dist.is_initialized()
            self.rank = (dist.get_rank(group) if on else 0) if rank is None else rank
            self.world = (dist.get_world_size(group) if on else 1) if world is None else world
            W, V, _, _, Zd, Zc, Q, Qt, _ = function_spaces(mesh, 2)
            self.global_spaces = dict(W=W, V=V, Zd=Zd, Zc=Zc, Q=Q, Qt=Qt)
            self.lp = LocalPartition(mesh, self.global_spaces, self.world, self.rank)
            if issubclass(base, EwmJBP) and kw.get("dt") is None:  # dt = c hmin^2 of the GLOBAL mesh (:164, :461-462)
                kw["dt"] = (1.0 if kw.get("prepass") else 10.0) * mesh.hmin() ** 2
            super().__init__(self.lp.mesh, dofmaps=self.lp.dofmaps(), **kw)
            self.ctx = dist_la.DistContext(self.plan, self.lp, group, fused=fused)
            check(lib.fx_plan_set_owned_cells(self.plan.h, self.lp.nc_owned))
            self.group = group
            # the small P1 systems are solved redundantly on every rank (dist_la.ReplicatedSolver)
            rep = (self.world > 1 and self.ctx.fused) if replicate_q is None else replicate_q
            self.rep_q = dist_la.ReplicatedSolver(self.ctx, mesh, Q, "Q") if rep else None

        def enable_pressure_coarse(self, k=None):
            """two-level preconditioner of the CG pressure solve: available on a partition only through the
            replicated Q solves (aggregates of the GLOBAL P1 space)."""
            if self.rep_q is None:
                return False
            Q = self.global_spaces["Q"]
            k = self._coarse_size(Q.dim(), k)
            ok = k >= 32 and self.rep_q.gplan.set_aggregates("Q", rcb_partition(Q.tabulate_dof_coordinates(), k), k)
            self.pressure_pc = "jacobi_coarse" if ok else None
            return bool(ok)

        def _solve(self, A, x, b, method=None, pc=None, rtol=None, label=""):
            method = method or self.solver[0]
            pc = pc or self.solver[1]
            rtol = (self.solve_rtol if rtol is None else rtol) or 1e-5
            self.solve_labels.append("%s:%s:%s" % (label, A.space, method))
            # check = False: the solves are only enqueued (fused kernels / replicated solves) and their outcomes fetched
            # once per step -- the host no longer waits for every solve before it can enqueue the next assembly
            lazy = not self.check and self.ctx.fused
            if self._ticket == 0:
                self._order, self._qticket = [], 0
            with self._phase("solve:%s:%s:%s" % (label, A.space, method)):
                if self.rep_q is not None and A.space == "Q":
                    info = self.rep_q.solve(A, x.vector(), b, method, pc, rtol, ticket=self._qticket if lazy else None)
                    if lazy:
                        self._order.append(("q", self._qticket))
                        self._qticket += 1
                else:
                    info = dist_la.krylov_solve(self.ctx, A, x.vector(), b, method, pc, rtol=rtol,
                                                ticket=len(self._order) - self._qticket if lazy else None)
                    if lazy:
                        self._order.append(("l", len(self._order) - self._qticket))
                if lazy:
                    self._ticket += 1   # solves enqueued this step (base class: reset at the start of step())
                else:
                    self.last_info.append(info)

        def _finish_step(self, allow_pipeline=True):
            if self.world > 1:
                dist.all_reduce(self.scal, group=self.group)  # functionals: owned cells summed over ranks
            if not self.check and self.ctx.fused and self._ticket:
                nq = self._qticket
                got = {"l": la.solve_results(self.plan, 0, self._ticket - nq) if self._ticket > nq else [],
                       "q": la.solve_results(self.rep_q.gplan, 0, nq) if nq else []}
                self.last_info = [got[k][i] for k, i in self._order]
                self._check_infos(self.last_info, self.solve_labels)
            self._scal_host.copy_(self.scal, non_blocking=False)
            return self._scal_host.numpy()

        def owned_fields(self):
            """{name: (global dof ids, values)} of this rank's owned entries -- for gathering / checking."""
            out = {}
            for name in self.RESULT_FIELDS or ("w1", "p1", "tau1_vec", "rho1"):
                sp = _RESULT_SPACE[name]
                m = self.lp.owned[sp]
                out[name] = (self.lp.l2g[sp][m], getattr(self, name).vector().get_local()[m])
            return out

    Dist.__name__ = Dist.__qualname__ = "Dist" + base.__name__
    Dist.__doc__ = "%s on a mesh partition (see drivers/ldc_dist.py)." % base.__name__
    return Dist


DistCompLDC = distributed(CompLDC)
DistEwmJBP = distributed(EwmJBP)
DistFenepmpJBP = distributed(FenepmpJBP)
DistCompDGP = distributed(CompDGP)
DistIncompDGP = distributed(IncompDGP)  # singular pressure system: the replicated solve keeps the Jacobi-Krylov gauge
