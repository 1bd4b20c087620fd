"""Config 2 on a mesh partition: one process per GPU, owned/ghost dofs, NCCL halo exchange
(SURVEY.md 8e.2).  Same loop body as drivers/ldc.py::CompLDC -- only the linear solves (distributed
Krylov + ghost refresh) and the functionals (owned cells + all-reduce) differ."""
import torch
import torch.distributed as dist

from .. import dist_la
from .._lib import check, lib
from ..partition import LocalPartition
from ..spaces import function_spaces
from .ldc import CompLDC


class DistCompLDC(CompLDC):
    def __init__(self, mesh, rank=None, world=None, group=None, **kw):
        on = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank(group) if on else 0) if rank is None else rank
        self.world = (dist.get_world_size(group) if on else 1) if world is None else world
        W, V, _, _, Zd, Zc, Q, Qt, _ = function_spaces(mesh, 2)
        self.global_spaces = dict(W=W, V=V, Zd=Zd, Zc=Zc, Q=Q, Qt=Qt)
        self.lp = LocalPartition(mesh, self.global_spaces, self.world, self.rank)
        super().__init__(self.lp.mesh, dofmaps=self.lp.dofmaps(), **kw)
        self.ctx = dist_la.DistContext(self.plan, self.lp, group)
        check(lib.fx_plan_set_owned_cells(self.plan.h, self.lp.nc_owned))
        self.group = group

    def _solve(self, A, x, b, method=None, pc=None, rtol=None, label=""):
        method = method or self.solver[0]
        pc = pc or self.solver[1]
        rtol = (self.solve_rtol if rtol is None else rtol) or 1e-5
        self.solve_labels.append("%s:%s:%s" % (label, A.space, method))
        with self._phase("solve:%s:%s:%s" % (label, A.space, method)):
            self.last_info.append(dist_la.krylov_solve(self.ctx, A, x.vector(), b, method, pc, rtol=rtol))

    def _finish_step(self):
        if self.world > 1:
            dist.all_reduce(self.scal, group=self.group)  # functionals: owned cells summed over ranks
        self._scal_host.copy_(self.scal, non_blocking=False)
        return self._scal_host.numpy()

    def owned_fields(self):
        """{name: (global dof ids, values)} of this rank's owned entries -- for gathering / checking."""
        out = {}
        for name, fn, sp in (("w1", self.w1, "W"), ("p1", self.p1, "Q"), ("tau1", self.tau1_vec, "Zc"),
                             ("rho1", self.rho1, "Q")):
            m = self.lp.owned[sp]
            out[name] = (self.lp.l2g[sp][m], fn.vector().get_local()[m])
        return out
