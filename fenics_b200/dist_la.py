"""Distributed (mesh-partitioned) Krylov solvers: SURVEY.md 8e.2.

Per rank: the local matrix (owned rows complete), local vectors with owned + ghost entries.  Two paths:

* **fused** (default on a CUDA/NCCL process group): every rank's Krylov work vectors live in a symmetric
  memory region all peers map over NVLink (`fx_plan_set_peer_regions`, `fx_plan_set_halo`); the SAME
  persistent kernels as the single-GPU solves then store interface values straight into the neighbours'
  ghost slots, close every grid barrier with a cross-GPU flag exchange and sum the dot products across
  ranks -- one kernel launch per solve per rank, no NCCL call, no host read inside a solve.
* **host-driven** (`fused=False`; the cross-check, and what the gloo CPU tests of the host logic mirror):
  ghost rows replaced by identity rows, one iteration = halo exchange of the SpMV input (all_to_all of packed interface
values), the rank-local CSR-stream SpMV (`fx_spmv`), owner-masked dot products (`fx_dotw_dev`) all-reduced
as ONE small tensor per reduction point, and fused vector updates whose scalar coefficients stay in device
memory (`fx_lincomb3_dev`) -- the host only reads the residual norm once per iteration for the stopping
test.  Jacobi preconditioning (identical to the serial solver, so iteration counts match the 1-GPU run).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _abi as abi
from . import la
from ._lib import SolveInfo, check, lib


class DistContext:
    """owner masks / halos of one LocalPartition on a device, plus scratch."""

    def __init__(self, plan, lp, group=None, fused=None):
        self.plan, self.lp, self.group = plan, lp, group
        dev = "cuda:%d" % plan.device
        on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if fused is None:
            fused = on and dist.get_backend(group) == "nccl" and os.environ.get("FX_DIST_FUSED", "1") != "0"
        self.fused = bool(fused)
        if self.fused:
            self._setup_fused(dev)
        self.mask = {k: torch.as_tensor(lp.owned[k].astype("float64"), device=dev) for k in lp.owned}
        self.ghost_rows = {k: torch.as_tensor((~lp.owned[k]).nonzero()[0].astype("int32"), device=dev) for k in lp.owned}
        self.zeros = {k: torch.zeros(len(self.ghost_rows[k]), dtype=torch.float64, device=dev) for k in lp.owned}
        self.scal = torch.zeros(16, dtype=torch.float64, device=dev)
        self.coef = torch.zeros(3, dtype=torch.float64, device=dev)
        self._work = {}

    def _setup_fused(self, dev):
        """symmetric (peer-mapped) region per rank + push lists of every space with a CSR pattern."""
        import numpy as np
        import torch.distributed._symmetric_memory as symm_mem
        lp, group = self.lp, self.group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        # stride of the work vectors inside a region: the Krylov kernels keep them in the internal node-major numbering,
        # which pads the 3 stress components of a node to 4 doubles (fx_nb_host.h)
        # (a multiple of 4 doubles: the kernels gather node blocks with 16- and 32-byte loads)
        nmax = ((4 * int(lp.n_local_max) + 2) // 3 + 8 + 3) // 4 * 4
        nbytes = int(lib.fx_peer_region_bytes(nmax))
        self._region = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self._region.zero_()
        pg = group if group is not None else dist.group.WORLD
        self._hdl = symm_mem.rendezvous(self._region, group=pg.group_name if hasattr(pg, "group_name") else pg)
        torch.cuda.synchronize()
        dist.barrier(group)  # every region is zero before anybody's first solve
        ptrs = (C.c_void_p * world)(*[int(p) for p in self._hdl.buffer_ptrs])
        check(lib.fx_plan_set_peer_regions(self.plan.h, rank, world, ptrs, nmax))
        for sp in ("W", "Zc", "Q"):
            if sp not in lp.halo:
                continue
            h = lp.halo[sp]
            peer = np.ascontiguousarray(np.repeat(np.arange(world), h.send_counts), dtype=np.int32)
            local = np.ascontiguousarray(h.send_idx_np, dtype=np.int32)
            remote = np.ascontiguousarray(h.send_remote_np, dtype=np.int32)
            # 0 = owned, else 1 + owner rank: the solver layout numbers ghost nodes per owner, in the owner's order
            ghost = np.ascontiguousarray(np.where(lp.owned[sp], 0, lp.owner[sp][lp.l2g[sp]] + 1), dtype=np.uint8)
            check(lib.fx_plan_set_halo(self.plan.h, abi.SPACE_ID[sp], len(peer), peer.ctypes.data, local.ctypes.data,
                                       remote.ctypes.data, ghost.ctypes.data))
            # the neighbours' INTERNAL (node-major) index of every ghost entry I push to: every rank publishes its
            # ext -> internal map, one all-gather per space at set-up (a collective: all ranks take part even when a
            # rank's space has no node-block layout -- then nobody uses it for this space)
            idx = np.full(nmax + 1, -1, dtype=np.int32)
            n_int = C.c_int(0)
            n_loc = len(ghost)
            mine = np.empty(n_loc, dtype=np.int32)
            rc = lib.fx_space_solver_index(self.plan.h, abi.SPACE_ID[sp], mine.ctypes.data, C.byref(n_int))
            if rc == 0 and os.environ.get("FX_NB", "1") != "0":
                idx[:n_loc] = mine
                idx[nmax] = 1
            mine_t = torch.as_tensor(idx, device=dev)
            allidx = [torch.empty_like(mine_t) for _ in range(world)]
            dist.all_gather(allidx, mine_t, group=group)
            allidx = torch.stack(allidx).cpu().numpy()
            if (allidx[:, nmax] == 1).all() and len(peer):
                rint = np.ascontiguousarray(allidx[peer, remote], dtype=np.int32)
                check(lib.fx_plan_set_halo_solver_index(self.plan.h, abi.SPACE_ID[sp], len(peer), rint.ctypes.data))

    def close(self):
        """detach the plan from the peer regions (before the symmetric memory goes away)."""
        if self.fused and self.plan.h:
            lib.fx_plan_set_peer_regions(self.plan.h, 0, 0, None, 0)
            self.fused = False

    def work(self, space, n, count):
        key = (space, count)
        if key not in self._work:
            self._work[key] = [torch.zeros(n, dtype=torch.float64, device=self.scal.device) for _ in range(count)]
        return self._work[key]

    def halo_update(self, space, vec):
        self.lp.halo[space].update(vec.t if hasattr(vec, "t") and not torch.is_tensor(vec) else vec, self.group)

    def seal_ghost_rows(self, A, b):
        """ghost rows -> identity rows, b_ghost = 0 (keeps every iterate finite; dots ignore ghosts)."""
        rows = self.ghost_rows[A.space]
        if len(rows):
            check(lib.fx_apply_bc(self.plan.h, abi.SPACE_ID[A.space], la._ptr(A.val), la._ptr(b.t) if b is not None else None,
                                  la._ptr(rows), la._ptr(self.zeros[A.space]), len(rows), la._stream()))

    def dots(self, space, pairs):
        """all-reduced owner-masked dot products of the given (x, y) tensor pairs -> device tensor view."""
        n = self.mask[space].numel()
        for i, (x, y) in enumerate(pairs):
            check(lib.fx_dotw_dev(self.plan.h, n, la._ptr(x), la._ptr(y), la._ptr(self.mask[space]),
                                  C.c_void_p(self.scal.data_ptr() + 8 * i), la._stream()))
        out = self.scal[:len(pairs)]
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(out, group=self.group)
        return out


class ReplicatedSolver:
    """Small systems of a partitioned run, solved redundantly on every GPU.

    The P1 pressure / density / temperature space is ~27x smaller than the whole problem (37k of 1M dofs at
    n=192) and its CG solve is latency bound: ~10^3 iterations of microseconds, each with grid barriers that
    would cross NVLink on a partition.  So every rank contributes its OWNED rows (values, right-hand side,
    initial guess) to one zero-padded buffer, ONE all-reduce (NCCL; x + 0 = x exactly, so all ranks hold
    bitwise identical systems) replicates the global system, and every rank runs the single-GPU solver on
    it -- including the shared-memory-resident CG kernel and its two-level preconditioner, which a
    partitioned solve cannot use -- then keeps its owned + ghost entries.  No halo exchange is needed
    afterwards.  The classic "replicate the coarse problem" device of domain decomposition."""

    def __init__(self, ctx, global_mesh, global_space, space="Q"):
        import numpy as np
        self.ctx, self.space = ctx, space
        lp, dev = ctx.lp, ctx.scal.device
        self.gplan = la.Plan(global_mesh, [global_space], ctx.plan.device)
        rp_l, col_l = ctx.plan.pattern_host(space)
        rp_g, col_g = self.gplan.pattern_host(space)
        l2g = lp.l2g[space]
        own = np.flatnonzero(lp.owned[space])
        g = l2g[own]
        cnt = (rp_l[own + 1] - rp_l[own]).astype(np.int64)
        assert np.array_equal(cnt, rp_g[g + 1] - rp_g[g]), "owned rows are not complete on the sub-mesh"
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        lpos = np.repeat(rp_l[own].astype(np.int64), cnt) + within
        gpos = np.repeat(rp_g[g].astype(np.int64), cnt) + within
        assert np.array_equal(l2g[col_l[lpos]], col_g[gpos]), "local and global column orders differ"
        self.nnz, self.n = len(col_g), len(rp_g) - 1
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int64, device=dev)  # noqa: E731
        self.lpos, self.own, self.l2g = T(lpos), T(own), T(l2g)
        self.dst = T(np.concatenate([gpos, self.nnz + g, self.nnz + self.n + g]))
        self.buf = torch.zeros(self.nnz + 2 * self.n, dtype=torch.float64, device=dev)
        self.A = la.Matrix(self.gplan, space)
        self.A.val = self.buf[:self.nnz]
        self.b = la.Vector(self.gplan, space, self.buf[self.nnz:self.nnz + self.n])
        self.x = la.Vector(self.gplan, space, self.buf[self.nnz + self.n:])

    def solve(self, A, x, b, method, pc, rtol, atol=None, maxit=None, ticket=None):
        self.buf.zero_()
        self.buf.index_copy_(0, self.dst, torch.cat([A.val.index_select(0, self.lpos), b.t.index_select(0, self.own),
                                                     x.t.index_select(0, self.own)]))
        if dist.is_initialized() and dist.get_world_size(self.ctx.group) > 1:
            dist.all_reduce(self.buf, group=self.ctx.group)
        info = la.krylov_solve(self.A, self.x, self.b, method, pc, rtol=rtol, atol=atol, maxit=maxit, ticket=ticket)
        torch.index_select(self.x.t, 0, self.l2g, out=x.t)
        return info


def _lincomb(n, out, c0, x, c1, y, c2, z, coef):
    """out = c0 x + c1 y + c2 z with 0-d device tensors (or floats) as coefficients."""
    coef[0], coef[1], coef[2] = c0, c1, c2
    check(lib.fx_lincomb3_dev(n, la._ptr(out), la._ptr(coef), la._ptr(x), la._ptr(y), la._ptr(z), la._stream()))


def _spmv(ctx, A, x, y):
    ctx.lp.halo[A.space].update(x, ctx.group)
    check(lib.fx_spmv(A.plan.h, abi.SPACE_ID[A.space], la._ptr(A.val), la._ptr(x), la._ptr(y), la._stream()))


def _diag_inv(ctx, A, pc):
    """Jacobi: 1/diag(A) by one SpMV-free pass: diag = A * e_i is not available, so read it from CSR."""
    n = A.n
    if pc in ("none",):
        return torch.ones(n, dtype=torch.float64, device=A.val.device)
    key = ("diagpos", A.space)
    if key not in ctx._work:
        rp, col = A.plan.pattern_host(A.space)
        import numpy as np
        rows = np.repeat(np.arange(n), np.diff(rp))
        pos = np.flatnonzero(col == rows)
        assert len(pos) == n, "pattern lacks a diagonal entry"
        ctx._work[key] = torch.as_tensor(pos, dtype=torch.int64, device=A.val.device)
    d = A.val.index_select(0, ctx._work[key])
    return torch.where(d != 0, 1.0 / d, torch.ones_like(d))


def krylov_solve(ctx, A, x, b, method="bicgstab", pc="default", rtol=1e-5, atol=1e-50, maxit=10000, ticket=None):
    """Distributed solve(A, x, b, method, pc); x, b are la.Vector on the local (owned+ghost) dofs.
    Same stopping rule as the 1-GPU kernels: ||M^-1 r|| <= max(rtol ||M^-1 b||, atol).  Returns SolveInfo;
    raises RuntimeError like dolfin when not converged."""
    if pc == "default":
        pc = "jacobi"
    if method not in ("bicgstab", "cg") or pc not in ("jacobi", "none"):
        raise ValueError("distributed solver supports bicgstab/cg with jacobi/none")
    sp, n = A.space, A.n
    if ctx.fused:
        if ticket is not None:  # enqueue only: the outcome lands in the plan's result slot (la.solve_results)
            check(lib.fx_krylov_solve_async(ctx.plan.h, abi.SPACE_ID[sp], la._ptr(A.val), la._ptr(x.t), la._ptr(b.t),
                                            la._KSP[method], la._PC[pc], float(rtol), float(atol), int(maxit), int(ticket),
                                            la._stream()))
            return None
        info = SolveInfo()
        rc = lib.fx_krylov_solve(ctx.plan.h, abi.SPACE_ID[sp], la._ptr(A.val), la._ptr(x.t), la._ptr(b.t),
                                 la._KSP[method], la._PC[pc], float(rtol), float(atol), int(maxit), C.byref(info), la._stream())
        if rc == abi.FX_ERR_NOCONV and not la.parameters["krylov_solver"]["error_on_nonconvergence"]:
            return info
        check(rc)
        return info
    ctx.seal_ghost_rows(A, b)
    dinv = _diag_inv(ctx, A, pc)
    xt, bt, coef = x.t, b.t, ctx.coef
    info = SolveInfo()
    if method == "cg":
        r, u, p, q = ctx.work(sp, n, 4)
        _spmv(ctx, A, xt, q)
        torch.sub(bt, q, out=r)
        torch.mul(dinv, r, out=u)
        p.copy_(u)
        bu = dinv * bt
        d = ctx.dots(sp, [(r, u), (u, u), (bu, bu)]).clone()
        gamma, un, bn = d[0].clone(), float(d[1].sqrt()), float(d[2].sqrt())
        tol = max(rtol * bn, atol)
        it = 0
        while un > tol and it < maxit:
            it += 1
            _spmv(ctx, A, p, q)
            pq = ctx.dots(sp, [(p, q)])[0].clone()
            alpha = gamma / pq
            _lincomb(n, xt, 1.0, xt, alpha, p, 0.0, p, coef)
            _lincomb(n, r, 1.0, r, -alpha, q, 0.0, q, coef)
            torch.mul(dinv, r, out=u)
            d = ctx.dots(sp, [(r, u), (u, u)]).clone()
            beta = d[0] / gamma
            gamma = d[0].clone()
            _lincomb(n, p, 1.0, u, beta, p, 0.0, p, coef)
            un = float(d[1].sqrt())  # the one host read per iteration
            if not (un == un):
                break
        info.iterations, info.residual, info.residual0 = it, un, bn
        info.converged = 1 if un <= tol else (0 if it >= maxit else -1)
    else:
        r, rh, p, v, s, t = ctx.work(sp, n, 6)
        _spmv(ctx, A, xt, v)
        torch.sub(bt, v, out=r)
        r.mul_(dinv)
        rh.copy_(r)
        p.zero_()
        v.zero_()
        bu = dinv * bt
        d = ctx.dots(sp, [(r, r), (bu, bu)]).clone()
        rho, rn, bn = d[0].clone(), float(d[0].sqrt()), float(d[1].sqrt())
        tol = max(rtol * bn, atol)
        one = torch.ones((), dtype=torch.float64, device=xt.device)
        rho_old, alpha, omega = one.clone(), one.clone(), one.clone()
        rhn, it, restarts = rn, 0, 0
        while rn > tol and it < maxit:
            it += 1
            if abs(float(rho)) <= 1e-14 * rhn * rn or float(omega) == 0.0:  # breakdown -> restart with rhat = r
                restarts += 1
                if restarts > 50:
                    break
                rh.copy_(r)
                p.zero_()
                v.zero_()
                rho = torch.as_tensor(rn * rn, dtype=torch.float64, device=xt.device)
                rhn = rn
                rho_old, alpha, omega = one.clone(), one.clone(), one.clone()
            beta = (rho / rho_old) * (alpha / omega)
            _lincomb(n, p, 1.0, r, beta, p, -beta * omega, v, coef)
            _spmv(ctx, A, p, v)
            v.mul_(dinv)
            rv = ctx.dots(sp, [(rh, v)])[0].clone()
            alpha = rho / rv
            _lincomb(n, s, 1.0, r, -alpha, v, 0.0, v, coef)
            _spmv(ctx, A, s, t)
            t.mul_(dinv)
            d = ctx.dots(sp, [(t, s), (t, t)]).clone()
            omega = d[0] / d[1]
            _lincomb(n, xt, 1.0, xt, alpha, p, omega, s, coef)
            _lincomb(n, r, 1.0, s, -omega, t, 0.0, t, coef)
            d = ctx.dots(sp, [(rh, r), (r, r)]).clone()
            rho_old, rho = rho, d[0].clone()
            rn = float(d[1].sqrt())  # the one host read per iteration
            if not (rn == rn):
                break
        info.iterations, info.residual, info.residual0 = it, rn, bn
        info.converged = 1 if rn <= tol else (0 if it >= maxit else -1)
    ctx.lp.halo[sp].update(xt, ctx.group)  # the solution's ghost entries are coefficients of the next forms
    if info.converged != 1 and la.parameters["krylov_solver"]["error_on_nonconvergence"]:
        raise RuntimeError("*** Error: Unable to solve linear system (distributed %s): %d iterations, residual %.3e"
                           % (method, info.iterations, info.residual))
    return info
