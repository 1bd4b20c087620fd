"""Mesh-partitioned multi-GPU mode: host-side partitioning and halo maps.

One process per GPU.  Cells are split by recursive coordinate bisection (no METIS header in the image);
a dof is owned by the lowest-numbered part among the cells that contain it.  A rank's sub-mesh holds its
owned cells first, then every ghost cell that touches one of its owned dofs, so

  * assembling over the sub-mesh gives COMPLETE rows for every owned dof with zero communication
    (redundant ghost-cell compute), and
  * the columns of those rows are owned-or-ghost dofs of the same sub-mesh.

Ghost entries of a vector are refreshed from their owners by `Halo.update` (one all_to_all of packed
interface values per exchange: NCCL over NVLink on the box, gloo in the CPU tests).  Every rank computes
the same global tables from the same global mesh, so no set-up communication is needed.
"""
import numpy as np
import torch
import torch.distributed as dist

from .mesh import Mesh
from .spaces import ELEMENTS, FunctionSpace

SPACES = ("W", "Zc", "Q", "Zd", "Qt", "V")


def rcb_partition(points, nparts):
    """recursive coordinate bisection: split along the longer extent at the weighted median."""
    part = np.zeros(len(points), dtype=np.int32)

    def rec(idx, lo, n):
        if n == 1:
            part[idx] = lo
            return
        p = points[idx]
        axis = int(np.argmax(p.max(0) - p.min(0)))
        order = idx[np.argsort(p[:, axis], kind="stable")]
        nl = n // 2
        cut = (len(order) * nl) // n
        rec(order[:cut], lo, nl)
        rec(order[cut:], lo + nl, n - nl)

    rec(np.arange(len(points)), 0, int(nparts))
    return part


class Halo:
    """ghost <- owner exchange of one space.  send_idx / recv_idx are local dof indices, concatenated in
    rank order; send_counts[q] values go to rank q, recv_counts[q] come from rank q."""

    def __init__(self, send_idx, send_counts, recv_idx, recv_counts, send_remote=None):
        self.send_idx_np, self.recv_idx_np = send_idx, recv_idx
        # send_remote[k]: index of send_idx[k]'s dof in the DESTINATION rank's local numbering (its ghost
        # slot) -- what the fused Krylov kernels store to over NVLink (fx_plan_set_halo)
        self.send_remote_np = send_remote
        self.send_counts, self.recv_counts = [int(c) for c in send_counts], [int(c) for c in recv_counts]
        self._dev = {}

    def _on(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.as_tensor(self.send_idx_np, dtype=torch.int64, device=device),
                              torch.as_tensor(self.recv_idx_np, dtype=torch.int64, device=device))
        return self._dev[key]

    def update(self, vec, group=None):
        """vec: 1-D float64 tensor of the rank's local dofs (owned + ghost), updated in place."""
        if sum(self.send_counts) == 0 and sum(self.recv_counts) == 0:
            return
        si, ri = self._on(vec.device)
        send = vec.index_select(0, si)
        recv = torch.empty(len(ri), dtype=vec.dtype, device=vec.device)
        dist.all_to_all_single(recv, send, self.recv_counts, self.send_counts, group=group)
        vec.index_copy_(0, ri, recv)


class LocalPartition:
    """Everything rank `rank` needs: its sub-mesh, local cell-dof tables, owned masks and halos."""

    def __init__(self, mesh, spaces, nparts, rank, part=None):
        cells = mesh.cells().astype(np.int64)
        nc = len(cells)
        self.rank, self.nparts = rank, nparts
        if part is None:
            part = rcb_partition(mesh.coordinates()[cells].mean(axis=1), nparts)
        self.part = part
        G = {k: np.asarray(s.cell_dofs, dtype=np.int64) for k, s in spaces.items()}
        dims = {k: s.dim() for k, s in spaces.items()}
        owner = {}
        for k, g in G.items():
            o = np.full(dims[k], nparts, dtype=np.int32)
            np.minimum.at(o, g.ravel(), np.repeat(part, g.shape[1]))
            owner[k] = o
        self.owner = owner

        def local_cells(r):
            touch = np.zeros(nc, dtype=bool)
            for k in ("W", "Zc", "Q"):
                if k in G:
                    touch |= (owner[k][G[k]] == r).any(axis=1)
            owned = np.flatnonzero(part == r)
            ghost = np.flatnonzero(touch & (part != r))
            return owned, ghost

        cell_sets = [local_cells(r) for r in range(nparts)]
        owned_c, ghost_c = cell_sets[rank]
        self.cells_global = np.concatenate([owned_c, ghost_c])
        self.nc_owned = len(owned_c)
        lc = cells[self.cells_global]
        self.verts_global = np.unique(lc)
        lcells = np.searchsorted(self.verts_global, lc)  # monotone map keeps UFC (ascending) order
        g2l_cell = np.full(nc, -1, dtype=np.int64)
        g2l_cell[self.cells_global] = np.arange(len(self.cells_global))
        keep = g2l_cell[mesh.bfacet_cell] >= 0
        bf = (g2l_cell[mesh.bfacet_cell[keep]], mesh.bfacet_local[keep], mesh.bfacet_marker[keep])
        self.mesh = Mesh(mesh.coordinates()[self.verts_global], lcells, boundary_facets=bf)
        assert np.array_equal(self.mesh.cells(), lcells)
        # spaces, owned masks, halos
        self.l2g, self.owned, self.halo, self.spaces = {}, {}, {}, {}
        dofsets = {k: [np.unique(G[k][np.concatenate(cs)]) for cs in cell_sets] for k in G}
        for k in G:
            l2g = dofsets[k][rank]
            self.l2g[k] = l2g
            self.owned[k] = owner[k][l2g] == rank
            self.spaces[k] = FunctionSpace(self.mesh, k, cell_dofs=np.searchsorted(l2g, G[k][self.cells_global]),
                                           comps=ELEMENTS.get(k))
            send_idx, send_counts, recv_idx, recv_counts, send_remote = [], [], [], [], []
            for q in range(nparts):
                if q == rank:
                    send_counts.append(0)
                    recv_counts.append(0)
                    continue
                # what q needs from me: dofs in q's sub-mesh that I own (sorted by global id on both sides)
                need_q = dofsets[k][q][owner[k][dofsets[k][q]] == rank]
                send_idx.append(np.searchsorted(l2g, need_q))
                send_remote.append(np.searchsorted(dofsets[k][q], need_q))
                send_counts.append(len(need_q))
                mine_q = l2g[owner[k][l2g] == q]
                recv_idx.append(np.searchsorted(l2g, mine_q))
                recv_counts.append(len(mine_q))
            cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.zeros(0, dtype=np.int64)  # noqa: E731
            self.halo[k] = Halo(cat(send_idx), send_counts, cat(recv_idx), recv_counts, cat(send_remote))
        self.n_local_max = max(len(d) for ds in dofsets.values() for d in ds)  # same on every rank

    def dofmaps(self):
        return {k: s.cell_dofs for k, s in self.spaces.items()}

    def scatter(self, space, global_vec):
        """global numpy vector -> this rank's local (owned + ghost) values."""
        return np.asarray(global_vec)[self.l2g[space]]

    def n_owned(self, space):
        return int(self.owned[space].sum())


def gather_owned(parts_local, partitions, space, n_global):
    """test helper: assemble a global vector from every rank's owned entries."""
    out = np.full(n_global, np.nan)
    for lp, v in zip(partitions, parts_local):
        m = lp.owned[space]
        out[lp.l2g[space][m]] = np.asarray(v)[m]
    return out
