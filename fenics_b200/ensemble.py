"""Independent-ensemble mode: the reference's parameter sweeps -- the (We, Ma) loops that
parameters-fenep-mp.csv drives in fenepmp_jbp.py:153-154,291-297,1080-1094 -- are embarrassingly parallel.
One process per GPU (torchrun); members are dealt round-robin over the ranks, every rank advances its own
members on its own device with no data-path collective, and rank 0 gathers the diagnostics at the end
(host-side `gather_object`).  Works under any torch.distributed backend (nccl on the box, gloo in the
CPU tests of the host logic)."""
import torch.distributed as dist


def assign(members, rank, world):
    """round-robin deal: member i -> rank i % world."""
    return [(i, m) for i, m in enumerate(members) if i % world == rank]


def run_ensemble(members, run_member, rank=None, world=None):
    """run_member(index, member) -> picklable result, executed on this rank for its share of `members`.
    Returns the full result list (ordered by member index) on rank 0, None elsewhere."""
    distributed = dist.is_available() and dist.is_initialized()
    rank = (dist.get_rank() if distributed else 0) if rank is None else rank
    world = (dist.get_world_size() if distributed else 1) if world is None else world
    mine = [(i, run_member(i, m)) for i, m in assign(members, rank, world)]
    if not distributed or world == 1:
        return [r for _, r in sorted(mine)]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    flat = sorted((i, r) for part in gathered for i, r in part)
    assert [i for i, _ in flat] == list(range(len(members))), "ensemble members lost or duplicated"
    return [r for _, r in flat]


def run_fenepmp_sweep(steps, input_csv=None, device=None, all_lambda=False, **driver_kw):
    """The config-4 ensemble: every (j, jjj) member of the We x Ma sweep for `steps` time steps; each result is
    the member's parameters plus its last diagnostics (chi = F_x/F_y is the script's stability measure, :768)."""
    from .drivers import jbp

    def run_member(i, m):
        drv = jbp.FenepmpJBP(Re=m["Re"], We=m["We"], Ma=m["Ma"], lambda_d=m["lambda_d"], device=device, **driver_kw)
        drv.check = False
        out = None
        for _ in range(steps):
            out = drv.step()
        out = dict(out)
        out["chi"] = out["F_x"] / out["F_y"] if out["F_y"] != 0.0 else float("nan")
        out.update(j=m["j"], jjj=m["jjj"], We=m["We"], Ma=m["Ma"], Re=m["Re"], lambda_d=m["lambda_d"])
        return out

    return run_ensemble(jbp.sweep(jbp.read_parameters(input_csv), all_lambda=all_lambda), run_member)
