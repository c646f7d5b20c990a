"""Batch sharding across the GPUs of one box: QPs are independent, so the batch is cut into
contiguous problem-index ranges, one per rank (one process per GPU), with NO collective on the solve path.
The only exchange is the optional final gather of the results onto rank 0 ("final host gather").

Used under ``torchrun`` / ``torch.distributed`` (backend nccl on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous index range [lo, hi) of rank `rank`: the first B % world ranks get one extra problem."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(inputs, B, rank, world):
    """Slice every [B, ...] array of `inputs` to this rank's range (None entries pass through)."""
    lo, hi = shard_range(B, rank, world)
    return {k: (None if v is None else v[lo:hi]) for k, v in inputs.items()}


def gather_results(local, B, dist=None, dst=0, group=None):
    """Gather per-rank result arrays (dict of numpy arrays whose leading dimension is the shard size) onto rank
    `dst` in problem order.  Returns the full dict on `dst`, None elsewhere.  With dist=None (single process)
    the local dict is returned unchanged.  `group`: the process group of the gather — on a GPU box pass a gloo group
    (``dist.new_group(backend="gloo")``): the results are host arrays and this is a host-side gather; with the default
    (NCCL) group the pickled shards would take a detour through device memory."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    host = {k: np.ascontiguousarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v)
            for k, v in local.items() if not k.startswith("_")}
    parts = [None] * world if rank == dst else None
    dist.gather_object(host, parts, dst=dst, group=group)
    if rank != dst:
        return None
    out = {}
    for k in host:
        out[k] = np.concatenate([p[k] for p in parts], axis=0)
        if out[k].shape[0] != B:
            raise RuntimeError("gathered %r has %d rows, expected %d" % (k, out[k].shape[0], B))
    return out


def solve_sharded(solver, B, x0, dist=None, gather=True, group=None, **inputs):
    """Solve this rank's contiguous slice of a B-problem batch on this rank's GPU; optionally gather on rank 0.

    `solver` is this rank's ``BatchSolver`` (device = LOCAL_RANK).  Every rank passes the same full-batch host
    arrays (or already-sharded ones when ``len(x0) != B``)."""
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    if x0.shape[0] == B:
        lo, hi = shard_range(B, rank, world)
        x0 = x0[lo:hi]
        inputs = {k: (v[lo:hi] if (v is not None and hasattr(v, "shape") and v.shape[0] == B) else v) for k, v in inputs.items()}
    res = solver.solve(x0, **inputs)
    return gather_results(res, B, dist, group=group) if gather else res
