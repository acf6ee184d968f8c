"""Row-sharding plumbing for the multi-GPU path: one process per GPU,
`torch.distributed` for rendezvous and host-side scalar reductions, NCCL (inside
libconicip_b200.so) for the data path (Gram all-reduce, A'v all-reduce).

`shard_cones` is pure host logic (tested on CPU with gloo, world_size 2)."""
import numpy as np


def shard_cones(cone_dims, nranks):
    """Split the rows of A into `nranks` contiguous cone-aligned slabs of near-equal size.
    A large R cone may be cut anywhere (W is diagonal there); Q/S cones are never cut.
    Returns a list of (row_lo, row_hi, local_cone_dims)."""
    pieces = []                      # (type, size) atoms in row order; R cones are splittable
    for t, k in cone_dims:
        pieces.append((t, int(k)))
    m = sum(k for _, k in pieces)
    targets = [int(np.floor(m * (r + 1) / nranks + 0.5)) for r in range(nranks)]   # == llround in cip_shard_plan
    out, row, cur, lo, r = [], 0, [], 0, 0
    for t, k in pieces:
        while k > 0:
            if r == nranks - 1:
                take = k
            elif t == "R":
                take = min(k, max(targets[r] - row, 0))
                if take == 0:
                    out.append((lo, row, cur)); cur, lo, r = [], row, r + 1
                    continue
            else:
                # whole cone goes to the rank whose target it straddles least
                if row >= targets[r] or (row + k - targets[r] > targets[r] - row and cur):
                    out.append((lo, row, cur)); cur, lo, r = [], row, r + 1
                    continue
                take = k
            if cur and cur[-1][0] == "R" and t == "R":
                cur[-1] = ("R", cur[-1][1] + take)
            else:
                cur.append((t, take))
            row += take
            k -= take
    out.append((lo, row, cur))
    while len(out) < nranks:
        out.append((row, row, []))
    return out


class TorchReducer:
    """Scalar reductions over the row shards through torch.distributed (any backend)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self._t, self._d, self._g = torch, dist, group
        self.nranks = dist.get_world_size(group)
        self._dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"

    def _red(self, x, op):
        t = self._t.tensor([x], dtype=self._t.float64, device=self._dev)
        self._d.all_reduce(t, op=op, group=self._g)
        return float(t.item())

    def sum(self, x):
        return self._red(x, self._d.ReduceOp.SUM)

    def min(self, x):
        return self._red(x, self._d.ReduceOp.MIN)

    def sum_tensor_(self, t):
        """In-place sum of a small device (nccl) or host (gloo) tensor across ranks."""
        if t.device.type != self._dev:
            u = t.to(self._dev)
            self._d.all_reduce(u, op=self._d.ReduceOp.SUM, group=self._g)
            t.copy_(u)
        else:
            self._d.all_reduce(t, op=self._d.ReduceOp.SUM, group=self._g)
        return t

    def min_list(self, xs):
        t = self._t.tensor(list(xs), dtype=self._t.float64, device=self._dev)
        self._d.all_reduce(t, op=self._d.ReduceOp.MIN, group=self._g)
        return t.tolist()


def init_engine_comm(engine, group=None):
    """Create the NCCL communicator inside the engine: rank 0 makes the unique id, the host
    broadcasts it, every rank calls cip_comm_init."""
    import torch.distributed as dist
    from .engine import nccl_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.comm_init(world, rank, box[0])
