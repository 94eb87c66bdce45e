"""Host-side plumbing of the multi-GPU scan-to-map mode (BASELINE.json configs[4]): slab bounds per rank
and the exchange of the mailbox handles over torch.distributed.  The data path itself (the 28-double all-reduce of
every linearisation) runs inside the LM kernel over peer memory — nothing here is on it."""
import math


def slab_bounds(rank, world, x_min, x_max):
    """[lo, hi) of rank's slab: `world` equal slabs of [x_min, x_max) along map-frame x; the outer slabs are open-ended so
    that every stack point has exactly one owner."""
    if world < 1 or not (0 <= rank < world) or not (x_min < x_max):
        raise ValueError("bad slab request")
    w = (x_max - x_min) / world
    lo = -math.inf if rank == 0 else x_min + rank * w
    hi = math.inf if rank == world - 1 else x_min + (rank + 1) * w
    return lo, hi


def slab_with_halo(points, lo, hi, halo=1.0):
    """Rows of an (n, >=3) map cloud a rank needs for exact 5-NN of its own queries: its slab plus `halo` metres
    (the acceptance radius is 1 m: LM:1884 / LM:1952)."""
    x = points[:, 0]
    return points[(x >= lo - halo) & (x < hi + halo)]


def exchange_handles(handle, dist):
    """All-gather of the 64-byte mailbox handles (host side, any backend)."""
    world = dist.get_world_size()
    out = [None] * world
    dist.all_gather_object(out, bytes(handle))
    return out


def attach_all(ctx, dist, x_min, x_max):
    """Connects ctx to the contexts of all other ranks and gives it its slab.  Call right after creating the context."""
    rank, world = dist.get_rank(), dist.get_world_size()
    handles = exchange_handles(ctx.comm_export(), dist)
    ctx.comm_attach(rank, world, handles)
    lo, hi = slab_bounds(rank, world, x_min, x_max)
    ctx.map_set_slab(lo, hi)
    dist.barrier()
    return lo, hi
