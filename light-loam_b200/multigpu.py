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


# ---- BASELINE.json configs[3]: one long scan stream cut into contiguous segments, one per GPU ----------
# Every rank runs the fused pipeline on its own segment (+ 1 overlap scan so that the first increment of segment r+1 is
# the scan pair (last of r, first of r+1)), with its odometry starting at identity.  The only exchange is one all-gather
# of 7 doubles per rank — the segment's total transform — followed by a local prefix product that places the segment's
# poses in the global frame (q_w = q_prefix * q_local, t_w = t_prefix + q_prefix * t_local; LO:830-831 chaining).
def quat_mul(a, b):
    """Hamilton product, x,y,z,w storage (Eigen order), float64."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return [aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz]


def quat_rotate(q, v):
    """Eigen's q * v: uv = 2 (u x v); v + w uv + u x uv."""
    x, y, z, w = q
    ux, uy, uz = 2 * (y * v[2] - z * v[1]), 2 * (z * v[0] - x * v[2]), 2 * (x * v[1] - y * v[0])
    return [v[0] + w * ux + (y * uz - z * uy), v[1] + w * uy + (z * ux - x * uz), v[2] + w * uz + (x * uy - y * ux)]


def segment_ranges(n_scans, world):
    """Contiguous [begin, end) of every rank's segment of an n_scans stream; rank r > 0 also reads scan begin-1 (overlap)."""
    base, rem = divmod(n_scans, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def chain_segments(local_poses, dist):
    """local_poses: (n, 7) array q(4) t(3) of this rank's segment in the segment's own frame (first scan = identity).
    One all-gather of the segments' last poses, then the prefix product; returns the (n, 7) poses in the global frame."""
    import numpy as np
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    last = torch.tensor(np.asarray(local_poses[-1], dtype=np.float64))
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    last = last.to(dev)
    allp = [torch.zeros(7, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allp, last)
    q, t = [0.0, 0.0, 0.0, 1.0], [0.0, 0.0, 0.0]
    for r in range(rank):                      # prefix over the earlier segments
        p = allp[r].cpu().tolist()
        rt = quat_rotate(q, p[4:7])
        t = [t[0] + rt[0], t[1] + rt[1], t[2] + rt[2]]
        q = quat_mul(q, p[0:4])
    out = np.zeros((len(local_poses), 7))
    for k, p in enumerate(np.asarray(local_poses, dtype=np.float64)):
        rt = quat_rotate(q, p[4:7])
        out[k, 0:4] = quat_mul(q, p[0:4])
        out[k, 4:7] = [t[0] + rt[0], t[1] + rt[1], t[2] + rt[2]]
    return out
