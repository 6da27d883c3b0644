"""Multi-GPU plumbing of HP2: one process per GPU, torch.distributed for
the rendezvous, a private NCCL communicator inside the CUDA library for the per-PCG-step
all-reduce of the camera-side vector.  Points are sharded; cameras are replicated."""
import ctypes as C

import numpy as np

from . import _abi, _lib


def init_from_torch(dist, rank, world, device=None):
    """Create the library's NCCL communicator: rank 0 makes the unique id, the bytes travel
    through the caller's process group (any backend), every rank calls psfm_dist_init."""
    import torch
    lib = _lib.lib()
    uid = (C.c_uint8 * _abi.NCCL_UNIQUE_ID_BYTES)()
    if rank == 0:
        _lib.check(lib.psfm_dist_get_unique_id(uid), "psfm_dist_get_unique_id")
    t = torch.tensor(list(uid), dtype=torch.uint8, device=device if device is not None else "cpu")
    dist.broadcast(t, 0)
    uid = (C.c_uint8 * _abi.NCCL_UNIQUE_ID_BYTES)(*t.cpu().tolist())
    _lib.check(lib.psfm_dist_init(uid, rank, world), "psfm_dist_init")


def finalize():
    _lib.lib().psfm_dist_finalize()


def owned_points(problem):
    """Indices of the points this shard observes (and therefore updates)."""
    return np.unique(problem.obs_point)


def merge_points(problem, dist, world):
    """After a sharded solve every rank holds the refined cameras (replicated) but only its
    own points; gather the points so that every rank ends with the full model."""
    import torch
    mine = owned_points(problem)
    counts = [None] * world
    dist.all_gather_object(counts, int(mine.shape[0]))
    nmax = max(counts)
    idx = np.full(nmax, -1, np.int64); idx[:mine.shape[0]] = mine
    val = np.zeros((nmax, 3)); val[:mine.shape[0]] = problem.xyz[mine]
    ti, tv = torch.from_numpy(idx), torch.from_numpy(val)
    gi = [torch.empty_like(ti) for _ in range(world)]
    gv = [torch.empty_like(tv) for _ in range(world)]
    if dist.get_backend() == "nccl":
        ti, tv = ti.cuda(), tv.cuda()
        gi, gv = [g.cuda() for g in gi], [g.cuda() for g in gv]
    dist.all_gather(gi, ti)
    dist.all_gather(gv, tv)
    for a, b in zip(gi, gv):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        ok = a >= 0
        problem.xyz[a[ok]] = b[ok]
    return problem
