"""torchrun --nproc-per-node N tools/mgpu_check.py : sharded HP2 solve == single-GPU solve."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from particlesfm_b200 import _abi, _lib, ba, distributed, synthetic as syn
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    lib = _lib.lib()
    _lib.check(lib.psfm_set_device(local), "set_device")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ctypes as C
    prob, truth = syn.make_ba_problem(40, 20000, 9, seed=21)
    ok = True

    def windowed(pb):
        """points renumbered by their first image: contiguous point ranges = windows of the video, so a
        shard sees only part of the images (the observed-flag all-reduce matters)"""
        first = np.full(pb.num_points, 1 << 30)
        np.minimum.at(first, pb.obs_point, pb.obs_image)
        order = np.argsort(first, kind="stable")
        inv = np.empty_like(order); inv[order] = np.arange(order.size)
        return _abi.BAProblem(pb.qvec, pb.tvec, pb.xyz[order], pb.cam_params, pb.obs_image, inv[pb.obs_point].astype(np.int32),
                              pb.obs_xy, pb.image_camera, pb.pose_constant, pb.tvec_constant_mask, pb.camera_constant)

    def empty_last(pb, r, w):
        """the last rank owns no observation at all; the others split the problem"""
        if w == 1:
            return pb
        if r == w - 1:
            z = np.zeros(0, np.int32)
            return _abi.BAProblem(pb.qvec, pb.tvec, pb.xyz, pb.cam_params, z, z, np.zeros((0, 2)), pb.image_camera,
                                  pb.pose_constant, pb.tvec_constant_mask, pb.camera_constant)
        return pb.shard(r, w - 1)

    cases = [("uniform", prob, lambda pb: pb.shard(rank, world), _abi.SOLVER_ITERATIVE_SCHUR),
             ("uniform", prob, lambda pb: pb.shard(rank, world), _abi.SOLVER_EXACT_SCHUR),
             ("windowed", windowed(prob), lambda pb: pb.shard(rank, world), _abi.SOLVER_EXACT_SCHUR),
             ("empty-last-rank", prob, lambda pb: empty_last(pb, rank, world), _abi.SOLVER_EXACT_SCHUR)]
    for name, base, shard_fn, solver in cases:
        o = _abi.BAOptions()
        lib.psfm_ba_global_options(C.byref(o))
        o.refine_rotation = 1; o.refine_focal_length = 1; o.print_summary = 0; o.minimizer_progress_to_stdout = 0
        o.linear_solver = solver
        single = base.copy()
        s1 = ba.solve_problem(single, o)                 # world size 1: communicator not yet created
        distributed.init_from_torch(dist, rank, world, device="cuda")
        shard = shard_fn(base).copy()
        sN = ba.solve_problem(shard, o)
        distributed.merge_points(shard, dist, world)
        distributed.finalize()
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        errs = dict(q=rel(shard.qvec, single.qvec), t=rel(shard.tvec, single.tvec), X=rel(shard.xyz, single.xyz),
                    K=rel(shard.cam_params, single.cam_params))
        line = dict(case=name, rank=rank, solver=solver, it1=s1.num_iterations, itN=sN.num_iterations, lin1=s1.num_linear_iterations,
                    linN=sN.num_linear_iterations, cost1=s1.final_cost, costN=sN.final_cost, world=sN.world_size, **errs)
        print(line, flush=True)
        good = (s1.num_iterations == sN.num_iterations and abs(s1.final_cost - sN.final_cost) <= 1e-9 * s1.final_cost
                and max(errs.values()) < 1e-8 and sN.world_size == world)
        ok = ok and good
    dist.barrier()
    dist.destroy_process_group()
    print("MGPU_CHECK", "PASS" if ok else "FAIL", "rank", rank, flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
