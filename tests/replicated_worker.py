"""Worker of the replicated-pipeline multi-GPU test (torch.distributed.run): guided sampling by row blocks and
refinement with all-gathered votes must equal the single-GPU pipeline bit for bit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import mega_nerf_viewer_b200 as mnv

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    MG = mnv.multigpu
    tree = mnv.synth.make_tree(depth=6)
    subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(2)]
    grid, mn, mx = (1, 2), (-1, -1, -1), (1, 1, 1)
    w, h = 320, 176
    cam = mnv.synth.default_camera(w, h, pose=2)
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True, appearance_embedding=0)
    ropt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_splitting=True, appearance_embedding=0,
                               split_batch_size=256)
    cap = tree.capacity + 4000
    pipe = MG.ReplicatedPipeline(tree, subs, grid, mn, mx, rank=rank, world=world, device=local, dist=dist, max_capacity=cap)
    solo = MG.ReplicatedPipeline(tree, subs, grid, mn, mx, device=local, max_capacity=cap) if rank == 0 else None
    def gather_frame(pieces):
        """pieces: [(first_row, rows tensor)] or a full-size frame with only this rank's bands filled (the rest zero):
        the partitions are disjoint, so the sum over ranks is the frame."""
        full = torch.zeros((h, w, 4), dtype=torch.uint8, device=f"cuda:{local}")
        if isinstance(pieces, list):
            for first, img in pieces:
                full[first:first + img.shape[0]] = img
        else:
            full.copy_(pieces)
        dist.all_reduce(full)
        return full.cpu().numpy() if rank == 0 else None

    ok = {}
    blk, rows = pipe.guided_blocks(cam, gopt)
    frame = gather_frame(blk)
    if rank == 0:
        want, _ = solo.guided_block(cam, gopt)
        ok["guided_equal"] = bool(np.array_equal(frame, want.cpu().numpy()))
    added = 0
    for f in range(4):
        blk, k = pipe.refine_frame(cam, ropt)
        added += k
        frame = gather_frame(blk)
        if rank == 0:
            want, k1 = solo.refine_frame(cam, ropt)
            ok[f"refine_frame_{f}_equal"] = bool(np.array_equal(frame, want.cpu().numpy())) and k1 == k
    sums = [None] * world
    dist.all_gather_object(sums, pipe.tree_checksum())
    if rank == 0:
        ok["replicas_identical"] = len(set(sums)) == 1
        ok["same_as_single_gpu"] = sums[0] == solo.tree_checksum()
        print(json.dumps({"world": world, "added": added, "capacity": pipe.dt.capacity, **ok}), flush=True)
        solo.close()
    pipe.close()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
