"""Worker of the multi-process sub-module-split test / demo (run under torch.distributed.run):
every rank marches its cell, peers exchange partials over NVLink peer stores, rank 0 gathers the RGBA8
blocks and compares with the single-tree frame it renders itself.  With --backend gloo (no GPU) only
the host plumbing runs: partition arithmetic, tree restriction, handle exchange."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--depth", type=int, default=7)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--hybrid", type=int, default=0, help="cells per group of the row-block x cell mode (0: plain split)")
    ap.add_argument("--guided", action="store_true", help="guided sampling with the sub-modules sharded by cell")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import mega_nerf_viewer_b200 as mnv

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    gpu = args.backend == "nccl"
    if gpu:
        torch.cuda.set_device(local)
    dist.init_process_group(args.backend)
    tree = mnv.synth.make_tree(depth=args.depth)
    w, h = args.width, args.height
    P = w * h
    MG = mnv.multigpu
    boxes = mnv.synth.cell_boxes(mnv.synth.grid_for_world(world), world)
    first, n = MG.owner_range(P, world, rank)

    if not gpu:
        # host plumbing only: the ranges tile the frame, the restricted trees cover the full tree's leaves,
        # 64-byte handles survive the object all-gather
        sub = mnv.synth.restrict_tree(tree, boxes[rank])
        mine = (bytes([rank]) * 64, bytes([rank + 100]) * 64)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        ranges = [None] * world
        dist.all_gather_object(ranges, (first, n, sub.capacity))
        ok = all(everyone[r][0] == bytes([r]) * 64 and everyone[r][1] == bytes([r + 100]) * 64 for r in range(world))
        covered = sum(r[1] for r in ranges) == P and all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0]
                                                         for i in range(world - 1))
        rng = np.random.default_rng(rank)
        b = boxes[rank]
        pts = (rng.random((5000, 3)) * (b[3:] - b[:3]) + b[:3]).astype(np.float32)
        qa, qb = mnv.synth.brute_force_query(tree, pts), mnv.synth.brute_force_query(sub, pts)
        same = bool(np.array_equal(qa[:, 2], qb[:, 2]) and
                    np.array_equal(tree.data[qa[:, 0], qa[:, 1]].view(np.uint16), sub.data[qb[:, 0], qb[:, 1]].view(np.uint16)))
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok and covered and same))
        if rank == 0:
            print(json.dumps({"world": world, "ok": all(flags), "nodes": [r[2] for r in ranges],
                              "full_nodes": tree.capacity}), flush=True)
        dist.destroy_process_group()
        return 0 if all(flags) else 1

    if args.guided:
        return guided_main(args, mnv, MG, dist, tree, rank, world, local)
    if args.hybrid:
        sp = MG.HybridSplit(tree, w, h, rank, world, local, dist, cells=args.hybrid)
        first, n = sp.pixel_range()
        blk_cap = sp.split.block
    else:
        sp = MG.SubmoduleSplit(tree, w, h, rank=rank, world=world, device=local, dist=dist)
        blk_cap = sp.block
    ranges = [None] * world
    dist.all_gather_object(ranges, (first, n, blk_cap))
    cap_max = max(r[2] for r in ranges)
    opt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8])
    full = mnv.DeviceTree(tree, device=local) if rank == 0 else None
    worst, fracs, psnrs = 0, [], []
    # all frames back to back with no host synchronisation between the ranks (the device-side flags and the two
    # buffer parities order them), results checked afterwards
    blocks = []
    for f in range(args.frames):
        cam = mnv.synth.default_camera(w, h, pose=f)
        blocks.append(sp.render_block(cam, opt).clone())
    torch.cuda.synchronize()
    for f in range(args.frames):
        cam = mnv.synth.default_camera(w, h, pose=f)
        blk = blocks[f]
        padded = torch.zeros((cap_max, 4), dtype=torch.uint8, device=blk.device)
        padded[: blk.shape[0]] = blk
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, parts, dst=0)
        if rank == 0:
            frame = torch.zeros((P, 4), dtype=torch.uint8, device=blk.device)
            for (f0, fn, _), part in zip(ranges, parts):
                frame[f0:f0 + fn] = part[:fn]
            got = frame.view(h, w, 4).cpu().numpy()
            want = full.render(cam, opt).cpu().numpy()
            d = np.abs(got.astype(int) - want.astype(int))
            worst = max(worst, int(d.max()))
            fracs.append(float((d <= 1).mean()))
            mse = float(np.mean(d.astype(np.float64) ** 2))
            psnrs.append(99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse))
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"world": world, "frames": args.frames, "max_abs": worst, "frac_within_1": min(fracs),
                          "psnr": min(psnrs), "local_nodes": sp.split.local_nodes if args.hybrid else sp.local_nodes,
                          "full_nodes": tree.capacity}), flush=True)
        full.close()
    sp.close()
    dist.destroy_process_group()
    return 0


def guided_main(args, mnv, MG, dist, tree, rank, world, local):
    import torch

    w, h = args.width, args.height
    P = w * h
    grid = mnv.synth.grid_for_world(world)
    subs = [mnv.synth.make_mlp_weights(seed=11 + i) for i in range(world)]
    mn, mx = (-1, -1, -1), (1, 1, 1)
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True,
                               appearance_embedding=0)
    sh = MG.ShardedGuided(tree, subs, grid, mn, mx, w, h, rank=rank, world=world, device=local, dist=dist)
    solo = MG.ReplicatedPipeline(tree, subs, grid, mn, mx, device=local) if rank == 0 else None
    first, n = MG.owner_range(P, world, rank)
    blocks, rows = [], []
    for f in range(args.frames):  # back to back: flags and buffer parities order the frames
        blk, r = sh.guided_block(mnv.synth.default_camera(w, h, pose=f), gopt)
        blocks.append(blk.clone())
        rows.append(r)
    torch.cuda.synchronize()
    worst, fracs, psnrs, rows_all, rows_want = 0, [], [], 0, 0
    for f in range(args.frames):
        padded = torch.zeros((sh.block, 4), dtype=torch.uint8, device=f"cuda:{local}")
        padded[:n] = blocks[f]
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, parts, dst=0)
        rt = torch.tensor([rows[f]], device=f"cuda:{local}")
        dist.all_reduce(rt)
        if rank == 0:
            got = torch.cat(parts)[:P].view(h, w, 4).cpu().numpy()
            want, rw = solo.guided_block(mnv.synth.default_camera(w, h, pose=f), gopt)
            d = np.abs(got.astype(int) - want.cpu().numpy().astype(int))
            worst = max(worst, int(d.max()))
            fracs.append(float((d <= 1).mean()))
            mse = float(np.mean(d.astype(np.float64) ** 2))
            psnrs.append(99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse))
            rows_all += int(rt.item())
            rows_want += rw
    if rank == 0:
        print(json.dumps({"world": world, "frames": args.frames, "max_abs": worst, "frac_within_1": min(fracs),
                          "psnr": min(psnrs), "rows": rows_all, "rows_want": rows_want,
                          "local_nodes": sh.local_nodes, "full_nodes": tree.capacity}), flush=True)
        solo.close()
    sh.close()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
