"""Sub-module split across the GPUs of one box (second mode) — host plumbing.

One process per GPU.  Rank g owns cell g of the Mega-NeRF (y, z) grid: the restriction of the
octree to that cell (and, for guided sampling / refinement, that cell's sub-MLP).  Per frame every
rank marches all rays through its cell, the march kernel stores each ray's partial straight into the
memory of the rank that owns the pixel (peer stores over NVLink through CUDA-IPC mappings), raises a
flag there, and every rank composites its contiguous block of P / world pixels.  torch.distributed is
used ONCE, at set-up, to exchange the 64-byte IPC handles; the per-frame data path has no collective
and no host synchronisation between ranks (csrc/mnv_multigpu.cu).
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import synth


def block_pixels(n_pixels: int, world: int) -> int:
    """Pixels per owner: contiguous ranges of the row-major frame, the last one may be short."""
    return (n_pixels + world - 1) // world


def owner_range(n_pixels: int, world: int, rank: int):
    b = block_pixels(n_pixels, world)
    first = min(rank * b, n_pixels)
    return first, max(0, min(b, n_pixels - first))


class DeviceBuffer:
    """cudaMalloc'd (not torch-cached) device memory: its base address can be exported over CUDA IPC."""

    def __init__(self, nbytes: int, device: int):
        from . import _check, lib

        self.ptr = C.c_void_p()
        self.nbytes = nbytes
        _check(lib().mnv_malloc(C.byref(self.ptr), nbytes, device))
        _check(lib().mnv_memset(self.ptr, 0, nbytes, None))

    def handle(self) -> bytes:
        from . import _check, lib

        buf = C.create_string_buffer(64)
        _check(lib().mnv_ipc_export(self.ptr, buf))
        return buf.raw

    def free(self):
        from . import lib

        if self.ptr:
            lib().mnv_free(self.ptr)
            self.ptr = C.c_void_p()


def open_handle(handle: bytes, device: int) -> int:
    from . import _check, lib

    p = C.c_void_p()
    _check(lib().mnv_ipc_open(handle, C.byref(p), device))
    return p.value


class SubmoduleSplit:
    """Per-rank state of the split renderer.  `dist` is torch.distributed (initialised) or None for a
    single process that plays every rank in turn on one GPU (tests)."""

    def __init__(self, tree, width: int, height: int, rank: int = 0, world: int = 1, device: int = 0,
                 dist=None, grid_dim=None, restrict: bool = True, group=None):
        import torch

        from . import DeviceTree

        self.rank, self.world, self.device, self.dist, self.group = rank, world, device, dist, group
        self.P = width * height
        self.block = block_pixels(self.P, world)
        self.grid_dim = tuple(grid_dim) if grid_dim is not None else synth.grid_for_world(world)
        self.boxes = synth.cell_boxes(self.grid_dim, world)
        assert self.boxes.shape[0] == world, "one spatial cell per rank"
        self.single = dist is None
        cells = range(world) if self.single else [rank]
        self.trees = {}
        for c in cells:
            sub = synth.restrict_tree(tree, self.boxes[c]) if restrict and world > 1 else tree
            self.trees[c] = DeviceTree(sub, device=device)
        self.local_nodes = sum(t.capacity for t in self.trees.values())
        # receive side: [2 frame parities][world slots][block] float4 partials + [world] u32 flags.
        # Two parities make the frames self-synchronising: a producer can start frame k+1 only after its own
        # composite of frame k, which waited for every peer's frame-k flag, which each peer raised after its
        # composite of frame k-1 — so the buffer of parity (k+1) & 1 is no longer being read anywhere.
        self.stride = world * self.block * 16
        self.partials = DeviceBuffer(2 * self.stride, device)
        self.flags = DeviceBuffer(64, device)
        self.frame_id = 0
        if self.single:
            self.dst = [self.partials.ptr.value] * world
            self.flag_dst = [self.flags.ptr.value] * world
        else:
            mine = (self.partials.handle(), self.flags.handle())
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            self.dst, self.flag_dst, self._opened = [], [], []
            for r, (hp, hf) in enumerate(everyone):
                if r == rank:
                    self.dst.append(self.partials.ptr.value)
                    self.flag_dst.append(self.flags.ptr.value)
                else:
                    a, b = open_handle(hp, device), open_handle(hf, device)
                    self._opened += [a, b]
                    self.dst.append(a)
                    self.flag_dst.append(b)
            dist.barrier(group=group)
        self.out = torch.empty(self.block * 4, dtype=torch.uint8, device=f"cuda:{device}")

    def march(self, cam, opt, stream=None):
        """This rank's segment of every ray -> the owners' memories, then the flags."""
        from . import _check, _stream_ptr, lib

        self.frame_id += 1
        par = (self.frame_id & 1) * self.stride
        for cell, dt in self.trees.items():
            dt.render_partial(cam, opt, [a + par for a in self.dst], self.block, cell, cell_box=self.boxes[cell],
                              stream=stream)
            if not self.single:
                arr = (C.c_void_p * self.world)(*[C.c_void_p(a) for a in self.flag_dst])
                _check(lib().mnv_signal_peers(arr, self.world, cell, self.frame_id, _stream_ptr(stream)))

    def composite(self, cam, opt, owner=None, stream=None):
        """RGBA8 of this rank's pixel block (device tensor [n_pixels, 4])."""
        owner = self.rank if owner is None else owner
        first, n = owner_range(self.P, self.world, owner)
        dt = next(iter(self.trees.values()))
        dt.composite_partials(cam, opt, self.partials.ptr.value + (self.frame_id & 1) * self.stride, self.world,
                              self.block, self.boxes, first, n,
                              self.out, flags_ptr=0 if self.single else self.flags.ptr.value,
                              wait_value=self.frame_id, stream=stream)
        return self.out[: n * 4].view(n, 4)

    def render_block(self, cam, opt, stream=None):
        self.march(cam, opt, stream)
        return self.composite(cam, opt, stream=stream)

    def render_full_single(self, cam, opt):
        """Single-process mode: the whole frame, every owner's block in turn (tests)."""
        import torch

        assert self.single
        cam_w, cam_h = (cam["width"], cam["height"]) if isinstance(cam, dict) else (cam.width, cam.height)
        frame = torch.empty((self.P, 4), dtype=torch.uint8, device=f"cuda:{self.device}")
        # with one receive buffer the owners are processed one at a time: march fills block `o` of every slot
        for o in range(self.world):
            first, n = owner_range(self.P, self.world, o)
            self._march_for_owner(cam, opt, o)
            frame[first:first + n] = self.composite(cam, opt, owner=o).clone()
        return frame.view(cam_h, cam_w, 4)

    def _march_for_owner(self, cam, opt, owner):
        # every slot writes pixel p to dst[p // block]; in single mode all dst alias one buffer, so only the
        # pixels of `owner` may land: pass a scratch buffer for the other owners
        if not hasattr(self, "_scratch"):
            self._scratch = DeviceBuffer(self.world * self.block * 16, self.device)
        self.frame_id += 1
        dst = [self._scratch.ptr.value] * self.world
        dst[owner] = self.partials.ptr.value + (self.frame_id & 1) * self.stride
        for cell, dt in self.trees.items():
            dt.render_partial(cam, opt, dst, self.block, cell, cell_box=self.boxes[cell])

    def close(self):
        from . import lib

        import torch

        torch.cuda.synchronize()
        if not self.single:
            self.dist.barrier(group=self.group)
            for a in self._opened:
                lib().mnv_ipc_close(C.c_void_p(a))
            self.dist.barrier(group=self.group)
        for t in self.trees.values():
            t.close()
        self.partials.free()
        self.flags.free()
        if hasattr(self, "_scratch"):
            self._scratch.free()


# ---------------------------------------------------------------------------------------------------------
# Image-row blocks with the tree AND the sub-MLPs replicated (first mode) for the two paths
# that run the MLP: guided sampling and dynamic refinement.  The eight sub-modules are 1.2 MB of bf16 each —
# replicating them is free — so guided sampling needs no exchange at all: every rank emits, evaluates and
# composites the rays of its own rows (a windowed camera: same intrinsics, cy shifted by the first row).
# Refinement needs one collective per frame: the per-ray votes (the two [P, 3] tracker arrays) are
# all-gathered, after which every rank runs the identical, deterministic select -> split -> MLP -> commit
# sequence on its replica, so the replicas stay bit-identical without broadcasting payloads.


def row_block(height: int, world: int, rank: int):
    """Rows [first, first + n) of rank's block; blocks are multiples of 8 rows (the march kernel's tile height)."""
    per = ((height + world - 1) // world + 7) // 8 * 8
    first = min(rank * per, height)
    return first, max(0, min(per, height - first))


def row_blocks(height: int, world: int, rank: int, parts: int = 4):
    """[(first, n)]: the frame cut into parts * world row blocks (multiples of 8 rows), dealt round-robin."""
    m = parts * world
    per = ((height + m - 1) // m + 7) // 8 * 8
    out = []
    for b in range(rank, m, world):
        first = min(b * per, height)
        out.append((first, max(0, min(per, height - first))))
    return out


def window_camera(cam: dict, first_row: int, n_rows: int) -> dict:
    """The camera of rows [first_row, first_row + n_rows): identical rays, bit for bit, when cy is a multiple of 0.5."""
    return dict(cam, height=int(n_rows), cy=float(cam["cy"]) - float(first_row))


class ReplicatedPipeline:
    def __init__(self, tree, submodules, grid_dim, min_position, max_position, rank=0, world=1, device=0,
                 dist=None, max_capacity=0, seed=0x5eed):
        from . import DeviceTree, MlpModel

        self.rank, self.world, self.device, self.dist, self.seed = rank, world, device, dist, seed
        self.dt = DeviceTree(tree, max_capacity=max_capacity, device=device)
        self.model = MlpModel(submodules, grid_dim=grid_dim, min_position=min_position, max_position=max_position,
                              device=device)
        self.grid_dim = list(grid_dim)
        self.min_position = list(min_position)
        self.range = [float(b) - float(a) for a, b in zip(min_position, max_position)]
        self.data_dim = tree.data_dim
        self.steps = 0
        import os
        self.stage_timing = os.environ.get("MNV_STAGE_TIMING") == "1"
        self.stages = {}

    def guided_block(self, cam: dict, opt, capacity_rows=None):
        """RGBA8 [rows of this rank, W, 4] of the guided-sampling frame (one contiguous row block per rank)."""
        first, n = row_block(cam["height"], self.world, self.rank)
        out, total = self._guided_rows(cam, opt, first, n, capacity_rows)
        return out, total

    def _guided_rows(self, cam, opt, first, n, capacity_rows):
        import torch

        wc = window_camera(cam, first, n)
        cap = capacity_rows or max(1 << 16, n * cam["width"] * 24)
        timing = os.environ.get("MNV_STAGE_TIMING") == "1"  # dev: CUDA events around the three stages (adds synchronisations)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timing else None
        if timing:
            ev[0].record()
        g = self.dt.guided_samples(wc, opt, self.grid_dim, self.min_position, self.range, capacity_rows=cap)
        vals = torch.empty((max(g["total"], 1), self.data_dim + 1), device=f"cuda:{self.device}")
        if timing:
            ev[1].record()
        if g["total"]:
            self.model.query_submodules(g["cluster"], g["rows"], vals)
        if timing:
            ev[2].record()
        img = self.dt.render_nerf_results(wc, opt, vals, g["z_vals"], g["offsets"], sigma_col=self.data_dim - 1)
        if timing:
            ev[3].record()
            torch.cuda.synchronize()
            print(f"[guided rows, rank {self.rank}] {g['total']} rows: emission {ev[0].elapsed_time(ev[1]):.3f} ms, "
                  f"MLP {ev[1].elapsed_time(ev[2]):.3f} ms, compositing {ev[2].elapsed_time(ev[3]):.3f} ms", file=sys.stderr, flush=True)
        return img, g["total"]

    def guided_blocks(self, cam: dict, opt, capacity_rows=None, parts: int = 4):
        """This rank's share of the guided-sampling frame: list of (first_row, RGBA8 [rows, W, 4]) sub-blocks and the
        MLP rows evaluated.  The frame is cut into parts * world row blocks dealt round-robin (sky and ground rows
        cost very different amounts: contiguous halves left one GPU with most of the samples); each sub-block is
        emitted, evaluated and composited on its own through a windowed camera — no exchange."""
        out, total = [], 0
        for first, n in row_blocks(cam["height"], self.world, self.rank, parts if self.world > 1 else 1):
            if n == 0:
                continue
            img, rows = self._guided_rows(cam, opt, first, n, capacity_rows)
            out.append((first, img))
            total += rows
        return out, total

    def _stage(self, name, t0=None):
        """MNV_STAGE_TIMING=1: synchronising per-stage clock (dev; the timed numbers of bench.py run without it)."""
        import time
        import torch

        if not self.stage_timing:
            return None
        torch.cuda.synchronize()
        now = time.perf_counter()
        if t0 is not None:
            self.stages[name] = self.stages.get(name, 0.0) + (now - t0) * 1e3
        return now

    def refine_frame(self, cam: dict, opt, band_rows: int = 8):
        """One frame with refinement on.  Returns (RGBA8 [H, W, 4] with this rank's bands filled, nodes added).

        Per rank: its interleaved bands (b % world == rank, like the image-tile mode: balanced whatever the scene)
        marched with vote tracking -> the votes reduced to (leaf, priority, count) records (mnv_vote_reduce; a
        fraction of the rays) -> ONE all-gather of the records (NCCL) -> the identical selection on every replica ->
        children linked on every replica -> the n*8*c MLP rows SHARDED by child across the ranks -> each rank reduces
        its children to fp16 payload records -> ONE all-gather of the records (64 B per child) -> committed on every
        replica.  Replicas stay bit-identical to each other and to the one-GPU sequence (per-row MLP results do not
        depend on the batch they are evaluated in)."""
        import torch

        from . import select_candidates, select_from_votes, vote_reduce

        dev = f"cuda:{self.device}"
        W, H = cam["width"], cam["height"]
        P = W * H
        if getattr(self, "_ts", None) is None or self._ts.shape[0] != P:
            self._ts = torch.empty((P, 3), device=dev)
            self._tp = torch.empty((P, 3), device=dev)
            self._img = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
        ts, tp, img = self._ts, self._tp, self._img
        t = self._stage("begin")
        if self.world > 1:
            ts.fill_(-1.0)  # rows of the other ranks' bands: no candidate
            tp.fill_(-1.0)
            self.dt.render_tiles(cam, opt, img, ((W + 15) // 16) * 16, band_rows, self.world, self.rank, to_split=ts,
                                 to_sample=tp)
            t = self._stage("march (own bands, vote tracking)", t)
            rec = vote_reduce(ts, cap_records=P // self.world + W * band_rows + 1024)  # [r, 3] i32, r known on the host
            t = self._stage("vote reduce", t)
            cnt = torch.tensor([rec.shape[0]], device=dev, dtype=torch.int64)
            self.dist.all_reduce(cnt, op=self.dist.ReduceOp.MAX)
            rmax = max(int(cnt.item()), 1)
            send = torch.zeros((rmax, 3), dtype=torch.int32, device=dev)  # votes == 0: padding
            send[: rec.shape[0]] = rec
            allrec = torch.empty((self.world * rmax, 3), dtype=torch.int32, device=dev)
            self.dist.all_gather_into_tensor(allrec, send)
            self.vote_bytes = int(allrec.numel() * 4)
            t = self._stage("vote records all-gather", t)
            nodes, _ = select_from_votes(allrec, opt.split_batch_size, "split")
        else:
            self.dt.render(cam, opt, out=img, to_split=ts, to_sample=tp)
            t = self._stage("march (own bands, vote tracking)", t)
            nodes, _ = select_candidates(ts, opt.split_batch_size, "split")
        t = self._stage("selection", t)
        k = nodes.shape[0]
        self.steps += 1
        if k == 0 or self.dt.capacity + k > self.dt.max_capacity:
            return img, 0
        c = opt.samples_per_corner
        rd = 3 + (3 if opt.need_viewdir else 0) + (1 if opt.appearance_embedding != -1 else 0)
        g = torch.Generator(device=dev).manual_seed(self.seed + self.steps)  # the same numbers on every replica
        samples = torch.rand((k * 8, c, rd), device=dev, generator=g)
        cluster = torch.zeros((k * 8, c), dtype=torch.int16, device=dev)
        self.dt.add_children(opt, nodes, samples, cluster, self.grid_dim, self.min_position, self.range)
        t = self._stage("rand + add children", t)
        if self.world == 1:
            results = torch.empty((k * 8 * c, self.data_dim + 1), device=dev)
            self.model.query_submodules(cluster.view(-1), samples.view(-1, rd), results)
            t = self._stage("query_submodules", t)
            self.dt.commit_children(opt, k, results.view(k * 8, c, -1))
            self._stage("commit", t)
            return img, k
        per = (k * 8 + self.world - 1) // self.world  # children per rank
        lo, hi = min(self.rank * per, k * 8), min((self.rank + 1) * per, k * 8)
        rb = self.dt.record_bytes
        mine = torch.zeros((per, rb), dtype=torch.uint8, device=dev)
        if hi > lo:
            results = torch.empty(((hi - lo) * c, self.data_dim + 1), device=dev)
            self.model.query_submodules(cluster[lo:hi].reshape(-1), samples[lo:hi].reshape(-1, rd), results)
            mine[: hi - lo] = self.dt.reduce_children(opt, results.view(hi - lo, c, -1))
        t = self._stage("query_submodules", t)
        allp = torch.empty((self.world * per, rb), dtype=torch.uint8, device=dev)
        self.dist.all_gather_into_tensor(allp, mine)
        self.payload_bytes = int(allp.numel())
        self.dt.commit_children_records(opt, k, allp)
        self._stage("payload all-gather + commit", t)
        return img, k

    def tree_checksum(self) -> int:
        data, child, parent, counts = self.dt.download()
        import zlib

        h = 0
        for a in (child, parent, data.view(np.uint16), counts):
            h = zlib.crc32(np.ascontiguousarray(a).tobytes(), h)
        return h

    def close(self):
        self.model.close()
        self.dt.close()


NO_SEGMENT = 3.0e38


def segment_context_host(records, slot: int, stop_thresh: float, max_samples: int):
    """Host mirror of `segment_context` (csrc/mnv_guided.cu) for one ray, used by the tests of the carry rule.
    records[c] = (T of cell c's segment marched alone, samples it emitted alone, z of its first sample).
    Returns (T at the entry of cell `slot`, samples the ray already has there, z of the first sample of the next
    segment that emits anything or NO_SEGMENT)."""
    order = sorted((c for c in range(len(records)) if records[c][1] > 0), key=lambda c: (records[c][2], c))
    T_in, count_in, z_next = 1.0, 0, NO_SEGMENT
    if slot not in order:
        z_mine = (NO_SEGMENT, slot)
    else:
        z_mine = (records[slot][2], slot)
    T_after, count_after = 1.0, 0
    for c in order:
        T_c, n_c, z_c = records[c]
        if (z_c, c) < z_mine:
            T_in *= T_c
            count_in += int(n_c)
            T_after, count_after = T_in, count_in
        elif c == slot:
            T_after, count_after = T_in * T_c, count_in + int(n_c)
        else:
            if not (T_after < stop_thresh) and count_after < max_samples:
                z_next = z_c
                break
            T_after *= T_c
            count_after += int(n_c)
    return T_in, count_in, z_next


class ShardedGuided(SubmoduleSplit):
    """Guided sampling with the sub-modules SHARDED (BASELINE.json configs[4]): rank g holds cell g's subtree and
    sub-MLP g only.  Per frame and rank: probe march of the cell -> one all-gather of 16 B per ray (NCCL) ->
    emission of the cell's samples with the transmittance / sample count the unsharded march would carry into the
    cell -> sub-MLP g on those rows -> per-segment compositor whose result leaves as one 16-byte peer store per
    ray into the pixel owner's memory -> flags -> the owner's front-to-back compositor.  No sample row, MLP
    output or tree node ever crosses NVLink."""

    def __init__(self, tree, submodules, grid_dim, min_position, max_position, width, height, rank=0, world=1,
                 device=0, dist=None, group=None):
        from . import MlpModel

        super().__init__(tree, width, height, rank=rank, world=world, device=device, dist=dist, grid_dim=grid_dim,
                         group=group)
        assert len(submodules) == world
        self.model_grid = list(grid_dim)
        self.min_position = list(min_position)
        self.range = [float(b) - float(a) for a, b in zip(min_position, max_position)]
        self.data_dim = tree.data_dim
        # one single-sub-module container per cell this process plays
        self.models = {c: MlpModel([submodules[c]], device=device) for c in self.trees}
        if self.single:
            # one receive buffer per owner (the base class aliases them; here every owner's block is kept)
            self._owner_bufs = DeviceBuffer(world * self.stride, device)
            self.dst = [self._owner_bufs.ptr.value + o * self.stride for o in range(world)]

    def guided_block(self, cam: dict, opt, capacity_rows=None, stream=None):
        """RGBA8 [n_pixels, 4] of this rank's pixel block (single-process mode: the whole frame) and the number of
        MLP rows this process evaluated."""
        import torch

        from . import _check, _stream_ptr, lib

        dev = f"cuda:{self.device}"
        P, W = self.P, self.world
        self.frame_id += 1
        table = torch.empty((W, P, 4), dtype=torch.float32, device=dev)
        for c, dt in self.trees.items():
            dt.guided_segment_probe(cam, opt, self.boxes[c], out=table[c], stream=stream)
        if not self.single:
            self.dist.all_gather_into_tensor(table.view(-1), table[self.rank].reshape(-1).clone(), group=self.group)
        par = 0 if self.single else (self.frame_id & 1) * self.stride
        rows_done = 0
        for c, dt in self.trees.items():
            cap = capacity_rows or max(1 << 16, int(table[c, :, 1].sum().item()))
            g = dt.guided_samples_segment(cam, opt, self.boxes[c], self.model_grid, self.min_position, self.range, table, c,
                                          capacity_rows=cap, stream=stream)
            vals = torch.empty((max(g["total"], 1), self.data_dim + 1), device=dev)
            if g["total"]:
                self.models[c].forward(g["rows"], 0, out=vals, stream=stream)
            rows_done += g["total"]
            dt.render_nerf_results_partial(cam, opt, vals, g["z_vals"], g["offsets"], table, c,
                                           [a + par for a in self.dst], self.block, sigma_col=self.data_dim - 1,
                                           stream=stream)
            if not self.single:
                arr = (C.c_void_p * W)(*[C.c_void_p(a) for a in self.flag_dst])
                _check(lib().mnv_signal_peers(arr, W, c, self.frame_id, _stream_ptr(stream)))
        dt0 = next(iter(self.trees.values()))
        if self.single:
            frame = torch.empty((P, 4), dtype=torch.uint8, device=dev)
            for o in range(W):
                first, n = owner_range(P, W, o)
                dt0.composite_partials(cam, opt, self.dst[o], W, self.block, self.boxes, first, n, self.out, guided=True,
                                       stream=stream)
                frame[first:first + n] = self.out[: n * 4].view(n, 4)
            return frame, rows_done
        first, n = owner_range(P, W, self.rank)
        dt0.composite_partials(cam, opt, self.partials.ptr.value + par, W, self.block, self.boxes, first, n, self.out,
                               flags_ptr=self.flags.ptr.value, wait_value=self.frame_id, guided=True, stream=stream)
        return self.out[: n * 4].view(n, 4), rows_done

    def close(self):
        for m in self.models.values():
            m.close()
        if hasattr(self, "_owner_bufs"):
            import torch

            torch.cuda.synchronize()
            self._owner_bufs.free()
        super().close()


class HybridSplit:
    """Row blocks x spatial cells: the `world` GPUs form world / cells groups; group g renders the rows of block g
    (windowed camera) and, inside the group, each GPU marches one spatial cell and owns 1 / cells of the block's
    pixels (SubmoduleSplit over the group's sub-communicator).  Cuts both terms of the frame time — the bulk
    (pixels / groups, visits / cells) and the longest-ray latency floor (a ray's visits are spread over the cells it
    crosses) — at the pixel tolerance of the split mode."""

    def __init__(self, tree, width, height, rank, world, device, dist, cells=4):
        assert world % cells == 0
        self.groups = world // cells
        self.g, self.c = rank // cells, rank % cells
        self.first_row, self.rows = row_block(height, self.groups, self.g)
        subgroup = None
        for gi in range(self.groups):  # every rank must create every group
            grp = dist.new_group(list(range(gi * cells, (gi + 1) * cells)))
            if gi == self.g:
                subgroup = grp
        self.split = SubmoduleSplit(tree, width, max(self.rows, 1), rank=self.c, world=cells, device=device, dist=dist,
                                    group=subgroup)
        self.width, self.height = width, height

    def render_block(self, cam: dict, opt, stream=None):
        """RGBA8 of this rank's pixels: rows [first_row, first_row + rows) of the frame, pixel range `owner_range`."""
        return self.split.render_block(window_camera(cam, self.first_row, self.rows), opt, stream=stream)

    def pixel_range(self):
        first, n = owner_range(self.width * self.rows, self.split.world, self.c)
        return self.first_row * self.width + first, n

    def close(self):
        self.split.close()
