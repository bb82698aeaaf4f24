#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native octree render path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): one headless 1920x1080 frame of the
synthetic depth-10 SH-deg-2 N3Tree (mega-nerf-viewer_b200/synth.py), camera
orbiting 16 poses.  A "step" is one frame = one launch of the traversal kernel
over all rays of the frame.  Metric: Mrays/s (rays of the frame / device time).

  value     kernel-only throughput, tree resident in HBM, CUDA-event timed on
            the launching stream, L2 flushed between timed frames
  e2e       the same frames through the C-ABI host call
            (mnv_render_frame_host[_bands]): camera + options in from the host,
            RGBA8 frame read back into pinned host memory every step
  roofline  algorithmic bytes (6 B per empty leaf visit, 60 B per shaded visit,
            + 4 B / 28 B per pixel) over the measured kernel
            time, against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (port of the reference's device code) on the
            host cores, same frame

N > 1: the frame is split into interleaved 8-row bands, one process per GPU,
tree replicated, no data-path collective (strong scaling of one frame); `parity`
compares the union of the ranks' bands with the frame rank 0 renders alone.

Sections of the default line: `mlp` (the fused tcgen05 MLP on the 262 144-row
refinement batch, tensor roofline), `point_query`, `refinement` / `guided_sampling`
(configs 4 / 5 through the C++ driver), `target_4k` (the north-star case: 3840x2160
on the Mill-19-scale octree — tiles, refinement ON, guided sampling — with their own
parity and clock records), `cpu_baseline`.
--mode split | hybrid | guided | refine (N > 1; refine also N = 1) print their own
lines; split / guided carry an in-run `parity` (PSNR, max-abs) against the unsharded
/ replicated frame and the exchange bytes per GPU per frame.

--impl reference: the reference's own CUDA kernel rebuilt for sm_100
(oracle/_ref/libref_render.so, unmodified sources) on the same workload; falls
back to the CPU oracle when that library is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, DEPTH, FMT, N_POSES = 1920, 1080, 10, "SH9", 16
BAND_ROWS = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--depth", type=int, default=DEPTH)
    ap.add_argument("--workload", default="1080p", choices=["1080p", "mill19"],
                    help="1080p: BASELINE.json configs[1] (the driver's run); mill19: configs[2], a 3840x2160 frame of a "
                         "multi-GB octree of 8 spatial blocks (2x4 on y,z), depth <= 12")
    ap.add_argument("--mode", default="tiles", choices=["tiles", "split", "hybrid", "guided", "refine"],
                    help="N > 1: image tiles with the tree replicated (default) or one spatial cell per GPU with "
                         "partials composited over NVLink peer stores; guided: the guided-sampling frame (configs[4]) "
                         "with the sub-modules sharded by cell, timed next to the row-block / replicated variant")
    ap.add_argument("--max-nodes", type=int, default=24_000_000, help="node budget of the mill19 tree")
    ap.add_argument("--no-target", action="store_true",
                    help="skip the target_4k sections (BASELINE.json configs[2..4] at 3840x2160 on the Mill-19-scale tree)")
    ap.add_argument("--target-max-nodes", type=int, default=24_000_000, help="node budget of the target_4k tree")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-headless", action="store_true", help="skip the config 4 / 5 sections (C++ driver)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the run (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            # nvidia-smi needs a moment to start streaming: a short timed region must not be over before the first row
            t_wait = time.time() + 3.0
            while not self.rows and time.time() < t_wait:
                time.sleep(0.01)
        except Exception:
            self.proc = None
        self.marks = []

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        allsm = []
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk = float(f[0]); mx = float(f[1])
            except ValueError:
                continue
            allsm.append(clk)
            if t0 <= ts <= t1 + 0.05:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        use = sm if sm else allsm[-3:]
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(stats: dict, pixels: int, trackers: bool) -> int:
    """6 B per empty visit (cell word + sigma), 60 B per shaded
    visit (SH9), + 4 B RGBA8 per pixel (+ 24 B when candidate tracking is on)."""
    empty = stats["visits"] - stats["shaded_visits"]
    return int(empty * 6 + stats["shaded_visits"] * 60 + pixels * (4 + (24 if trackers else 0)))


def ncu_traffic(name="traversal_ncu_summary.json"):
    """dram bytes per launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", name)
    try:
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measured_tensor_peak():
    """(burst, sustained) dense bf16 TFLOP/s of this pool's B200s (driver-written), else the nominal fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return float(j["bf16_tflops"]), float(j["bf16_tflops_sustained"]), "measured"
    except Exception:
        return 2250.0, 2250.0, "fallback (nominal dense bf16)"


def mlp_section(mnv, torch, dev, iters=20):
    """Config 4's dense contraction: one refinement batch = 4096 splits x 8 children x
    8 samples = 262144 rows through the fused tcgen05 MLP (8x256 trunk, appearance embedding, SH9 head),
    random-init weights; CUDA events around each launch, inputs resident; algorithmic FLOPs = 2 * sum in*out."""
    model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3)], device=dev.index)
    rows = 4096 * 8 * 8
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.rand((rows, model.in_dim), device=dev, generator=g) * 2 - 1
    x[:, -1] = 0
    out = torch.empty((rows, model.out_dim + 1), device=dev)  # the reference's [rows, data_dim + 1] result buffer
    for _ in range(3):
        model.forward(x, out=out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.forward(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.mean(ms))
    tf = rows * model.flops_per_row / (t * 1e-3) / 1e12
    burst, sustained, src = measured_tensor_peak()
    sec = {"workload": f"{rows} rows (4096 splits x 8 children x 8 samples) through one 8x256 Mega-NeRF sub-MLP, "
                       "appearance embedding on, SH9 head, random-init weights",
           "rows": rows, "ms_per_launch": t, "mrows_per_s": rows / t / 1e3, "flops_per_row": model.flops_per_row,
           "dtype": "bf16 operands, fp32 accumulate (TMEM)", "gpu_launches": iters,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst,
                        "frac_of_sustained_peak": tf / sustained, "peak_source": src + " cuBLAS bf16, burst (kernel timed alone)",
                        "kernel": "mnv::mlp_forward_kernel", "traffic": ncu_traffic("mlp_ncu_summary.json")}}
    model.close()
    # the reference's evaluation mode for the same contraction: the TorchScript module under fp16 autocast through LibTorch
    # (cuda_renderer.cpp:188-193) — here the PyTorch statement of the named shapes (tests/mlp_reference.py, test infrastructure),
    # eager, same rows; reported beside the fused kernel, never on the product path
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from mlp_reference import MegaNerfMLP
        torch.manual_seed(3)
        ref = MegaNerfMLP().to(dev).eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            for _ in range(3):
                ref(x)
            torch.cuda.synchronize()
            tms = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ref(x)
                e1.record()
                torch.cuda.synchronize()
                tms.append(e0.elapsed_time(e1))
        sec["vs_torch_autocast"] = {"torch_fp16_autocast_ms": float(np.mean(tms)), "speedup": float(np.mean(tms)) / t,
                                    "what": "torch eager fp16 autocast of the same module (the reference's evaluation mode, "
                                            "cuda_renderer.cpp:188-193), same rows"}
    except Exception as e:  # the comparison is informational
        sec["vs_torch_autocast"] = {"error": str(e)[:200]}
    return sec


def point_query_section(mnv, torch, dev, dt, tree, n=4_000_000):
    """Config 1 of BASELINE.json on this tree: query_single_from_root (rt_core.cuh:117-159) for n uniform points —
    the native integer-cell descent on the GPU next to the CPU port on the host cores; results must be equal."""
    from oracle import oracle_py as O

    rng = np.random.default_rng(2)
    xyz = rng.random((n, 3), dtype=np.float32)
    x = torch.from_numpy(xyz).to(dev)
    for _ in range(3):  # warm-up: the output tensor's first allocation is not the kernel
        out = dt.query_points(x)
    torch.cuda.synchronize()
    ms = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = dt.query_points(x)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = min(n, 1_000_000)
    t0 = time.perf_counter()
    cpu = O.query_points(tree, xyz[:m])
    t_cpu = time.perf_counter() - t0
    equal = bool(np.array_equal(out[:m].cpu().numpy(), cpu))
    return {"points": n, "gpu_mqueries_per_s": n / float(np.median(ms)) / 1e3, "cpu_port_mqueries_per_s": m / t_cpu / 1e6,
            "cpu_cores": int(O.lib().oracle_num_threads()), "equal_to_cpu_port": equal}


def headless_sections(mnv, tree, W, H):
    """Configs 4 and 5 of BASELINE.json through the C++ API (viewer::VolumeRenderer via bin/mnv_headless): frames
    with dynamic refinement on (4192-leaf split batches x 8 children x 8 samples through 8 sub-MLPs per frame) and
    with guided sampling (per-sample MLP evaluation, 16-pose orbit so every frame re-samples).  Wall clock per
    frame including the RGBA8 read-back.  Returns {} when the driver binary is missing."""
    import subprocess
    import tempfile

    if not os.path.exists(mnv.HEADLESS_BIN):
        return {}
    out = {}
    with tempfile.TemporaryDirectory() as d:
        tp, mp = os.path.join(d, "tree.npz"), os.path.join(d, "model.npz")
        tree.save_npz(tp)
        subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)]
        mnv.save_model_container(mp, subs, grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))

        def run(*extra):
            r = subprocess.run([mnv.HEADLESS_BIN, tp, "--model", mp, *map(str, extra)], capture_output=True,
                               text=True, timeout=600)
            if r.returncode != 0:
                return {"error": r.stderr[-300:]}
            j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
            return {k: j[k] for k in ("width", "height", "frames", "ms_per_frame_median", "fps_median",
                                      "mrays_per_s_median", "capacity", "guided_rows", "nodes_added")}

        out["refinement"] = dict(run("--width", W, "--height", H, "--frames", 32, "--use_splitting",
                                     "--max_tree_capacity", tree.capacity + 400000),
                                 workload="config 4: frame + vote aggregation + 4192 splits + 268288-row MLP batch over "
                                          "8 sub-modules + commit, every frame")
        gw, gh = W // 2, H // 2
        g = run("--width", gw, "--height", gh, "--frames", 8, "--use_guided_sampling")
        if "guided_rows" in g and g.get("frames"):
            g["mlp_rows_per_frame"] = g["guided_rows"] / g["frames"]
            g["mrows_per_s"] = g["mlp_rows_per_frame"] / g["ms_per_frame_median"] / 1e3
        out["guided_sampling"] = dict(g, workload=f"config 5 on one GPU at {gw}x{gh}: sample emission + per-sample MLP "
                                                  "over 8 sub-modules + per-ray compositing, every frame")
    return out


MILL19_DEPTHS = [12, 11, 11, 12, 12, 11, 11, 12]


def mill19_params(max_nodes):
    """BASELINE.json configs[2]: 8 spatial blocks on a 2x4 (y, z) grid, depth <= 12, ~1.8*10^7 nodes (8-9 GB of
    leaves).  The height field is tilted across the whole z extent so that every block holds surface (round 1's
    flat field left four of the eight cells empty)."""
    return dict(depth=12, data_format=FMT, blocks_yz=(2, 4), block_depths=MILL19_DEPTHS, max_nodes=int(max_nodes),
                tilt=0.75, amp_scale=0.4, fast_data=True)


def shared_tree(params, rank, world, dist, mnv):
    """Multi-GB trees are generated once per box — by rank 0, into /dev/shm, keyed by their parameters — and
    mapped by every rank (and by later bench.py runs on the same box) from there."""
    import hashlib
    import shutil
    key = hashlib.sha1(json.dumps(params, sort_keys=True).encode()).hexdigest()[:12]
    d = f"/dev/shm/mnv_bench_tree_{key}"
    done = os.path.join(d, "meta.json")
    if rank == 0 and not os.path.exists(done):
        for old in [p for p in os.listdir("/dev/shm") if p.startswith("mnv_bench_tree")]:
            shutil.rmtree(os.path.join("/dev/shm", old), ignore_errors=True)
        os.makedirs(d)
        t = mnv.synth.make_tree(**params)
        for k in ("child", "parent", "depth", "data", "scale", "offset"):
            np.save(os.path.join(d, k + ".npy"), getattr(t, k))
        with open(done + ".tmp", "w") as f:
            json.dump({"data_dim": t.data_dim, "data_format": t.data_format, "params": params}, f)
        os.replace(done + ".tmp", done)
        del t
    if dist is not None:
        dist.barrier()
    with open(done) as f:
        meta = json.load(f)
    a = {k: np.load(os.path.join(d, k + ".npy"), mmap_mode="r") for k in ("child", "parent", "depth", "data", "scale", "offset")}
    return mnv.HostTree(N=2, data_dim=meta["data_dim"], data_format=meta["data_format"], child=a["child"],
                        parent=a["parent"], depth=a["depth"], data=a["data"], scale=np.array(a["scale"]),
                        offset=np.array(a["offset"]))


def frame_parity(got: np.ndarray, want: np.ndarray) -> dict:
    """RGBA8 frames -> {bit_exact, max_abs, frac_within_1, psnr} (north_star: <= 1/255 max-abs, >= 50 dB)."""
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    mse = float(np.mean(d.astype(np.float64) ** 2))
    return {"bit_exact": bool(d.max() == 0), "max_abs": int(d.max()), "frac_within_1": float((d <= 1).mean()),
            "psnr": 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))}


def gather_frame(torch, dist, dev, blk, first: int, n_pixels: int, world: int):
    """Every rank contributes the RGBA8 pixels [first, first + len(blk)) it owns; returns the whole frame [n_pixels, 4]
    as numpy (meaningful on every rank; parity checks use rank 0's)."""
    blk = blk.reshape(-1, 4)
    meta = torch.tensor([first, blk.shape[0]], device=dev, dtype=torch.int64)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu().tolist() for m in metas]
    cap = max(m[1] for m in metas)
    buf = torch.zeros((max(cap, 1), 4), dtype=torch.uint8, device=dev)
    buf[: blk.shape[0]] = blk
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    frame = np.zeros((n_pixels, 4), np.uint8)
    for (f, n), b in zip(metas, bufs):
        frame[f:f + n] = b[:n].cpu().numpy()
    return frame


def worst_parity(pars):
    """The worst of several frame_parity records."""
    w = None
    for pr in pars:
        if w is None or pr["max_abs"] > w["max_abs"] or pr["psnr"] < w["psnr"]:
            w = dict(pr, frac_within_1=min(pr["frac_within_1"], (w or pr)["frac_within_1"]), psnr=min(pr["psnr"], (w or pr)["psnr"]),
                     max_abs=max(pr["max_abs"], (w or pr)["max_abs"]), bit_exact=pr["bit_exact"] and (w or pr)["bit_exact"])
    return w


def cpu_baseline(tree, cams, O, opt_kw, seconds_budget=20.0):
    """CPU oracle (port) on all host cores; bounded sample: whole frames of the
    orbit until ~seconds_budget elapsed (at least one)."""
    opt = O.default_options(**opt_kw)
    cores = O.lib().oracle_num_threads()
    t_all, rays, n = 0.0, 0, 0
    O.render_voxels(tree, cams[0], opt, stats=False, row_step=8)  # warm-up (1/8 frame)
    while n < len(cams) and (n == 0 or t_all < seconds_budget):
        t0 = time.perf_counter()
        O.render_voxels(tree, cams[n], opt, stats=False)
        t_all += time.perf_counter() - t0
        rays += cams[n]["width"] * cams[n]["height"]
        n += 1
    return {"value": rays / t_all / 1e6, "unit": "Mrays/s", "cores": int(cores), "kind": "port",
            "sample": f"{n} full {cams[0]['width']}x{cams[0]['height']} frame(s) of the orbit, {t_all:.1f} s"}


def target_sections(args, mnv, torch, dist, rank, world, local_rank):
    """The north-star target case inside the default line: 3840x2160 on the Mill-19-scale octree (BASELINE.json
    configs[2..4]) — image tiles, frames with dynamic refinement ON, and guided sampling — each timed on the device /
    host as its own section, with its own clock record, at the N GPUs of this run.  Rank 0 returns the dict."""
    W, H = 3840, 2160
    P = W * H
    dev = torch.device("cuda", local_rank)
    tree = shared_tree(mill19_params(args.target_max_nodes), rank, world, dist, mnv)
    cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=N_POSES) for i in range(N_POSES)]
    opt_kw = dict(background_brightness=0.0, basis_minmax=[0, 8])
    opt = mnv.default_options(**opt_kw)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"workload": f"Mill-19-scale: {W}x{H}, 8 spatial blocks (2x4 on y,z), depth <= 12, {tree.capacity} nodes, "
                       f"{tree.nbytes() / 1e9:.2f} GB AoS, {N_POSES}-pose orbit", "tree_nodes": int(tree.capacity)}

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxed(ms):
        t = torch.tensor(ms, device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    # ---- (a) image tiles, tree replicated (configs[2]) ----
    dt = mnv.DeviceTree(tree, device=local_rank)
    img = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
    ts = torch.empty((P, 3), device=dev)
    tp = torch.empty((P, 3), device=dev)
    tile = (((W + 15) // 16) * 16, BAND_ROWS, world, rank)

    def launch(i):
        if world == 1:
            dt.render(cams[i % N_POSES], opt, out=img, to_split=ts, to_sample=tp)
        else:
            dt.render_tiles(cams[i % N_POSES], opt, img, *tile, to_split=ts, to_sample=tp)

    for i in range(3):
        launch(i)
    sync()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.time()
    evs = []
    steps = 24
    for i in range(steps):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(i)
        e1.record()
        evs.append((e0, e1))
    sync()
    clocks = sampler.stop(t0, time.time()) if sampler else None
    ms = maxed([a.elapsed_time(b) for a, b in evs])
    par = None
    if world > 1:  # union of the ranks' bands against rank 0's own full frame
        img.zero_()
        launch(5)
        torch.cuda.synchronize()
        dist.all_reduce(img)  # bands are disjoint, the rest is zero
        if rank == 0:
            full = torch.zeros_like(img)
            dt.render(cams[5], opt, out=full)
            torch.cuda.synchronize()
            par = frame_parity(img.cpu().numpy(), full.cpu().numpy())
            del full
    m = float(ms.mean())
    out["tiles"] = {"ms_per_frame": m, "fps": 1e3 / m, "mrays_per_s": P / m / 1e3, "steps": steps, "candidate_tracking": True,
                    "timing": "CUDA events on the launching stream, max over ranks per frame, L2 flushed between frames",
                    "parallelism": "one GPU" if world == 1 else f"interleaved {BAND_ROWS}-row bands over {world} GPUs, tree replicated",
                    "parity_vs_one_gpu": par, "clocks": clocks}
    dt.close()
    del img, ts, tp
    torch.cuda.empty_cache()

    # ---- (b) dynamic refinement ON (configs[3]) and guided sampling (configs[4]) on replicas ----
    subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)]
    n_ref = 16
    ropt = mnv.default_options(use_splitting=True, appearance_embedding=0, split_batch_size=4096, **opt_kw)
    pipe = mnv.multigpu.ReplicatedPipeline(tree, subs, (2, 4), (-1, -1, -1), (1, 1, 1), rank=rank, world=world,
                                           device=local_rank, dist=dist, max_capacity=tree.capacity + (n_ref + 4) * 4096)
    cap0 = tree.capacity
    added = 0
    for i in range(3):
        added += pipe.refine_frame(cams[i % N_POSES], ropt)[1]
    sync()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.time()
    ms = []
    for i in range(n_ref):
        flush.fill_(i & 0xff)
        sync()
        t1 = time.perf_counter()
        _, k = pipe.refine_frame(cams[i % N_POSES], ropt)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t1) * 1e3)
        added += k
    clocks = sampler.stop(t0, time.time()) if sampler else None
    ms = maxed(ms)
    # replicas must hold the same refined tree: checksum of every node added since the start
    import zlib
    crc = 0
    for a in pipe.dt.download(first=cap0):
        crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
    same = True
    if dist is not None:
        c = torch.tensor([crc], device=dev, dtype=torch.int64)
        lst = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(lst, c)
        same = all(int(x.item()) == crc for x in lst)
    m = float(np.median(ms))
    out["refinement"] = {"ms_per_frame_median": m, "fps": 1e3 / m, "mrays_per_s": P / m / 1e3, "steps": n_ref,
                         "nodes_added": int(added), "mlp_rows_per_frame": 4096 * 8 * 8,
                         "refined_nodes_crc32": f"{crc:08x}", "replicas_identical": bool(same),
                         "timing": "host clock around the frame (march with vote tracking -> vote exchange -> selection -> "
                                   "4096 splits x 8 children x 8 samples -> 8 sub-MLPs -> commit), max over ranks, median",
                         "parallelism": "one GPU" if world == 1 else
                         f"interleaved {BAND_ROWS}-row bands over {world} GPUs, tree + sub-MLPs replicated; vote records "
                         "and fp16 payloads all-gathered (NCCL), MLP rows sharded by child",
                         "stage_ms_per_frame": ({k: v / (n_ref + 3) for k, v in pipe.stages.items()} if pipe.stages else None),
                         "exchange_bytes_per_frame": None if world == 1 else
                         {"vote_records_all_gather": getattr(pipe, "vote_bytes", None),
                          "payload_records_all_gather": getattr(pipe, "payload_bytes", None)},
                         "clocks": clocks}
    # guided sampling, row blocks, replicated (no exchange)
    gopt = mnv.default_options(use_guided_sampling=True, appearance_embedding=0, **opt_kw)
    first, n = mnv.multigpu.row_block(H, world, rank)
    cap = int(max(1 << 20, n * W * 14))
    n_g = 4
    rows = 0
    while True:  # sample-row capacity: the densest pose of the orbit decides
        try:
            for i in range(n_g):
                pipe.guided_blocks(cams[i], gopt, capacity_rows=cap)
            break
        except mnv.MnvError as e:
            if e.code != 7 or cap > n * W * 40:
                raise
            cap = int(cap * 1.5)
    sync()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0 = time.time()
    ms = []
    for i in range(n_g):
        flush.fill_(i & 0xff)
        sync()
        t1 = time.perf_counter()
        _, r = pipe.guided_blocks(cams[i % N_POSES], gopt, capacity_rows=cap)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t1) * 1e3)
        rows += r
    clocks = sampler.stop(t0, time.time()) if sampler else None
    ms = maxed(ms)
    rt = torch.tensor([rows], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(rt)
    m = float(np.median(ms))
    rows_frame = float(rt.item()) / n_g
    out["guided_sampling"] = {"ms_per_frame_median": m, "fps": 1e3 / m, "mrays_per_s": P / m / 1e3, "steps": n_g,
                              "mlp_rows_per_frame": rows_frame, "mrows_per_s": rows_frame / m / 1e3,
                              "mlp_tflops": rows_frame * 1210624.0 / (m * 1e-3) / 1e12,
                              "timing": "host clock around the frame (emission -> per-sample MLP over 8 sub-modules -> "
                                        "per-ray compositing) incl. the row-count sync, max over ranks, median",
                              "parallelism": "one GPU" if world == 1 else
                              f"{4 * world} row blocks dealt round-robin over {world} GPUs, tree + sub-MLPs replicated, no exchange",
                              "clocks": clocks}
    pipe.close()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, H = args.width, args.height
    P = W * H

    if args.impl == "reference" and rank != 0:
        return 0  # the reference arm is single-GPU; other ranks exit without work

    import torch
    import mega_nerf_viewer_b200 as mnv

    have_gpu = torch.cuda.is_available()
    if have_gpu:
        torch.cuda.set_device(local_rank)
    dist = None
    if world > 1 and args.impl == "native":
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl" if have_gpu else "gloo")

    if args.workload == "mill19":
        if (args.width, args.height) == (WIDTH, HEIGHT):
            W, H = 3840, 2160
            P = W * H
        tree = shared_tree(mill19_params(args.max_nodes), rank, world, dist, mnv)
        wl = f"Mill-19-scale: headless {W}x{H} frame, 8 spatial blocks (2x4 on y,z), depth <= 12"
    else:
        tree = mnv.synth.make_tree(depth=args.depth, data_format=FMT)
        wl = f"headless {W}x{H} frame, synthetic depth-{args.depth}"
    cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=N_POSES) for i in range(N_POSES)]
    opt_kw = dict(background_brightness=0.0, basis_minmax=[0, 8])  # CLI bg default; set() basis range
    config = {"workload": f"{wl} {FMT} N3Tree "
                          f"({tree.capacity} nodes, {tree.nbytes() / 1e9:.2f} GB AoS), {N_POSES}-pose orbit",
              "resolution": [W, H], "tree_nodes": tree.capacity, "data_format": FMT,
              "options": "RenderOptions defaults, background_brightness=0",
              "l2_policy": "L2 flushed (256 MiB write) between timed frames; camera pose changes every frame"}

    if args.impl == "reference":
        return run_reference(args, tree, cams, opt_kw, config, have_gpu)
    if not have_gpu:
        raise SystemExit("bench.py: no CUDA device — the native path has no CPU fallback")

    dev = torch.device("cuda", local_rank)
    if args.mode == "refine":
        return run_refine(args, mnv, torch, dist, tree, cams, config, W, H, rank, world, local_rank)
    if args.mode == "guided" and world == 1:
        return run_guided_single(args, mnv, torch, tree, local_rank)
    if args.mode == "guided" and world > 1:
        return run_guided(args, mnv, torch, dist, tree, rank, world, local_rank)
    if args.mode in ("split", "hybrid") and world > 1:
        return run_split(args, mnv, torch, dist, tree, cams, opt_kw, config, W, H, rank, world, local_rank)
    dt = mnv.DeviceTree(tree, device=local_rank)
    opt = mnv.default_options(**opt_kw)
    out = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
    ts = torch.empty((P, 3), device=dev)
    tp = torch.empty((P, 3), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
    tile = (((W + 15) // 16) * 16, BAND_ROWS, world, rank)

    def launch(i, trackers=True):
        cam = cams[i % N_POSES]
        if world == 1:
            dt.render(cam, opt, out=out, to_split=ts if trackers else None, to_sample=tp if trackers else None)
        else:
            dt.render_tiles(cam, opt, out, *tile, to_split=ts if trackers else None,
                            to_sample=tp if trackers else None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        evs = []
        for i in range(steps):
            flush.fill_(i & 0xff)  # evict L2 (untimed)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = np.array([a.elapsed_time(b) for a, b in evs], np.float64)
        if dist is not None:
            t = torch.tensor(ms, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # per-step max over ranks
            ms = t.cpu().numpy()
        return ms

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_start = time.time()
    ms = timed(lambda i: launch(i, True), args.steps, args.warmup)
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if sampler else None
    ms_nt = timed(lambda i: launch(i, False), max(args.steps // 4, 3), 3)

    # e2e: the host-buffer C-ABI call, wall clock around K synchronous frames.  Like `value` (and like the
    # reference arm, which fills and writes both candidate trackers every frame as cuda_renderer.cpp:97-98,141
    # does) the frame is marched WITH candidate tracking (opt.use_splitting routes the tree-owned trackers into
    # the launch); the tracker-less call is timed next to it as e2e_no_trackers.
    opt_track = mnv.default_options(use_splitting=True, **opt_kw)

    def e2e_run(o):
        def frame(i):
            cam = cams[i % N_POSES]
            if world == 1:
                dt.render_frame_host(cam, o, host)
            else:
                dt.render_frame_host(cam, o, host, bands=(BAND_ROWS, world, rank))

        for i in range(max(args.warmup, 3)):
            frame(i)
        barrier()
        out_ms = []
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            frame(i)
            out_ms.append((time.perf_counter() - t0) * 1e3)
        barrier()
        out_ms = np.array(out_ms)
        if dist is not None:
            t = torch.tensor(out_ms, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out_ms = t.cpu().numpy()
        return out_ms

    e2e_ms = e2e_run(opt_track)
    e2e_nt_ms = e2e_run(opt)

    # multi-GPU parity, verified by the run that is timed: the union of the ranks' bands against the frame rank 0
    # renders alone (bit-exact by construction: rays are independent)
    parity = None
    if dist is not None:
        worst = None
        for pose in (0, 7):
            out.zero_()
            dt.render_tiles(cams[pose], opt, out, *tile)
            torch.cuda.synchronize()
            dist.all_reduce(out)  # bands are disjoint, the rest is zero
            if rank == 0:
                full = torch.zeros_like(out)
                dt.render(cams[pose], opt, out=full)
                torch.cuda.synchronize()
                pr = frame_parity(out.cpu().numpy(), full.cpu().numpy())
                if worst is None or pr["max_abs"] > worst["max_abs"]:
                    worst = pr
        parity = worst and dict(worst, against="the whole frame rendered by rank 0 alone, poses 0 and 7")

    target = None
    if not args.no_target and args.workload == "1080p" and args.mode == "tiles":
        dt_keep = dt
        target = target_sections(args, mnv, torch, dist, rank, world, local_rank)
        dt = dt_keep

    if rank != 0:
        dist.destroy_process_group()
        return 0

    # algorithmic bytes: visit statistics of the timed frames (whole frame, all ranks' tiles)
    alg = []
    for i in range(min(args.steps, N_POSES)):
        _, st = dt.render_frame_host(cams[i % N_POSES], opt, stats=True)
        alg.append((algorithmic_bytes(st, P, True), st))
    alg_bytes = float(np.mean([a for a, _ in alg]))
    visits = float(np.mean([s["visits"] for _, s in alg]))
    shaded = float(np.mean([s["shaded_visits"] for _, s in alg]))
    ms_step = float(ms.mean())
    peak, peak_src = measured_peak()
    achieved = alg_bytes / world / (ms_step * 1e-3) / 1e9  # per-GPU kernel
    line = {
        "metric": "Mrays/s", "value": P / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (fp16 storage, fp64 ray setup)", "data": "synthetic", "config": config,
        "fps": 1e3 / ms_step,
        "value_no_trackers": P / (float(ms_nt.mean()) * 1e-3) / 1e6,
        "e2e": {"value": P / (float(e2e_ms.mean()) * 1e-3) / 1e6, "unit": "Mrays/s",
                "ms_per_step": float(e2e_ms.mean()),
                "h2d_bytes_per_step": 72 + 104,  # mnv_camera + mnv_render_options (kernel params)
                "d2h_bytes_per_step": P * 4 // world,
                "api": "mnv_render_frame_host (C-ABI), candidate tracking on (use_splitting) like the reference arm"},
        "e2e_no_trackers": {"value": P / (float(e2e_nt_ms.mean()) * 1e-3) / 1e6, "unit": "Mrays/s",
                            "ms_per_step": float(e2e_nt_ms.mean())},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                     "kernel": "mnv::render_voxels_kernel<9,track>",
                     "algorithmic_bytes_per_launch": alg_bytes / world,
                     "leaf_visits_per_frame": visits, "shaded_visits_per_frame": shaded,
                     "gvisits_per_s": visits / (ms_step * 1e-3) / 1e9},
        "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    if target is not None:
        line["target_4k"] = target
    line["mlp"] = mlp_section(mnv, torch, dev)
    if world == 1 and not args.no_cpu_baseline:
        line["point_query"] = point_query_section(mnv, torch, dev, dt, tree)
    if world == 1 and not args.no_headless:
        dt.close()
        del flush, ts, tp, out
        torch.cuda.empty_cache()
        line.update(headless_sections(mnv, tree, W, H))
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_py as O
        line["cpu_baseline"] = cpu_baseline(tree, cams, O, opt_kw)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_split(args, mnv, torch, dist, tree, cams, opt_kw, config, W, H, rank, world, local_rank):
    """N > 1, --mode split: one spatial cell (restricted subtree) per GPU, every GPU marches every ray through
    its cell, partials land in the pixel owner's memory as NVLink peer stores, device-side flags, front-to-back
    composite of each owner's P / N pixels (csrc/mnv_multigpu.cu).  Step = march + signal + composite."""
    P = W * H
    dev = torch.device("cuda", local_rank)
    hybrid = args.mode == "hybrid" and world >= 4
    if hybrid:  # row blocks x spatial cells (4 cells per group; 2 on 4 GPUs)
        cells = 4 if world >= 8 else 2
        sp = mnv.multigpu.HybridSplit(tree, W, H, rank, world, local_rank, dist, cells=cells)
        first, n = sp.pixel_range()
    else:
        sp = mnv.multigpu.SubmoduleSplit(tree, W, H, rank=rank, world=world, device=local_rank, dist=dist)
        first, n = mnv.multigpu.owner_range(P, world, rank)
    opt = mnv.default_options(**opt_kw)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    host = torch.empty((max(n, 1), 4), dtype=torch.uint8).pin_memory()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(steps, warmup, e2e):
        for i in range(warmup):
            sp.render_block(cams[i % N_POSES], opt)
        barrier()
        ms = []
        for i in range(steps):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            if e2e:
                t0 = time.perf_counter()
                blk = sp.render_block(cams[i % N_POSES], opt)
                host[:n].copy_(blk, non_blocking=True)
                torch.cuda.synchronize()
                ms.append((time.perf_counter() - t0) * 1e3)
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sp.render_block(cams[i % N_POSES], opt)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
        barrier()
        t = torch.tensor(ms, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_start = time.time()
    ms = timed(args.steps, args.warmup, False)
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if sampler else None
    e2e_ms = timed(args.steps, 3, True)
    # parity, verified in the run that is timed: the sharded frame (every owner's block) against the frame rank 0 renders
    # alone from the WHOLE tree.  Tolerance parity by construction: early termination acts per segment.
    full_dt = mnv.DeviceTree(tree, device=local_rank) if rank == 0 else None
    pars = []
    for pose in (0, 7):
        blk = sp.render_block(cams[pose], opt)
        torch.cuda.synchronize()
        frame = gather_frame(torch, dist, dev, blk[:n], first, P, world)
        if rank == 0:
            want = full_dt.render(cams[pose], opt).cpu().numpy().reshape(P, 4)
            pars.append(frame_parity(frame, want))
    parity = worst_parity(pars) if rank == 0 else None
    if full_dt is not None:
        full_dt.close()
    nodes = torch.tensor([sp.split.local_nodes if hybrid else sp.local_nodes], device=dev)
    dist.all_reduce(nodes, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = float(ms.mean())
        desc = (f"hybrid: {sp.groups} row blocks x {sp.split.world} spatial cells (grid {sp.split.grid_dim} on (y,z))" if hybrid
                else f"sub-module split: {world} spatial cells, grid {sp.grid_dim} on (y,z)")
        config = dict(config, parallelism=f"{desc}; largest per-GPU subtree {int(nodes.item())} nodes of {tree.capacity}")
        line = {"metric": "Mrays/s", "value": P / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 storage, fp64 ray setup)",
                "data": "synthetic", "config": config, "fps": 1e3 / ms_step,
                "e2e": {"value": P / (float(e2e_ms.mean()) * 1e-3) / 1e6, "unit": "Mrays/s",
                        "ms_per_step": float(e2e_ms.mean()), "h2d_bytes_per_step": 176,
                        "d2h_bytes_per_step": int(n) * 4, "api": "multigpu.SubmoduleSplit.render_block + D2H of the owner's block"},
                "gpu_launches": args.steps * 3,
                "exchange": {"bytes_per_gpu_per_step": (P // sp.groups if hybrid else P) * 16, "transport": "NVLink peer stores from the march kernel "
                             "(CUDA IPC mappings), no collective"},
                "parity": dict(parity, against="the unsharded frame rendered by rank 0 alone from the whole tree, poses 0 and 7; "
                                               "tolerance parity by construction (early termination acts per segment)"),
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    sp.close()
    dist.destroy_process_group()
    return 0


def run_refine(args, mnv, torch, dist, tree, cams, config, W, H, rank, world, local_rank):
    """--mode refine (any N): frames with dynamic refinement ON (BASELINE.json configs[3]; with --workload mill19 the
    north-star target case: 3840x2160 on the multi-GB octree).  Per frame: march with vote tracking (row block of
    this rank) -> NCCL all-gather of the votes when N > 1 -> candidate selection -> 4096 splits x 8 children x 8
    samples -> 8 sub-MLPs -> commit into the tree, on every replica (multigpu.ReplicatedPipeline).  Host clock
    around the frame, max over ranks, median over the timed frames; the tree grows by 4096 nodes x 8 per frame."""
    P = W * H
    dev = torch.device("cuda", local_rank)
    subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)]
    steps = min(args.steps, 48)
    ropt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_splitting=True, appearance_embedding=0,
                               split_batch_size=4096)
    pipe = mnv.multigpu.ReplicatedPipeline(tree, subs, (2, 4), (-1, -1, -1), (1, 1, 1), rank=rank, world=world,
                                           device=local_rank, dist=dist, max_capacity=tree.capacity + (steps + 4) * 4096)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    added = 0
    for i in range(3):
        added += pipe.refine_frame(cams[i % N_POSES], ropt)[1]
    sync()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_start = time.time()
    ms = []
    for i in range(steps):
        flush.fill_(i & 0xff)
        sync()
        t0 = time.perf_counter()
        _, k = pipe.refine_frame(cams[i % N_POSES], ropt)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
        added += k
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if sampler else None
    t = torch.tensor(ms, device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.cpu().numpy()
    if rank == 0:
        m = float(np.median(ms))
        line = {"metric": "Mrays/s", "value": P / (m * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": steps,
                "warmup": 3, "ms_per_step": m, "ms_per_step_mean": float(ms.mean()), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 march (fp16 storage), bf16 MLP operands with f32 accumulate", "data": "synthetic",
                "config": dict(config, refinement="ON: 4096 leaf splits x 8 children x 8 samples = 262144 MLP rows over 8 "
                                                  "sub-modules + commit, every frame",
                               parallelism="one GPU" if world == 1 else
                               f"row blocks over {world} GPUs, tree + sub-MLPs replicated, votes all-gathered (NCCL)"),
                "fps": 1e3 / m, "nodes_added": int(added), "capacity_end": int(pipe.dt.capacity),
                "gpu_launches": steps * 12, "clocks": clocks}
        print(json.dumps(line), flush=True)
    pipe.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_guided_single(args, mnv, torch, tree, local_rank):
    """--mode guided on ONE GPU: the guided-sampling frame (emission + per-sample MLP over 8 sub-modules on a 2x4 (y,z)
    grid + per-ray compositing) on the workload's tree, in-process.  960x540 unless --width/--height are given."""
    W, H = (960, 540) if (args.width, args.height) == (WIDTH, HEIGHT) else (args.width, args.height)
    P = W * H
    dev = torch.device("cuda", local_rank)
    subs = [mnv.synth.make_mlp_weights(seed=11 + i) for i in range(8)]
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True,
                               appearance_embedding=0)
    cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=N_POSES) for i in range(N_POSES)]
    rp = mnv.multigpu.ReplicatedPipeline(tree, subs, (2, 4), (-1, -1, -1), (1, 1, 1), device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps = min(args.steps, 32)
    cap = P * 12
    for i in range(2):
        rp.guided_block(cams[i], gopt, capacity_rows=cap)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    t_start = time.time()
    ms, rows = [], 0
    for i in range(steps):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, r = rp.guided_block(cams[i % N_POSES], gopt, capacity_rows=cap)
        torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
        rows += r
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end)
    m = float(np.median(ms))
    line = {"metric": "Mrays/s", "value": P / (m * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": 1, "steps": steps, "warmup": 2,
            "ms_per_step": m, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 MLP operands, f32 accumulate / compositing", "data": "synthetic",
            "config": {"workload": f"guided-sampling frame {W}x{H}, {args.workload} tree ({tree.capacity} nodes), 8 sub-modules "
                                   "on a 2x4 (y,z) grid, one GPU",
                       "timing": "host clock around the frame incl. the row-count sync, median of steps; L2 flushed between frames"},
            "fps": 1e3 / m, "mlp_rows_per_frame": rows / steps, "mrows_per_s": rows / steps / m / 1e3,
            "gpu_launches": steps * 12, "clocks": clocks}
    print(json.dumps(line), flush=True)
    rp.close()
    return 0


def run_guided(args, mnv, torch, dist, tree, rank, world, local_rank):
    """N > 1, --mode guided: BASELINE.json configs[4] — a guided-sampling frame (sample emission + per-sample MLP +
    compositing) at 960x540 unless --width/--height say otherwise, two ways on the same tree and weights:
    'sharded'  one spatial cell + its sub-MLP per GPU, every GPU handles its segment of every ray, one all-gather
               of 16 B per ray and one 16 B peer store per ray (multigpu.ShardedGuided);
    'rows'     tree and all sub-MLPs replicated, each GPU does the rays of its row block, no exchange
               (multigpu.ReplicatedPipeline)."""
    W, H = (960, 540) if (args.width, args.height) == (WIDTH, HEIGHT) else (args.width, args.height)
    P = W * H
    dev = torch.device("cuda", local_rank)
    grid = mnv.synth.grid_for_world(world)
    subs = [mnv.synth.make_mlp_weights(seed=11 + i) for i in range(world)]
    mn, mx = (-1, -1, -1), (1, 1, 1)
    gopt = mnv.default_options(background_brightness=0.0, basis_minmax=[0, 8], use_guided_sampling=True,
                               appearance_embedding=0)
    cams = [mnv.synth.default_camera(W, H, pose=i, n_poses=N_POSES) for i in range(N_POSES)]
    sh = mnv.multigpu.ShardedGuided(tree, subs, grid, mn, mx, W, H, rank=rank, world=world, device=local_rank, dist=dist)
    rp = mnv.multigpu.ReplicatedPipeline(tree, subs, grid, mn, mx, rank=rank, world=world, device=local_rank, dist=dist)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps = min(args.steps, 32)

    def timed(fn, cap):
        rows = 0
        for i in range(3):
            fn(cams[i % N_POSES], gopt, capacity_rows=cap)
        dist.barrier()
        torch.cuda.synchronize()
        ms = []
        for i in range(steps):
            flush.fill_(i & 0xff)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, r = fn(cams[i % N_POSES], gopt, capacity_rows=cap)
            torch.cuda.synchronize()
            ms.append((time.perf_counter() - t0) * 1e3)
            rows += r
        t = torch.tensor(ms, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rt = torch.tensor([rows], device=dev, dtype=torch.float64)
        rmax = rt.clone()
        dist.all_reduce(rt)
        dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
        return t.cpu().numpy(), float(rt.item()) / steps, float(rmax.item()) / steps

    cap = P * 12  # rows: generous for either partition (a ray averages ~10 samples on this tree)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_start = time.time()
    ms_sh, rows_sh, rows_sh_max = timed(sh.guided_block, cap)
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if sampler else None
    ms_rp, rows_rp, rows_rp_max = timed(rp.guided_block, cap)
    # parity in the run that is timed: the sharded frame against the row-block frame of the replicated pipeline (which the
    # tests hold bit-identical to one GPU)
    pars = []
    for pose in (0, 7):
        blk, _ = sh.guided_block(cams[pose], gopt, capacity_rows=cap)
        torch.cuda.synchronize()
        f_sh, n_sh = mnv.multigpu.owner_range(P, world, rank)
        got = gather_frame(torch, dist, dev, blk[:n_sh], f_sh, P, world)
        rows_img, _ = rp.guided_block(cams[pose], gopt, capacity_rows=cap)
        torch.cuda.synchronize()
        r0, _nr = mnv.multigpu.row_block(H, world, rank)
        want = gather_frame(torch, dist, dev, rows_img, r0 * W, P, world)
        pars.append(frame_parity(got, want))
    parity = worst_parity(pars)
    nodes = torch.tensor([sh.local_nodes], device=dev)
    dist.all_reduce(nodes, op=dist.ReduceOp.MAX)
    if rank == 0:
        m = float(np.median(ms_sh))
        line = {"metric": "Mrays/s", "value": P / (m * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": steps,
                "warmup": 3, "ms_per_step": m, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "bf16 MLP operands, f32 accumulate / compositing", "data": "synthetic",
                "config": {"workload": f"guided-sampling frame {W}x{H}, synthetic depth-{args.depth} SH9 N3Tree "
                                       f"({tree.capacity} nodes), {world} sub-modules sharded by (y,z) cell {grid}",
                           "parallelism": f"one cell + sub-MLP per GPU; largest subtree {int(nodes.item())} nodes",
                           "timing": "host clock around the frame incl. the row-count sync, max over ranks, median of steps; "
                                     "L2 flushed between frames"},
                "fps": 1e3 / m, "mlp_rows_per_frame": rows_sh, "mlp_rows_busiest_gpu": rows_sh_max,
                "exchange": {"all_gather_bytes_per_gpu": P * 16, "peer_store_bytes_per_gpu": P * 16,
                             "transport": "NCCL all-gather of probe records + NVLink peer stores of segment partials"},
                "rows_replicated": {"ms_per_step": float(np.median(ms_rp)), "fps": 1e3 / float(np.median(ms_rp)),
                                    "mlp_rows_per_frame": rows_rp, "mlp_rows_busiest_gpu": rows_rp_max,
                                    "parallelism": "row blocks, tree + all sub-MLPs replicated, no exchange"},
                "parity": dict(parity, against="the same frame from the replicated row-block pipeline (bit-identical to one GPU), "
                                               "poses 0 and 7; tolerance parity by construction (transmittance carried per segment)"),
                "gpu_launches": steps * 7, "clocks": clocks}
        print(json.dumps(line), flush=True)
    sh.close()
    rp.close()
    dist.destroy_process_group()
    return 0


def run_reference(args, tree, cams, opt_kw, config, have_gpu):
    """The reference arm: its own CUDA kernel (unmodified sources rebuilt for
    sm_100, oracle/_ref) when available, else the CPU port of its device code."""
    from oracle import oracle_py as O
    W, H = args.width, args.height
    P = W * H
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(tree, cams, O, opt_kw, seconds_budget=10.0)
    if have_gpu and O.ref_available():
        import torch
        npz = "/tmp/mnv_bench_tree.npz"
        tree.save_npz(npz)
        sys.stdout.flush()
        saved = os.dup(1)  # the reference's loader prints to stdout: keep this arm's stdout to the one JSON line
        os.dup2(2, 1)
        try:
            ref = O.RefRenderer(npz)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        opt = O.default_options(**opt_kw)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        sampler = ClockSampler(0)
        t_start = time.time()
        for i in range(args.warmup):
            ref.render(cams[i % N_POSES], opt, iters=1, trackers=False)
        ms = []
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            ms.append(float(ref.render(cams[i % N_POSES], opt, iters=1, trackers=False)["ms"][0]))
        t_end = time.time()
        clocks = sampler.stop(t_start, t_end)
        host = np.empty((H, W, 4), np.uint8)
        for i in range(3):
            ref.render_frame_host(cams[i % N_POSES], opt, host)
        e2e = []
        for i in range(args.steps):
            flush.fill_(i & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref.render_frame_host(cams[i % N_POSES], opt, host)
            e2e.append((time.perf_counter() - t0) * 1e3)
        ms_step = float(np.mean(ms))
        line = {"impl": "reference", "reference_kind": "reference CUDA kernel (render_voxels_kernel, "
                "unmodified sources rebuilt for sm_100) on the same GPU",
                "metric": "Mrays/s", "value": P / (ms_step * 1e-3) / 1e6, "unit": "Mrays/s",
                "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (fp16 storage, fp64 ray setup)", "data": "synthetic", "config": config,
                "fps": 1e3 / ms_step,
                "e2e": {"value": P / (float(np.mean(e2e)) * 1e-3) / 1e6, "unit": "Mrays/s",
                        "ms_per_step": float(np.mean(e2e)), "h2d_bytes_per_step": 48,
                        "d2h_bytes_per_step": P * 4,
                        "api": "Camera::_update + render_voxels + cudaMemcpy2DFromArray"},
                "gpu_launches": args.steps, "clocks": clocks}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
        return 0
    if cpu is None:
        cpu = cpu_baseline(tree, cams, O, opt_kw, seconds_budget=10.0)
    line = {"impl": "reference", "reference_kind": "CPU port of the reference's device code "
            "(the reference CUDA kernel needs a GPU and oracle/_ref)",
            "metric": "Mrays/s", "value": cpu["value"], "unit": "Mrays/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": P / cpu["value"] / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
