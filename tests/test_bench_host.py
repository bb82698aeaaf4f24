"""CPU tests of bench.py's host logic and of the multi-process (world_size 2,
gloo) band partition used for the N > 1 path."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_definition():
    sys.path.insert(0, ROOT)
    import bench

    st = dict(visits=100, shaded_visits=10)
    # 90 empty * 6 B + 10 shaded * 60 B + pixels * (4 [+24])
    assert bench.algorithmic_bytes(st, 50, False) == 90 * 6 + 10 * 60 + 50 * 4
    assert bench.algorithmic_bytes(st, 50, True) == 90 * 6 + 10 * 60 + 50 * 28


def test_reference_arm_runs_on_cpu_and_prints_contract_line():
    """Without a GPU the reference arm times the CPU port on a bounded sample."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--depth", "6", "--width", "320", "--height", "180"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}


WORKER = r"""
import os, sys, numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
from oracle import oracle_py as O
import mega_nerf_viewer_b200 as mnv
tree = mnv.synth.make_tree(depth=5)
cam = mnv.synth.default_camera(96, 70, pose=3)
opt = O.default_options(background_brightness=0.0)
H, W, band = cam["height"], cam["width"], 8
# every rank renders only its interleaved 8-row bands (same rule as mnv_render_frame_host_bands)
img = np.zeros((H, W, 4), np.uint8)
for b in range(rank, (H + band - 1) // band, world):
    y0, y1 = b * band, min(H, (b + 1) * band)
    part = O.render_voxels(tree, cam, opt, y0=y0, y1=y1, stats=False)["rgba"]
    img[y0:y1] = part[y0:y1]
t = torch.from_numpy(img.astype(np.int32))
dist.all_reduce(t)                       # bands are disjoint: the sum is the union
ms = torch.tensor([1.0 + rank]); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    full = O.render_voxels(tree, cam, opt, stats=False)["rgba"]
    assert np.array_equal(t.numpy().astype(np.uint8), full), "band union != whole frame"
    assert ms.item() == float(world)
    print("OK")
dist.destroy_process_group()
"""


def test_band_partition_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout


def test_frame_parity_and_worst_of():
    sys.path.insert(0, ROOT)
    import bench

    a = np.zeros((4, 4), np.uint8)
    b = a.copy()
    p0 = bench.frame_parity(a, b)
    assert p0 == {"bit_exact": True, "max_abs": 0, "frac_within_1": 1.0, "psnr": 99.0}
    b[0, 0] = 3
    b[1, 1] = 1
    p1 = bench.frame_parity(a, b)
    assert p1["max_abs"] == 3 and not p1["bit_exact"] and p1["frac_within_1"] == 15 / 16
    assert abs(p1["psnr"] - 10 * np.log10(255.0 ** 2 / (10 / 16))) < 1e-9
    w = bench.worst_parity([p0, p1, p0])
    assert w["max_abs"] == 3 and not w["bit_exact"] and w["frac_within_1"] == 15 / 16 and w["psnr"] == p1["psnr"]


_GATHER_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
P = 37  # ragged: the ranks own 19 and 18 pixels
first = 0 if rank == 0 else 19
n = 19 if rank == 0 else 18
blk = (torch.arange(first, first + n, dtype=torch.int64)[:, None] * 4 + torch.arange(4)[None]).to(torch.uint8)
frame = bench.gather_frame(torch, dist, torch.device("cpu"), blk, first, P, world)
want = (np.arange(P)[:, None] * 4 + np.arange(4)[None]).astype(np.uint8)
assert np.array_equal(frame, want), (rank, frame[:3], want[:3])
dist.destroy_process_group()
print("OK", rank)
"""


def test_gather_frame_two_ranks_gloo(tmp_path):
    """bench.gather_frame (the in-run parity of --mode split / guided): ragged owner blocks from 2 ranks -> one frame."""
    script = tmp_path / "gather_worker.py"
    script.write_text(_GATHER_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29571", str(script), ROOT], capture_output=True, text=True, timeout=300,
                       cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert r.stdout.count("OK") == 2
