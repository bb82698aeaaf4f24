#!/usr/bin/env bash
# Round-2 GPU session G (2 GPUs): balanced refinement / guided partitions at N=2, stage timing, C++ group on 2 devices.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_viewer_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r2g_pytest.log | cut -c1-200
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 60 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err ) 2> gpurun_out/r2g_time_n2.txt
tail -3 gpurun_out/r2g_bench_n2.err | cut -c1-300; cat gpurun_out/r2g_time_n2.txt
MNV_STAGE_TIMING=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 > gpurun_out/r2g_bench_n2_stages.json 2> gpurun_out/r2g_bench_n2_stages.err
MNV_STAGE_TIMING=1 timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-headless > gpurun_out/r2g_bench_n1_stages.json 2> gpurun_out/r2g_bench_n1_stages.err
python - <<'PY'
import mega_nerf_viewer_b200 as mnv, numpy as np, subprocess, json, tempfile, os
tree = mnv.synth.make_tree(depth=9)
d = tempfile.mkdtemp(); p = os.path.join(d, "t.npz"); tree.save_npz(p)
mp = os.path.join(d, "m.npz")
mnv.save_model_container(mp, [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)], grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))
for extra in ([], ["--model", mp, "--use_splitting"]):
    for g in (1, 2):
        env = dict(os.environ, MNV_TIMING="1") if (g == 1 and extra) else os.environ
        r = subprocess.run([mnv.HEADLESS_BIN, p, "--width", "3840", "--height", "2160", "--frames", "16", "--gpus", str(g)] + extra, capture_output=True, text=True, env=env)
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]) if r.returncode == 0 else r.stderr[-400:]
        print("headless gpus", g, extra[-1:] , {k: j[k] for k in ("ms_per_frame_median", "fps_median", "frame_hash", "nodes_added")} if isinstance(j, dict) else j)
        if g == 1 and extra: print(r.stderr[-1500:])
PY
