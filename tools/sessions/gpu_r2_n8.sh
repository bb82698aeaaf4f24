#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err ) 2> gpurun_out/r2_time_n8.txt
tail -c 3000 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err | cut -c1-300; grep real gpurun_out/r2_time_n8.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err ) 2> gpurun_out/r2_time_n4.txt
tail -c 600 gpurun_out/r2_bench_n4.json; grep real gpurun_out/r2_time_n4.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2_headless_group_n8.log
import mega_nerf_viewer_b200 as mnv, subprocess, json, tempfile, os
tree = mnv.synth.make_tree(depth=10)
d = tempfile.mkdtemp(); p = os.path.join(d, "t.npz"); tree.save_npz(p); mp = os.path.join(d, "m.npz")
mnv.save_model_container(mp, [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)], grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))
for extra in ([], ["--model", mp, "--use_splitting"]):
    for g in (1, 4, 8):
        r = subprocess.run([mnv.HEADLESS_BIN, p, "--width", "3840", "--height", "2160", "--frames", "24", "--gpus", str(g)] + extra, capture_output=True, text=True)
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]) if r.returncode == 0 else r.stderr[-400:]
        print("mnv_headless 3840x2160 depth-10 tree, gpus", g, extra[-1:], {k: j[k] for k in ("ms_per_frame_median", "fps_median", "frame_hash", "nodes_added", "capacity")} if isinstance(j, dict) else j)
PY
