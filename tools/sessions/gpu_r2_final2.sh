#!/usr/bin/env bash
# Round-2 final measurements on two GPUs.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_2gpu.log | cut -c1-200
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err ) 2> gpurun_out/r2_time_n2.txt
tail -2 gpurun_out/r2_bench_n2.err | cut -c1-300; cat gpurun_out/r2_time_n2.txt | grep real
MNV_STAGE_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 > gpurun_out/r2_bench_n2_stages.json 2> gpurun_out/r2_bench_n2_stages.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --mode split --workload mill19 --steps 24 > gpurun_out/r2_bench_split_mill19_n2.json 2> gpurun_out/r2_bench_split_mill19_n2.err
tail -c 1500 gpurun_out/r2_bench_split_mill19_n2.json; tail -3 gpurun_out/r2_bench_split_mill19_n2.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --mode guided --workload mill19 --width 1920 --height 1080 --steps 8 > gpurun_out/r2_bench_guided_mill19_n2.json 2> gpurun_out/r2_bench_guided_mill19_n2.err
tail -c 1500 gpurun_out/r2_bench_guided_mill19_n2.json; tail -3 gpurun_out/r2_bench_guided_mill19_n2.err | cut -c1-300
python - <<'PY' 2>&1 | tee gpurun_out/r2_headless_group.log
import mega_nerf_viewer_b200 as mnv, subprocess, json, tempfile, os
tree = mnv.synth.make_tree(depth=10)
d = tempfile.mkdtemp(); p = os.path.join(d, "t.npz"); tree.save_npz(p); mp = os.path.join(d, "m.npz")
mnv.save_model_container(mp, [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)], grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))
for extra in ([], ["--model", mp, "--use_splitting"]):
    for g in (1, 2):
        r = subprocess.run([mnv.HEADLESS_BIN, p, "--width", "3840", "--height", "2160", "--frames", "24", "--gpus", str(g)] + extra, capture_output=True, text=True)
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]) if r.returncode == 0 else r.stderr[-400:]
        print("mnv_headless 3840x2160 depth-10 tree, gpus", g, extra[-1:], {k: j[k] for k in ("ms_per_frame_median", "fps_median", "frame_hash", "nodes_added", "capacity")} if isinstance(j, dict) else j)
PY
