#!/usr/bin/env bash
# Round-2 GPU session F (2 GPUs): multi-process tests, default bench line at N=2 (parity + target_4k), C++ group on 2 devices.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py tests/test_viewer_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2f_pytest.log | cut -c1-200
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err ) 2> gpurun_out/r2f_time_n2.txt
tail -c 5000 gpurun_out/r2f_bench_n2.json; tail -8 gpurun_out/r2f_bench_n2.err; cat gpurun_out/r2f_time_n2.txt
python - <<'PY'
import mega_nerf_viewer_b200 as mnv, numpy as np, subprocess, json, tempfile, os
tree = mnv.synth.make_tree(depth=9)
d = tempfile.mkdtemp(); p = os.path.join(d, "t.npz"); tree.save_npz(p)
for g in (1, 2):
    r = subprocess.run([mnv.HEADLESS_BIN, p, "--width", "3840", "--height", "2160", "--frames", "16", "--gpus", str(g)], capture_output=True, text=True)
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]) if r.returncode == 0 else r.stderr[-400:]
    print("headless gpus", g, {k: j[k] for k in ("ms_per_frame_median", "fps_median", "frame_hash")} if isinstance(j, dict) else j)
PY
