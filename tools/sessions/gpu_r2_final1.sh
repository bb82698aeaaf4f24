#!/usr/bin/env bash
# Round-2 final measurements on one GPU: smoke, default bench line, reference arm, launch list of the bench command.
set -x
mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2> gpurun_out/r2_time_n1.txt
tail -c 1500 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_time_n1.txt
( time python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2> gpurun_out/r2_time_ref.txt
tail -c 800 gpurun_out/r2_bench_ref.json; cat gpurun_out/r2_time_ref.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
MNV_TIMING=1 python - <<'PY' 2>&1 | tail -12 | tee gpurun_out/r2_headless_stages.log
import mega_nerf_viewer_b200 as mnv, subprocess, json, tempfile, os
tree = mnv.synth.make_tree(depth=10)
d = tempfile.mkdtemp(); p = os.path.join(d, "t.npz"); tree.save_npz(p); mp = os.path.join(d, "m.npz")
mnv.save_model_container(mp, [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)], grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))
r = subprocess.run([mnv.HEADLESS_BIN, p, "--model", mp, "--width", "1920", "--height", "1080", "--frames", "32", "--use_splitting", "--max_tree_capacity", str(tree.capacity + 400000)], capture_output=True, text=True)
print(r.stdout.splitlines()[-1][:400]); print(r.stderr[-1200:])
PY
