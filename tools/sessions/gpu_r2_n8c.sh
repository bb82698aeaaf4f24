#!/usr/bin/env bash
# C++ replica group with refinement at 4K on 4 and 8 GPUs after moving the whole exchange onto one host thread per replica.
set -x
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/r2_headless_group_n8c.log
import mega_nerf_viewer_b200 as mnv, subprocess, json, tempfile, os
tree = mnv.synth.make_tree(depth=10)
d = tempfile.mkdtemp(); p = os.path.join(d, "t.npz"); tree.save_npz(p); mp = os.path.join(d, "m.npz")
mnv.save_model_container(mp, [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)], grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))
for extra in (["--model", mp, "--use_splitting"],):
    for g in (1, 4, 8):
        r = subprocess.run([mnv.HEADLESS_BIN, p, "--width", "3840", "--height", "2160", "--frames", "24", "--gpus", str(g)] + extra, capture_output=True, text=True)
        j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1]) if r.returncode == 0 else r.stderr[-400:]
        print("mnv_headless 3840x2160 depth-10 tree, gpus", g, extra[-1:], {k: j[k] for k in ("ms_per_frame_median", "fps_median", "frame_hash", "nodes_added", "capacity")} if isinstance(j, dict) else j)
PY
