#!/usr/bin/env bash
# Dev: the C++ VolumeRenderer through mnv_headless on the bench tree (config 2 / 4 / 5 shapes).
set -x
python - <<'PY'
import sys; sys.path.insert(0, '.')
import mega_nerf_viewer_b200 as mnv
t = mnv.synth.make_tree(depth=10); t.save_npz('/tmp/t10.npz')
subs = [mnv.synth.make_mlp_weights(seed=3 + i) for i in range(8)]
mnv.save_model_container('/tmp/m8.npz', subs, grid_dim=(2, 4), min_position=(-1, -1, -1), max_position=(1, 1, 1))
print("nodes", t.capacity)
PY
B=mega-nerf-viewer_b200/bin/mnv_headless
#
MNV_TIMING=1 $B /tmp/t10.npz --frames 32 --model /tmp/m8.npz --use_splitting --max_tree_capacity 4000000 2>&1 | tail -10
MNV_TIMING=1 $B /tmp/t10.npz --frames 8 --model /tmp/m8.npz --use_guided_sampling --width 960 --height 540 2>&1 | tail -8
