#!/usr/bin/env python
"""TorchScript Mega-NeRF container (what the reference's --model_path takes) -> .npz container for mnv_headless /
viewer::VolumeRenderer::load_model.   usage: export_model.py container.pt model.npz [--sigma-activation relu]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mega_nerf_viewer_b200 as mnv  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("container")
    ap.add_argument("out")
    ap.add_argument("--sigma-activation", default="softplus", choices=["softplus", "relu"])
    a = ap.parse_args()
    mnv.export.export_torchscript_container(a.container, a.out, a.sigma_activation)
    print("wrote", a.out)
