"""Mega-NeRF checkpoint / TorchScript container -> the flat `.npz` model container viewer::VolumeRenderer::load_model
reads (csrc/viewer/model.hpp).  The reference loads the TorchScript archive itself with LibTorch
(src/renderer/cuda_renderer.cpp:518-543); here the tensors are pulled out once, offline, with PyTorch.

Parameter names follow cmusatyalab/mega-nerf's `NeRF` module (the artefact is not part of the viewer repository, so
this mapping is checked against a synthetic state dict only — tests/test_export.py):

    xyz_encodings.<i>.0.{weight,bias}   trunk layer i (Linear + ReLU)
    sigma.{weight,bias}                 density head
    xyz_encoding_final.{weight,bias}    feature layer
    embedding_a.weight                  appearance embedding table (optional)
    dir_encoding.0.{weight,bias}        head 1: cat(feature, [dir PE], [appearance]) -> width / 2
    rgb.{weight,bias}                   head 2: -> 3 (RGB) or 3 * basis (SH)
"""
from __future__ import annotations

import numpy as np


def submodule_from_state_dict(sd: dict, sigma_activation: str = "softplus") -> dict:
    """-> the dict mega_nerf_viewer_b200.MlpModel / save_model_container take.  Hyper-parameters are read off the
    tensor shapes: positional-encoding octaves from the first layer's fan-in, the skip layer from the layer whose
    fan-in is width + PE, view directions / appearance from head 1's fan-in."""
    f = lambda k: np.ascontiguousarray(np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k]), np.float32)
    n = 0
    while f"xyz_encodings.{n}.0.weight" in sd:
        n += 1
    if n < 2:
        raise ValueError("no xyz_encodings.<i>.0.weight tensors: not a Mega-NeRF NeRF state dict")
    trunk_w = [f(f"xyz_encodings.{i}.0.weight") for i in range(n)]
    trunk_b = [f(f"xyz_encodings.{i}.0.bias") for i in range(n)]
    width, pe = trunk_w[0].shape
    if (pe - 3) % 6:
        raise ValueError(f"first layer fan-in {pe} is not 3 + 6 * octaves")
    skips = [i for i in range(1, n) if trunk_w[i].shape[1] == width + pe]
    if len(skips) != 1 or any(trunk_w[i].shape[1] != width for i in range(1, n) if i not in skips):
        raise ValueError("expected exactly one skip layer with fan-in width + PE")
    emb = f("embedding_a.weight") if "embedding_a.weight" in sd else None
    head1_w = f("dir_encoding.0.weight")
    extra = head1_w.shape[1] - width - (0 if emb is None else emb.shape[1])
    if extra < 0 or (extra and (extra - 3) % 6):
        raise ValueError(f"head fan-in {head1_w.shape[1]} does not decompose into feature + dir PE + appearance")
    return dict(trunk_w=trunk_w, trunk_b=trunk_b, sigma_w=f("sigma.weight"), sigma_b=f("sigma.bias"),
                final_w=f("xyz_encoding_final.weight"), final_b=f("xyz_encoding_final.bias"), embedding=emb,
                head1_w=head1_w, head1_b=f("dir_encoding.0.bias"), head2_w=f("rgb.weight"), head2_b=f("rgb.bias"),
                skip_layer=skips[0], pe_xyz_freqs=(pe - 3) // 6, pe_dir_freqs=(extra - 3) // 6 if extra else 4,
                need_viewdir=bool(extra), sigma_activation=1 if sigma_activation == "softplus" else 0)


def export_torchscript_container(script_path: str, out_path: str, sigma_activation: str = "softplus") -> None:
    """The container the reference loads (attrs grid_dim / min_position / max_position / centroids / need_viewdir /
    need_appearance_embedding / sub_module_<i>, cuda_renderer.cpp:525-539) -> `.npz`."""
    import torch

    from . import save_model_container

    c = torch.jit.load(script_path, map_location="cpu")
    centroids = c.centroids.detach().cpu().numpy()
    subs = [submodule_from_state_dict(getattr(c, f"sub_module_{i}").state_dict(), sigma_activation)
            for i in range(centroids.shape[0])]
    save_model_container(out_path, subs, grid_dim=c.grid_dim.detach().cpu().numpy(),
                         min_position=c.min_position.detach().cpu().numpy(),
                         max_position=c.max_position.detach().cpu().numpy(), centroids=centroids,
                         need_viewdir=bool(c.need_viewdir), need_appearance_embedding=bool(c.need_appearance_embedding))
