"""CPU tests of the model-container writer / exporter: Mega-NeRF parameter names ->
flat container, round trip through the .npz the C++ loader reads, and a real TorchScript container."""
import numpy as np
import pytest


def megenerf_state_dict(ref):
    """The names cmusatyalab/mega-nerf's NeRF module gives the same tensors."""
    sd = {}
    for i, l in enumerate(ref.trunk):
        sd[f"xyz_encodings.{i}.0.weight"], sd[f"xyz_encodings.{i}.0.bias"] = l.weight, l.bias
    sd["sigma.weight"], sd["sigma.bias"] = ref.sigma.weight, ref.sigma.bias
    sd["xyz_encoding_final.weight"], sd["xyz_encoding_final.bias"] = ref.final.weight, ref.final.bias
    if ref.embedding is not None:
        sd["embedding_a.weight"] = ref.embedding.weight
    sd["dir_encoding.0.weight"], sd["dir_encoding.0.bias"] = ref.head1.weight, ref.head1.bias
    sd["rgb.weight"], sd["rgb.bias"] = ref.head2.weight, ref.head2.bias
    return sd


@pytest.mark.parametrize("need_viewdir,app_dim,basis", [(False, 48, 9), (True, 48, 1), (True, 0, 4), (False, 0, 9)])
def test_state_dict_mapping_recovers_the_module(mnv, need_viewdir, app_dim, basis):
    import torch
    from mlp_reference import MegaNerfMLP

    torch.manual_seed(1)
    ref = MegaNerfMLP(basis_dim=basis, appearance_dim=app_dim, need_viewdir=need_viewdir)
    got = mnv.export.submodule_from_state_dict(megenerf_state_dict(ref))
    want = ref.export()
    assert got["skip_layer"] == 4 and got["pe_xyz_freqs"] == 12 and got["need_viewdir"] == need_viewdir
    if need_viewdir:
        assert got["pe_dir_freqs"] == 4
    assert (got["embedding"] is None) == (app_dim == 0)
    for k in ("sigma_w", "sigma_b", "final_w", "final_b", "head1_w", "head1_b", "head2_w", "head2_b"):
        assert np.array_equal(got[k], want[k]), k
    for a, b in zip(got["trunk_w"] + got["trunk_b"], want["trunk_w"] + want["trunk_b"]):
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        mnv.export.submodule_from_state_dict({"foo": np.zeros(3)})


def test_container_round_trip(mnv, tmp_path):
    subs = [mnv.synth.make_mlp_weights(seed=s) for s in (1, 2, 3)]
    p = tmp_path / "m.npz"
    mnv.save_model_container(str(p), subs, grid_dim=(1, 3), min_position=(-1, -2, -3), max_position=(1, 2, 3),
                             centroids=np.arange(9, dtype=np.float32).reshape(3, 3), compressed=True)
    z = np.load(p)
    assert z["grid_dim"].tolist() == [1, 3] and z["centroids"].shape == (3, 3)
    assert bool(z["need_appearance_embedding"]) and not bool(z["need_viewdir"])
    assert z["sub_module_2/config"].tolist() == [8, 4, 12, 4, 1]
    assert np.array_equal(z["sub_module_1/trunk_w_4"], subs[1]["trunk_w"][4]) and z["sub_module_1/trunk_w_4"].shape == (256, 331)
    assert np.array_equal(z["sub_module_0/embedding"], subs[0]["embedding"])


def test_torchscript_container_export(mnv, tmp_path):
    """A TorchScript container with the attributes the reference reads (cuda_renderer.cpp:525-539)."""
    import torch
    from torch import nn

    class Sub(nn.Module):  # parameter names of Mega-NeRF's NeRF module
        def __init__(self):
            super().__init__()
            pe, w = 75, 256
            self.xyz_encodings = nn.ModuleList(
                [nn.Sequential(nn.Linear(pe if i == 0 else (w + pe if i == 4 else w), w), nn.ReLU()) for i in range(8)])
            self.sigma = nn.Linear(w, 1)
            self.xyz_encoding_final = nn.Linear(w, w)
            self.embedding_a = nn.Embedding(3, 48)
            self.dir_encoding = nn.Sequential(nn.Linear(w + 48, 128), nn.ReLU())
            self.rgb = nn.Linear(128, 27)

        def forward(self, x: torch.Tensor, sigma_only: bool = False) -> torch.Tensor:
            return x

    class Container(nn.Module):
        def __init__(self):
            super().__init__()
            self.sub_module_0, self.sub_module_1 = Sub(), Sub()
            self.register_buffer("grid_dim", torch.tensor([1, 2]))
            self.register_buffer("min_position", torch.tensor([-1.0, -1.0, -1.0]))
            self.register_buffer("max_position", torch.tensor([1.0, 1.0, 1.0]))
            self.register_buffer("centroids", torch.zeros(2, 3))
            self.need_viewdir = False
            self.need_appearance_embedding = True

        def forward(self, x: torch.Tensor) -> torch.Tensor:
            return self.sub_module_0(x) + self.sub_module_1(x)

    torch.manual_seed(0)
    c = Container()
    pt, out = tmp_path / "c.pt", tmp_path / "m.npz"
    torch.jit.script(c).save(str(pt))
    mnv.export.export_torchscript_container(str(pt), str(out))
    z = np.load(out)
    assert z["grid_dim"].tolist() == [1, 2] and bool(z["need_appearance_embedding"]) and not bool(z["need_viewdir"])
    assert np.array_equal(z["sub_module_1/head2_w"], c.sub_module_1.rgb.weight.detach().numpy())
    assert np.array_equal(z["sub_module_0/trunk_w_4"], c.sub_module_0.xyz_encodings[4][0].weight.detach().numpy())
    assert z["sub_module_0/config"].tolist() == [8, 4, 12, 4, 1]
