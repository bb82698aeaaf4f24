"""GPU tests of the fused tcgen05 MLP against the PyTorch
statement of the named shapes (tests/mlp_reference.py).  Tolerance from
BASELINE.json: 1e-3 relative with bf16 operands / fp32 accumulation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    """relative error of the whole output block (Frobenius norm)."""
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b.astype(np.float64)) + 1e-30))


def _row_rel(a, b):
    """per-row max |a-b| relative to the rms magnitude of the reference outputs."""
    scale = np.sqrt((b.astype(np.float64) ** 2).mean()) + 1e-30
    return np.abs(a.astype(np.float64) - b).max(1) / scale


def _check_close(got, want_bf16, want_fp32):
    """bf16 operands / fp32 accumulation, tolerance 1e-3 relative (BASELINE.json).

    Activations are re-rounded to bf16 between layers, so two correct
    implementations that differ only in fp32 summation order disagree on a small
    fraction of rows where one activation lands on the other side of a bf16
    rounding boundary (measured: torch fp32-accumulate vs fp64-accumulate of the
    same bf16 emulation differ by >1e-3 on 0.3 % of elements).  Hence: relative L2
    error <= 1e-3 (measured 5e-5), >= 98 % of rows within 1e-3, no row beyond 5e-2;
    against the pure-fp32 module only the bf16 operand rounding shows (~2e-3 L2)."""
    assert np.isfinite(got).all()
    assert _rel_l2(got, want_bf16) < 1e-3, _rel_l2(got, want_bf16)
    rr = _row_rel(got, want_bf16)
    assert (rr < 1e-3).mean() >= 0.98, (rr < 1e-3).mean()
    assert rr.max() < 5e-2, rr.max()
    assert _rel_l2(got, want_fp32) < 1e-2, _rel_l2(got, want_fp32)


@pytest.mark.parametrize("need_viewdir,app_dim,basis,rows", [
    (False, 48, 9, 128 * 5), (False, 48, 9, 1000), (True, 48, 9, 777), (False, 0, 1, 300), (True, 0, 4, 129)])
def test_mlp_matches_torch_reference(need_viewdir, app_dim, basis, rows, mnv):
    import torch
    from mlp_reference import MegaNerfMLP, flops_per_row

    torch.manual_seed(3)
    ref = MegaNerfMLP(basis_dim=basis, appearance_dim=app_dim, need_viewdir=need_viewdir).cuda().eval()
    model = mnv.MlpModel([ref.export()])
    assert model.in_dim == 3 + 3 * need_viewdir + (app_dim > 0) and model.out_dim == 3 * basis + 1
    assert model.flops_per_row == pytest.approx(flops_per_row(ref))
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((rows, model.in_dim), device="cuda", generator=g) * 2 - 1
    if need_viewdir:
        x[:, 3:6] = torch.nn.functional.normalize(x[:, 3:6], dim=-1)
    if app_dim > 0:
        x[:, -1] = torch.randint(0, 4, (rows,), device="cuda", generator=g).float()
    with torch.no_grad():
        want_bf16 = ref(x, emulate_bf16=True).cpu().numpy()
        want_fp32 = ref(x, emulate_bf16=False).cpu().numpy()
    got = model.forward(x)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    _check_close(got, want_bf16, want_fp32)
    model.close()


def test_mlp_config4_shape_and_determinism(mnv):
    """Config 4: 4096 splits x 8 children x 8 samples = 262144 rows per frame."""
    import torch
    from mlp_reference import MegaNerfMLP

    torch.manual_seed(3)
    ref = MegaNerfMLP().cuda().eval()
    model = mnv.MlpModel([ref.export(), ref.export()])
    x = torch.rand((262144, model.in_dim), device="cuda") * 2 - 1
    x[:, -1] = 0
    a = model.forward(x, submodule=0)
    b = model.forward(x, submodule=1)
    torch.cuda.synchronize()
    assert torch.equal(a, b)  # identical weights, deterministic kernel
    with torch.no_grad():
        want = ref(x[:4096], emulate_bf16=True).cpu().numpy()
        want32 = ref(x[:4096], emulate_bf16=False).cpu().numpy()
    _check_close(a[:4096].cpu().numpy(), want, want32)
    model.close()


_MODE_SCRIPT = r"""
import sys, zlib, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import mega_nerf_viewer_b200 as mnv
from mlp_reference import MegaNerfMLP
torch.manual_seed(3)
crc = 0
for need_viewdir, rows in ((False, 128 * 5 + 77), (True, 1000), (False, 40000)):
    ref = MegaNerfMLP(need_viewdir=need_viewdir).cuda().eval()
    model = mnv.MlpModel([ref.export()])
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((rows, model.in_dim), device="cuda", generator=g) * 2 - 1
    x[:, -1] = torch.randint(0, 4, (rows,), device="cuda", generator=g).float()
    out = model.forward(x)
    torch.cuda.synchronize()
    crc = zlib.crc32(out.cpu().numpy().tobytes(), crc)
    model.close()
print("CRC", crc)
"""


def test_mlp_launch_shapes_bit_identical(tmp_path):
    """The MLP's launch shapes — CTA pairs with cta_group::2 MMAs (M = 256 over two SMs) or one CTA (M = 128), 1 / 2 / 4 MMAs
    per weight-ring stage — reorder nothing inside a row's dot products: outputs are equal bit for bit (odd group counts and
    a partial last group included, so the pair's second CTA also runs a group without rows)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "mlp_mode.py"
    script.write_text(_MODE_SCRIPT)
    crcs = {}
    for pair, per in ((1, 0), (1, 1), (0, 1), (0, 2)):
        env = dict(os.environ, MNV_MLP_PAIR=str(pair), MNV_MLP_PER=str(per))
        r = subprocess.run([sys.executable, str(script), root], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        crcs[(pair, per)] = [l for l in r.stdout.splitlines() if l.startswith("CRC")][-1]
    assert len(set(crcs.values())) == 1, crcs
