"""Dev: per-column / per-tile error of the fused MLP vs the bf16-emulated torch reference."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mega_nerf_viewer_b200 as mnv
from mlp_reference import MegaNerfMLP
torch.manual_seed(3)
ref = MegaNerfMLP().cuda().eval()
model = mnv.MlpModel([ref.export()])
rows = 4096
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.rand((rows, model.in_dim), device="cuda", generator=g) * 2 - 1; x[:, -1] = torch.randint(0, 4, (rows,), device="cuda", generator=g).float()
with torch.no_grad():
    wb = ref(x, emulate_bf16=True)
got = model.forward(x); torch.cuda.synchronize()
d = (got - wb).abs().cpu().numpy(); scale = float(wb.pow(2).mean().sqrt())
print("scale", scale)
print("per-column max err / scale:", np.round(d.max(0) / scale, 5))
print("per-column mean err / scale:", np.round(d.mean(0) / scale, 6))
rr = d.max(1) / scale
for t in range(4):
    sel = (np.arange(rows) // 128) % 2 == (t % 2)
    print("tile parity", t % 2, "frac rows >1e-3:", (rr[sel] > 1e-3).mean())
    if t == 1: break
print("frac rows > 1e-3:", (rr > 1e-3).mean(), " rel L2:", float((got - wb).norm() / wb.norm()))
