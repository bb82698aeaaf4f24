import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mega_nerf_viewer_b200 as mnv
from mlp_reference import MegaNerfMLP
torch.manual_seed(3)
ref = MegaNerfMLP().cuda().eval()
model = mnv.MlpModel([ref.export()])
rows = 65536
x = torch.rand((rows, model.in_dim), device="cuda") * 2 - 1; x[:, -1] = 0
with torch.no_grad(): wb = ref(x, emulate_bf16=True)
got = model.forward(x); torch.cuda.synchronize()
d = (got - wb).abs().double(); scale = wb.double().pow(2).mean(0).sqrt()
print("col scale", scale.cpu().numpy().round(3))
rel = d / (scale + 1e-6)
print("per-col max rel", rel.max(0).values.cpu().numpy().round(4))
rowmax = rel.max(1).values
print("rows with rel>1e-2:", (rowmax > 1e-2).sum().item(), " >1e-3:", (rowmax>1e-3).sum().item(), "of", rows)
bad = torch.nonzero(rowmax > 1e-2).flatten()[:20]
print("bad rows", bad.cpu().numpy(), "mod128", (bad % 128).cpu().numpy())
# is the error row-wise "all columns" (cascade) or single-column?
for r in bad[:3]:
    print(int(r), rel[r].cpu().numpy().round(3))
# distribution
q = torch.quantile(rowmax.float(), torch.tensor([0.5,0.9,0.99,0.999,0.9999], device="cuda"))
print("rowmax quantiles 50/90/99/99.9/99.99:", q.cpu().numpy())
# compare with torch reference computed with a different accumulation order (bf16 emulation in float64)
with torch.no_grad():
    ref64 = MegaNerfMLP().cuda().double().eval(); ref64.load_state_dict(ref.state_dict())
    def lin(layer, xx):
        w = layer.weight.float().to(torch.bfloat16).double(); xb = xx.float().to(torch.bfloat16).double()
        return xb @ w.t() + layer.bias
    ref64._lin = lambda layer, xx, bf: lin(layer, xx)
    w64 = ref64(x.double(), emulate_bf16=True).float()
d2 = (wb - w64).abs().double(); rel2 = d2/(scale+1e-6)
print("torch-fp32-accum vs torch-fp64-accum (both bf16-emulated): max rel", rel2.max().item(), "frac>1e-3", (rel2>1e-3).double().mean().item(), "rms", rel2.pow(2).mean().sqrt().item())
d3 = (got - w64).abs().double(); rel3 = d3/(scale+1e-6)
print("kernel vs torch-fp64-accum: max rel", rel3.max().item(), "frac>1e-3", (rel3>1e-3).double().mean().item(), "rms", rel3.pow(2).mean().sqrt().item())
