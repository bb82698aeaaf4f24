"""Dev: error statistics + timing of the fused MLP vs the torch reference."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mega_nerf_viewer_b200 as mnv
from mlp_reference import MegaNerfMLP
torch.manual_seed(3)
ref = MegaNerfMLP().cuda().eval()
model = mnv.MlpModel([ref.export()])
rows = int(os.environ.get("ROWS", 262144))
x = torch.rand((rows, model.in_dim), device="cuda") * 2 - 1; x[:, -1] = 0
with torch.no_grad():
    wb = ref(x[:65536], emulate_bf16=True); wf = ref(x[:65536], emulate_bf16=False)
got = model.forward(x); torch.cuda.synchronize()
g = got[:65536]
for name, w in (("bf16-emulated", wb), ("fp32", wf)):
    d = (g - w).double(); scale = w.double().pow(2).mean(0).sqrt() + 1e-6
    rel = (d.abs() / scale)
    print(f"vs {name}: max rel-to-scale {rel.max().item():.2e}  rms rel {rel.pow(2).mean().sqrt().item():.2e}  rel L2 {(d.norm()/w.double().norm()).item():.2e}  frac>1e-3 {(rel>1e-3).double().mean().item():.2e}")
# timing
for _ in range(3): model.forward(x)
torch.cuda.synchronize()
ms = []
for _ in range(20):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); model.forward(x); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
ms = np.array(ms); tf = rows * model.flops_per_row / (ms.mean() * 1e-3) / 1e12
print(f"rows {rows}: {ms.mean():.3f} ms (min {ms.min():.3f}) -> {tf:.1f} TFLOP/s bf16 = {tf/1378.9*100:.1f}% of sustained measured peak; {rows/ms.mean()/1e3:.1f} Mrows/s")
# torch eager fp16 autocast (how the reference runs the model) for comparison
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for _ in range(2): ref(x)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ref(x)
    e1.record(); torch.cuda.synchronize()
print(f"torch eager fp16 autocast: {e0.elapsed_time(e1)/5:.3f} ms")
