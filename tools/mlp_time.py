"""Dev: time mlp_forward_kernel on the refinement batch (262 144 rows, one sub-module) and larger / smaller batches.
   python tools/mlp_time.py [--lib path] [--rows 4096,32768,262144,2097152]"""
import argparse, os, sys
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--rows", default="4096,32768,262144,2097152")
ap.add_argument("--tag", default="")
args = ap.parse_args()
if args.lib:
    os.environ["MNV_B200_LIB"] = os.path.abspath(args.lib)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mega_nerf_viewer_b200 as mnv
dev = torch.device("cuda", 0)
model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3)], device=0)
res = []
for rows in [int(r) for r in args.rows.split(",")]:
    x = torch.rand((rows, model.in_dim), device=dev) * 2 - 1
    x[:, -1] = 0
    out = torch.empty((rows, model.out_dim + 1), device=dev)
    for _ in range(3):
        model.forward(x, out=out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); model.forward(x, out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    res.append(f"{rows} rows: {med:.4f} ms ({rows * model.flops_per_row / med / 1e9:.0f} TFLOP/s)")
print(args.tag, os.path.basename(mnv.LIB_PATH), "|", "; ".join(res))
