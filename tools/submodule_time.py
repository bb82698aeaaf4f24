"""Dev: time query_submodules (bucketed launches, sizes decided on the device) for S sub-modules and V rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mega_nerf_viewer_b200 as mnv
for S, V in ((2, 2_300_000), (8, 2_300_000), (8, 262_144)):
    model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3 + i) for i in range(S)], device=0)
    x = torch.rand((V, model.in_dim), device="cuda") * 2 - 1
    x[:, -1] = 0
    cl = torch.randint(0, S, (V,), device="cuda").to(torch.int16)
    out = torch.empty((V, model.out_dim + 1), device="cuda")
    for _ in range(3):
        model.query_submodules(cl, x, out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); model.query_submodules(cl, x, out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    print(f"pair={os.environ.get('MNV_MLP_PAIR', '1')} per={os.environ.get('MNV_MLP_PER', '0')}: {S} sub-modules, {V} rows: {ms[len(ms)//2]:.3f} ms "
          f"({V * model.flops_per_row / ms[len(ms)//2] / 1e9:.0f} TFLOP/s)", flush=True)
    model.close()
