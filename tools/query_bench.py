import sys, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
import mega_nerf_viewer_b200 as mnv
tree = mnv.synth.make_tree(depth=8); dt = mnv.DeviceTree(tree)
for n in (1_000_000, 4_000_000):
    x = torch.rand((n,3), device='cuda')
    out = dt.query_points(x); torch.cuda.synchronize()
    for rep in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        t0=time.perf_counter(); e0.record(); out = dt.query_points(x); e1.record(); torch.cuda.synchronize(); t1=time.perf_counter()
        print(n, "event ms", e0.elapsed_time(e1), "wall ms", (t1-t0)*1e3)
    d = out[:,2].float(); print("mean depth", d.mean().item(), "max", d.max().item())
