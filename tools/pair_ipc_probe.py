"""Dev (torchrun, 2 ranks): does the MLP's CTA-pair kernel slow down once the process holds CUDA-IPC peer mappings?
Times one 2.3 M-row launch of the fused MLP before and after a ShardedGuided object (IPC-mapped exchange buffers) exists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import mega_nerf_viewer_b200 as mnv
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
dev = torch.device("cuda", lr)
model = mnv.MlpModel([mnv.synth.make_mlp_weights(seed=3)], device=lr)
V = 2_300_000
x = torch.rand((V, model.in_dim), device=dev) * 2 - 1
x[:, -1] = 0
out = torch.empty((V, model.out_dim + 1), device=dev)


def t(tag):
    for _ in range(3):
        model.forward(x, out=out)
    torch.cuda.synchronize()
    ms = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); model.forward(x, out=out); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    print(f"rank {rank} pair={os.environ.get('MNV_MLP_PAIR', '1')} {tag}: {ms[len(ms) // 2]:.3f} ms", flush=True)


t("before any peer mapping")
dist.barrier(); torch.cuda.synchronize()
t("after an NCCL barrier")
tree = mnv.synth.make_tree(depth=8)
W, H = 960, 540
sh = mnv.multigpu.ShardedGuided(tree, [mnv.synth.make_mlp_weights(seed=11 + i) for i in range(world)], mnv.synth.grid_for_world(world),
                                (-1, -1, -1), (1, 1, 1), W, H, rank=rank, world=world, device=lr, dist=dist)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t("with ShardedGuided (CUDA-IPC peer mappings) alive")
sh.close()
dist.barrier(); torch.cuda.synchronize()
t("after closing it")
dist.destroy_process_group()
