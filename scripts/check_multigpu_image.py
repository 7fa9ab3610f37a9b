"""N-GPU check of the spp-sharded path (run under torchrun, NCCL): the reduced image equals the single-GPU
render of the same seed up to the fp32 order of the N partial sums.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_multigpu_image.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import vecchio_b200 as vb
from vecchio_b200.sharding import reduce_sums_to_root, spp_slice

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for name, W, spp, depth in (("cornell_box", 300, 200, 100), ("final_scene", 200, 64, 100)):
    scene = vb.Scene(name, seed=1); cam = scene.next_camera(); H = scene.height_for(W)
    ctx = vb.Context(local); ctx.upload(scene)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    n = W * H * 3
    d_sum = torch.empty(n, dtype=torch.float32, device=dev); d_rgb = torch.empty_like(d_sum)
    b, c = spp_slice(rank, world, spp)
    st = ctx.render_device(cam, vb.render_params(W, H, spp, depth, seed=7, spp_begin=b, spp_count=c), d_sum.data_ptr())
    reduce_sums_to_root(d_sum, world)
    rays = torch.tensor([st.rays], dtype=torch.int64, device=dev); dist.all_reduce(rays)
    if rank == 0:
        ctx.finalize_device(d_sum.data_ptr(), d_rgb.data_ptr(), n, spp)
        stream.synchronize()
        sharded = d_rgb.cpu().numpy().reshape(H, W, 3)
        whole, _, sw = ctx.render(cam, vb.render_params(W, H, spp, depth, seed=7))
        err = np.abs(sharded - whole).max() / max(whole.max(), 1e-9)
        ok = np.allclose(sharded, whole, rtol=1e-5, atol=1e-7) and int(rays.item()) == sw.rays
        print(f"{name}: {world} GPUs, {W}x{H}x{spp}: sharded vs single-GPU max rel diff {err:.2e}, rays {int(rays.item())} vs {sw.rays}: {'OK' if ok else 'MISMATCH'}", flush=True)
        assert ok
    dist.barrier()
    ctx.close()
dist.destroy_process_group()
