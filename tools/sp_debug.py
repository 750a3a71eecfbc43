import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist, flux2b
from oracle import flux2_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); device = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=device)
cfg = O.DiTConfig(num_layers=1, num_single_layers=1, num_attention_heads=8, joint_attention_dim=256, guidance_embeds=False)
ctx = flux2b.Context(dit=cfg, device=rank, options={"keep_raw_weights": 0, "sp_mode": int(os.environ.get("SP_MODE", "0"))})
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=device).manual_seed(0)
for k, (o, i) in O.dit_weight_shapes(cfg).items():
    b = 1.0 / math.sqrt(i)
    ctx.set_tensor(k, torch.empty(o, i, device=device, dtype=torch.float32).uniform_(-b, b, generator=g).to(torch.bfloat16))
ctx.finalize()
ids = [flux2b.sp_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0, device=device)
print(rank, "id len", len(ids[0]), flush=True)
ctx.sp_init(ids[0], rank, world)
S_img, HW = 256, 256
lat = torch.randn(1, S_img, 128, generator=torch.Generator().manual_seed(42)).to(device)
enc = torch.randn(1, 64, 256, generator=torch.Generator().manual_seed(43)).to(torch.bfloat16).to(device)
ctx.prof_enable(True); ctx.prof_reset()
x = lat.clone()
ctx.denoise(x, enc, [1.0, 0.5], HW, HW)
torch.cuda.synchronize()
print(rank, "comm", ctx.prof_get(flux2b.PROF_COMM), "gemm", ctx.prof_get(flux2b.PROF_GEMM)["launches"], "x", float(x.abs().sum()), flush=True)
dist.barrier(); dist.destroy_process_group()
