"""What the column exchange of bench.py --gpus N costs, piece by piece: torchrun ... scripts/a2a_probe.py
(strided slicing + concatenation of the send buffer, the NCCL all-to-all, an all-gather of the same bytes for comparison)."""
import os, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
rows, cols = 226853, 189
local_m = torch.randint(0, 1000, (rows, cols), dtype=torch.int32, device=dev)
ncol_of = [len(range(r, cols, world)) for r in range(world)]
cnt = ncol_of[rank]
def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1000.0 * (time.perf_counter() - t0) / reps
send = torch.cat([local_m[:, r::world].reshape(-1) for r in range(world)])
full = torch.empty((rows * world, cnt), dtype=torch.int32, device=dev)
gath = torch.empty((rows * world, cols), dtype=torch.int32, device=dev)
res = {
    "cat_ms": t(lambda: torch.cat([local_m[:, r::world].reshape(-1) for r in range(world)])),
    "a2a_ms": t(lambda: dist.all_to_all_single(full.view(-1), send, output_split_sizes=[rows * cnt] * world, input_split_sizes=[rows * c for c in ncol_of])),
    "allgather_rows_ms": t(lambda: dist.all_gather_into_tensor(gath.view(-1), local_m.view(-1))),
    "bytes_sent_per_rank_MB": round(send.numel() * 4 * (world - 1) / world / 1e6, 1),
}
if rank == 0:
    print(res, {k: v for k, v in os.environ.items() if k.startswith("NCCL_")})
dist.destroy_process_group()
