"""Times the dibit all-gather alone (run under torchrun): transport sanity check for the multi-GPU bench."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
for nbytes in (1 << 16, 1 << 20, 1 << 22, 1 << 26):
    src = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.zeros(nbytes * world, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dist.all_gather_into_tensor(dst, src)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(dst, src)
    ev1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"all_gather {nbytes} B/rank: {ev0.elapsed_time(ev1) / 10:.3f} ms/call (wall {(time.perf_counter() - t0) * 100:.3f} ms)", flush=True)
dist.destroy_process_group()
