"""Does torch symmetric memory (P2P-mapped peer buffers) work on this box?  2 ranks."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm_mem.empty(1024, 256, dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024, 256), torch.bfloat16)
    print(rank, "peer value", float(peer[0, 0]), "ptr", hex(peer.data_ptr()), flush=True)
    # write into the peer with one of OUR kernels (GEMM epilogue -> peer memory)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from more4d_b200 import ops
    a = torch.randn(1024, 512, device=dev, dtype=torch.bfloat16)
    w = torch.randn(256, 512, device=dev, dtype=torch.bfloat16) * 0.05
    hdl.barrier()
    ops.linear(a, w, None, out=peer)
    ref = ops.linear(a, w, None)
    hdl.barrier()
    torch.cuda.synchronize()
    # what the peer wrote into MY buffer must equal what the peer computed locally
    other = [torch.empty_like(ref) for _ in range(world)]
    dist.all_gather(other, ref)
    print(rank, "gemm->peer exact:", bool(torch.equal(t, other[(rank - 1) % world])), flush=True)
    s = time.perf_counter()
    for _ in range(100):
        hdl.barrier()
    torch.cuda.synchronize()
    print(rank, "barrier us", (time.perf_counter() - s) * 1e4, flush=True)
except Exception as e:          # noqa: BLE001
    print(rank, "SYMM FAILED", type(e).__name__, str(e)[:300], flush=True)
dist.destroy_process_group()
