"""probe: torch symmetric memory (peer pointers over NVLink) + latency of small collectives on this pod.
torchrun --nproc-per-node 2 tests/multi_gpu/probe_symm.py"""
import os, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)

def timeit(name, f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {1e6*(time.perf_counter()-t0)/n:.1f} us", flush=True)

x = torch.zeros(1, dtype=torch.float64, device=dev)
timeit("all_reduce 1 double", lambda: dist.all_reduce(x))
outs = [torch.empty_like(x) for _ in range(world)]
timeit("all_gather 1 double", lambda: dist.all_gather(outs, x))
for cnt in (0, 1000, 100000):
    send = torch.zeros(cnt * 4 * (world - 1), dtype=torch.float64, device=dev)
    recv = torch.empty_like(send)
    sp = [0 if r == rank else cnt * 4 for r in range(world)]
    timeit(f"all_to_all_single {cnt} particles/peer", lambda: dist.all_to_all_single(recv, send, sp, sp))
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 1 << 20
    t = symm_mem.empty(n, dtype=torch.float64, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(float(rank))
    hdl.barrier()
    peer = (rank + 1) % world
    pbuf = hdl.get_buffer(peer, (n,), torch.float64)
    if rank == 0: print("symm ok; peer value", float(pbuf[0]), "ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    src = torch.full((1000,), 100.0 + rank, dtype=torch.float64, device=dev)
    def peer_write():
        pbuf[:1000].copy_(src)
    timeit("peer copy 1000 doubles", peer_write)
    timeit("symm barrier", lambda: hdl.barrier())
    hdl.barrier()
    torch.cuda.synchronize()
    if rank == 0: print("after peer write, my buffer holds", float(t[0]), "(expect", 100.0 + (rank - 1) % world, ")", flush=True)
    big = torch.zeros(n, dtype=torch.float64, device=dev)
    timeit("peer copy 8 MB", lambda: pbuf.copy_(big))
except Exception as e:  # noqa: BLE001
    print("symmetric memory unavailable:", repr(e), flush=True)
dist.destroy_process_group()
