import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for mb in (1, 9, 64):
    a = torch.zeros(mb * 1024 * 1024 // 4, device=dev); b = torch.zeros_like(a)
    def xchg():
        ops = [dist.P2POp(dist.isend, a, (rank - 1) % world), dist.P2POp(dist.irecv, b, (rank + 1) % world)]
        for r in dist.batch_isend_irecv(ops): r.wait()
    for _ in range(5): xchg()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): xchg()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("ring exchange %d MB: %.3f ms each -> %.1f GB/s" % (mb, e0.elapsed_time(e1) / 20, mb / 1024 / (e0.elapsed_time(e1) / 20e3)), flush=True)
print("can_access_peer", rank, torch.cuda.can_device_access_peer(local, (local + 1) % world))
dist.destroy_process_group()
