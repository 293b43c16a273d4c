"""Is NVLink-switch multicast (NVLS) usable from this container?  torchrun --nproc-per-node N scripts/nvls_probe.py
Allocates symmetric memory through torch.distributed._symmetric_memory (plumbing only) and prints whether a multicast
address came back; then runs multimem.ld_reduce / multimem.st through torch's own one-shot all-reduce as a smoke test."""
import os, sys
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(f"rank {rank}: buffer_ptrs={[hex(p) for p in hdl.buffer_ptrs][:4]} multicast_ptr={hex(hdl.multicast_ptr)} "
          f"signal_pads={len(hdl.signal_pad_ptrs)} has_multicast={hdl.multicast_ptr != 0}", flush=True)
    t.fill_(rank + 1)
    dist.barrier()
    if hdl.multicast_ptr != 0:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(f"rank {rank}: multimem_all_reduce_ -> {out[:4].float().tolist()} (want {sum(range(1, world + 1))})", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory probe failed: {type(e).__name__}: {str(e)[:300]}", flush=True)
dist.barrier()
dist.destroy_process_group()
