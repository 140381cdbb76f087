"""Device->host copy rate of every rank while all ranks copy at once (torchrun): what the end-to-end leg of bench.py can
at best sustain per GPU on this host. Prints one line per rank: alone (rank 0 only) and contended GB/s for image-sized
(16 x 1.344 MB) copies into page-locked memory, plus the NUMA node of the GPU and of the CPU the rank runs on."""
import os
import time

import torch
import torch.distributed as dist


def rate(dst, src, n, stream):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(n):
            dst.copy_(src, non_blocking=True)
    stream.synchronize()
    return n * src.numel() / (time.perf_counter() - t0) / 1e9


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    src = torch.zeros(16 * 3360 * 400, dtype=torch.uint8, device="cuda")
    dst = torch.empty(16 * 3360 * 400, dtype=torch.uint8, pin_memory=True)
    st = torch.cuda.Stream()
    rate(dst, src, 20, st)
    alone = None
    for r in range(world):                               # every rank alone, one after the other
        if world > 1:
            dist.barrier()
        if r == rank:
            alone = rate(dst, src, 100, st)
    if world > 1:
        dist.barrier()
    both = rate(dst, src, 200, st)                       # all ranks at once
    bus = torch.cuda.get_device_properties(lr).pci_bus_id if hasattr(torch.cuda.get_device_properties(lr), "pci_bus_id") else None
    numa_gpu = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(lr)
        busid = pynvml.nvmlDeviceGetPciInfo(h).busId
        busid = busid.decode() if isinstance(busid, bytes) else busid
        p = "/sys/bus/pci/devices/%s/numa_node" % busid.lower()[-12:]
        numa_gpu = open(p).read().strip() if os.path.exists(p) else "?"
    except Exception as e:
        numa_gpu = "err %s" % e
    cpu = os.sched_getcpu() if hasattr(os, "sched_getcpu") else -1
    print("rank %d: D2H alone %.1f GB/s, all %d ranks at once %.1f GB/s; gpu numa %s, running on cpu %d of %d (affinity %d cpus)" % (
        rank, alone, world, both, numa_gpu, cpu, os.cpu_count(), len(os.sched_getaffinity(0))), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
