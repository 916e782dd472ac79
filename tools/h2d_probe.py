#!/usr/bin/env python
"""What bounds the per-rank pinned host->device rate when all 8 ranks of a box copy at once?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/h2d_probe.py

Every rank reports: its CPU affinity, the NUMA node of its GPU (sysfs, through the PCI bus id NVML gives), the memory nodes the
container may allocate from (cgroup cpuset.mems), and the H2D rate of a 1 GiB pinned buffer
  (a) alone (ranks take turns),
  (b) all ranks at once, buffers pinned by torch after binding the process to the GPU-local CPUs (what bench.py does),
  (c) all ranks at once, buffers first-touched under set_mempolicy(MPOL_BIND, <GPU's NUMA node>) and then cudaHostRegister'ed,
  (d) all ranks at once with half the bytes (what a narrower host staging format would buy).
Rank 0 prints one JSON object. The e2e leg of bench.py needs 1.7536 GB per 65 ms step = 27 GB/s per rank."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NBYTES = 1 << 30


def gpu_numa_node(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        link = f"gen{pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h)} x{pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)}"
        return node, bus, link
    except Exception as e:
        return None, str(e)[:80], None


def set_mempolicy(mode, node):
    libc = ctypes.CDLL("libc.so.6", use_errno=True)
    if node is None:
        return libc.syscall(238, 0, None, 0)
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(238, mode, ctypes.byref(mask), 64)
    return rc if rc == 0 else -ctypes.get_errno()


def rate(dst, src, reps=4):
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return reps * src.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bench = __import__("bench")
    info = {"rank": rank, "cpus_before": len(os.sched_getaffinity(0))}
    info["bind"] = bench.bind_to_gpu_cpus(local)
    info["cpus_after"] = sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]
    node, bus, link = gpu_numa_node(local)
    info.update({"gpu_numa_node": node, "pci": bus, "pcie": link})
    for pth in ("/sys/fs/cgroup/cpuset.mems.effective", "/sys/fs/cgroup/cpuset/cpuset.mems"):
        if os.path.exists(pth):
            info["cgroup_mems"] = open(pth).read().strip()
            break
    try:
        info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
    except OSError:
        pass
    dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    src = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    solo = 0.0
    for r in range(world):
        barrier()
        if r == rank:
            solo = rate(dst, src)
    info["solo_gbs"] = solo
    barrier()
    info["concurrent_gbs"] = rate(dst, src, 8)
    barrier()
    info["concurrent_half_bytes_gbs"] = rate(dst[:NBYTES // 2], src[:NBYTES // 2], 16)
    # (c) NUMA-bound first touch + cudaHostRegister
    if node is not None and node >= 0:
        rc = set_mempolicy(2, node)
        info["set_mempolicy_rc"] = rc
        if rc == 0:
            buf = torch.empty(NBYTES, dtype=torch.uint8)
            buf.fill_(1)
            set_mempolicy(0, None)
            er = torch.cuda.cudart().cudaHostRegister(buf.data_ptr(), NBYTES, 0)
            info["host_register_rc"] = int(er)
            barrier()
            info["concurrent_numa_bound_gbs"] = rate(dst, buf, 8)
            barrier()
            torch.cuda.cudart().cudaHostUnregister(buf.data_ptr())
        else:
            barrier(); barrier()
    else:
        barrier(); barrier()
    allinfo = [None] * world
    if world > 1:
        dist.all_gather_object(allinfo, info)
    else:
        allinfo = [info]
    if rank == 0:
        agg = {k: sum(i.get(k) or 0 for i in allinfo) for k in ("solo_gbs", "concurrent_gbs", "concurrent_numa_bound_gbs", "concurrent_half_bytes_gbs")}
        print(json.dumps({"world": world, "aggregate_gbs": agg, "need_per_rank_gbs": 1.7536 / 0.065, "ranks": allinfo}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
