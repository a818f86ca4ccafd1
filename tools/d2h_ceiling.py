"""Host-side ceiling of the end-to-end path: pinned device->host (and host->device) copy bandwidth per GPU, alone and with
all ranks copying at once.  bench.py's `e2e` moves 62.9 MB of fp32 mel per step and rank to the host; this says what the
box can absorb, independent of any kernel.

    python tools/d2h_ceiling.py                                              (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/d2h_ceiling.py

Prints one JSON line (rank 0): per-rank GB/s alone / concurrent, the aggregate, CPU affinity, NUMA nodes visible to the
process, and -- when more than one NUMA node is visible -- the same D2H copy into buffers bound to each node
(mbind + cudaHostRegister), which is what a NUMA-aware pinned allocation could gain.
"""
import ctypes
import json
import mmap
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

MB = 62_914_560          # bytes of one step's mel: 256 x 768 x 80 fp32


def numa_nodes():
    base = "/sys/devices/system/node"
    try:
        return sorted(int(d[4:]) for d in os.listdir(base) if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return []


def timed_copy(dst, src, reps, stream):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
        stream.synchronize()
        ev0.record(stream)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        ev1.record(stream)
    stream.synchronize()
    return src.numel() * src.element_size() * reps / (ev0.elapsed_time(ev1) * 1e-3) / 1e9


def bound_host_buffer(nbytes, node):
    """Anonymous mapping bound to one NUMA node (mbind MPOL_BIND), page-locked with cudaHostRegister."""
    libc = ctypes.CDLL(None, use_errno=True)
    mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    mask = ctypes.c_ulong(1 << node)
    SYS_mbind = 237                                     # x86_64
    rc = libc.syscall(SYS_mbind, ctypes.c_void_p(addr), ctypes.c_ulong(nbytes), 2, ctypes.byref(mask), ctypes.c_ulong(64), 0)
    if rc != 0:
        return None, f"mbind(node {node}) failed: errno {ctypes.get_errno()}"
    ctypes.memset(addr, 0, nbytes)                      # first touch on the bound node
    rt = torch.cuda.cudart()
    err = rt.cudaHostRegister(addr, nbytes, 0)
    if int(err) != 0:
        return None, f"cudaHostRegister failed: {err}"
    t = torch.frombuffer(mm, dtype=torch.uint8)
    return (t, mm, addr), None


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    d = torch.empty(MB, dtype=torch.uint8, device=dev)
    h = torch.empty(MB, dtype=torch.uint8).pin_memory()
    reps = 20
    res = {"rank": rank}
    # alone: ranks take turns
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            res["d2h_alone_gbs"] = timed_copy(h, d, reps, stream)
            res["h2d_alone_gbs"] = timed_copy(d, h, reps, stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    res["d2h_concurrent_gbs"] = timed_copy(h, d, reps, stream)
    if world > 1:
        dist.barrier()
    res["h2d_concurrent_gbs"] = timed_copy(d, h, reps, stream)
    nodes = numa_nodes()
    res["numa_nodes_visible"] = nodes
    res["cpu_affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
    if len(nodes) > 1:
        per_node = {}
        for n in nodes:
            buf, err = bound_host_buffer(MB, n)
            if buf is None:
                per_node[str(n)] = err
                continue
            if world > 1:
                dist.barrier()
            per_node[str(n)] = timed_copy(buf[0], d, reps, stream)       # every rank copies into node n at once
            torch.cuda.cudart().cudaHostUnregister(buf[2])
        res["d2h_concurrent_by_numa_node_gbs"] = per_node
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    else:
        allres = [res]
    if rank == 0:
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception as e:
            topo = str(e)
        out = {"n_gpus": world, "bytes_per_copy": MB,
               "aggregate_d2h_concurrent_gbs": sum(r["d2h_concurrent_gbs"] for r in allres),
               "aggregate_h2d_concurrent_gbs": sum(r["h2d_concurrent_gbs"] for r in allres),
               "mean_d2h_alone_gbs": sum(r["d2h_alone_gbs"] for r in allres) / world,
               "ranks": allres, "topo": topo.splitlines()[:14]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
