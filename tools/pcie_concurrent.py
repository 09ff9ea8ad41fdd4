"""Host-to-device bandwidth of 1 / 2 / 4 / 8 GPUs copying at the same time (the ceiling of bench.py's e2e number at N > 1).

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_concurrent.py

Every rank pins a 185 MB frame buffer (one bench step) and a 32 MB result buffer and copies H2D (+ D2H on a second
stream) to / from its own GPU; phases with N = 1, 2, 4, 8 active ranks are separated by barriers, the other ranks idle.
Each phase runs twice: pinned memory wherever the process happens to run ("default"), and pinned memory bound to the
NUMA node the GPU hangs off (set_mempolicy(MPOL_BIND) before cudaHostAlloc, the thread pinned to that node's CPUs when
the cpuset allows it).  Rank 0 prints the topology it can see and one JSON line per phase."""
import ctypes
import glob
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gpu_numa_node(idx):
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(idx), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True,
                             text=True).stdout.strip().lower()
        bus = bus[4:] if bus.startswith("0000") and len(bus) > 12 else bus   # 00000000:1B:00.0 -> 0000:1b:00.0
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
    except (OSError, ValueError):
        return -1


def node_cpus(node):
    try:
        out = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except OSError:
        return []


def bind_memory(node):
    """set_mempolicy(MPOL_BIND, {node}): pages pinned from here on come from that node.  Returns an error string or None."""
    if node < 0:
        return "GPU reports no NUMA node"
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(238, 2, ctypes.byref(mask), ctypes.c_ulong(64))  # __NR_set_mempolicy (x86-64), MPOL_BIND
    return None if rc == 0 else "set_mempolicy: errno %d" % ctypes.get_errno()


def unbind_memory():
    ctypes.CDLL(None).syscall(238, 0, None, ctypes.c_ulong(0))  # MPOL_DEFAULT


def measure(active, label):
    n_in, n_out = 512 * 752 * 480, 32 * 1024 * 1024
    h = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n_in, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(n_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for both in (False, True):
        barrier()
        dt = 0.0
        if rank < active:
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    with torch.cuda.stream(s1):
                        d.copy_(h, non_blocking=True)
                    if both:
                        with torch.cuda.stream(s2):
                            h2.copy_(d2, non_blocking=True)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 10
        barrier()
        t = torch.tensor([n_in / dt / 1e9 if dt else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            tl = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(tl, t)
            vals = [float(x.item()) for x in tl][:active]
        else:
            vals = [float(t.item())]
        res["h2d_with_d2h" if both else "h2d_alone"] = {"per_gpu_gbs": [round(v, 1) for v in vals], "total_gbs": round(sum(vals), 1)}
    if rank == 0:
        print(json.dumps({"active_gpus": active, "pinned_memory": label, **res}), flush=True)
    del h, d, h2, d2


node = gpu_numa_node(local)
if rank == 0:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout, flush=True)
    print("NUMA nodes visible:", sorted(glob.glob("/sys/devices/system/node/node[0-9]*")), flush=True)
    print("allowed CPUs of this process:", sorted(os.sched_getaffinity(0)), flush=True)
    try:
        print("cpuset.mems:", open("/sys/fs/cgroup/cpuset.mems.effective").read().strip(), flush=True)
    except OSError:
        pass
info = torch.tensor([node], dtype=torch.int64, device=dev)
if world > 1:
    il = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(il, info)
    if rank == 0:
        print("NUMA node of GPU 0..%d: %s" % (world - 1, [int(x.item()) for x in il]), flush=True)

sizes = [n for n in (1, 2, 4, 8) if n <= world]
for n in sizes:
    measure(n, "default")
err = bind_memory(node)
cpus = [c for c in node_cpus(node) if c in os.sched_getaffinity(0)]
if cpus:
    os.sched_setaffinity(0, cpus)
if rank == 0:
    print("binding pinned memory to the GPU's NUMA node: %s; CPU affinity narrowed: %s" % (err or "ok", bool(cpus)), flush=True)
for n in sizes:
    measure(n, "bound to the GPU's NUMA node" if not err else "bind failed (%s)" % err)
unbind_memory()
if world > 1:
    dist.destroy_process_group()
