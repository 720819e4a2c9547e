"""Diagnostic (not a test): host<->device copy bandwidth of all ranks at once, for several placements of the pinned
buffers.  torchrun --nproc-per-node N tools/h2d_scaling.py"""
import ctypes, os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(local)
pci = nv.nvmlDeviceGetPciInfo(h).busId
pci = pci.decode() if isinstance(pci, bytes) else pci
def sysfs(p):
    try: return open(p).read().strip()
    except Exception as e: return "?"
busid = pci.lower()[4:] if len(pci) > 12 else pci.lower()
numa = sysfs("/sys/bus/pci/devices/%s/numa_node" % busid)
cpul = sysfs("/sys/bus/pci/devices/%s/local_cpulist" % busid)
allowed = sorted(os.sched_getaffinity(0))
if rank == 0:
    print("allowed cpus: %d (%d..%d)  cpuset.mems: %s  nodes: %s" % (len(allowed), allowed[0], allowed[-1], sysfs("/sys/fs/cgroup/cpuset.mems.effective"),
          [ (n, sysfs("/sys/devices/system/node/%s/cpulist" % n)) for n in sorted(os.listdir("/sys/devices/system/node")) if n.startswith("node")]), flush=True)
for r in range(world):
    dist.barrier()
    if r == rank: print("rank %d gpu %s numa_node %s local_cpulist %s" % (rank, pci, numa, cpul), flush=True)
libnuma = None
try:
    libnuma = ctypes.CDLL("libnuma.so.1")
except OSError:
    pass
if rank == 0: print("libnuma:", bool(libnuma), flush=True)

def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part.strip().isdigit(): out.append(int(part))
    return out

def bench(tag, cpus):
    if cpus:
        cp = [c for c in cpus if c in allowed]
        if cp: os.sched_setaffinity(0, cp)
    n = 46080000 // 4
    hx = [torch.empty(n).pin_memory() for _ in range(2)]
    for t in hx: t.normal_()            # first touch under the affinity
    ho = torch.empty(4980736 // 4).pin_memory(); ho.zero_()
    dx = torch.empty(n, device=dev); do = torch.empty(4980736 // 4, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3): dx.copy_(hx[0], non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i in range(40):
        with torch.cuda.stream(s1): dx.copy_(hx[i % 2], non_blocking=True)
        with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = 40 * (46080000 + 4980736) / dt / 1e9
    t = torch.tensor([gbs], device=dev); lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    if rank == 0: print("%-28s per-rank GB/s: %s  total %.1f" % (tag, " ".join("%.1f" % v.item() for v in lst), sum(v.item() for v in lst)), flush=True)
    os.sched_setaffinity(0, allowed)
    del hx, ho

bench("default", None)
bench("gpu local_cpulist", parse_cpulist(cpul) if cpul != "?" else None)
nodes = sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node"))
nodecpus = [parse_cpulist(sysfs("/sys/devices/system/node/%s/cpulist" % n)) for n in nodes]
bench("spread rank %% nodes", nodecpus[rank % len(nodecpus)])
bench("all on last node", nodecpus[-1])
dist.destroy_process_group()
