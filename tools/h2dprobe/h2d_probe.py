"""Host->device copy bandwidth from pinned memory, per NUMA node of the allocating thread (diagnostic for bench.py's e2e leg)."""
import glob, os, subprocess, time
import torch

def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return "ERR %r" % (e,)

print("cpus allowed:", sorted(os.sched_getaffinity(0)))
print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
print(sh("nvidia-smi topo -m"))
for f in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    v = open(f).read().strip()
    cls = open(os.path.dirname(f) + "/class").read().strip()
    if cls.startswith("0x0302") or cls.startswith("0x0300"):
        print(f, v)
for n in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    print(n, open(n).read().strip())
print(sh("cat /proc/meminfo | head -3"))

torch.cuda.set_device(0)
allowed = sorted(os.sched_getaffinity(0))
nodes = {}
for n in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    cpus = set()
    for part in open(n).read().strip().split(","):
        if "-" in part:
            a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    cpus &= set(allowed)
    if cpus:
        nodes[os.path.basename(os.path.dirname(n))] = cpus
if not nodes:
    nodes = {"all": set(allowed)}
GB = 8
dev = torch.empty(GB << 30, dtype=torch.uint8, device="cuda")
for name, cpus in nodes.items():
    os.sched_setaffinity(0, cpus)
    t0 = time.perf_counter()
    host = torch.empty(GB << 30, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    t_alloc = time.perf_counter() - t0
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("node %s cpus %s: pin+touch %.2fs, H2D %d GB in %.3f s = %.1f GB/s" % (name, sorted(cpus)[:4], t_alloc, GB, dt, GB * 1.073741824 / dt))
    # two streams, halves
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = (GB << 30) // 2
    with torch.cuda.stream(s1):
        dev[:h].copy_(host[:h], non_blocking=True)
    with torch.cuda.stream(s2):
        dev[h:].copy_(host[h:], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("  two streams: %.1f GB/s" % (GB * 1.073741824 / dt))
    del host
    os.sched_setaffinity(0, set(allowed))
