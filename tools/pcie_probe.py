"""Host <-> device copy ceiling of the box, per rank and in aggregate: pinned H2D alone, D2H alone, both at once -- with
the pinned buffers placed by the default policy and placed on the NUMA node of the rank's GPU (CPU affinity set to the
GPU's local_cpulist before the buffers are allocated and first touched).  bench.py's e2e numbers are bounded by these.

usage: python tools/pcie_probe.py                      (one GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py
Prints one JSON line per (rank, placement) and an aggregate line from rank 0."""
import json
import os
import subprocess
import time

import torch


def gpu_pci_bdf(index: int) -> str:
    out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                         capture_output=True, text=True).stdout.strip()
    return out.lower().replace("00000000:", "0000:")


def local_cpus(bdf: str):
    try:
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
    except OSError:
        return None, None
    cpus = []
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus, node


def measure(nbytes, reps=3):
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)   # first touch on the current CPU set
    h_out.fill_(2)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    run(True, True)
    return {"h2d_alone": run(True, False), "d2h_alone": run(False, True), "duplex_per_direction": run(True, True)}


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = (2 << 30) if world == 1 else (1 << 30)
    bdf = gpu_pci_bdf(local)
    cpus, node = local_cpus(bdf)
    results = {}
    for placement in ("default", "gpu_numa_node"):
        if placement == "gpu_numa_node":
            if not cpus:
                continue
            os.sched_setaffinity(0, cpus)
        r = measure(nbytes)
        r.update(rank=rank, placement=placement, gpu_bdf=bdf, numa_node=node, affinity=len(os.sched_getaffinity(0)))
        results[placement] = r
        print(json.dumps(r), flush=True)
    if world > 1:
        for placement, r in results.items():
            t = torch.tensor([r["h2d_alone"], r["d2h_alone"], r["duplex_per_direction"]], dtype=torch.float64, device="cuda")
            torch.distributed.all_reduce(t)
            if rank == 0:
                print(json.dumps({"aggregate_GBps": True, "world": world, "placement": placement, "h2d_alone": t[0].item(),
                                  "d2h_alone": t[1].item(), "duplex_per_direction": t[2].item()}), flush=True)
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        print(topo, flush=True)
        print("cpus:", os.cpu_count(), "numa nodes:", sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")))


if __name__ == "__main__":
    main()
