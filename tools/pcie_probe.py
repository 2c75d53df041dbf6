"""PCIe ceiling of the box: pinned host<->device copy bandwidth, one direction and both at once
(the bound of bench.py's e2e numbers).  usage: python tools/pcie_probe.py"""
import torch, time
n = 1 << 30  # bytes... use 2 GiB buffers
h_in = torch.empty(2 << 30, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(2 << 30, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
d_out = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return 3 * (2 << 30) / dt / 1e9
run(True)
print("H2D alone GB/s", run(False)); print("H2D with concurrent D2H, per direction GB/s", run(True))
