"""Measures pinned-memory PCIe copy rates on the box (one stream, and two concurrent streams):
the floor of bench.py's e2e arm is result bytes / this D2H rate.    python profiles/pcie_probe.py"""
import torch


def rate(mb, direction, streams):
    n = (mb << 20) // streams
    d = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(streams)]
    h = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for k, s in enumerate(ss):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                if direction == "d2h":
                    h[k].copy_(d[k], non_blocking=True)
                else:
                    d[k].copy_(h[k], non_blocking=True)
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n * streams / best / 1e6


if __name__ == "__main__":
    for mb in (16, 128, 1024):
        for direction in ("d2h", "h2d"):
            for streams in (1, 2, 4):
                print(f"{direction} {mb:5d} MB x{streams} stream(s): {rate(mb, direction, streams):6.1f} GB/s")
