"""Concurrent pinned D2H / H2D rates of the box's GPUs: what bounds the end-to-end call at N > 1 (every GPU
returns its CSR slab to host memory at the same time).

    python profiles/pcie_probe_multi.py --gpus 8                      # ONE process driving all GPUs
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        profiles/pcie_probe_multi.py --per-process                     # one process per GPU

Prints one JSON line: per-GPU and aggregate GB/s with k = 1, 2, 4, ... GPUs copying at once, for D2H into
separate pinned buffers, D2H into slices of ONE pinned buffer (the single host CSR of
c2b_visibility_graph_multi) and H2D.
"""
import argparse
import json
import os
import time

import torch

MB = 1 << 20


def one_process(G, nbytes, reps):
    devs = [torch.device("cuda", g) for g in range(G)]
    dbuf = [torch.empty(nbytes, dtype=torch.uint8, device=d).fill_(g + 1) for g, d in enumerate(devs)]
    hsep = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(G)]
    hone = torch.empty(nbytes * G, dtype=torch.uint8).pin_memory()
    streams = [torch.cuda.Stream(device=d) for d in devs]
    out = {}

    def run(kind, k):
        def go():
            for g in range(k):
                with torch.cuda.stream(streams[g]):
                    if kind == "d2h_separate":
                        hsep[g].copy_(dbuf[g], non_blocking=True)
                    elif kind == "d2h_one_buffer":
                        hone[g * nbytes:(g + 1) * nbytes].copy_(dbuf[g], non_blocking=True)
                    else:
                        dbuf[g].copy_(hsep[g], non_blocking=True)
            for g in range(k):
                streams[g].synchronize()
        go()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            go()
            best = min(best, time.perf_counter() - t0)
        return k * nbytes / best / 1e9

    for kind in ("d2h_separate", "d2h_one_buffer", "h2d"):
        out[kind] = {}
        k = 1
        while k <= G:
            agg = run(kind, k)
            out[kind][str(k)] = {"aggregate_GBps": round(agg, 1), "per_gpu_GBps": round(agg / k, 1)}
            k *= 2
    return out


def per_process(nbytes, reps):
    import torch.distributed as dist
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    world, rank = dist.get_world_size(), dist.get_rank()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda").fill_(1)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    res = {}
    for kind in ("d2h", "h2d"):
        best = 1e9
        for r in range(reps + 1):
            dist.barrier()
            t0 = time.perf_counter()
            if kind == "d2h":
                h.copy_(d, non_blocking=True)
            else:
                d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if r:
                best = min(best, float(dt))
        res[kind] = {"aggregate_GBps": round(world * nbytes / best / 1e9, 1), "per_gpu_GBps": round(nbytes / best / 1e9, 1)}
    if rank == 0:
        print(json.dumps({"mode": f"{world} processes, one GPU each", "bytes_per_gpu": nbytes, **res}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--per-process", action="store_true")
    a = ap.parse_args()
    if a.per_process:
        return per_process(a.mb * MB, a.reps)
    print(json.dumps({"mode": f"one process, {a.gpus} GPUs", "bytes_per_gpu": a.mb * MB, **one_process(a.gpus, a.mb * MB, a.reps)}))


if __name__ == "__main__":
    main()
