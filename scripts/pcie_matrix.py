#!/usr/bin/env python
"""Host <-> device copy bandwidth of every GPU of the box, alone, in pairs and all together: which link is the
end-to-end ceiling of the frame pipeline (bench.py `e2e`)?

    python scripts/pcie_matrix.py [--mb 256] [--iters 12] [--out profiles/r02_pcie_matrix.json]

One process per GPU (like bench.py's ranks), page-locked host buffers, copies issued with cudaMemcpyAsync semantics
(torch non_blocking copies on two streams), timed with CUDA events inside each process, all active processes released
together by a barrier. Rounds:
  alone       every GPU by itself: H2D, D2H, both directions at once
  pair        every pair of GPUs at once, both directions (shared root port / switch uplink shows as a drop
              against the two `alone` figures)
  all         every GPU at once: H2D, D2H, both
  all_bound   the same with every process first pinned to the CPUs NVML reports as local to its GPU (NUMA placement
              of the page-locked buffers by first touch)
  threads     one PROCESS driving all GPUs from threads (one context each), both directions: is the ceiling a
              per-process effect?
Prints one JSON document (and writes it to --out)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.multiprocessing as mp


def local_cpus(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        return sorted(cpus & os.sched_getaffinity(0))
    except Exception:
        return []


class Copier:
    def __init__(self, dev, mb, touch=True):
        torch.cuda.set_device(dev)
        n = mb << 20
        self.h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        self.h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        if touch:
            self.h_in.fill_(1); self.h_out.fill_(2)
        self.d_in = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}")
        self.d_out = torch.empty(n, dtype=torch.uint8, device=f"cuda:{dev}")
        self.up, self.down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.dev, self.bytes = dev, n

    def run(self, direction, iters):
        """GB/s per direction (h2d, d2h); a direction not exercised reports 0."""
        torch.cuda.set_device(self.dev)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(2)]
        for k in range(iters + 1):  # first iteration untimed
            if k == 1:
                ev[0][0].record(self.up); ev[1][0].record(self.down)
            if direction in ("h2d", "both"):
                with torch.cuda.stream(self.up):
                    self.d_in.copy_(self.h_in, non_blocking=True)
            if direction in ("d2h", "both"):
                with torch.cuda.stream(self.down):
                    self.h_out.copy_(self.d_out, non_blocking=True)
        ev[0][1].record(self.up); ev[1][1].record(self.down)
        torch.cuda.synchronize(self.dev)
        gb = self.bytes * iters / 1e9
        up = gb / (ev[0][0].elapsed_time(ev[0][1]) * 1e-3) if direction in ("h2d", "both") else 0.0
        dn = gb / (ev[1][0].elapsed_time(ev[1][1]) * 1e-3) if direction in ("d2h", "both") else 0.0
        return up, dn


def worker(rank, world, mb, iters, rounds, barrier, results, bind):
    if bind:
        cpus = local_cpus(rank)
        if cpus:
            os.sched_setaffinity(0, cpus)
    c = Copier(rank, mb)
    for i, (active, direction) in enumerate(rounds):
        barrier.wait()
        if rank in active:
            up, dn = c.run(direction, iters)
            results[(i * world + rank) * 2] = up
            results[(i * world + rank) * 2 + 1] = dn
        barrier.wait()


def run_rounds(world, mb, iters, rounds, bind):
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(world)
    results = ctx.Array("d", len(rounds) * world * 2)
    procs = [ctx.Process(target=worker, args=(r, world, mb, iters, rounds, barrier, results, bind)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
        assert p.exitcode == 0
    out = []
    for i, (active, direction) in enumerate(rounds):
        out.append({"gpus": list(active), "direction": direction,
                    "h2d_gbs": [round(results[(i * world + r) * 2], 1) for r in active],
                    "d2h_gbs": [round(results[(i * world + r) * 2 + 1], 1) for r in active]})
    return out


def threads_round(world, mb, iters):
    import threading
    copiers = [Copier(d, mb) for d in range(world)]
    res = [None] * world
    gate = threading.Barrier(world)

    def go(d):
        gate.wait()
        res[d] = copiers[d].run("both", iters)
    th = [threading.Thread(target=go, args=(d,)) for d in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return {"gpus": list(range(world)), "direction": "both", "h2d_gbs": [round(r[0], 1) for r in res], "d2h_gbs": [round(r[1], 1) for r in res]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    world = torch.cuda.device_count()
    alone = [((r,), d) for r in range(world) for d in ("h2d", "d2h", "both")]
    pairs = [((i, j), "both") for i in range(world) for j in range(i + 1, world)]
    everyone = [(tuple(range(world)), d) for d in ("h2d", "d2h", "both")]
    t0 = time.time()
    doc = {"gpus": world, "buffer_mb": a.mb, "iters": a.iters, "cpus": len(os.sched_getaffinity(0)),
           "local_cpus_per_gpu": [len(local_cpus(r)) for r in range(world)],
           "numa_nodes": sorted(n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")) if os.path.isdir("/sys/devices/system/node") else None}
    rounds = run_rounds(world, a.mb, a.iters, alone + pairs + everyone, bind=False)
    doc["alone"] = rounds[: len(alone)]
    doc["pair"] = rounds[len(alone): len(alone) + len(pairs)]
    doc["all"] = rounds[len(alone) + len(pairs):]
    doc["all_bound"] = run_rounds(world, a.mb, a.iters, everyone, bind=True)
    doc["threads"] = threads_round(world, a.mb, a.iters)

    def total(r):
        return round(sum(r["h2d_gbs"]) + sum(r["d2h_gbs"]), 1)
    doc["summary"] = {
        "alone_both_sum_of_directions_gbs": [total(r) for r in doc["alone"] if r["direction"] == "both"],
        "pair_both_sum_gbs": {f"{r['gpus'][0]}-{r['gpus'][1]}": total(r) for r in doc["pair"]},
        "all_h2d_total_gbs": round(sum(doc["all"][0]["h2d_gbs"]), 1), "all_d2h_total_gbs": round(sum(doc["all"][1]["d2h_gbs"]), 1),
        "all_both_total_gbs": total(doc["all"][2]), "all_bound_both_total_gbs": total(doc["all_bound"][2]),
        "threads_both_total_gbs": total(doc["threads"]),
    }
    doc["seconds"] = round(time.time() - t0, 1)
    text = json.dumps(doc)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    sys.exit(main())
