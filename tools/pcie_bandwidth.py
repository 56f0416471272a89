"""Pinned-memory PCIe bandwidth of the box: every visible GPU alone, then all of them at once (one process, one
copy stream per device) -- the ceiling of the end-to-end record at N GPUs.  One JSON line per measurement."""
import json
import time

import torch

NBYTES = 112_500_000            # the 9 B per microbe record of a 12.5 M-microbe shard
REPS = 10


def run(devs, direction):
    bufs = []
    for d in devs:
        with torch.cuda.device(d):
            bufs.append((torch.empty(NBYTES, dtype=torch.uint8, device="cuda:%d" % d),
                         torch.empty(NBYTES, dtype=torch.uint8).pin_memory(), torch.cuda.Stream(device=d)))

    def go(reps):
        for dv, hv, st in bufs:
            with torch.cuda.stream(st):
                for _ in range(reps):
                    (hv if direction == "d2h" else dv).copy_(dv if direction == "d2h" else hv, non_blocking=True)

    def sync():
        for d in devs:
            torch.cuda.synchronize(d)

    go(2); sync()
    t0 = time.perf_counter()
    go(REPS); sync()
    dt = time.perf_counter() - t0
    return REPS * NBYTES * len(devs) / dt / 1e9


n = torch.cuda.device_count()
for direction in ("d2h", "h2d"):
    for d in range(n):
        print(json.dumps({"direction": direction, "gpus": [d], "aggregate_GBps": round(run([d], direction), 2)}), flush=True)
    k = 2
    while k <= n:
        devs = list(range(k))
        agg = run(devs, direction)
        print(json.dumps({"direction": direction, "gpus": devs, "aggregate_GBps": round(agg, 2), "per_gpu_GBps": round(agg / k, 2)}), flush=True)
        k *= 2
