#!/usr/bin/env python
"""End to end over every GPU of the box from ONE process: crispy_ns_multi_process_streams_host (MultiDenoiser) on
pinned host buffers, f32 and PCM16 on the host link, next to the box's aggregate host<->device copy ceiling measured
the same way (every device copying both ways at once from this process).  Prints one JSON line.
usage: python scripts/multi_e2e.py [streams_per_gpu] [seconds]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import crispy_b200 as cb  # noqa: E402
from crispy_b200.synth import synth_chunk  # noqa: E402

per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ndev = torch.cuda.device_count()
n = per_gpu * ndev
ns = seconds * 48000


def copy_ceiling(nbytes=1 << 30, reps=4):
    """GB/s per direction with every device copying H2D and D2H at once (pinned memory, one stream per direction)."""
    bufs = []
    for d in range(ndev):
        with torch.cuda.device(d):
            bufs.append((torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                         torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"),
                         torch.cuda.Stream(d), torch.cuda.Stream(d)))
    out = {}
    for name, h2d, d2h in (("h2d_alone", True, False), ("d2h_alone", False, True), ("both", True, True)):
        for d in range(ndev):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(reps):
            for d, (hi, ho, di, do, s1, s2) in enumerate(bufs):
                if h2d:
                    with torch.cuda.stream(s1):
                        di.copy_(hi, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        ho.copy_(do, non_blocking=True)
        for d in range(ndev):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        out[name + "_gbs_per_direction_all_gpus"] = ndev * nbytes * reps / dt / 1e9
    return out


ceiling = copy_ceiling()
x = torch.empty((n, ns), dtype=torch.float32).pin_memory()
for d in range(ndev):  # the synthetic workload, generated on the devices block by block
    blk = slice(d * per_gpu, (d + 1) * per_gpu)
    for s0 in range(0, ns, 480000):
        s1 = min(ns, s0 + 480000)
        x[blk, s0:s1].copy_(synth_chunk(per_gpu, s1 - s0, first_stream=d * per_gpu, start_sample=s0, device=f"cuda:{d}"))
out = torch.empty_like(x).pin_memory()
vad = torch.empty((n, ns // 480), dtype=torch.float32).pin_memory()
den = cb.MultiDenoiser(n)
res = {"n_gpus": ndev, "streams": n, "seconds_per_stream": seconds, "partition": den.ranges, "copy_ceiling": ceiling}


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        den.reset()
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


t = timed(lambda: den.process_streams_host(x, unit_scale=True, out=out, vad=vad))
res["e2e_f32"] = {"stream_seconds_per_s": n * seconds / t, "wall_s": t, "bytes_each_way": x.numel() * 4,
                  "gbs_each_way": x.numel() * 4 / t / 1e9}
res["e2e_f32"]["fraction_of_copy_ceiling"] = res["e2e_f32"]["gbs_each_way"] / ceiling["both_gbs_per_direction_all_gpus"]
xi = (x * 32767.0).round().clamp(-32768, 32767).to(torch.int16).pin_memory()
oi = torch.empty((n, ns), dtype=torch.int16).pin_memory()
t = timed(lambda: den.process_streams_host(xi, unit_scale=True, out=oi, vad=vad, out_i16=True))
res["e2e_pcm16"] = {"stream_seconds_per_s": n * seconds / t, "wall_s": t, "bytes_each_way": xi.numel() * 2,
                    "gbs_each_way": xi.numel() * 2 / t / 1e9}
res["e2e_pcm16"]["fraction_of_copy_ceiling"] = res["e2e_pcm16"]["gbs_each_way"] / ceiling["both_gbs_per_direction_all_gpus"]
print(json.dumps(res))
