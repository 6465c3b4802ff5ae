"""NVLink baselines, single process driving every visible GPU (no NCCL, no libdfft):
  (a) unidirectional cudaMemcpyPeerAsync 0 -> 1,
  (b) every GPU copies one block to every other GPU at once (the all-to-all-v traffic pattern of the slab
      transposition), per-GPU egress GB/s = bytes sent / time (max over GPUs).
Writes gpurun_out/peer_bw_<N>.json.  These are the measured ceilings quoted beside the exchange passes."""
import json
import os
import sys

import torch


def main():
    n = torch.cuda.device_count()
    if n < 2:
        print("needs >= 2 GPUs")
        return
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nbytes = mib << 20
    src = [[torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}") for _ in range(n)] for i in range(n)]
    dst = [[torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{j}") for _ in range(n)] for j in range(n)]  # dst[j][i]: block from i on j
    streams = [[torch.cuda.Stream(device=i) for _ in range(n)] for i in range(n)]
    res = {"gpus": n, "block_MiB": mib}

    def sync():
        for i in range(n):
            torch.cuda.synchronize(i)

    # (a) one pair, one direction
    for _ in range(3):
        dst[1][0].copy_(src[0][1], non_blocking=True)
    sync()
    with torch.cuda.device(0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst[1][0].copy_(src[0][1], non_blocking=True)
        e1.record()
    sync()
    res["pair_unidirectional_gbs"] = 10 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # (b) all-to-all: GPU i sends block q to GPU q, all at once, each transfer on its own stream of the sender
    def a2a():
        for i in range(n):
            with torch.cuda.device(i):
                for q in range(n):
                    if q != i:
                        with torch.cuda.stream(streams[i][q]):
                            dst[q][i].copy_(src[i][q], non_blocking=True)
    for _ in range(3):
        a2a()
    sync()
    evs = []
    for i in range(n):
        with torch.cuda.device(i):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(torch.cuda.current_stream(i))
            for q in range(n):
                if q != i:
                    streams[i][q].wait_event(a)
            evs.append((a, b))
    reps = 5
    for _ in range(reps):
        a2a()
    for i in range(n):
        with torch.cuda.device(i):
            for q in range(n):
                if q != i:
                    torch.cuda.current_stream(i).wait_stream(streams[i][q])
            evs[i][1].record(torch.cuda.current_stream(i))
    sync()
    ms = max(a.elapsed_time(b) for a, b in evs) / reps
    res["all_to_all_ms"] = ms
    res["all_to_all_egress_gbs_per_gpu"] = (n - 1) * nbytes / (ms * 1e-3) / 1e9
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/peer_bw_{n}.json", "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
