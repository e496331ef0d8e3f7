"""PCIe ceiling for the host-buffer entry points (bench.py `e2e`): pinned H2D alone, D2H alone, both at once (two
streams), for the transfer size of one L-vector of the headline workload (139 MB) and for the chunk size the pipelined
lpf_apply_T_host uses (1/32 of it).  Prints GB/s per direction."""
import sys
import time

import torch


def bw(n_bytes, chunks, mode, reps=10):
    dev = torch.device("cuda:0")
    h_in = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    cs = n_bytes // chunks

    def run():
        for i in range(chunks):
            sl = slice(i * cs, (i + 1) * cs)
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)

    run(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n_bytes / dt / 1e9, dt * 1e3


if __name__ == "__main__":
    n = 17369088 * 8
    for chunks in (1, 32):
        for mode in ("h2d", "d2h", "both"):
            g, ms = bw(n, chunks, mode)
            print(f"{mode:5s} {n / 1e6:.0f} MB in {chunks:2d} chunk(s): {g:6.1f} GB/s per direction, {ms:.3f} ms", flush=True)
