#!/usr/bin/env python
"""Host <-> device copy rates of the box (pinned memory, CUDA events): one direction at a time and both at once.
The host-buffer apply moves 8 B/dof each way; its floor is max(one-way time, both-way time under contention).

    python tools/pcie_probe.py [--mb 65]      -> one JSON line
"""
import argparse
import json

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=64.96)
    a = ap.parse_args()
    n = int(a.mb * 1e6 / 8)
    hx, hy = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
    hx.fill_(1.0)
    dx, dy = torch.empty(n, dtype=torch.float64, device="cuda"), torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def h2d():
        with torch.cuda.stream(s1):
            dx.copy_(hx, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)

    def d2h():
        with torch.cuda.stream(s2):
            hy.copy_(dy, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s2)

    def both():
        with torch.cuda.stream(s1):
            dx.copy_(hx, non_blocking=True)
        with torch.cuda.stream(s2):
            hy.copy_(dy, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    t_up, t_down, t_both = timed(h2d), timed(d2h), timed(both)
    gb = 8 * n / 1e9
    print(json.dumps({"mb_each_way": 8 * n / 1e6, "h2d_ms": t_up, "h2d_gbs": gb / t_up * 1e3, "d2h_ms": t_down, "d2h_gbs": gb / t_down * 1e3,
                      "both_at_once_ms": t_both, "both_total_gbs": 2 * gb / t_both * 1e3,
                      "floor_ms_of_a_host_buffer_apply": t_both,
                      "gdof_per_s_at_that_floor": n / (t_both * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
