#!/usr/bin/env python
"""Timeline of the pipelined host-buffer apply at configs[1] (B200PA_PIPE_TRACE=1: per chunk, when its x tiles were up, its
kernels done, its y tiles down), for a few plans.   python tools/e2e_trace.py 2> trace.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))
os.environ["B200PA_PIPE_TRACE"] = "1"
import torch  # noqa: E402

import b200pa  # noqa: E402

p, n = 2, 100
ctx = b200pa.Context(0)
m = b200pa.hex_build(n, n, n, p, want=("gather_map", "elem_vertices", "vertices"))
bas = b200pa.basis(p)
nq = m["ne"] * (p + 2) ** 3
kq = 0.5 + np.random.default_rng(0).random(nq)
xp = ctx.pinned(m["ndofs"], fill=np.random.default_rng(1).random(m["ndofs"]))
yp = ctx.pinned(m["ndofs"])
for C, TS in [(8, 32768), (16, 32768), (4, 65536)]:
    os.environ["B200PA_PIPE_CHUNKS"], os.environ["B200PA_PIPE_TILE"] = str(C), str(TS)
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], bas["B"], bas["G"])
    sp.geometry_from_vertices(bas["W"], m["vertices"], m["elem_vertices"])
    f = b200pa.Form(sp)
    f.assemble_diffusion(kq)
    f.assemble_mass(np.array([3.6]))
    f.set_essential(None)
    print(f"--- chunks {C}, tile {TS}", file=sys.stderr, flush=True)
    for _ in range(4):
        f.mult_host(xp, yp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.torch_stream)
    for _ in range(5):
        f.mult_host(xp, yp)
    e1.record(ctx.torch_stream)
    torch.cuda.synchronize()
    print(f"    {e0.elapsed_time(e1) / 5:.3f} ms per call (with the trace's timed events)", file=sys.stderr, flush=True)
    f.close()
    sp.close()
