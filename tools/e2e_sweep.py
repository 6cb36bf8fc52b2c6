#!/usr/bin/env python
"""b200pa_form_mult_host on configs[1] for several pipeline plans (element chunks x dofs per tile) and the serial route:
ms per call with pinned host vectors (CUDA events around 10 calls).   python tools/e2e_sweep.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))
import torch  # noqa: E402

import b200pa  # noqa: E402

p, n = 2, 100
ctx = b200pa.Context(0)
m = b200pa.hex_build(n, n, n, p, want=("gather_map", "elem_vertices", "vertices"))
bas = b200pa.basis(p)
nq = m["ne"] * (p + 2) ** 3
kq = 0.5 + np.random.default_rng(0).random(nq)
xh = np.random.default_rng(1).random(m["ndofs"])
xp = torch.from_numpy(xh).pin_memory()
yp = torch.empty(m["ndofs"], dtype=torch.float64).pin_memory()
rows = []
for C, TS in [(0, 0), (4, 65536), (8, 32768), (8, 131072), (16, 32768), (16, 65536), (16, 131072), (16, 262144), (32, 32768), (32, 131072), (64, 65536)]:
    if C:
        os.environ["B200PA_PIPE_CHUNKS"], os.environ["B200PA_PIPE_TILE"] = str(C), str(TS)
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], bas["B"], bas["G"])
    sp.geometry_from_vertices(bas["W"], m["vertices"], m["elem_vertices"])
    f = b200pa.Form(sp)
    f.assemble_diffusion(kq)
    f.assemble_mass(np.array([3.6]))
    f.set_essential(None)
    if C == 0:
        # the serial route for comparison: a communicator-free form below the size threshold cannot be forced, so time the
        # three steps by hand (H2D, device apply, D2H on one stream) through torch
        x = ctx.empty(m["ndofs"]); y = ctx.empty(m["ndofs"])
        def call():
            x.copy_(xp, non_blocking=True); f.mult(x, y); yp.copy_(y, non_blocking=True)
    else:
        def call():
            f.mult_host(xp, yp)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.torch_stream)
    for _ in range(10):
        call()
    e1.record(ctx.torch_stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows.append({"chunks": C, "tile_dofs": TS, "ms": round(ms, 4), "gdof_per_s": round(m["ndofs"] / ms / 1e6, 3)})
    f.close(); sp.close()
print(json.dumps(rows))
