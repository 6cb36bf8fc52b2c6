#!/usr/bin/env python
"""a few launches of every set-up / q-point kernel of the implicit step at configs[1] size, for ncu captures:
    ncu --set full --import-source on -k regex:"k_diag_sf|pa_element_kernel" -c 12 -o out python tools/prof_setup.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))
import b200pa  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctx = b200pa.Context(0)
m = b200pa.hex_build(n, n, n, p, want=("gather_map", "elem_vertices", "vertices", "bdr_attr", "lattice"))
bas = b200pa.basis(p)
sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], bas["B"], bas["G"])
sp.geometry_from_vertices(bas["W"], m["vertices"], m["elem_vertices"])
nq = m["ne"] * (p + 2) ** 3
T = ctx.to_dev(37.0 + np.random.default_rng(0).random(m["ndofs"]))
kq = sp.coeff_linear(0.5, 0.02, 37.0, T)
f = b200pa.Form(sp)
f.assemble_diffusion(kq)
f.assemble_mass(np.array([3.6]))
f.set_essential(None)
diag = ctx.empty(m["ndofs"])
src = ctx.empty(nq)
for _ in range(2):
    f.assemble_diagonal(diag)
    f.assemble_diffusion_with_diagonal(kq, diag)
    sp.coeff_linear(0.5, 0.02, 37.0, T, out=kq)
    sp.joule(T, kq, 1.0, out=src)
    sp.domain_lf(src, out=diag)
ctx.sync()
print("ok")
