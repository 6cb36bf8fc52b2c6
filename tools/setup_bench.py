#!/usr/bin/env python
"""Set-up / q-point kernels of the implicit bioheat step on one GPU, each timed alone (CUDA events, 10 launches
after 2 warm-ups) against the HBM roofline of ITS OWN algorithmic bytes (what it must read + write once).

    python tools/setup_bench.py [--order 2] [--elems 100] [--skew]      -> one JSON line
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--elems", type=int, default=100)
    ap.add_argument("--skew", action="store_true", help="non-affine-free sheared mesh stays affine; this moves vertices: trilinear path")
    a = ap.parse_args()
    import torch
    import b200pa
    import bench
    p, n = a.order, a.elems
    ctx = b200pa.Context(0)
    m = b200pa.hex_build(n, n, n, p, want=("gather_map", "elem_vertices", "vertices", "bdr_attr", "lattice"))
    bas = b200pa.basis(p)
    v = m["vertices"]
    if a.skew:
        v = v.copy().reshape(-1, 3)
        v += 0.2 / n * np.sin(7.0 * v[:, [1, 2, 0]])
        v = v.ravel()
    nd, ne = m["ndofs"], m["ne"]
    D3, Q3 = (p + 1) ** 3, (p + 2) ** 3
    nq, nE = ne * Q3, ne * D3
    sp = b200pa.Space(ctx, p + 1, p + 2, ne, nd, m["gather_map"], bas["B"], bas["G"])
    sp.geometry_from_vertices(bas["W"], v, m["elem_vertices"])
    peak, _ = bench.measured_peak_hbm()
    T = ctx.to_dev(37.0 + np.random.default_rng(0).random(nd))
    kq = sp.coeff_linear(0.5, 0.02, 37.0, T)
    mqf = ctx.to_dev(3.0 + np.random.default_rng(1).random(nq))
    f = b200pa.Form(sp)
    f.assemble_diffusion(kq)
    f.assemble_mass(mqf)
    f.set_essential(None)
    ff = b200pa.Form(sp)
    if sp.affine:
        ff.set_factorised(True)
        ff.assemble_diffusion(kq)
        ff.assemble_mass(mqf)
        ff.set_essential(None)
    diag = ctx.empty(nd)
    src = ctx.empty(nq)
    out_q = ctx.empty(nq)
    g3 = None
    lf = ctx.empty(nd)
    const_m = np.array([3.6])

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.torch_stream)
        for _ in range(reps):
            fn()
        e1.record(ctx.torch_stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    idx = 4 * nE                       # one int32 index stream over the E-entries
    rows = {}

    def row(name, fn, nbytes, note=""):
        ms = timed(fn)
        rows[name] = {"ms": round(ms, 4), "GB": round(nbytes / 1e9, 3), "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / peak, 3), "note": note}

    row("coeff_linear k(T)  [gather + B-interp -> Q^3]", lambda: sp.coeff_linear(0.5, 0.02, 37.0, T, out=kq), 8 * nd + idx + 8 * nq)
    row("assemble_diffusion (stored)", lambda: f.assemble_diffusion(kq), 8 * nq + 48 * nq + (48 * ne if sp.affine else 0),
        "affine: streaming" if sp.affine else "trilinear: J rebuilt per q-point")
    row("assemble_mass (q-field)", lambda: f.assemble_mass(mqf), 8 * nq * 3)
    row("assemble_mass (constant)", lambda: f.assemble_mass(const_m), 8 * nq * 2)
    f.assemble_mass(mqf)
    row("assemble_diagonal (stored: diag kernel + segmented sum)", lambda: f.assemble_diagonal(diag), 56 * nq + idx + 16 * nE + 4 * nd + 8 * nd,
        "q-data read + slot stream + slot-order scratch write/read + offsets + diag")
    row("assemble_diffusion + assemble_diagonal (two passes)", lambda: (f.assemble_diffusion(kq), f.assemble_diagonal(diag)),
        8 * nq + 48 * nq + 56 * nq + idx + 16 * nE + 12 * nd)
    row("assemble_diffusion_with_diagonal (one pass)", lambda: f.assemble_diffusion_with_diagonal(kq, diag),
        8 * nq + 48 * nq + 8 * nq + idx + 16 * nE + 12 * nd, "coefficient + mass q-data read, q-data write, slot stream, scratch, diag")
    if sp.affine:
        row("factorised: assemble_diffusion", lambda: ff.assemble_diffusion(kq), 16 * nq)
        row("factorised: assemble_diagonal", lambda: ff.assemble_diagonal(diag), 16 * nq + 48 * ne + idx + 16 * nE + 12 * nd)
        row("factorised: assemble_diffusion_with_diagonal", lambda: ff.assemble_diffusion_with_diagonal(kq, diag), 24 * nq + 48 * ne + idx + 16 * nE + 12 * nd)
    row("jacobi_setup (dinv = 1/diag)", lambda: ctx.jacobi_setup(diag, None), 16 * nd)
    row("joule  [gather + grad + sigma|grad phi|^2 -> Q^3]", lambda: sp.joule(T, kq, 1.0, out=src), 8 * nd + idx + 16 * nq + (0 if not sp.affine else 0),
        "J rebuilt from vertices per q-point")
    row("domain_lf  [Q^3 -> slot scratch -> L]", lambda: sp.domain_lf(src, out=lf), 8 * nq + 8 * nq + idx + 16 * nE + 12 * nd, "f_q + detJ read")
    row("qvalues  [gather + B-interp -> Q^3]", lambda: sp.qvalues(T), 8 * nd + idx + 8 * nq, "includes a torch allocation of the output")
    x = ctx.to_dev(np.random.default_rng(2).random(nd))
    y = ctx.empty(nd)
    tot, _ = bench.algorithmic_bytes_per_dof(p, 7)
    row("apply (stored, whole L->L)", lambda: f.mult(x, y), tot * nd)
    print(json.dumps({"order": p, "elems": n, "dofs": nd, "affine": bool(sp.affine), "peak_gbs": peak, "rows": rows}))


if __name__ == "__main__":
    main()
