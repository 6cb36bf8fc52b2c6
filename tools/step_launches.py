#!/usr/bin/env python
"""One implicit bioheat step (k(T) q-data, PA set-up fused with the Jacobi diagonal, PCG to 1e-8) at configs[1] between
cudaProfilerStart/Stop, for a per-launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/step_launches.csv python tools/step_launches.py [--factorised]

Without ncu it prints the CUDA-event time of the step and of its parts (each part timed alone, 5 repetitions)."""
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
    ap.add_argument("--factorised", action="store_true")
    a = ap.parse_args()
    import torch
    import b200pa
    import bench
    p, n = a.order, a.elems
    ctx = b200pa.Context(0)
    m = b200pa.hex_build(n, n, n, p, want=("gather_map", "elem_vertices", "vertices", "bdr_attr", "lattice"))
    bas = b200pa.basis(p)
    nd, ne = m["ndofs"], m["ne"]
    nq = ne * (p + 2) ** 3
    sp = b200pa.Space(ctx, p + 1, p + 2, ne, nd, m["gather_map"], bas["B"], bas["G"])
    sp.geometry_from_vertices(bas["W"], m["vertices"], m["elem_vertices"])
    lat = m["lattice"].reshape(-1, 3)
    xyz = (lat // p + bas["gll"][lat % p]) / float(n) if "gll" in bas else lat / float(p * n)
    T0 = ctx.to_dev(37.0 + 20.0 * np.exp(-40.0 * ((xyz - 0.5) ** 2).sum(1)))
    P = bench.PHYS
    kq = ctx.empty(nq)
    mq = ctx.coeff_eval(1, nq, P["rc"] / P["dt"] + P["wbcb"], 0.0, 0.0)
    f = b200pa.Form(sp)
    f.set_factorised(a.factorised)
    f.assemble_diffusion(sp.coeff_linear(P["k0"], P["ak"], 37.0, T0, out=kq))
    f.assemble_mass(mq)
    f.set_essential(None)
    diag = ctx.empty(nd)
    rhs = f.mult(T0)
    rhs = ctx.add(rhs, 0.01, ctx.to_dev(np.random.default_rng(0).random(nd)))
    T1 = T0.clone()

    def step():
        k2 = sp.coeff_linear(P["k0"], P["ak"], 37.0, T0, out=kq)
        f.assemble_mass(mq)
        d2 = f.jacobi_from(f.assemble_diffusion_with_diagonal(k2, diag))
        T1.copy_(T0)
        return f.pcg(d2, rhs, T1, 1e-8, 0.0, 500, want_norms=False)[0]

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.torch_stream)
        for _ in range(reps):
            fn()
        e1.record(ctx.torch_stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    step()
    res = step()
    out = {"order": p, "elems": n, "dofs": nd, "factorised": a.factorised, "pcg_iters": int(res.final_iter),
           "step_ms": timed(step),
           "coeff_linear_ms": timed(lambda: sp.coeff_linear(P["k0"], P["ak"], 37.0, T0, out=kq)),
           "assemble_mass_ms": timed(lambda: f.assemble_mass(mq)),
           "assemble_diffusion_with_diagonal_ms": timed(lambda: f.assemble_diffusion_with_diagonal(kq, diag)),
           "jacobi_ms": timed(lambda: f.jacobi_from(diag)),
           "apply_ms": timed(lambda: f.mult(T0, T1))}
    d2 = f.jacobi_from(diag)
    for its in (1, 7, 8, 9):
        def solve(its=its):
            T1.copy_(T0)
            f.pcg(d2, rhs, T1, 0.0, 0.0, its, want_norms=False)
        out[f"pcg_{its}_its_ms"] = timed(solve)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
