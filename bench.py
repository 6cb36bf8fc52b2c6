#!/usr/bin/env python
"""bench.py — GDOF/s of the FP64 partial-assembly diffusion+mass apply (L->L, A.Mult semantics incl.
gather and scatter; ≙ BK3/BK1PARTIAL of the reference's tests/benchmarks/bench_assembly_levels.cpp:258-317)
on BASELINE.json configs[1]: the Pennes-bioheat operator k(T) grad + (rho c/dt + perfusion) mass on a
synthetic hex slab, order 2, N=100 per GPU (8,120,601 dofs per GPU), plus the PCG-iteration and
implicit-step times of the same configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--order P] [--n N]

One JSON line on rank 0.  A "step" is one operator apply.  N>1 (torchrun, one rank per GPU): the
global mesh is a PX x PY x PZ grid of N^3-element boxes (weak scaling), every apply includes the
shared-dof exchange over NCCL; value = global true dofs x K / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))

METRIC = "GDOF/s of FP64 PA diffusion+mass apply"
PHYS = dict(dt=0.5, rc=3.6e6, wbcb=4.0e4, Ta=37.0, k0=0.5, ak=0.02, s0=0.3, as_=0.015, V=30.0)


def algorithmic_bytes_per_dof(p, ncomp=7, factorised=False):
    """x_L read + y_L write + q-data + gather map + scatter indices + offsets.
    factorised q-data (affine meshes): the six diffusion components per q-point become one scalar per q-point
    and six doubles per element"""
    D, Q = p + 1, p + 2
    rq, rd = Q ** 3 / p ** 3, D ** 3 / p ** 3
    qbytes = 8 * ncomp * rq if not factorised else 8 * (ncomp - 5) * rq + 48 / p ** 3
    total = 8 + 8 + qbytes + 4 * rd + 4 * rd + 4
    elem = total - 8 - 4          # the element kernel's share: everything but the y_L write and the offsets
    return total, elem


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 10.0:   # nvidia-smi takes ~1 s to print its first sample
                time.sleep(0.02)
            self.rows.clear()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(p, n, ops):
    """per-launch DRAM traffic of the dominant kernel from the committed ncu capture of this workload"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[f"p{p}_n{n}_{ops}"]["pa_apply_kernel"]
    except Exception:
        return None


def ref_driver_path():
    p = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    return p if os.path.exists(p) else None


def run_reference_cpu(p, n, reps, warm):
    """the UNMODIFIED reference (oracle/_ref/ref_driver = stock MFEM sources + a driver) on all host
    cores (OpenMP device), same mesh / operator / input vector; returns its JSON record"""
    drv = ref_driver_path()
    if drv is None:
        return None
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close")
    out = subprocess.run([drv, "time_apply", str(p), str(n), str(reps), str(warm), "omp"], env=env, capture_output=True,
                         text=True, timeout=1500)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            r = json.loads(line)
            r["cores"] = cores
            return r
    raise RuntimeError("ref_driver failed: " + out.stderr[-400:])


def run_reference_bioheat(p, n, iters):
    """the reference's own composition of the coupled RF + bioheat step (oracle/ref_driver.cpp `bioheat`)"""
    drv = ref_driver_path()
    if drv is None:
        return None
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close")
    out = subprocess.run([drv, "time_bioheat", str(p), str(n), str(iters), "omp"], env=env, capture_output=True, text=True,
                         timeout=1500)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            r = json.loads(line)
            r["cores"] = cores
            return r
    raise RuntimeError("ref_driver time_bioheat failed: " + out.stderr[-400:])


def oracle_port_cpu(p, n):
    """fallback CPU baseline when oracle/_ref did not travel: the C restatement on a bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import b200pa
    import orc
    m = b200pa.hex_build(n, n, n, p)
    b = b200pa.basis(p)
    nq = m["ne"] * (p + 2) ** 3
    rng = np.random.default_rng(1)
    op = orc.Operator(p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"], rng.random(6 * nq), rng.random(nq))
    x = rng.random(m["ndofs"])
    op.mult(x)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        op.mult(x)
    dt = (time.perf_counter() - t0) / reps
    return {"ndofs": m["ndofs"], "t_apply_mean": dt, "N": n, "cores": 1}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    p, n = args.order, args.n
    r = run_reference_cpu(p, n, args.steps, args.warmup)
    if r is not None:
        kind, sample = "reference", f"full workload: p={p}, N={n}, {r['ndofs']} dofs, {args.steps} applies, OpenMP device"
    else:
        n_s = min(n, 40)
        r = oracle_port_cpu(p, n_s)
        kind, sample = "port", f"oracle port, p={p}, N={n_s} ({r['ndofs']} dofs), 3 applies, 1 thread"
    v = r["ndofs"] / r["t_apply_mean"] / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["t_apply_mean"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"bioheat PA diffusion+mass apply, hex N={n}^3, order {p}, {r['ndofs']} dofs (CPU, one node)"},
            "cpu_baseline": {"value": v, "unit": "GDOF/s", "cores": r["cores"], "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


_JSON_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--elems", "--n", dest="n", type=int, default=100, help="elements per direction per GPU")
    ap.add_argument("--ops", default="both", choices=["both", "diff"], help="diffusion+mass (headline) or diffusion only (configs[3] sweep)")
    ap.add_argument("--qdata", default="stored", choices=["stored", "factorised"],
                    help="diffusion q-data of the timed apply: the reference's six components per q-point (headline) or the "
                         "factorised form for affine meshes (b200pa_form_set_factorised)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip PCG / implicit-step extras (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    # rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's "NCCL version ..." banner at any
    # NCCL_DEBUG level >= VERSION): keep the real stdout aside for the JSON line and point fd 1 at stderr meanwhile
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist

    import b200pa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p, n, K, W = args.order, args.n, args.steps, args.warmup
    from b200pa import partition
    grid = partition.GRIDS.get(world)
    if grid is None:
        raise SystemExit(f"unsupported world size {world}")
    GN = (n * grid[0], n * grid[1], n * grid[2])

    # ---- problem set-up (host builder -> device handles); not timed
    t_setup = time.perf_counter()
    m = partition.build_part(GN, grid, rank, p, want=("gather_map", "elem_vertices", "vertices", "bdr_attr", "lattice"))
    bas = b200pa.basis(p)
    ctx = b200pa.Context(local)
    nd, ne = m["ndofs"], m["ne"]
    sp = b200pa.Space(ctx, p + 1, p + 2, ne, nd, m["gather_map"], bas["B"], bas["G"])
    sp.geometry_from_vertices(bas["W"], m["vertices"], m["elem_vertices"])
    lat = m["lattice"].reshape(-1, 3)
    gll = bas["gll"]
    xyz = (lat // p + gll[lat % p]) / np.array(GN, dtype=np.float64)
    T0h = 37.0 + 20.0 * np.exp(-40.0 * ((xyz - 0.5) ** 2).sum(1))
    T0 = ctx.to_dev(T0h)
    nq = ne * (p + 2) ** 3
    kq = sp.coeff_linear(PHYS["k0"], PHYS["ak"], 37.0, T0)
    mq = ctx.coeff_eval(1, nq, PHYS["rc"] / PHYS["dt"] + PHYS["wbcb"], 0.0, 0.0)
    fact = args.qdata == "factorised"
    form = b200pa.Form(sp)
    form.set_factorised(fact)
    form.assemble_diffusion(kq)
    if args.ops == "both":
        form.assemble_mass(mq)
    form.set_essential(None)
    comm = None
    if world > 1:
        ids = [b200pa.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = b200pa.Comm(ctx, ids[0], rank, world)
        comm.set_tables(nd, *partition.shared_tables(m, grid, p))
        form.set_comm(comm)
    global_dofs = (GN[0] * p + 1) * (GN[1] * p + 1) * (GN[2] * p + 1)
    xh = b200pa.randomize(nd, 1) if world == 1 else np.random.default_rng(1).random(nd)
    x = ctx.to_dev(xh)
    if comm is not None:
        comm.bcast(x)          # consistent L-vector
    y = ctx.empty(nd)
    ctx.sync()
    t_setup = time.perf_counter() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        """CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.torch_stream)
        for _ in range(reps):
            fn()
        e1.record(ctx.torch_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- the timed region: K applies, inputs resident in HBM
    for _ in range(W):
        form.mult(x, y)
    with ClockSampler(local) as clk:
        l0 = b200pa.launch_count()
        ms_total = timed(lambda: form.mult(x, y), K)
        launches = b200pa.launch_count() - l0
        # dominant kernel alone (same stream, same inputs) for the roofline
        ms_elem = timed(lambda: form.mult_phases(x, y, 1), K)
        ms_seg = timed(lambda: form.mult_phases(x, y, 2), K)
    clocks = clk.summary()
    form.mult(x, y)
    ynorm2 = ctx.dot(y, y) if world == 1 else None

    # ---- e2e: the same apply through the host-buffer C-ABI entry point (pinned host x, y)
    xp = torch.from_numpy(xh).pin_memory()
    yp = torch.empty(nd, dtype=torch.float64).pin_memory()
    for _ in range(3):
        form.mult_host(xp, yp)
    Ke = max(3, min(K, 20))
    ms_e2e = timed(lambda: form.mult_host(xp, yp), Ke)

    bytes_total, bytes_elem = algorithmic_bytes_per_dof(p, 7 if args.ops == "both" else 6, fact)
    peak, peak_src = measured_peak_hbm()
    t_elem = ms_elem / K * 1e-3
    achieved = bytes_elem * nd / t_elem / 1e9
    value = global_dofs * K / (ms_total * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"configs[1]: Pennes bioheat operator k(T) diffusion + (rho c/dt + perfusion) mass, PA apply L->L, "
                               f"hex {GN[0]}x{GN[1]}x{GN[2]} (N={n}^3 per GPU), order {p}, {global_dofs} dofs",
                   "order": p, "ops": args.ops, "qdata": args.qdata, "elements_per_gpu": ne, "dofs_per_gpu": nd, "global_dofs": global_dofs,
                   "partition": "x".join(map(str, grid)),
                   "exchange": ("none" if comm is None else ("peer-memory stores + flags over NVLink (CUDA IPC)" if comm.p2p_enabled()
                                                             else "NCCL send/recv + all-reduce")),
                   "l2": f"q-data + index streams = {bytes_total * nd / 1e9:.2f} GB per step >> 126 MB L2, no flush needed"},
        "e2e": {"value": global_dofs * Ke / (ms_e2e * 1e-3) / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * nd,
                "d2h_bytes_per_step": 8 * nd, "ms_per_step": ms_e2e / Ke,
                "api": "b200pa_form_mult_host (pinned host x -> H2D -> apply -> D2H -> pinned host y)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None if fact else measured_traffic(p, n, args.ops), "algorithmic_bytes": bytes_elem * nd,
                     "kernel": "pa_apply_kernel (gather + diffusion + mass + slot-order write)",
                     "bytes_per_dof": bytes_elem, "ms_per_launch": ms_elem / K, "peak_source": peak_src},
        "roofline_apply": {"achieved": bytes_total * nd / (ms_total / K * 1e-3) / 1e9, "frac": bytes_total * nd / (ms_total / K * 1e-3) / 1e9 / peak,
                           "bytes_per_dof": bytes_total, "ms_element_kernel": ms_elem / K, "ms_segment_sum": ms_seg / K},
        "clocks": clocks, "setup_s": t_setup,
    }

    # ---- the same operator with the factorised diffusion q-data (the mesh is affine): what a caller who opts in gets
    if not fact and sp.affine and not args.no_extras:
        f2 = b200pa.Form(sp)
        f2.set_factorised(True)
        f2.assemble_diffusion(kq)
        if args.ops == "both":
            f2.assemble_mass(mq)
        f2.set_essential(None)
        if comm is not None:
            f2.set_comm(comm)
        y2 = ctx.empty(nd)
        for _ in range(W):
            f2.mult(x, y2)
        ms2 = timed(lambda: f2.mult(x, y2), K)
        ms2_elem = timed(lambda: f2.mult_phases(x, y2, 1), K)
        bt2, be2 = algorithmic_bytes_per_dof(p, 7 if args.ops == "both" else 6, True)
        f2.mult(x, y2)
        form.mult(x, y)
        diff = float((y2 - y).abs().max().item() / y.abs().max().item())
        d2 = f2.jacobi()
        lf2 = sp.domain_lf(ctx.coeff_eval(1, nq, PHYS["wbcb"] * PHYS["Ta"], 0.0, 0.0))
        if comm is not None:
            comm.exchange_sum(lf2)
        rhs2 = ctx.add(lf2, 1.0, f2.mult(T0))
        Tf = T0.clone()
        f2.pcg(d2, rhs2, Tf, 0.0, 0.0, 3, want_norms=False)
        Tf.copy_(T0)
        msp1 = timed(lambda: f2.pcg(d2, rhs2, Tf, 0.0, 0.0, 20, want_norms=False), 1)
        Tf.copy_(T0)
        msp3 = timed(lambda: f2.pcg(d2, rhs2, Tf, 0.0, 0.0, 60, want_norms=False), 1)

        def implicit_step2():
            k2 = sp.coeff_linear(PHYS["k0"], PHYS["ak"], 37.0, T0, out=kq)
            f2.assemble_diffusion(k2)
            if args.ops == "both":
                f2.assemble_mass(mq)
            dd = f2.jacobi()
            Tf.copy_(T0)
            return f2.pcg(dd, rhs2, Tf, 1e-8, 0.0, 500, want_norms=False)[0]

        implicit_step2()
        r2 = [None]
        ms_step2 = timed(lambda: r2.__setitem__(0, implicit_step2()), 1)
        line["factorised_qdata"] = {
            "what": "same operator, diffusion q-data stored as w_q k_q per q-point + adj(J)adj(J)^T/detJ per element "
                    "(b200pa_form_set_factorised; the mesh is affine)",
            "value": global_dofs * K / (ms2 * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": ms2 / K, "ms_element_kernel": ms2_elem / K,
            "bytes_per_dof": bt2, "hbm_frac_own_bytes": bt2 * nd / (ms2 / K * 1e-3) / 1e9 / peak,
            "max_rel_diff_vs_stored": diff, "pcg_ms_per_iter_marginal": (msp3 - msp1) / 40,
            "bioheat_step_ms": ms_step2, "bioheat_step_pcg_iters": r2[0].final_iter}
        f2.close()

    # ---- extras on the same configuration: PCG iteration time and the implicit bioheat step
    if not args.no_extras:
        dinv = form.jacobi()
        lf = sp.domain_lf(ctx.coeff_eval(1, nq, PHYS["wbcb"] * PHYS["Ta"], 0.0, 0.0))
        if comm is not None:
            comm.exchange_sum(lf)      # local partial sums -> consistent L-vector
        rhs = ctx.add(lf, 1.0, form.mult(T0))
        its = 20
        T1 = T0.clone()
        form.pcg(dinv, rhs, T1, 0.0, 0.0, 3, want_norms=False)
        T1.copy_(T0)
        ms_pcg = timed(lambda: form.pcg(dinv, rhs, T1, 0.0, 0.0, its, want_norms=False), 1)
        T1.copy_(T0)
        ms_pcg3 = timed(lambda: form.pcg(dinv, rhs, T1, 0.0, 0.0, 3 * its, want_norms=False), 1)
        marginal = (ms_pcg3 - ms_pcg) / (2 * its)       # without the two set-up applies and the final read-back
        line["pcg"] = {"ms_per_iter": ms_pcg / its, "iters": its, "gdof_per_s": global_dofs * its / (ms_pcg * 1e-3) / 1e9,
                       "ms_per_iter_marginal": marginal, "gdof_per_s_marginal": global_dofs / (marginal * 1e-3) / 1e9}

        def implicit_step():
            k2 = sp.coeff_linear(PHYS["k0"], PHYS["ak"], 37.0, T0, out=kq)
            form.assemble_diffusion(k2)
            if args.ops == "both":
                form.assemble_mass(mq)
            d2 = form.jacobi()
            T1.copy_(T0)
            return form.pcg(d2, rhs, T1, 1e-8, 0.0, 500, want_norms=False)[0]

        implicit_step()
        res = [None]
        ms_step = timed(lambda: res.__setitem__(0, implicit_step()), 1)
        line["bioheat_step"] = {"ms": ms_step, "pcg_iters": res[0].final_iter, "converged": bool(res[0].converged),
                                "what": "k(T) q-data + PA setup + Jacobi diagonal + PCG to rel 1e-8"}

        # the whole RF-ablation coupled step (electrostatics + Joule + bioheat), fixed 20 + 20 PCG iterations
        from b200pa.bioheat import CoupledStep
        cs = CoupledStep(ctx, sp, m, GN, comm=comm)
        cs.step(T0, 2, 2)
        out = [None]
        ms_rf = timed(lambda: out.__setitem__(0, cs.step(T0, 20, 20)), 1)
        o = out[0]
        line["rf_step"] = {"ms": ms_rf, "pcg_iters": [20, 20], "what": "sigma(T),k(T) q-data + 2 PA set-ups + 2 Jacobi diagonals + "
                           "EliminateRHS + 20 PCG its (phi) + Joule q-data + RHS + 20 PCG its (T)"}
        if world == 1:
            line["rf_step"].update(phi_norm=float(np.sqrt(ctx.dot(o["phi"], o["phi"]))), T1_norm=float(np.sqrt(ctx.dot(o["T1"], o["T1"]))),
                                   src_sum=float(o["src"].sum().item()))
        cs.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference itself on the host cores
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = run_reference_cpu(p, n, 3, 1)
            if r is not None:
                cpu_v = r["ndofs"] / r["t_apply_mean"] / 1e9
                line["cpu_baseline"] = {"value": cpu_v, "unit": "GDOF/s", "cores": r["cores"], "kind": "reference",
                                        "sample": f"full workload (p={p}, N={n}, {r['ndofs']} dofs), 3 applies after 1 warm-up, "
                                                  f"OpenMP device on {r['cores']} threads"}
                # full-size parity: same mesh, same numbering, same x = Randomize(1), same operator
                line["parity_vs_reference_cpu"] = {"ref_y_norm": r["y_norm"], "gpu_y_norm": float(np.sqrt(ynorm2)),
                                                   "rel_diff": abs(np.sqrt(ynorm2) - r["y_norm"]) / r["y_norm"]}
                if "rf_step" in line and args.ops == "both":
                    rb = run_reference_bioheat(p, n, 20)
                    t_ref = sum(rb[k] for k in ("t_coef", "t_asm_e", "t_cg_e", "t_joule", "t_asm_t", "t_rhs", "t_cg_t"))
                    g = line["rf_step"]
                    line["rf_step"]["reference_cpu"] = {
                        "ms": t_ref * 1e3, "cores": rb["cores"], "phi_norm": rb["phi_norm"], "T1_norm": rb["T1_norm"], "src_sum": rb["src_sum"],
                        "rel_diff": {"phi_norm": abs(g["phi_norm"] - rb["phi_norm"]) / rb["phi_norm"],
                                     "T1_norm": abs(g["T1_norm"] - rb["T1_norm"]) / rb["T1_norm"],
                                     "src_sum": abs(g["src_sum"] - rb["src_sum"]) / abs(rb["src_sum"])}}
            else:
                r = oracle_port_cpu(p, min(n, 40))
                line["cpu_baseline"] = {"value": r["ndofs"] / r["t_apply_mean"] / 1e9, "unit": "GDOF/s", "cores": 1, "kind": "port",
                                        "sample": f"oracle port on p={p}, N={r['N']} ({r['ndofs']} dofs), 3 applies"}
        except Exception as e:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": "GDOF/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    form.close()
    sp.close()
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
