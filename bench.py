#!/usr/bin/env python
"""bench.py — GDOF/s of the FP64 partial-assembly diffusion+mass apply (L->L, A.Mult semantics incl.
gather and scatter; ≙ BK3/BK1PARTIAL of the reference's tests/benchmarks/bench_assembly_levels.cpp:258-317)
on BASELINE.json configs[1]: the Pennes-bioheat operator k(T) grad + (rho c/dt + perfusion) mass on a
synthetic hex slab, order 2, N=100 per GPU (8,120,601 dofs per GPU), plus the PCG-iteration and
implicit-step times of the same configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--order P] [--elems N]

One JSON line on rank 0.  A "step" is one operator apply.  N>1 (torchrun, one rank per GPU): the global mesh
is a PX x PY x PZ grid of N^3-element boxes (weak scaling), every apply includes the shared-dof exchange over
NVLink; value = global true dofs x K / max-over-ranks time.

Beside the headline the line carries (each leg can be switched off, `--legs` lists the ones to run):
  N = 1   order_sweep (configs[3]: p = 1..6 at ~8 M dofs, diffusion+mass and diffusion only), c3 (configs[2]: the
          coupled RF step at ~30 M dofs), c5_one_gpu (the per-GPU size of configs[4] on one GPU), pcg /
          bioheat_step / rf_step / factorised_qdata on configs[1], cpu_baseline + parity against the reference CPU run;
  N > 1   parity_multi (partitioned == serial on the communicator of the timed region, global ||Ax|| against a
          one-GPU run of the same global mesh), weak_efficiency (the same per-GPU work without the exchange),
          c5 (configs[4]: N=199 per GPU, 506 M dofs at 8 GPUs, weak), strong (the 398^3 mesh of configs[4] split
          over the ranks).
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
GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
SWEEP_N = {1: 200, 2: 100, 3: 67, 4: 50, 5: 40, 6: 34}       # ~8 M dofs at every order (configs[3])
C5_N = 199                                                   # configs[4]: 398^3 elements, 506 M dofs over 2x2x2
T_START = time.perf_counter()


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


def global_dofs_of(GN, p):
    return (GN[0] * p + 1) * (GN[1] * p + 1) * (GN[2] * p + 1)


def workload_config(p, n, world, ops="both", qdata="stored"):
    """the `config` both arms print (identical for the same command line)"""
    grid = GRIDS[world]
    GN = (n * grid[0], n * grid[1], n * grid[2])
    gd = global_dofs_of(GN, p)
    tot, _ = algorithmic_bytes_per_dof(p, 7 if ops == "both" else 6, qdata == "factorised")
    what = "k(T) diffusion + (rho c/dt + perfusion) mass" if ops == "both" else "k(T) diffusion"
    return {"workload": f"configs[1]: Pennes bioheat operator {what}, PA apply L->L, hex {GN[0]}x{GN[1]}x{GN[2]} "
                        f"(N={n}^3 per GPU), order {p}, {gd} dofs",
            "order": p, "ops": ops, "qdata": qdata, "elements_per_gpu": n ** 3, "dofs_per_gpu": global_dofs_of((n, n, n), p),
            "global_dofs": gd, "partition": "x".join(map(str, grid)),
            "l2": f"q-data + index streams = {tot * global_dofs_of((n, n, n), p) / 1e9:.2f} GB per step per GPU >> 126 MB L2, no flush needed"}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled from a thread every
    ~1 ms (nvidia-smi -lms cannot go below ~20 ms: a 13 ms timed region would get one sample), nvidia-smi as the fallback"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.index, self.proc, self.stop, self.t, self.h, self.nv = [], index, None, False, None, None, None
        self.marks = {}

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(self._nvml_index(nv))
            self.nv = nv
            self.smmax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nv = None
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

    def _nvml_index(self, nv):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self):
        nv = self.nv
        # nvml.h: nvmlClocksEventReason{SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40}
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(get_reasons(self.h))
                self.rows.append((time.perf_counter(), sm, [n for n, b in bits if rs & b]))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.smmax = float(r[1])
                self.rows.append((time.perf_counter(), float(r[0]), [n for n, v in zip(names, r[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def mark(self, name):
        self.marks[name] = time.perf_counter()

    def __exit__(self, *a):
        self.stop = True
        if self.proc:
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)

    def summary(self, t0=None, t1=None):
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        sm = sorted(r[1] for r in rows)
        reasons = sorted({n for r in rows for n in r[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": getattr(self, "smmax", None), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nv else "nvidia-smi"}


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


def run_reference_cuda(p, n, reps, warm):
    """the reference's OWN CUDA backend on the same GPU (oracle/_ref/ref_driver_cuda = the unmodified sources compiled with
    nvcc -x cu for sm_100 by `make -C oracle refcuda`; optional - None when it was not built): a second baseline reported
    inside the cpu_baseline leg, never on the product path"""
    drv = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "ref_driver_cuda")
    if not os.path.exists(drv):
        return None
    out = subprocess.run([drv, "time_apply", str(p), str(n), str(reps), str(warm), "cuda"], capture_output=True, text=True, timeout=600)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    return None


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
    """--impl reference: the reference's own CPU implementation of the path, all host threads.  At --gpus N > 1 the
    workload is the N-box mesh; one host runs a bounded sample of it: ONE box (the reference's int32 q-data indexing
    ends at 5.6 M elements per rank at p=2), GDOF/s being a rate."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    p, n = args.order, args.n
    world = args.gpus if args.gpus in GRIDS else 1
    r = run_reference_cpu(p, n, args.steps, args.warmup)
    if r is not None:
        kind = "reference"
        sample = (f"{'full workload' if world == 1 else 'one of the ' + str(world) + ' boxes of the workload'}: p={p}, N={n}, "
                  f"{r['ndofs']} dofs, {args.steps} applies, OpenMP device on {r['cores']} threads")
    else:
        n_s = min(n, 40)
        r = oracle_port_cpu(p, n_s)
        kind, sample = "port", f"oracle port, p={p}, N={n_s} ({r['ndofs']} dofs), 3 applies, 1 thread"
    v = r["ndofs"] / r["t_apply_mean"] / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "GDOF/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["t_apply_mean"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(p, n, world, args.ops, "stored"),
            "cpu_baseline": {"value": v, "unit": "GDOF/s", "cores": r["cores"], "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


_JSON_OUT = sys.stdout


def hash01(gid):
    """deterministic pseudo-random value in [0,1) per GLOBAL dof id: the same global vector on any partition"""
    import numpy as np
    z = gid.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


class Env:
    """process-wide state: torch, ranks, context, communicator, timing helpers"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import b200pa
        self.torch, self.dist, self.b = torch, dist, b200pa
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        if self.world not in GRIDS:
            raise SystemExit(f"unsupported world size {self.world}")
        self.grid = GRIDS[self.world]
        self.ctx = b200pa.Context(self.local)
        self.comm = None
        if self.world > 1:
            ids = [b200pa.Comm.unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            self.comm = b200pa.Comm(self.ctx, ids[0], self.rank, self.world)
        self.peak, self.peak_src = measured_peak_hbm()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, reps, collective=True):
        """CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
        collective=False: fn has no exchange inside (every rank on its own), still max over ranks"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.ctx.torch_stream)
        for _ in range(reps):
            fn()
        e1.record(self.ctx.torch_stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        self.barrier()
        return ms

    def allsum(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.synchronize()
        self.torch.cuda.empty_cache()

    def elapsed(self):
        return time.perf_counter() - T_START


class Problem:
    """one partitioned bioheat operator: mesh part, space, stored-q-data form, consistent input vector"""

    def __init__(self, env, GN, p, ops="both", qdata="stored", x_kind="hash", grid=None, lean=False):
        import numpy as np
        b200pa, ctx = env.b, env.ctx
        from b200pa import partition
        self.env, self.p, self.ops, self.GN = env, p, ops, GN
        self.grid = grid = env.grid if grid is None else grid
        serial = grid == (1, 1, 1)
        t0 = time.perf_counter()
        rank = 0 if serial else env.rank
        m = partition.build_part(GN, grid, rank, p, want=("gather_map", "elem_vertices", "vertices", "bdr_attr", "lattice"))
        self.m, self.bas = m, b200pa.basis(p)
        bas = self.bas
        self.nd, self.ne = m["ndofs"], m["ne"]
        self.sp = sp = b200pa.Space(ctx, p + 1, p + 2, self.ne, self.nd, m["gather_map"], bas["B"], bas["G"])
        sp.geometry_from_vertices(bas["W"], m["vertices"], m["elem_vertices"])
        m["gather_map"] = m["elem_vertices"] = m["vertices"] = None        # host copies are no longer needed
        lat = m["lattice"].reshape(-1, 3)
        xyz = (lat // p + bas["gll"][lat % p]) / np.array(GN, dtype=np.float64)
        self.T0 = ctx.to_dev(37.0 + 20.0 * np.exp(-40.0 * ((xyz - 0.5) ** 2).sum(1)))
        del xyz
        self.nq = self.ne * (p + 2) ** 3
        self.kq = sp.coeff_linear(PHYS["k0"], PHYS["ak"], 37.0, self.T0)
        self.mq = np.array([PHYS["rc"] / PHYS["dt"] + PHYS["wbcb"]])       # constant coefficient (CoefficientVector of size 1)
        self.comm = None if serial else env.comm
        self.owner = None
        if self.comm is not None:
            tabs = partition.shared_tables(m, grid, p)
            self.comm.set_tables(self.nd, *tabs)
            self.owner = b200pa.comm_build_tables(env.rank, self.nd, *tabs)[3].astype(bool)
        self.form = self.make_form(qdata == "factorised")
        if lean:                 # the coefficient q-data is only needed again by a re-assembly (8 B per q-point)
            self.kq = None
            env.free()
        self.global_dofs = global_dofs_of(GN, p)
        if x_kind == "randomize":                                        # the reference's Vector::Randomize(1) (one rank)
            self.xh = b200pa.randomize(self.nd, 1)
        else:
            self.xh = hash01(partition.global_ids(m, GN, p))
        self.x = ctx.to_dev(self.xh)
        self.y = ctx.empty(self.nd)
        self.diag = ctx.empty(self.nd)
        ctx.sync()
        self.setup_s = time.perf_counter() - t0

    def make_form(self, factorised=False, comm="default"):
        f = self.env.b.Form(self.sp)
        f.set_factorised(factorised)
        f.assemble_diffusion(self.kq)
        if self.ops == "both":
            f.assemble_mass(self.mq)
        f.set_essential(None)
        c = self.comm if comm == "default" else comm
        if c is not None:
            f.set_comm(c)
        return f

    def global_sqnorm(self, v):
        """sum over owned dofs of v^2, all ranks"""
        import numpy as np
        h = self.env.ctx.to_host(v)
        s = float(np.sum(h[self.owner] ** 2)) if self.owner is not None else float(np.sum(h ** 2))
        return self.env.allsum(s) if self.comm is not None else s

    # ---------------------------------------------------------------- measurements
    def apply_times(self, form, K, W, phases=True, collective=True):
        x, y, env = self.x, self.y, self.env
        for _ in range(W):
            form.mult(x, y)
        out = {"ms_apply": env.timed(lambda: form.mult(x, y), K, collective) / K}
        if phases:
            out["ms_elem"] = env.timed(lambda: form.mult_phases(x, y, 1), K, False) / K
            out["ms_seg"] = env.timed(lambda: form.mult_phases(x, y, 2), K, collective) / K
        return out

    def roofline(self, t, factorised=False):
        bt, be = algorithmic_bytes_per_dof(self.p, 7 if self.ops == "both" else 6, factorised)
        peak = self.env.peak
        r = {"gdof_per_s": self.global_dofs / (t["ms_apply"] * 1e-3) / 1e9, "ms_apply": t["ms_apply"],
             "apply_frac": bt * self.nd / (t["ms_apply"] * 1e-3) / 1e9 / peak, "bytes_per_dof": bt}
        if "ms_elem" in t:
            r.update(ms_element_kernel=t["ms_elem"], ms_segment_sum=t["ms_seg"],
                     kernel_frac=be * self.nd / (t["ms_elem"] * 1e-3) / 1e9 / peak)
        return r

    def rhs(self, form):
        env, sp = self.env, self.sp
        lf = sp.domain_lf(env.ctx.to_dev(__import__("numpy").array([PHYS["wbcb"] * PHYS["Ta"]])))
        if form is not None and getattr(form, "_comm", None) is not None:
            form._comm.exchange_sum(lf)      # local partial sums -> consistent L-vector
        return env.ctx.add(lf, 1.0, form.mult(self.T0))

    def pcg_times(self, form, rhs, its=20, collective=True):
        env = self.env
        dinv = form.jacobi()
        T1 = self.T0.clone()
        form.pcg(dinv, rhs, T1, 0.0, 0.0, 3, want_norms=False)
        T1.copy_(self.T0)
        ms1 = env.timed(lambda: form.pcg(dinv, rhs, T1, 0.0, 0.0, its, want_norms=False), 1, collective)
        T1.copy_(self.T0)
        ms3 = env.timed(lambda: form.pcg(dinv, rhs, T1, 0.0, 0.0, 3 * its, want_norms=False), 1, collective)
        marginal = (ms3 - ms1) / (2 * its)       # without the two set-up applies and the final read-back
        return {"ms_per_iter": ms1 / its, "iters": its, "gdof_per_s": self.global_dofs * its / (ms1 * 1e-3) / 1e9,
                "ms_per_iter_marginal": marginal, "gdof_per_s_marginal": self.global_dofs / (marginal * 1e-3) / 1e9}

    def implicit_step(self, form, rhs, rel_tol, max_iter, collective=True, with_rhs=False):
        """k(T) q-data + PA set-up + Jacobi diagonal + PCG: (ms, result).  rel_tol = 0: exactly max_iter iterations.
        with_rhs: the right-hand side is assembled inside the timed region as well: the perfusion source
        through the load-vector kernel, its shared-dof sums, and (rho c / dt) M T^n through the mass-only form"""
        env, sp = self.env, self.sp
        T1 = self.T0.clone()
        fm = None
        if with_rhs:
            import numpy as np
            fm = env.b.Form(sp)
            fm.assemble_mass(np.array([PHYS["rc"] / PHYS["dt"]]))
            fm.set_essential(None)
            if getattr(form, "_comm", None) is not None:
                fm.set_comm(form._comm)
            src1 = env.ctx.to_dev(np.array([PHYS["wbcb"] * PHYS["Ta"]]))
            lf, mt, rhs2 = env.ctx.empty(self.nd), env.ctx.empty(self.nd), env.ctx.empty(self.nd)

        def step():
            k2 = sp.coeff_linear(PHYS["k0"], PHYS["ak"], 37.0, self.T0, out=self.kq)
            if self.ops == "both":
                form.assemble_mass(self.mq)
            # diffusion q-data and the Jacobi diagonal in one pass over the q-points (two passes on non-affine meshes)
            d2 = form.jacobi_from(form.assemble_diffusion_with_diagonal(k2, self.diag))
            b = rhs
            if with_rhs:
                sp.domain_lf(src1, out=lf)
                if getattr(form, "_comm", None) is not None:
                    form._comm.exchange_sum(lf)
                fm.mult(self.T0, mt)
                b = env.ctx.add(lf, 1.0, mt, out=rhs2)
            T1.copy_(self.T0)
            return form.pcg(d2, b, T1, rel_tol, 0.0, max_iter, want_norms=False)[0]

        step()
        res = [None]
        ms = env.timed(lambda: res.__setitem__(0, step()), 1, collective)
        if fm is not None:
            fm.close()
        return ms, res[0]

    def close(self):
        self.form.close()
        self.sp.close()
        self.T0 = self.kq = self.x = self.y = self.diag = self.m = None
        self.env.free()


def leg_bioheat(P, form, collective=True, fixed_iters=10, with_rhs=False):
    """PCG iteration time + the implicit bioheat step to tolerance and at a fixed iteration count"""
    rhs = P.rhs(form)
    out = {"pcg": P.pcg_times(form, rhs, 20, collective)}
    ms, res = P.implicit_step(form, rhs, 1e-8, 500, collective)
    out["bioheat_step"] = {"ms": ms, "pcg_iters": res.final_iter, "converged": bool(res.converged),
                           "what": "k(T) q-data + PA set-up fused with the Jacobi diagonal + PCG to rel 1e-8"}
    ms, res = P.implicit_step(form, rhs, 0.0, fixed_iters, collective)
    out["bioheat_step_fixed"] = {"ms": ms, "pcg_iters": res.final_iter,
                                 "what": f"the same step with exactly {fixed_iters} PCG iterations (comparable across GPU counts: "
                                         "the iteration count to tolerance grows with the global mesh)"}
    if not with_rhs:
        return out
    ms, res = P.implicit_step(form, rhs, 1e-8, 500, collective, with_rhs=True)
    out["bioheat_step_with_rhs"] = {"ms": ms, "pcg_iters": res.final_iter, "converged": bool(res.converged),
                                    "what": "SURVEY 8(d)(ii): k(T) q-data + PA set-up + Jacobi diagonal + right-hand side (perfusion source "
                                            "load vector + (rho c/dt) M T^n) + PCG to rel 1e-8"}
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="headline only (profiling / tuning runs)")
    ap.add_argument("--legs", default="all", help="comma list of: factorised,bioheat,rf,sweep,c3,c5,parity,strong,mg (default all)")
    ap.add_argument("--c5-elems", type=int, default=C5_N, help="per-GPU N of the configs[4] legs (tests use a small one)")
    ap.add_argument("--budget-s", type=float, default=600.0, help="optional legs are skipped once the run is older than this")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    legs = set("factorised,bioheat,rf,sweep,c3,c5,parity,strong,mg".split(",")) if args.legs == "all" else set(args.legs.split(","))
    if args.no_extras:
        legs = set()

    # rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL's "NCCL version ..." banner at any
    # NCCL_DEBUG level >= VERSION): keep the real stdout aside for the JSON line and point fd 1 at stderr meanwhile
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import numpy as np

    env = Env(args)
    torch, b200pa, ctx = env.torch, env.b, env.ctx
    rank, world = env.rank, env.world
    args.gpus = world
    p, n, K, W = args.order, args.n, args.steps, args.warmup
    grid = env.grid
    GN = (n * grid[0], n * grid[1], n * grid[2])
    line_extra = {}

    def guarded(name, fn):
        """optional legs never take the headline down; all ranks take the same decision"""
        skip = env.elapsed() > args.budget_s
        if world > 1:
            t = torch.tensor([1.0 if skip else 0.0], device="cuda")
            env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
            skip = bool(t.item() > 0)
        if skip:
            line_extra[name] = {"skipped": f"run older than --budget-s {args.budget_s:.0f} s"}
            return
        t0 = time.perf_counter()
        try:
            r = fn()
            if isinstance(r, dict):
                r["leg_s"] = time.perf_counter() - t0
            line_extra[name] = r
        except Exception as e:  # noqa: BLE001
            msg = f"{type(e).__name__}: {e}"
            # a rank that drops out of a collective leg would hang the others: fail loudly - except for running out of
            # device memory, which every rank of an evenly partitioned problem does at the same allocation
            if world > 1 and "out of memory" not in msg.lower():
                raise
            line_extra[name] = {"failed": msg[:400]}
        if isinstance(line_extra.get(name), dict) and "failed" in line_extra[name]:
            del fn
            env.free()

    # ---- multi-GPU correctness FIRST, on the communicator / transport the timed region uses
    if world > 1 and "parity" in legs:
        from b200pa import selfcheck
        r = selfcheck.partitioned_vs_serial(ctx, env.comm, rank, world, p=2, GN=(16, 12, 8), full=False)
        r["tolerances"] = {"apply": 1e-12, "pcg_fixed_iters": 1e-10, "iteration_counts": "+-1"}
        line_extra["parity_multi"] = r

    # ---- the headline problem (set-up not timed)
    P = Problem(env, GN, p, args.ops, args.qdata, x_kind="randomize" if world == 1 else "hash")
    form, x, y, nd, ne = P.form, P.x, P.y, P.nd, P.ne
    comm = P.comm
    fact = args.qdata == "factorised"
    global_dofs = P.global_dofs

    # ---- the timed region: K applies, inputs resident in HBM
    for _ in range(W):
        form.mult(x, y)
    with ClockSampler(env.local) as clk:
        l0 = b200pa.launch_count()
        clk.mark("t0")
        ms_total = env.timed(lambda: form.mult(x, y), K)
        clk.mark("t1")
        launches = b200pa.launch_count() - l0
        # dominant kernel alone (same stream, same inputs) for the roofline
        ms_elem = env.timed(lambda: form.mult_phases(x, y, 1), K, False)
        ms_seg = env.timed(lambda: form.mult_phases(x, y, 2), K)
    # clocks of the K-step timed region itself (NVML polled every ~1 ms), and of the three timing loops together
    clocks = clk.summary(clk.marks["t0"], clk.marks["t1"])
    clocks["all_timing_loops"] = {k: v for k, v in clk.summary().items() if k in ("sm_mhz", "reasons", "samples")}
    form.mult(x, y)
    ynorm2 = P.global_sqnorm(y)

    # ---- e2e: the same apply through the host-buffer C-ABI entry point (pinned host x, y)
    xp = ctx.pinned(nd, fill=P.xh)          # page-locked, on the NUMA node this rank's GPU hangs off
    yp = ctx.pinned(nd)
    host_node = ctx.host_node(xp)
    for _ in range(3):
        form.mult_host(xp, yp)
    Ke = max(3, min(K, 20))
    ms_e2e = env.timed(lambda: form.mult_host(xp, yp), Ke)

    bytes_total, bytes_elem = algorithmic_bytes_per_dof(p, 7 if args.ops == "both" else 6, fact)
    peak, peak_src = env.peak, env.peak_src
    t_elem = ms_elem / K * 1e-3
    achieved = bytes_elem * nd / t_elem / 1e9
    value = global_dofs * K / (ms_total * 1e-3) / 1e9
    traffic = None if fact else measured_traffic(p, n, args.ops)
    line = {
        "metric": METRIC, "value": value, "unit": "GDOF/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(p, n, world, args.ops, args.qdata),
        "transport": ("none (one GPU)" if comm is None else ("peer-memory stores + flags over NVLink (CUDA IPC mailboxes), no NCCL call in the loop"
                                                             if comm.p2p_enabled() else "NCCL send/recv + all-reduce")),
        "e2e": {"value": global_dofs * Ke / (ms_e2e * 1e-3) / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 8 * nd,
                "d2h_bytes_per_step": 8 * nd, "ms_per_step": ms_e2e / Ke,
                "api": "b200pa_form_mult_host (pinned host x -> H2D -> apply -> D2H -> pinned host y)",
                "host_buffers": "b200pa_host_alloc: page-locked, " + (f"first-touched on NUMA node {host_node} of this rank's GPU" if host_node >= 0
                                                                       else "placement left to the kernel (no NUMA information for the GPU)")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_source": None if traffic is None else "ncu --set full capture of this command committed under profiles/ (profiles/traffic.json); not measured in this run",
                     "algorithmic_bytes": bytes_elem * nd,
                     "kernel": "pa_apply_kernel (gather + diffusion + mass + slot-order write)",
                     "bytes_per_dof": bytes_elem, "ms_per_launch": ms_elem / K, "peak_source": peak_src},
        "roofline_apply": {"achieved": bytes_total * nd / (ms_total / K * 1e-3) / 1e9, "frac": bytes_total * nd / (ms_total / K * 1e-3) / 1e9 / peak,
                           "bytes_per_dof": bytes_total, "ms_element_kernel": ms_elem / K, "ms_segment_sum": ms_seg / K},
        "clocks": clocks, "setup_s": P.setup_s,
        "apply_sqnorm_global": ynorm2,
    }

    # ---- e2e of what an application calls: one PCG solve through the host-buffer entry point (b, x cross PCIe once per solve)
    if legs:
        rhs0 = P.rhs(form)
        bp = ctx.pinned(nd, fill=rhs0.cpu().numpy())
        xs = ctx.pinned(nd)
        dinv0 = form.jacobi()
        its = 20

        xs[:] = 0.0
        form.pcg(dinv0, bp, xs, 0.0, 0.0, its, want_norms=False, host=True)
        xs[:] = 0.0
        ms_sh = env.timed(lambda: form.pcg(dinv0, bp, xs, 0.0, 0.0, its, want_norms=False, host=True), 1)
        line["e2e_pcg"] = {"value": global_dofs * its / (ms_sh * 1e-3) / 1e9, "unit": "GDOF/s (dofs x PCG iterations / s)", "iters": its,
                           "ms_per_solve": ms_sh, "h2d_bytes_per_solve": 16 * nd, "d2h_bytes_per_solve": 8 * nd,
                           "api": "b200pa_pcg_solve_host (pinned host b, x -> H2D -> 20 Jacobi-PCG iterations -> D2H x)"}
        del bp, xs, rhs0, dinv0

    # ---- multi-GPU: the same per-GPU work with the exchange switched off (every rank alone) = the weak-scaling reference
    if world > 1 and legs:
        fs = P.make_form(fact, comm=None)
        t_solo = P.apply_times(fs, K, W, phases=False, collective=False)
        line["weak_efficiency"] = {"what": "this rank's box as a one-GPU problem (no exchange, no all-reduce), max over ranks, against the "
                                           "partitioned run of the same command", "ms_apply_no_exchange": t_solo["ms_apply"],
                                   "apply": t_solo["ms_apply"] / (ms_total / K)}
        if "bioheat" in legs:
            solo = leg_bioheat(P, fs, collective=False)
            line_extra["_solo"] = solo
        fs.close()
        # global ||Ax|| against a one-GPU run of the SAME global mesh and the same global input vector (where it fits)
        if "parity" in legs and GN[0] * GN[1] * GN[2] <= 8_200_000:
            def serial_norm():
                r = {"what": "||A x||_2 over the global mesh: partitioned (owned dofs, all ranks) vs the same mesh on ONE GPU (rank 0)"}
                if rank == 0:
                    S = Problem(env, GN, p, args.ops, args.qdata, x_kind="hash", grid=(1, 1, 1))
                    S.form.mult(S.x, S.y)
                    s2 = S.global_sqnorm(S.y)
                    S.close()
                    r.update(partitioned=float(np.sqrt(ynorm2)), one_gpu=float(np.sqrt(s2)),
                             rel_diff=abs(np.sqrt(ynorm2) - np.sqrt(s2)) / np.sqrt(s2), tolerance=1e-12)
                env.barrier()
                return r
            guarded("parity_global_norm", serial_norm)

    # ---- the same operator with the factorised diffusion q-data (the mesh is affine): what a caller who opts in gets
    if not fact and P.sp.affine and "factorised" in legs:
        def leg_fact():
            f2 = P.make_form(True)
            t2 = P.apply_times(f2, K, W, phases=True)
            bt2, _ = algorithmic_bytes_per_dof(p, 7 if args.ops == "both" else 6, True)
            y2 = ctx.empty(nd)
            f2.mult(x, y2)
            form.mult(x, y)
            diff = float((y2 - y).abs().max().item() / y.abs().max().item())
            bh = leg_bioheat(P, f2)
            out = {"what": "same operator, diffusion q-data stored as w_q k_q per q-point + adj(J)adj(J)^T/detJ per element "
                           "(b200pa_form_set_factorised; the mesh is affine)",
                   "value": global_dofs / (t2["ms_apply"] * 1e-3) / 1e9, "unit": "GDOF/s", "ms_per_step": t2["ms_apply"],
                   "ms_element_kernel": t2["ms_elem"], "bytes_per_dof": bt2,
                   "hbm_frac_own_bytes": bt2 * nd / (t2["ms_apply"] * 1e-3) / 1e9 / peak, "max_rel_diff_vs_stored": diff,
                   "pcg_ms_per_iter_marginal": bh["pcg"]["ms_per_iter_marginal"], "bioheat_step_ms": bh["bioheat_step"]["ms"],
                   "bioheat_step_pcg_iters": bh["bioheat_step"]["pcg_iters"], "bioheat_step_fixed_ms": bh["bioheat_step_fixed"]["ms"]}
            f2.close()
            return out
        guarded("factorised_qdata", leg_fact)

    # ---- PCG iteration time and the implicit bioheat step on the headline configuration
    if "bioheat" in legs:
        bh = leg_bioheat(P, form, with_rhs=True)
        line.update(bh)
        solo = line_extra.pop("_solo", None)
        if solo is not None:
            we = line["weak_efficiency"]
            we["pcg_iteration"] = solo["pcg"]["ms_per_iter_marginal"] / bh["pcg"]["ms_per_iter_marginal"]
            we["bioheat_step_fixed"] = solo["bioheat_step_fixed"]["ms"] / bh["bioheat_step_fixed"]["ms"]
            we["ms_no_exchange"] = {"pcg_iteration": solo["pcg"]["ms_per_iter_marginal"], "bioheat_step_fixed": solo["bioheat_step_fixed"]["ms"]}

    # ---- the whole RF-ablation coupled step (electrostatics + Joule + bioheat), fixed 20 + 20 PCG iterations
    def rf_leg(PP, GNN):
        from b200pa.bioheat import CoupledStep
        cs = CoupledStep(ctx, PP.sp, PP.m, GNN, comm=PP.comm)
        cs.step(PP.T0, 2, 2)
        out = [None]
        ms_rf = env.timed(lambda: out.__setitem__(0, cs.step(PP.T0, 20, 20)), 1)
        o = out[0]
        r = {"ms": ms_rf, "pcg_iters": [20, 20], "what": "sigma(T),k(T) q-data + 2 PA set-ups + 2 Jacobi diagonals + "
             "EliminateRHS + 20 PCG its (phi) + Joule q-data + RHS + 20 PCG its (T)"}
        if world == 1:
            r.update(phi_norm=float(np.sqrt(ctx.dot(o["phi"], o["phi"]))), T1_norm=float(np.sqrt(ctx.dot(o["T1"], o["T1"]))),
                     src_sum=float(o["src"].sum().item()))
        cs.close()
        return r
    if "rf" in legs and args.ops == "both":
        line["rf_step"] = rf_leg(P, GN)

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
                try:
                    rc = run_reference_cuda(p, n, 10, 3)
                    if rc is not None:
                        line["cpu_baseline"]["reference_cuda_backend_same_gpu"] = {
                            "value": rc["ndofs"] / rc["t_apply_mean"] / 1e9, "unit": "GDOF/s", "ms_per_apply": rc["t_apply_mean"] * 1e3,
                            "y_norm": rc["y_norm"], "rel_diff_vs_this_library": abs(np.sqrt(ynorm2) - rc["y_norm"]) / rc["y_norm"],
                            "what": "the reference's own CUDA kernels (unmodified sources, nvcc -x cu, sm_100; BilinearForm(PARTIAL)::Mult "
                                    "under Device(\"cuda\")) on this GPU, same workload and input vector; 10 applies after 3 warm-ups"}
                except Exception as e:  # noqa: BLE001
                    line["cpu_baseline"]["reference_cuda_backend_same_gpu"] = {"failed": str(e)[:200]}
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
    del form, x, y, xp, yp
    P.close()

    # ---- configs[3]: order sweep on one GPU, ~8 M dofs per order, diffusion+mass and diffusion only
    if world == 1 and "sweep" in legs:
        def sweep():
            rows = []
            for pp in range(1, 7):
                nn = SWEEP_N[pp]
                for ops in ("both", "diff"):
                    if ops == "both":
                        S = Problem(env, (nn, nn, nn), pp, "both")
                        f = S.form
                    else:
                        S.ops = "diff"
                        f = S.make_form(False)
                    with ClockSampler(env.local) as ck:
                        t = S.apply_times(f, 20, 3)
                    r = S.roofline(t)
                    r.update(order=pp, ops=ops, elems=nn, dofs=S.nd, clocks=ck.summary())
                    rows.append(r)
                    if ops == "diff":
                        f.close()
                        S.close()
            return {"what": "configs[3]: PA apply roofline sweep on one B200, ~8 M dofs per order; apply_frac = SURVEY §8(d) bytes / "
                            "time / measured HBM peak for the whole L->L apply, kernel_frac = the element kernel's share alone",
                    "rows": rows}
        guarded("order_sweep", sweep)

    # ---- configs[2]: the coupled RF step at ~30 M dofs on one GPU
    if world == 1 and "c3" in legs and args.ops == "both":
        def c3():
            nn = 155
            S = Problem(env, (nn, nn, nn), 2, "both")
            t = S.apply_times(S.form, 10, 3)
            r = {"what": "configs[2]: RF-ablation coupled step, order 2, hex 155^3, 30,080,231 dofs, one GPU", "dofs": S.nd,
                 "apply": S.roofline(t), "setup_s": S.setup_s}
            r.update(leg_bioheat(S, S.form, with_rhs=True))
            r["rf_step"] = rf_leg(S, (nn, nn, nn))
            S.close()
            return r
        guarded("c3", c3)

    # ---- the electrostatic solve (pure diffusion, Dirichlet on two faces) to a tolerance: Jacobi-PCG against PCG
    #      preconditioned by the p-multigrid cycle (orders 1 -> 2, Chebyshev smoothers, CG coarse solve: examples/ex26.cpp)
    if "mg" in legs and p == 2:
        def mg_leg():
            from b200pa import partition, selfcheck
            nn = min(n, 64)
            GNm = (nn * grid[0], nn * grid[1], nn * grid[2])
            levels, lcomms = [], []
            for pp in (1, 2):
                mm = partition.build_part(GNm, grid, rank, pp, want=("gather_map", "elem_vertices", "vertices", "bdr_attr", "lattice"))
                bb = b200pa.basis(pp)
                spp = b200pa.Space(ctx, pp + 1, pp + 2, mm["ne"], mm["ndofs"], mm["gather_map"], bb["B"], bb["G"])
                spp.geometry_from_vertices(bb["W"], mm["vertices"], mm["elem_vertices"])
                lat = mm["lattice"].reshape(-1, 3)
                xyz = (lat // pp + bb["gll"][lat % pp]) / np.array(GNm, dtype=np.float64)
                Tl = ctx.to_dev(37.0 + 20.0 * np.exp(-40.0 * ((xyz - 0.5) ** 2).sum(1)))
                ff = b200pa.Form(spp)
                ff.assemble_diffusion(spp.coeff_linear(PHYS["s0"], PHYS["as_"], 37.0, Tl))
                ess = b200pa.essential_dofs(mm["bdr_attr"], [1, 6])
                ff.set_essential(ess)
                if comm is not None:      # one communicator per level (the shared-dof tables differ per order), same transport
                    lc = selfcheck._level_comm(ctx, comm, rank, world)
                    lc.set_tables(mm["ndofs"], *partition.shared_tables(mm, grid, pp))
                    ff.set_comm(lc)
                    lcomms.append(lc)
                levels.append((spp, ff, mm, ess, lat))
            spf, ffn, mf, ess, lat = levels[1]
            T = b200pa.Transfer(levels[0][1], ffn, b200pa.basis_transfer(1, 2))
            mg = b200pa.Multigrid([levels[0][1], ffn], [T])
            mg.set_coarse_solver(1e-2, 0.0, 200, jacobi=True)
            mg.setup()
            phi_bc = np.zeros(mf["ndofs"])
            phi_bc[ess] = PHYS["V"] * (1.0 - lat[ess, 2] / (2 * GNm[2]))
            out = {"what": f"electrostatic solve div sigma(T) grad phi = 0 to rel 1e-8, order 2, hex {GNm[0]}x{GNm[1]}x{GNm[2]}, "
                           f"{global_dofs_of(GNm, 2)} dofs on {world} GPU(s): OperatorJacobiSmoother against the p-multigrid V-cycle "
                           "(orders 1-2, Chebyshev smoothing, CG on the order-1 level to 1e-2 / 200 iterations as in ex26) as the CG "
                           "preconditioner; at order 2 the order-1 coarse problem is only 8x smaller, so the cycle buys iterations, "
                           "not time"}
            for name in ("jacobi", "p_multigrid"):
                def solve():
                    phi = ctx.to_dev(phi_bc)
                    rhs = ctx.zeros(mf["ndofs"])
                    ffn.eliminate_rhs(phi, rhs)
                    if name == "jacobi":
                        return ffn.pcg(ffn.jacobi(), rhs, phi, 1e-8, 0.0, 5000, want_norms=False)[0]
                    return mg.pcg(rhs, phi, 1e-8, 0.0, 500, want_norms=False)[0]
                solve()
                rr = [None]
                ms = env.timed(lambda: rr.__setitem__(0, solve()), 1)
                out[name] = {"ms": ms, "iters": int(rr[0].final_iter), "converged": bool(rr[0].converged)}
            out["coarse_cg_iterations_total"] = mg.coarse_iterations()
            mg.close(); T.close()
            for spp, ff, *_ in levels:
                ff.close(); spp.close()
            for lc in lcomms:
                lc.check_p2p()
                lc.close()
            return out
        guarded("p_multigrid", mg_leg)

    # ---- configs[4]: N=199 per GPU (398^3 elements, 506 M dofs at 8 GPUs): weak leg at any N, incl. the one-GPU base
    c5n = args.c5_elems
    if "c5" in legs and args.ops == "both":
        def c5():
            GN5 = (c5n * grid[0], c5n * grid[1], c5n * grid[2])
            S = Problem(env, GN5, 2, "both")
            t = S.apply_times(S.form, 10, 3)
            r = {"what": f"configs[4] weak: N={c5n}^3 elements per GPU, order 2, global hex {GN5[0]}x{GN5[1]}x{GN5[2]}, "
                         f"{S.global_dofs} dofs on {world} GPU(s)", "global_dofs": S.global_dofs, "dofs_per_gpu": S.nd,
                 "apply": S.roofline(t), "setup_s": S.setup_s}
            r.update(leg_bioheat(S, S.form))
            if world > 1:
                fs = S.make_form(False, comm=None)
                ts = S.apply_times(fs, 10, 3, phases=False, collective=False)
                solo = leg_bioheat(S, fs, collective=False)
                fs.close()
                r["one_gpu_same_size"] = {"what": "this rank's box as a one-GPU problem (no exchange), max over ranks",
                                          "ms_apply": ts["ms_apply"], "pcg_ms_per_iter_marginal": solo["pcg"]["ms_per_iter_marginal"],
                                          "bioheat_step_fixed_ms": solo["bioheat_step_fixed"]["ms"], "bioheat_step_ms": solo["bioheat_step"]["ms"],
                                          "bioheat_step_pcg_iters": solo["bioheat_step"]["pcg_iters"]}
                r["parallel_efficiency"] = {"apply": ts["ms_apply"] / t["ms_apply"],
                                            "pcg_iteration": solo["pcg"]["ms_per_iter_marginal"] / r["pcg"]["ms_per_iter_marginal"],
                                            "bioheat_step_fixed": solo["bioheat_step_fixed"]["ms"] / r["bioheat_step_fixed"]["ms"],
                                            "bioheat_step_to_tolerance": solo["bioheat_step"]["ms"] / r["bioheat_step"]["ms"]}
            S.close()
            return r
        guarded("c5", c5)

    # ---- configs[4] strong: the 398^3 mesh split over 2 / 4 / 8 GPUs (at 8 it is the weak leg's problem)
    if world in (2, 4) and "strong" in legs and args.ops == "both":
        def strong():
            GNs = (2 * c5n, 2 * c5n, 2 * c5n)
            # two ranks hold 31.5 M elements each: 113 GB of stored q-data per GPU - the coefficient array is dropped after
            # the assembly and the step with its re-assembly is only timed from four ranks on
            S = Problem(env, GNs, 2, "both", lean=(world == 2))
            t = S.apply_times(S.form, 5, 3)
            r = {"what": f"configs[4] strong: global hex {GNs[0]}^3, order 2, {S.global_dofs} dofs split over {world} GPUs "
                         "(compare the same key across the --gpus 2 / 4 lines and c5 of the --gpus 8 line)",
                 "global_dofs": S.global_dofs, "dofs_per_gpu": S.nd, "apply": S.roofline(t), "setup_s": S.setup_s,
                 "gpu_mem_gb": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}
            rhs = S.rhs(S.form)
            r["pcg"] = S.pcg_times(S.form, rhs, 10)
            if world > 2 or c5n < 150:
                ms, res = S.implicit_step(S.form, rhs, 0.0, 10)
                r["bioheat_step_fixed"] = {"ms": ms, "pcg_iters": res.final_iter}
            S.close()
            return r
        guarded("strong", strong)
    if world == 8 and "strong" in legs and isinstance(line_extra.get("c5"), dict) and "apply" in line_extra["c5"]:
        c = line_extra["c5"]
        line_extra["strong"] = {"what": "configs[4] strong at 8 GPUs = the c5 problem (398^3 elements over 2x2x2)", "global_dofs": c["global_dofs"],
                                "dofs_per_gpu": c["dofs_per_gpu"], "apply": c["apply"], "pcg": c["pcg"], "bioheat_step_fixed": c["bioheat_step_fixed"]}

    # ---- the headline apply kept running for ~0.3 s (last: a sustained FP64 + HBM load runs into the board's power cap,
    # which would then colour every leg after it)
    if legs:
        def sustained():
            S = Problem(env, GN, p, args.ops, args.qdata)
            reps = max(K, int(0.3 / max(ms_total / K * 1e-3, 1e-6)))
            for _ in range(W):
                S.form.mult(S.x, S.y)
            with ClockSampler(env.local) as ck:
                ms = env.timed(lambda: S.form.mult(S.x, S.y), reps)
            r = {"what": "the headline apply repeated back to back", "applies": reps, "ms_per_step": ms / reps,
                 "gdof_per_s": S.global_dofs / (ms / reps * 1e-3) / 1e9, "clocks": ck.summary()}
            S.close()
            return r
        guarded("sustained_apply", sustained)

    line.update(line_extra)
    line["bench_wall_s"] = env.elapsed()
    if rank == 0:
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if env.comm is not None:
        env.comm.close()
    ctx.close()
    if world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
