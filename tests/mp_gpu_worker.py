"""torchrun worker for tests/test_gpu_multi.py: partitioned operator / PCG over NCCL on N GPUs
against the same problem solved serially on this rank's own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "cardiac-ablation-ecm2_b200"))
import b200pa  # noqa: E402
from b200pa import partition  # noqa: E402


def setup(ctx, m, p, ess_attrs, comm=None, grid=None):
    b = b200pa.basis(p)
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    lat = m["lattice"].reshape(-1, 3)
    T = ctx.to_dev(37.0 + 5.0 * np.sin(0.37 * lat[:, 0]) * np.cos(0.21 * lat[:, 1]) + 0.1 * lat[:, 2])
    kq = sp.coeff_linear(0.5, 0.02, 37.0, T)
    f = b200pa.Form(sp)
    f.assemble_diffusion(kq)
    f.assemble_mass(np.array([3.6]))
    f.set_essential(b200pa.essential_dofs(m["bdr_attr"], ess_attrs))
    if comm is not None:
        f.set_comm(comm)
    return sp, f


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    GN = (8, 6, 4)
    grid = partition.GRIDS[world]
    ctx = b200pa.Context(local)
    ids = [b200pa.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = b200pa.Comm(ctx, ids[0], rank, world)
    m = partition.build_part(GN, grid, rank, p, size=(1.0, 0.7, 0.4), skew=True)
    comm.set_tables(m["ndofs"], *partition.shared_tables(m, grid, p))
    sp, f = setup(ctx, m, p, [1, 6], comm)
    gid = partition.global_ids(m, GN, p)
    # the same problem, serial, on this GPU
    ms = b200pa.hex_build(*GN, p, 1.0, 0.7, 0.4, skew=True)
    sps, fs = setup(ctx, ms, p, [1, 6])
    gs = partition.global_ids(dict(lattice=ms["lattice"]), GN, p)
    nglob = ms["ndofs"]
    rng = np.random.default_rng(11)
    xg, bg = rng.random(nglob), rng.random(nglob)

    def ser(v):  # lattice-indexed -> serial numbering
        return ctx.to_dev(v[gs])

    def cmp(loc, serial, tol, what):
        a = ctx.to_host(loc)
        s = np.empty(nglob)
        s[gs] = ctx.to_host(serial)
        err = np.max(np.abs(a - s[gid])) / np.max(np.abs(s))
        assert err <= tol, f"rank {rank}: {what}: rel err {err:.3e} > {tol:.1e}"
        return err

    e1 = cmp(f.constrained_mult(ctx.to_dev(xg[gid])), fs.constrained_mult(ser(xg)), 1e-12, "constrained apply")
    e2 = cmp(f.assemble_diagonal(), fs.assemble_diagonal(), 1e-12, "diagonal")
    X, Xs = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
    res, norms = f.pcg(f.jacobi(), ctx.to_dev(bg[gid]), X, 0.0, 0.0, 15)
    ress, normss = fs.pcg(fs.jacobi(), ser(bg), Xs, 0.0, 0.0, 15)
    e3 = cmp(X, Xs, 1e-10, "PCG solution after 15 iterations")
    assert res.final_iter == ress.final_iter == 15
    assert np.max(np.abs(norms - normss) / normss) <= 1e-9
    X2, Xs2 = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
    r2, _ = f.pcg(f.jacobi(), ctx.to_dev(bg[gid]), X2, 1e-8, 0.0, 2000)
    rs2, _ = fs.pcg(fs.jacobi(), ser(bg), Xs2, 1e-8, 0.0, 2000)
    assert abs(r2.final_iter - rs2.final_iter) <= 1 and r2.converged and rs2.converged
    # Chebyshev-preconditioned PCG (order 3), eigenvalue estimate from the serial power method, and the factorised
    # q-data: partitioned == serial
    lam = fs.power_method(fs.jacobi(), ctx.to_dev(b200pa.randomize(nglob, 12345)))
    X3, Xs3 = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
    r3, n3 = f.pcg_chebyshev(f.jacobi(), 3, lam, ctx.to_dev(bg[gid]), X3, 0.0, 0.0, 6)
    rs3, ns3 = fs.pcg_chebyshev(fs.jacobi(), 3, lam, ser(bg), Xs3, 0.0, 0.0, 6)
    e4 = cmp(X3, Xs3, 1e-10, "Chebyshev-PCG solution after 6 iterations")
    assert r3.final_iter == rs3.final_iter == 6 and np.max(np.abs(n3 - ns3) / ns3) <= 1e-9
    assert sp.affine
    for g in (f, fs):
        g.set_factorised(True)
    lat, lats = m["lattice"].reshape(-1, 3), ms["lattice"].reshape(-1, 3)
    tf = lambda L: ctx.to_dev(37.0 + 5.0 * np.sin(0.37 * L[:, 0]) * np.cos(0.21 * L[:, 1]) + 0.1 * L[:, 2])
    f.assemble_diffusion(sp.coeff_linear(0.5, 0.02, 37.0, tf(lat)))
    fs.assemble_diffusion(sps.coeff_linear(0.5, 0.02, 37.0, tf(lats)))
    e5 = cmp(f.constrained_mult(ctx.to_dev(xg[gid])), fs.constrained_mult(ser(xg)), 1e-12, "factorised constrained apply")
    X4, Xs4 = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
    r4, _ = f.pcg(f.jacobi(), ctx.to_dev(bg[gid]), X4, 0.0, 0.0, 15)
    rs4, _ = fs.pcg(fs.jacobi(), ser(bg), Xs4, 0.0, 0.0, 15)
    cmp(X4, Xs4, 1e-10, "factorised PCG solution after 15 iterations")
    cmp(X4, Xs, 1e-10, "factorised vs stored PCG solution")
    # bcast: owner value wins
    v = ctx.to_dev(xg[gid] + rank)
    comm.bcast(v)
    own = np.zeros(nglob)
    # expected: value of the lowest sharing rank; reconstruct from every rank's gid
    gl = [None] * world
    dist.all_gather_object(gl, gid)
    low = np.full(nglob, world, int)
    for r in range(world):
        low[gl[r]] = np.minimum(low[gl[r]], r)
    assert np.array_equal(ctx.to_host(v), xg[gid] + low[gid])
    dist.barrier()
    comm.check_p2p()
    if rank == 0:
        print(f"MULTI_OK world={world} p={p} p2p={comm.p2p_enabled()} apply={e1:.2e} diag={e2:.2e} pcg={e3:.2e} cheb={e4:.2e} fact={e5:.2e} its={r2.final_iter}/{rs2.final_iter}", flush=True)
    f.close(); sp.close(); fs.close(); sps.close(); comm.close(); ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
