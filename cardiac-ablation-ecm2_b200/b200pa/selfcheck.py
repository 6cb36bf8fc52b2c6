"""Partitioned == serial: the multi-GPU path checked against the one-GPU path of the SAME library on the same
global mesh (the reference's MPI path cannot be built in this image - SURVEY.md §8c - so this is how the
shared-dof exchange and the all-reduced dots are pinned; the one-GPU path itself is pinned against the
reference by tests/test_gpu_parity.py and bench.py's parity_vs_reference_cpu).

Used by tests/mp_gpu_worker.py (pytest -m gpu, >= 2 ranks) and by bench.py's world > 1 leg, over the very
communicator object / transport the timed region uses.  Every rank solves the serial problem on its own GPU.
"""
import numpy as np

from . import Comm, Form, Multigrid, Space, Transfer, basis, basis_transfer, essential_dofs, hex_build, partition, randomize


def _field(lat):
    return 37.0 + 5.0 * np.sin(0.37 * lat[:, 0]) * np.cos(0.21 * lat[:, 1]) + 0.1 * lat[:, 2]


def _setup(ctx, m, p, ess_attrs, comm=None, factorised=False):
    b = basis(p)
    sp = Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    T = ctx.to_dev(_field(m["lattice"].reshape(-1, 3)))
    f = Form(sp)
    f.set_factorised(factorised)
    f.assemble_diffusion(sp.coeff_linear(0.5, 0.02, 37.0, T))
    f.assemble_mass(np.array([3.6]))
    f.set_essential(essential_dofs(m["bdr_attr"], ess_attrs))
    if comm is not None:
        f.set_comm(comm)
    return sp, f


def partitioned_vs_serial(ctx, comm, rank, world, p=2, GN=(8, 6, 4), full=True):
    """Returns a dict of relative errors / iteration counts; raises AssertionError beyond the north-star
    tolerances (1e-12 per apply, 1e-10 after a fixed number of PCG iterations, iteration counts +-1).
    `comm` must be a b200pa.Comm of `world` ranks; its tables are (re)set here for the check mesh."""
    grid = partition.GRIDS[world]
    m = partition.build_part(GN, grid, rank, p, size=(1.0, 0.7, 0.4), skew=True)
    comm.set_tables(m["ndofs"], *partition.shared_tables(m, grid, p))
    sp, f = _setup(ctx, m, p, [1, 6], comm)
    gid = partition.global_ids(m, GN, p)
    ms = hex_build(*GN, p, 1.0, 0.7, 0.4, skew=True)
    sps, fs = _setup(ctx, ms, p, [1, 6])
    gs = partition.global_ids(dict(lattice=ms["lattice"]), GN, p)
    nglob = ms["ndofs"]
    rng = np.random.default_rng(11)
    xg, bg = rng.random(nglob), rng.random(nglob)
    out = {"world": world, "order": p, "global_mesh": list(GN), "global_dofs": int(nglob), "transport": "peer-memory" if comm.p2p_enabled() else "nccl"}

    def ser(v):  # lattice-indexed -> serial numbering
        return ctx.to_dev(v[gs])

    def cmp(loc, serial, tol, what):
        a = ctx.to_host(loc)
        s = np.empty(nglob)
        s[gs] = ctx.to_host(serial)
        err = float(np.max(np.abs(a - s[gid])) / np.max(np.abs(s)))
        assert err <= tol, f"rank {rank}: {what}: rel err {err:.3e} > {tol:.1e}"
        return err

    out["apply_rel_err"] = cmp(f.constrained_mult(ctx.to_dev(xg[gid])), fs.constrained_mult(ser(xg)), 1e-12, "constrained apply")
    out["diag_rel_err"] = cmp(f.assemble_diagonal(), fs.assemble_diagonal(), 1e-12, "diagonal")
    X, Xs = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
    res, norms = f.pcg(f.jacobi(), ctx.to_dev(bg[gid]), X, 0.0, 0.0, 15)
    ress, normss = fs.pcg(fs.jacobi(), ser(bg), Xs, 0.0, 0.0, 15)
    out["pcg15_rel_err"] = cmp(X, Xs, 1e-10, "PCG solution after 15 iterations")
    assert res.final_iter == ress.final_iter == 15
    assert np.max(np.abs(norms - normss) / normss) <= 1e-9
    X2, Xs2 = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
    r2, _ = f.pcg(f.jacobi(), ctx.to_dev(bg[gid]), X2, 1e-8, 0.0, 2000)
    rs2, _ = fs.pcg(fs.jacobi(), ser(bg), Xs2, 1e-8, 0.0, 2000)
    assert abs(r2.final_iter - rs2.final_iter) <= 1 and r2.converged and rs2.converged
    out["pcg_iters_to_1e-8"] = [int(r2.final_iter), int(rs2.final_iter)]
    if full:
        # Chebyshev-preconditioned PCG (order 3), eigenvalue estimate from the serial power method, and the
        # factorised q-data: partitioned == serial
        lam = fs.power_method(fs.jacobi(), ctx.to_dev(randomize(nglob, 12345)))
        X3, Xs3 = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
        r3, n3 = f.pcg_chebyshev(f.jacobi(), 3, lam, ctx.to_dev(bg[gid]), X3, 0.0, 0.0, 6)
        rs3, ns3 = fs.pcg_chebyshev(fs.jacobi(), 3, lam, ser(bg), Xs3, 0.0, 0.0, 6)
        out["cheb_pcg6_rel_err"] = cmp(X3, Xs3, 1e-10, "Chebyshev-PCG solution after 6 iterations")
        assert r3.final_iter == rs3.final_iter == 6 and np.max(np.abs(n3 - ns3) / ns3) <= 1e-9
        assert sp.affine
        for g in (f, fs):
            g.set_factorised(True)
        f.assemble_diffusion(sp.coeff_linear(0.5, 0.02, 37.0, ctx.to_dev(_field(m["lattice"].reshape(-1, 3)))))
        fs.assemble_diffusion(sps.coeff_linear(0.5, 0.02, 37.0, ctx.to_dev(_field(ms["lattice"].reshape(-1, 3)))))
        out["factorised_apply_rel_err"] = cmp(f.constrained_mult(ctx.to_dev(xg[gid])), fs.constrained_mult(ser(xg)), 1e-12,
                                              "factorised constrained apply")
        X4, Xs4 = ctx.zeros(m["ndofs"]), ctx.zeros(nglob)
        f.pcg(f.jacobi(), ctx.to_dev(bg[gid]), X4, 0.0, 0.0, 15)
        fs.pcg(fs.jacobi(), ser(bg), Xs4, 0.0, 0.0, 15)
        cmp(X4, Xs4, 1e-10, "factorised PCG solution after 15 iterations")
        cmp(X4, Xs, 1e-10, "factorised vs stored PCG solution")
    # bcast: the owner's (lowest sharing rank's) value wins
    import torch.distributed as dist
    v = ctx.to_dev(xg[gid] + rank)
    comm.bcast(v)
    gl = [None] * world
    dist.all_gather_object(gl, gid)
    low = np.full(nglob, world, int)
    for r in range(world):
        low[gl[r]] = np.minimum(low[gl[r]], r)
    assert np.array_equal(ctx.to_host(v), xg[gid] + low[gid]), "bcast: owner value did not win"
    comm.check_p2p()
    dist.barrier()   # nobody re-sets the communicator's tables while a peer is still inside its last exchange
    for h in (f, sp, fs, sps):
        h.close()
    return out


def _level_comm(ctx, like, rank, world):
    """one communicator per multigrid level (the shared-dof tables differ per order), on the transport `like` uses"""
    import torch.distributed as dist
    if like.p2p_enabled():
        return Comm(ctx, None, rank, world)          # peer memory only
    ids = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return Comm(ctx, ids[0], rank, world)


def multigrid_partitioned_vs_serial(ctx, comm, rank, world, orders=(1, 2, 3), GN=(6, 4, 4)):
    """p-multigrid across ranks == the same hierarchy on one GPU (transfers, one V-cycle, MG-preconditioned CG), with the
    serial run's eigenvalue estimates handed to both so that the two cycles are the same polynomial.  `comm` only selects
    the transport; every level gets a communicator of its own."""
    import torch.distributed as dist
    grid = partition.GRIDS[world]
    L = len(orders)
    par, ser, comms = [], [], []
    for p in orders:
        m = partition.build_part(GN, grid, rank, p, size=(1.0, 0.7, 0.4), skew=True)
        c = _level_comm(ctx, comm, rank, world)
        c.set_tables(m["ndofs"], *partition.shared_tables(m, grid, p))
        comms.append(c)
        par.append((m,) + _setup(ctx, m, p, [1, 6], c))
        ms = hex_build(*GN, p, 1.0, 0.7, 0.4, skew=True)
        ser.append((ms,) + _setup(ctx, ms, p, [1, 6]))
    gid = [partition.global_ids(x[0], GN, p) for x, p in zip(par, orders)]
    gs = [partition.global_ids(dict(lattice=x[0]["lattice"]), GN, p) for x, p in zip(ser, orders)]
    T = [Transfer(par[l][2], par[l + 1][2], basis_transfer(orders[l], orders[l + 1])) for l in range(L - 1)]
    Ts = [Transfer(ser[l][2], ser[l + 1][2], basis_transfer(orders[l], orders[l + 1])) for l in range(L - 1)]
    out = {"orders": list(orders), "global_mesh": list(GN)}

    def cmp(l, loc, serial, tol, what):
        a = ctx.to_host(loc)
        s = np.empty(ser[l][0]["ndofs"])
        s[gs[l]] = ctx.to_host(serial)
        err = float(np.max(np.abs(a - s[gid[l]])) / max(np.max(np.abs(s)), 1e-300))
        assert err <= tol, f"rank {rank}: multigrid {what}: rel err {err:.3e} > {tol:.1e}"
        return err

    rng = np.random.default_rng(5)
    errs = []
    for l in range(L - 1):
        xc, xf = rng.random(ser[l][0]["ndofs"]), rng.random(ser[l + 1][0]["ndofs"])      # lattice-indexed global vectors
        errs.append(cmp(l + 1, T[l].mult(ctx.to_dev(xc[gid[l]])), Ts[l].mult(ctx.to_dev(xc[gs[l]])), 1e-13, f"prolongation {l}"))
        errs.append(cmp(l, T[l].mult_transpose(ctx.to_dev(xf[gid[l + 1]])), Ts[l].mult_transpose(ctx.to_dev(xf[gs[l + 1]])), 1e-12,
                        f"restriction {l}"))
    out["transfer_rel_err"] = max(errs)
    mgs = Multigrid([x[2] for x in ser], Ts)
    mgs.set_coarse_solver(1e-10, 0.0, 500)
    mgs.setup()
    eig = [mgs.max_eig(l) for l in range(L)]
    mg = Multigrid([x[2] for x in par], T)
    mg.set_coarse_solver(1e-10, 0.0, 500)
    mg.setup()                                     # its own power method across the ranks ...
    own = [mg.max_eig(l) for l in range(1, L)]
    assert all(abs(a - b) <= 0.05 * b for a, b in zip(own, eig[1:])), f"rank {rank}: eigenvalue estimates {own} vs serial {eig[1:]}"
    gl = [None] * world
    dist.all_gather_object(gl, own)
    assert all(g == gl[0] for g in gl), "eigenvalue estimates differ between ranks"
    mg.setup(max_eig=eig)                          # ... then the serial estimates, so that both cycles are the same operator
    n, ns = par[-1][0]["ndofs"], ser[-1][0]["ndofs"]
    b = rng.random(ns)
    ess_s = np.zeros(ns, bool)
    ess_s[essential_dofs(ser[-1][0]["bdr_attr"], [1, 6])] = True
    b[gs[-1][ess_s]] = 0.0                          # a right-hand side of the eliminated system
    out["vcycle_rel_err"] = cmp(L - 1, mg.mult(ctx.to_dev(b[gid[-1]])), mgs.mult(ctx.to_dev(b[gs[-1]])), 1e-8, "V-cycle")
    X, Xs = ctx.zeros(n), ctx.zeros(ns)
    r, nrm = mg.pcg(ctx.to_dev(b[gid[-1]]), X, 1e-10, 0.0, 100)
    rs, nrms = mgs.pcg(ctx.to_dev(b[gs[-1]]), Xs, 1e-10, 0.0, 100)
    assert r.converged and rs.converged and abs(r.final_iter - rs.final_iter) <= 1, (r.final_iter, rs.final_iter)
    out["mgpcg_iters"] = [int(r.final_iter), int(rs.final_iter)]
    out["mgpcg_rel_err"] = cmp(L - 1, X, Xs, 1e-7, "MG-PCG solution")
    k = min(len(nrm), len(nrms), 4)
    assert np.max(np.abs(nrm[:k] - nrms[:k]) / nrms[0]) <= 1e-7
    for c in comms:
        c.check_p2p()
    dist.barrier()
    for h in [mg, mgs] + T + Ts + [x[2] for x in par + ser] + [x[1] for x in par + ser] + comms:
        h.close()
    return out
