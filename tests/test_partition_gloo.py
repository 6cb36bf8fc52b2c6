"""CPU, world_size 2 and 4 over gloo: the multi-GPU algorithm of csrc/comm.cu with the transport
swapped for gloo send/recv and the local operator for the CPU oracle.

What is pinned: the box partition + neighbour tables (b200pa/partition.py), the
host-side CSR of sources in ascending rank order and the owner mask (b200pa_comm_build_tables,
the same C++ that feeds the NCCL path), and the algorithm itself — PCG on consistent L-vectors with
one symmetric shared-dof exchange per apply and owner-masked dots — against the SERIAL oracle on
the same global mesh: operator apply to 1e-12, PCG solution to 1e-10, identical iteration counts.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "cardiac-ablation-ecm2_b200"))

GN, P_ORDER = (4, 3, 2), 2
CHEB_ORDER, CHEB_MAX_EIG, CHEB_ITS = 3, 1.45, 5


def kfun(xyz):
    return np.sin(8 * np.pi * xyz[:, 0]) * np.cos(6 * np.pi * xyz[:, 1]) * np.sin(4 * np.pi * xyz[:, 2]) + 2.0


def make_local(m, p):
    """oracle operator of one mesh (global or part): trilinear geometry -> q-data -> operator"""
    import b200pa
    import orc
    b = b200pa.basis(p)
    Q, ne = p + 2, m["ne"]
    # geometry on the host: J from vertices (same formula as k_geometry_trilinear)
    xi = np.polynomial.legendre.leggauss(Q)[0] * 0.5 + 0.5
    V = m["vertices"].reshape(-1, 3)[m["elem_vertices"].reshape(-1, 8)]          # [ne,8,3]
    ci = np.array([0, 1, 1, 0, 0, 1, 1, 0]); cj = np.array([0, 0, 1, 1, 0, 0, 1, 1]); ck = np.array([0, 0, 0, 0, 1, 1, 1, 1])
    qx, qy, qz = np.meshgrid(xi, xi, xi, indexing="ij")
    qx, qy, qz = [a.transpose(2, 1, 0).ravel() for a in (qx, qy, qz)]            # q = qx + Q(qy + Q qz)
    def n1(c, t): return np.where(c[:, None] == 1, t[None, :], 1 - t[None, :])  # [8,NQ]
    def g1(c): return np.where(c == 1, 1.0, -1.0)[:, None]
    N = n1(ci, qx) * n1(cj, qy) * n1(ck, qz)
    dN = np.stack([g1(ci) * n1(cj, qy) * n1(ck, qz), n1(ci, qx) * g1(cj) * n1(ck, qz), n1(ci, qx) * n1(cj, qy) * g1(ck)], 0)  # [3,8,NQ]
    Jm = np.einsum("evr,cvq->eqrc", V, dN)                                        # J[e,q,row,col]
    Xq = np.einsum("evr,vq->eqr", V, N).reshape(-1, 3)
    J = np.ascontiguousarray(Jm.transpose(0, 3, 2, 1)).ravel()                    # [e][col][row][q]
    detJ = np.linalg.det(Jm).ravel()
    kq = kfun(Xq)
    mq = 3.0 + Xq[:, 0] * Xq[:, 1]
    pa_d = orc.diffusion_setup(Q, ne, b["W"], J, kq)
    pa_m = orc.mass_setup(Q, ne, b["W"], detJ, mq)
    return b, pa_d, pa_m


def worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import b200pa
    import orc
    from b200pa import partition
    p = P_ORDER
    grid = partition.GRIDS[world]
    m = partition.build_part(GN, grid, rank, p, size=(1.0, 0.7, 0.4), skew=True)
    nd = m["ndofs"]
    nbr, offs, ldofs = partition.shared_tables(m, grid, p)
    sh_ldof, sh_off, sh_src, own = b200pa.comm_build_tables(rank, nd, nbr, offs, ldofs)
    gid = partition.global_ids(m, GN, p)
    b, pa_d, pa_m = make_local(m, p)
    ess = b200pa.essential_dofs(m["bdr_attr"], [1, 6])
    op = orc.Operator(p + 1, p + 2, m["ne"], nd, m["gather_map"], b["B"], b["G"], pa_d, pa_m, ess)

    def exchange_sum(y):
        """comm.cu::exchange with gloo as the transport: pack, send/recv per neighbour, unpack"""
        send = y[ldofs].copy()
        recv = np.zeros_like(send)
        reqs = []
        for k, r in enumerate(nbr):
            sl = slice(offs[k], offs[k + 1])
            reqs.append(dist.isend(torch.from_numpy(send[sl].copy()), int(r)))
        for k, r in enumerate(nbr):
            buf = torch.zeros(int(offs[k + 1] - offs[k]), dtype=torch.float64)
            dist.recv(buf, int(r))
            recv[offs[k]:offs[k + 1]] = buf.numpy()
        for q in reqs:
            q.wait()
        for i, l in enumerate(sh_ldof):                      # k_unpack: ascending rank order
            v = 0.0
            for j in range(sh_off[i], sh_off[i + 1]):
                v += y[l] if sh_src[j] < 0 else recv[sh_src[j]]
            y[l] = v
        return y

    def dot(a, c):
        t = torch.tensor([float(np.dot(a[own == 1], c[own == 1]))], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    # local pieces of ConstrainedOperator::Mult on a consistent L-vector
    un = orc.Operator(p + 1, p + 2, m["ne"], nd, m["gather_map"], b["B"], b["G"], pa_d, pa_m, None)
    essmask = np.zeros(nd, bool)
    essmask[ess] = True

    def A(x):
        z = x.copy()
        z[essmask] = 0.0
        y = exchange_sum(un.mult(z))
        y[essmask] = x[essmask]
        return y

    rng = np.random.default_rng(5)
    nglob = (GN[0] * p + 1) * (GN[1] * p + 1) * (GN[2] * p + 1)
    xg, bg = rng.random(nglob), rng.random(nglob)
    x = xg[gid]                                               # consistent by construction
    y = A(x)
    # Jacobi diagonal: local AbsMultTranspose + exchange (ParBilinearForm::AssembleDiagonal)
    diag = exchange_sum(un.diag())
    dinv = 1.0 / diag
    dinv[essmask] = 1.0
    # PCG on consistent L-vectors (csrc/b200pa.cu::b200pa_pcg_solve with a comm)
    rhs = bg[gid]
    X = np.zeros(nd)
    r = rhs - A(X)
    z = dinv * r
    d = z.copy()
    nom = dot(d, r)
    norms = [nom]
    zz = A(d)
    den = dot(zz, d)
    its = 12
    for i in range(1, its + 1):
        alpha = nom / den
        X += alpha * d
        r -= alpha * zz
        z = dinv * r
        betanom = dot(r, z)
        norms.append(betanom)
        if i == its:
            break
        d = z + (betanom / nom) * d
        zz = A(d)
        den = dot(d, zz)
        nom = betanom
    # the same loop preconditioned by the Chebyshev smoother (b200pa_pcg_solve_chebyshev with a comm): every term of
    # the polynomial applies the partitioned operator, dots stay owner-masked
    coeffs = orc.chebyshev_coeffs(CHEB_ORDER, CHEB_MAX_EIG)

    def cheb(rr):
        res, zc = rr.copy(), np.zeros(nd)
        for k in range(CHEB_ORDER):
            if k > 0:
                res = A(res)
            res = dinv * res
            zc = zc + coeffs[k] * res
        return zc

    Xc = np.zeros(nd)
    r = rhs - A(Xc)
    z = cheb(r)
    d = z.copy()
    nom = dot(d, r)
    cnorms = [nom]
    zz = A(d)
    den = dot(zz, d)
    for i in range(1, CHEB_ITS + 1):
        alpha = nom / den
        Xc += alpha * d
        r -= alpha * zz
        z = cheb(r)
        betanom = dot(r, z)
        cnorms.append(betanom)
        if i == CHEB_ITS:
            break
        d = z + (betanom / nom) * d
        zz = A(d)
        den = dot(d, zz)
        nom = betanom
    ret[rank] = dict(gid=gid, y=y, diag=diag, X=X, norms=np.array(norms), own=own, n_shared=len(sh_ldof), Xc=Xc,
                     cnorms=np.array(cnorms))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_matches_serial(world):
    import b200pa
    import orc
    port = 29600 + world + (os.getpid() % 200)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(worker, args=(world, port, ret), nprocs=world, join=True)
    # serial oracle on the global mesh
    p = P_ORDER
    m = b200pa.hex_build(*GN, p, 1.0, 0.7, 0.4, skew=True)
    b, pa_d, pa_m = make_local(m, p)
    ess = b200pa.essential_dofs(m["bdr_attr"], [1, 6])
    op = orc.Operator(p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"], pa_d, pa_m, ess)
    lat = m["lattice"].reshape(-1, 3).astype(np.int64)
    nx, ny = GN[0] * p + 1, GN[1] * p + 1
    g_of_l = lat[:, 0] + nx * (lat[:, 1] + ny * lat[:, 2])      # serial L-dof -> lattice id
    rng = np.random.default_rng(5)
    nglob = m["ndofs"]
    xg, bg = rng.random(nglob), rng.random(nglob)
    y_ser = np.empty(nglob); d_ser = np.empty(nglob); X_ser = np.empty(nglob)
    y_ser[g_of_l] = op.constrained_mult(xg[g_of_l])
    un = orc.Operator(p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"], pa_d, pa_m, None)
    d_ser[g_of_l] = un.diag()
    Xs, it, conv, fn, norms = op.pcg(op.jacobi_dinv(), bg[g_of_l], np.zeros(nglob), 0.0, 0.0, 12)
    X_ser[g_of_l] = Xs
    Xcs, itc, _, _, cnorms = op.pcg_chebyshev(op.jacobi_dinv(), CHEB_ORDER, CHEB_MAX_EIG, bg[g_of_l], np.zeros(nglob), 0.0, 0.0, CHEB_ITS)
    Xc_ser = np.empty(nglob)
    Xc_ser[g_of_l] = Xcs
    assert itc == CHEB_ITS
    covered = np.zeros(nglob, int)
    owned = np.zeros(nglob, int)
    for r in range(world):
        o = ret[r]
        gid = o["gid"]
        covered[gid] += 1
        owned[gid[o["own"] == 1]] += 1
        assert o["n_shared"] > 0
        assert np.max(np.abs(o["y"] - y_ser[gid])) <= 1e-12 * np.max(np.abs(y_ser))
        assert np.max(np.abs(o["diag"] - d_ser[gid])) <= 1e-12 * np.max(np.abs(d_ser))
        assert np.max(np.abs(o["X"] - X_ser[gid])) <= 1e-10 * np.max(np.abs(X_ser))
        assert np.max(np.abs(o["norms"] - norms) / norms) <= 1e-9
        assert np.max(np.abs(o["Xc"] - Xc_ser[gid])) <= 1e-10 * np.max(np.abs(Xc_ser))
        assert np.max(np.abs(o["cnorms"] - cnorms) / cnorms) <= 1e-9
    assert covered.min() >= 1 and np.all(owned == 1)            # every dof owned exactly once
    # bit-identical copies of shared dofs on all sharers (ascending-rank summation order)
    for r in range(world):
        for s in range(r + 1, world):
            a, c = ret[r], ret[s]
            common, ia, ic = np.intersect1d(a["gid"], c["gid"], return_indices=True)
            assert np.array_equal(a["y"][ia], c["y"][ic]) and np.array_equal(a["X"][ia], c["X"][ic])


def test_build_tables_rejects_bad_input():
    import b200pa
    with pytest.raises(b200pa.B200paError, match="ascending"):
        b200pa.comm_build_tables(0, 4, [2, 1], [0, 1, 2], [0, 1])
    with pytest.raises(b200pa.B200paError, match="itself"):
        b200pa.comm_build_tables(1, 4, [1], [0, 1], [0])
    with pytest.raises(b200pa.B200paError, match="out of range"):
        b200pa.comm_build_tables(0, 4, [1], [0, 1], [7])
    # no neighbours: everything owned, nothing shared
    sh_ldof, sh_off, sh_src, own = b200pa.comm_build_tables(0, 5, [], [0], [])
    assert len(sh_ldof) == 0 and own.tolist() == [1] * 5
