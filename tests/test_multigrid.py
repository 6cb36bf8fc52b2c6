"""p-multigrid (fem/multigrid.hpp, fem/transfer.cpp TensorProductPRefinementTransferOperator, examples/ex26.cpp)
against golden vectors produced by the unmodified reference (tests/golden/make_golden.py `multigrid`): order-refinement
transfers with and without essential-dof constraints, one V-cycle, and CG preconditioned by the cycle."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

CASES = sorted(os.path.basename(f)[3:-4] for f in glob.glob(os.path.join(GOLDEN, "mg_*.npz")))


def load(tag):
    d = dict(np.load(os.path.join(GOLDEN, f"mg_{tag}.npz")))
    for k in list(d):
        if k in ("nlevels", "NE", "mgpcg_tol_iters", "mgpcg_tol_converged", "jacobi_pcg_tol_iters") or k.startswith(("order", "ndofs")):
            d[k] = int(d[k][0])
    return d


def rel(a, r):
    return float(np.max(np.abs(a - r)) / max(np.max(np.abs(r)), 1e-300))


@pytest.mark.parametrize("tag", CASES)
def test_transfer_matrix_matches_reference(tag):
    """the product-side builder's 1-D transfer matrix (GLL-nodal basis of the coarse order at the fine GLL nodes) is the
    DofToQuad::B the reference operator builds"""
    import b200pa
    c = load(tag)
    for l in range(c["nlevels"] - 1):
        B = b200pa.basis_transfer(c[f"order{l}"], c[f"order{l + 1}"])
        assert rel(B, c[f"PB{l}"]) <= 1e-14


def build_levels(ctx, c):
    import b200pa
    sps, forms = [], []
    for l in range(c["nlevels"]):
        p = c[f"order{l}"]
        sp = b200pa.Space(ctx, p + 1, p + 2, c["NE"], c[f"ndofs{l}"], c[f"gather_map{l}"], c[f"B{l}"], c[f"G{l}"])
        sp.geometry_from_vertices(c[f"W{l}"], c["vertices"], c["elem_vertices"])
        f = b200pa.Form(sp)
        f.assemble_diffusion(c[f"kq{l}"])
        f.assemble_mass(c[f"mq{l}"])
        f.set_essential(c[f"ess{l}"])
        sps.append(sp)
        forms.append(f)
    return sps, forms


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_gpu_transfers_match_reference(ctx, tag):
    import b200pa
    c = load(tag)
    sps, forms = build_levels(ctx, c)
    # without constraints: forms that carry no essential dofs
    plain = []
    for l in range(c["nlevels"]):
        f = b200pa.Form(sps[l])
        f.assemble_mass(np.array([1.0]))
        f.set_essential(None)
        plain.append(f)
    for l in range(c["nlevels"] - 1):
        xc, xf = ctx.to_dev(c[f"xc{l}"]), ctx.to_dev(c[f"xf{l}"])
        T = b200pa.Transfer(plain[l], plain[l + 1], c[f"PB{l}"])
        assert rel(ctx.to_host(T.mult(xc)), c[f"P_xc{l}"]) <= 1e-13
        assert rel(ctx.to_host(T.mult_transpose(xf)), c[f"Pt_xf{l}"]) <= 1e-13
        T.close()
        Tc = b200pa.Transfer(forms[l], forms[l + 1], c[f"PB{l}"])
        assert rel(ctx.to_host(Tc.mult(xc)), c[f"Pc_xc{l}"]) <= 1e-13
        assert rel(ctx.to_host(Tc.mult_transpose(xf)), c[f"Pct_xf{l}"]) <= 1e-13
        Tc.close()
    for h in plain + forms + sps:
        h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_gpu_vcycle_and_mg_pcg_match_reference(ctx, tag):
    import b200pa
    c = load(tag)
    sps, forms = build_levels(ctx, c)
    Ts = [b200pa.Transfer(forms[l], forms[l + 1], c[f"PB{l}"]) for l in range(c["nlevels"] - 1)]
    mg = b200pa.Multigrid(forms, Ts)
    mg.set_cycle(False, 1, 1)
    mg.set_coarse_solver(1e-2, 0.0, 200, jacobi=False)      # ex26: CGSolver, rel tol sqrt(1e-4), no preconditioner
    mg.setup()                                              # Chebyshev order 2, power-method eigenvalue estimates
    for l in range(1, c["nlevels"]):
        assert abs(mg.max_eig(l) - float(c[f"max_eig{l}"][0])) <= 1e-10 * float(c[f"max_eig{l}"][0])
    y = ctx.to_host(mg.mult(ctx.to_dev(c["x"])))
    # the coarse CG stops on a tolerance: its iterate is the same to rounding as long as the iteration counts agree
    assert rel(y, c["vcycle_x"]) <= 1e-9
    b = ctx.to_dev(c["B_rhs"])
    X = ctx.zeros(len(c["B_rhs"]))
    res, norms = mg.pcg(b, X, 0.0, 0.0, 3)
    assert res.final_iter == 3
    assert rel(ctx.to_host(X), c["X_mgpcg3"]) <= 1e-8
    assert np.max(np.abs(norms - c["mgpcg3_norms"][:4]) / c["mgpcg3_norms"][0]) <= 1e-8
    X = ctx.zeros(len(c["B_rhs"]))
    res, _ = mg.pcg(b, X, 1e-8, 0.0, 500)
    assert res.converged and abs(res.final_iter - c["mgpcg_tol_iters"]) <= 1
    assert res.final_iter < c["jacobi_pcg_tol_iters"]
    assert rel(ctx.to_host(X), c["X_mgpcg_tol"]) <= 1e-6
    mg.close()
    for h in Ts + forms + sps:
        h.close()
