"""Pins oracle/pa_oracle.c against outputs of the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py from oracle/_ref/ref_driver).  CPU only.

The reference build has no FMA and the oracle is compiled with -ffp-contract=off and the same
loop order, so agreement is expected to be bit-for-bit; the asserts allow 2e-15 relative so
that a different libm / compiler version on another box does not turn this red.
"""
import numpy as np

import orc

RTOL = 2e-15


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.max(np.abs(b)), 1e-300)
    err = np.max(np.abs(a - b)) / scale
    assert err <= rtol, f"rel err {err:.3e}"


def make_op(c, diff=True, mass=True, ess=True):
    return orc.Operator(c["D1D"], c["Q1D"], c["NE"], c["ndofs"], c["gather_map"], c["B"], c["G"],
                        c["pa_diff"] if diff else None, c["pa_mass"] if mass else None,
                        c["ess"] if ess else None)


def test_restriction_tables(case):
    op = make_op(case)
    assert np.array_equal(op.offsets, case["offsets"])
    assert np.array_equal(op.indices, case["indices"])


def test_gather_scatter(case):
    nd = case["D1D"] ** 3
    xE = orc.restrict_mult(case["NE"], nd, case["gather_map"], case["x"])
    assert np.array_equal(xE, case["xE"])
    y = orc.restrict_mult_transpose(case["ndofs"], case["offsets"], case["indices"], case["yE_diff"])
    assert np.array_equal(y, case["y_diff"])


def test_setup(case):
    D = orc.diffusion_setup(case["Q1D"], case["NE"], case["W"], case["J"], case["kq"])
    close(D, case["pa_diff"])
    v = orc.mass_setup(case["Q1D"], case["NE"], case["W"], case["detJ"], case["mq"])
    close(v, case["pa_mass"])


def test_apply_E(case):
    c = case
    y = orc.diffusion_apply(c["NE"], c["D1D"], c["Q1D"], c["B"], c["G"], c["pa_diff"], c["xE"])
    close(y, c["yE_diff"])
    y = orc.mass_apply(c["NE"], c["D1D"], c["Q1D"], c["B"], c["pa_mass"], c["xE"])
    close(y, c["yE_mass"])
    y = orc.diffusion_apply(c["NE"], c["D1D"], c["Q1D"], c["B"], c["G"], c["pa_diff"], c["xE"])
    y = orc.mass_apply(c["NE"], c["D1D"], c["Q1D"], c["B"], c["pa_mass"], c["xE"], y)
    close(y, c["yE"])


def test_apply_L_and_diag(case):
    op = make_op(case)
    close(op.mult(case["x"]), case["y"])
    close(op.diag(), case["diag"])
    c = case
    close(orc.diffusion_diag(c["NE"], c["D1D"], c["Q1D"], c["B"], c["G"], c["pa_diff"]), c["dE_diff"])
    close(orc.mass_diag(c["NE"], c["D1D"], c["Q1D"], c["B"], c["pa_mass"]), c["dE_mass"])


def test_constrained_and_rhs(case):
    op = make_op(case)
    close(op.constrained_mult(case["x"]), case["y_constrained"])
    B = op.eliminate_rhs(case["x0_L"], case["b_L"])
    close(B, case["B_rhs"])
    dinv = op.jacobi_dinv()
    close(orc.jacobi_mult(dinv, case["x"]), case["jacobi_z"])


def test_pcg(case):
    op = make_op(case)
    dinv = op.jacobi_dinv()
    for k in (1, 2):
        x, it, conv, fn, norms = op.pcg(dinv, case["B_rhs"], case["X0"], 0.0, 0.0, k)
        close(x, case[f"X_pcg{k}"], 1e-14)
    kmax = len(case["pcg_norms"]) - 1
    x, it, conv, fn, norms = op.pcg(dinv, case["B_rhs"], case["X0"], 0.0, 0.0, kmax)
    assert it == kmax
    close(x, case[f"X_pcg{kmax}"], 1e-13)
    close(norms, case["pcg_norms"], 1e-12)
    x, it, conv, fn, norms = op.pcg(dinv, case["B_rhs"], case["X0"], 1e-8, 0.0, 5000)
    assert it == int(case["pcg_tol_iters"][0])
    assert conv == bool(case["pcg_tol_converged"][0])
    close(fn, case["pcg_tol_final_norm"][0], 1e-9)
    close(x, case["X_pcg_tol"], 1e-12)


def test_qpoint_ops(case):
    c = case
    close(orc.qvalues(c["NE"], c["D1D"], c["Q1D"], c["B"], c["xE"]), c["xq_values"])
    close(orc.qphysgrad(c["NE"], c["D1D"], c["Q1D"], c["B"], c["G"], c["J"], c["xE"]), c["xq_physgrad"], 1e-14)
    bE = orc.domain_lf(c["NE"], c["D1D"], c["Q1D"], c["B"], c["detJ"], c["W"], c["lf_fq"])
    b = orc.restrict_mult_transpose(c["ndofs"], c["offsets"], c["indices"], bE)
    close(b, c["lf_b"])


def test_chebyshev_and_power_method(case):
    """OperatorChebyshevSmoother (linalg/solvers.cpp:455-657) and the power method behind its
    eigenvalue estimate (linalg/operator.cpp:871-928) against the reference's own outputs"""
    op = make_op(case)
    dinv = op.jacobi_dinv()
    lam_ref = float(case["cheb_max_eig"][0])
    close(op.power_method(dinv, case["cheb_v0"]), lam_ref, 1e-13)
    for order in range(1, 6):
        close(op.chebyshev_mult(dinv, order, lam_ref, case["x"]), case[f"cheb_z{order}"], 1e-13)
    x, it, conv, fn, norms = op.pcg_chebyshev(dinv, 3, lam_ref, case["B_rhs"], case["X0"], 0.0, 0.0, 4)
    assert it == 4
    close(x, case["X_cheb3_pcg4"], 1e-12)
    close(norms, case["cheb3_pcg_norms"], 1e-11)
    x, it, conv, fn, norms = op.pcg_chebyshev(dinv, 3, lam_ref, case["B_rhs"], case["X0"], 1e-8, 0.0, 5000)
    assert it == int(case["cheb3_tol_iters"][0]) and conv == bool(case["cheb3_tol_converged"][0])
    close(x, case["X_cheb3_tol"], 1e-10)


def test_factorised_qdata_reproduces_reference_pa_data(case):
    """the factorisation behind b200pa_form_set_factorised, stated in numpy: on the (affine) golden meshes the
    reference's pa_data D[Q^3,6,NE] = (w_q c_q) x (adj(J) adj(J)^T / det J taken once per element), to rounding"""
    c = case
    NQ, NE = c["Q1D"] ** 3, c["NE"]
    J = c["J"].reshape(NE, 9, NQ)                      # J(q, row + 3 col, e), q fastest
    assert np.max(np.abs(J - J[:, :, :1])) <= 1e-13 * np.max(np.abs(J))      # affine: J constant over an element
    J0 = J[:, :, 0]
    J11, J21, J31, J12, J22, J32, J13, J23, J33 = (J0[:, k] for k in range(9))
    det = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13)
    A = np.array([[J22 * J33 - J23 * J32, J32 * J13 - J12 * J33, J12 * J23 - J22 * J13],
                  [J31 * J23 - J21 * J33, J11 * J33 - J13 * J31, J21 * J13 - J11 * J23],
                  [J21 * J32 - J31 * J22, J31 * J12 - J11 * J32, J11 * J22 - J12 * J21]])          # [3,3,NE]
    G = np.einsum("ike,jke->ije", A, A) / det                                                       # adj adj^T / det
    geo6 = np.stack([G[0, 0], G[0, 1], G[0, 2], G[1, 1], G[1, 2], G[2, 2]], axis=1)                 # [NE,6]
    kq = c["kq"] if c["kq"].size > 1 else np.full(NQ * NE, c["kq"][0])
    cq = (np.tile(c["W"], NE) * kq).reshape(NE, 1, NQ)
    D = (cq * geo6[:, :, None]).ravel()
    close(D, c["pa_diff"], 1e-13)
