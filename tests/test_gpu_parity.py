"""GPU parity (run on the B200 box: pytest -m gpu): every entry point of the C ABI against
 (a) the golden fixtures = outputs of the unmodified reference, and
 (b) the CPU oracle (oracle/pa_oracle.c) on the same inputs.
Tolerances are the north star's: 1e-12 relative per operator apply, 1e-10 on the solution
after a fixed number of PCG iterations, iteration counts to a tolerance within +-1."""
import numpy as np
import pytest

import b200pa
import orc

pytestmark = pytest.mark.gpu

TOL_APPLY = 1e-12
TOL_PCG = 1e-10


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def close(a, b, tol=TOL_APPLY):
    e = relerr(a, b)
    assert e <= tol, f"rel err {e:.3e} > {tol:.1e}"


class Dev:
    """a golden case uploaded to the GPU"""

    def __init__(self, ctx, c):
        self.c, self.ctx = c, ctx
        self.D, self.Q, self.NE, self.nd = c["D1D"], c["Q1D"], c["NE"], c["ndofs"]
        self._cache = {}

    def __getitem__(self, k):
        if k not in self._cache:
            self._cache[k] = self.ctx.to_dev(self.c[k])
        return self._cache[k]

    def space(self, geometry="given"):
        c = self.c
        sp = b200pa.Space(self.ctx, self.D, self.Q, self.NE, self.nd, c["gather_map"], c["B"], c["G"])
        if geometry == "given":
            sp.set_geometry(c["W"], self["J"], self["detJ"])
        elif geometry == "vertices":
            sp.geometry_from_vertices(c["W"], c["vertices"], c["elem_vertices"])
        return sp

    def form(self, sp, diff=True, mass=True, ess=True, assemble=False):
        f = b200pa.Form(sp)
        if assemble:
            f.assemble_diffusion(self.c["kq"] if diff else None)
            f.assemble_mass(self.c["mq"] if mass else None)
        else:
            f.set_pa_data(self["pa_diff"] if diff else None, self["pa_mass"] if mass else None)
        f.set_essential(self.c["ess"] if ess else None)
        return f


@pytest.fixture
def dev(ctx, case):
    return Dev(ctx, case)


def test_restriction(ctx, dev):
    c = dev.c
    nd3 = dev.D ** 3
    xE = ctx.restrict_mult(dev.NE, nd3, dev["gather_map"], dev["x"])
    assert np.array_equal(ctx.to_host(xE), c["xE"])                      # pure data movement: bit-exact
    y = ctx.restrict_mult_transpose(dev.nd, dev["offsets"], dev["indices"], dev["yE_diff"])
    assert np.array_equal(ctx.to_host(y), c["y_diff"])                   # same summation order: bit-exact
    ya = ctx.restrict_mult_transpose(dev.nd, dev["offsets"], dev["indices"], dev["dE_diff"], abs_=True)
    close(ctx.to_host(ya), orc.restrict_mult_transpose(dev.nd, c["offsets"], c["indices"], c["dE_diff"], True), 0.0)


def test_space_tables_match_reference(ctx, dev):
    sp = dev.space(geometry=None)
    assert np.array_equal(sp.offsets(), dev.c["offsets"])
    assert np.array_equal(sp.indices(), dev.c["indices"])
    sp.close()


def test_setup(ctx, dev):
    c = dev.c
    D = ctx.diffusion_setup(dev.Q, dev.NE, dev["W"], dev["J"], dev["kq"])
    close(ctx.to_host(D), c["pa_diff"])
    close(ctx.to_host(D), orc.diffusion_setup(dev.Q, dev.NE, c["W"], c["J"], c["kq"]))
    v = ctx.mass_setup(dev.Q, dev.NE, dev["W"], dev["detJ"], dev["mq"])
    close(ctx.to_host(v), c["pa_mass"])


def test_geometry_from_vertices(ctx, dev):
    sp = dev.space(geometry="vertices")
    close(sp.J(), dev.c["J"])
    close(sp.detJ(), dev.c["detJ"])
    sp.close()


def test_apply_E(ctx, dev):
    c = dev.c
    n = dev.NE * dev.D ** 3
    y = ctx.diffusion_apply(dev.NE, dev.D, dev.Q, c["B"], c["G"], dev["pa_diff"], dev["xE"], ctx.zeros(n))
    close(ctx.to_host(y), c["yE_diff"])
    y = ctx.mass_apply(dev.NE, dev.D, dev.Q, c["B"], dev["pa_mass"], dev["xE"], ctx.zeros(n))
    close(ctx.to_host(y), c["yE_mass"])
    # accumulate semantics (AddMultPA): y += ...
    y = ctx.diffusion_apply(dev.NE, dev.D, dev.Q, c["B"], c["G"], dev["pa_diff"], dev["xE"], ctx.zeros(n))
    y = ctx.mass_apply(dev.NE, dev.D, dev.Q, c["B"], dev["pa_mass"], dev["xE"], y)
    close(ctx.to_host(y), c["yE"])
    close(ctx.to_host(y), orc.mass_apply(dev.NE, dev.D, dev.Q, c["B"], c["pa_mass"], c["xE"],
                                         orc.diffusion_apply(dev.NE, dev.D, dev.Q, c["B"], c["G"], c["pa_diff"], c["xE"])))


def test_diag_E(ctx, dev):
    c = dev.c
    n = dev.NE * dev.D ** 3
    d = ctx.diffusion_diag(dev.NE, dev.D, dev.Q, c["B"], c["G"], dev["pa_diff"], ctx.zeros(n))
    close(ctx.to_host(d), c["dE_diff"])
    d = ctx.mass_diag(dev.NE, dev.D, dev.Q, c["B"], dev["pa_mass"], ctx.zeros(n))
    close(ctx.to_host(d), c["dE_mass"])


@pytest.mark.parametrize("assemble", [False, True])
def test_form_mult_and_diag(ctx, dev, assemble):
    c = dev.c
    sp = dev.space()
    f = dev.form(sp, assemble=assemble)
    close(ctx.to_host(f.mult(dev["x"])), c["y"])
    close(ctx.to_host(f.assemble_diagonal()), c["diag"])
    close(ctx.to_host(f.constrained_mult(dev["x"])), c["y_constrained"])
    # host-buffer entry point
    yh = np.zeros(dev.nd)
    f.mult_host(np.ascontiguousarray(c["x"]), yh)
    close(yh, c["y"])
    # single-integrator forms
    fd = dev.form(sp, diff=True, mass=False, assemble=assemble)
    close(ctx.to_host(fd.mult(dev["x"])), c["y_diff"])
    fm = dev.form(sp, diff=False, mass=True, assemble=assemble)
    close(ctx.to_host(fm.mult(dev["x"])), c["y_mass"])
    for h in (f, fd, fm):
        h.close()
    sp.close()


def test_rhs_and_jacobi(ctx, dev):
    c = dev.c
    sp = dev.space()
    f = dev.form(sp)
    b = ctx.to_dev(c["b_L"])
    f.eliminate_rhs(dev["x0_L"], b)
    close(ctx.to_host(b), c["B_rhs"])
    dinv = f.jacobi()
    close(ctx.to_host(ctx.jacobi_mult(dinv, dev["x"])), c["jacobi_z"])
    f.close()
    sp.close()


def test_blas1(ctx, dev):
    c = dev.c
    x, y = dev["x"], dev["y"]
    d = ctx.dot(x, y)
    ref = orc.dot(c["x"], c["y"])
    assert abs(d - ref) <= 1e-13 * np.sum(np.abs(c["x"] * c["y"]))
    assert d == ctx.dot(x, y)                                            # deterministic run to run
    close(ctx.to_host(ctx.add(x, -0.37, y)), c["x"] - 0.37 * c["y"], 1e-15)


def test_pcg(ctx, dev):
    c = dev.c
    sp = dev.space()
    f = dev.form(sp)
    dinv = f.jacobi()
    kmax = len(c["pcg_norms"]) - 1
    for k in (1, 2, kmax):
        x = ctx.to_dev(c["X0"])
        res, norms = f.pcg(dinv, dev["B_rhs"], x, 0.0, 0.0, k)
        assert res.final_iter == k and not res.converged
        close(ctx.to_host(x), c[f"X_pcg{k}"], TOL_PCG)
    close(norms, c["pcg_norms"], 1e-9)
    # to tolerance: iteration count within +-1 of the reference, same flags
    x = ctx.to_dev(c["X0"])
    res, norms = f.pcg(dinv, dev["B_rhs"], x, 1e-8, 0.0, 5000)
    assert abs(res.final_iter - int(c["pcg_tol_iters"][0])) <= 1
    assert bool(res.converged) == bool(c["pcg_tol_converged"][0])
    if res.final_iter == int(c["pcg_tol_iters"][0]):
        # the residual at the stopping iteration carries the accumulated rounding of all iterations
        assert abs(res.final_norm - c["pcg_tol_final_norm"][0]) <= 1e-3 * c["pcg_tol_final_norm"][0]
    close(ctx.to_host(x), c["X_pcg_tol"], 1e-7)   # both are 1e-8-converged iterates
    # host-buffer entry point, same answer as the device one
    xh = np.ascontiguousarray(c["X0"]).copy()
    res2, _ = f.pcg(dinv, np.ascontiguousarray(c["B_rhs"]), xh, 0.0, 0.0, kmax, host=True)
    assert res2.final_iter == kmax
    close(xh, c[f"X_pcg{kmax}"], TOL_PCG)
    # against the oracle with the same Jacobi
    op = orc.Operator(dev.D, dev.Q, dev.NE, dev.nd, c["gather_map"], c["B"], c["G"], c["pa_diff"], c["pa_mass"], c["ess"])
    xo, it, conv, fn, _ = op.pcg(op.jacobi_dinv(), c["B_rhs"], c["X0"], 0.0, 0.0, kmax)
    close(xh, xo, TOL_PCG)
    f.close()
    sp.close()


def test_pcg_edge_cases(ctx, dev):
    c = dev.c
    sp = dev.space()
    f = dev.form(sp)
    dinv = f.jacobi()
    # already converged: x = 1e-8-converged solution, absolute tolerance above its residual
    # -> (Br,r) <= r0 at iteration 0 (linalg/solvers.cpp:919-927)
    x = ctx.to_dev(c["X_pcg_tol"])
    res, norms = f.pcg(dinv, dev["B_rhs"], x, 0.0, 1.0, 50)
    assert res.final_iter == 0 and res.converged and len(norms) == 1
    # zero rhs and zero guess: (Br,r) = 0 <= r0 -> converged at iteration 0
    z = ctx.zeros(dev.nd)
    res, _ = f.pcg(dinv, ctx.zeros(dev.nd), z, 1e-12, 0.0, 10)
    assert res.final_iter == 0 and res.converged and res.final_norm == 0.0
    # max_iter smaller than needed: not converged, final_iter == max_iter
    x = ctx.to_dev(c["X0"])
    res, _ = f.pcg(dinv, dev["B_rhs"], x, 1e-14, 0.0, 3)
    assert res.final_iter == 3 and not res.converged
    f.close()
    sp.close()


def test_qpoint_ops(ctx, dev):
    c = dev.c
    close(ctx.to_host(ctx.qvalues(dev.NE, dev.D, dev.Q, c["B"], dev["xE"])), c["xq_values"])
    close(ctx.to_host(ctx.qphysgrad(dev.NE, dev.D, dev.Q, c["B"], c["G"], dev["J"], dev["xE"])), c["xq_physgrad"])
    n = dev.NE * dev.D ** 3
    bE = ctx.domain_lf(dev.NE, dev.D, dev.Q, c["B"], dev["detJ"], dev["W"], dev["lf_fq"], ctx.zeros(n))
    b = ctx.restrict_mult_transpose(dev.nd, dev["offsets"], dev["indices"], bE)
    close(ctx.to_host(b), c["lf_b"])
    # fused L-vector variants
    sp = dev.space()
    close(ctx.to_host(sp.qvalues(dev["x"])), c["xq_values"])
    close(ctx.to_host(sp.qphysgrad(dev["x"])), c["xq_physgrad"])
    close(ctx.to_host(sp.domain_lf(dev["lf_fq"])), c["lf_b"])
    k = ctx.to_host(sp.coeff_linear(0.5, 0.02, 37.0, dev["x"]))
    close(k, 0.5 * (1.0 + 0.02 * (c["xq_values"] - 37.0)), 1e-14)
    g = c["xq_physgrad"].reshape(-1, 3)
    s = ctx.to_dev(c["xq_values"])
    j = ctx.to_host(sp.joule(dev["x"], s, 1.5))
    close(j, c["xq_values"] * (g * g).sum(1) + 1.5)
    # unfused coefficient kernel
    cq = ctx.coeff_eval(2, dev.NE * dev.Q ** 3, 1.5, 0.0, 0.0, None, s, dev["xq_physgrad"])
    close(ctx.to_host(cq), c["xq_values"] * (g * g).sum(1) + 1.5, 1e-14)
    sp.close()


def test_unsupported_rejected(ctx):
    """no fallback kernel: (D1D,Q1D) outside the supported set is an error, not a slow path"""
    gm = np.zeros(8, np.int32)
    with pytest.raises(b200pa.B200paError, match="unsupported"):
        b200pa.Space(ctx, 2, 2, 1, 8, gm, np.zeros(4), np.zeros(4))
    with pytest.raises(b200pa.B200paError, match="unsupported"):
        b200pa.Space(ctx, 9, 10, 1, 8, gm, np.zeros(90), np.zeros(90))
    bad = np.array([0, 1, 2, 3, 4, 5, 6, -1], np.int32)
    with pytest.raises(b200pa.B200paError, match="negative"):
        b200pa.Space(ctx, 2, 3, 1, 8, bad, np.zeros(6), np.zeros(6))


def test_empty_space(ctx):
    """ragged/empty input: a rank that owns no elements"""
    b = b200pa.basis(2)
    sp = b200pa.Space(ctx, 3, 4, 0, 0, np.zeros(0, np.int32), b["B"], b["G"])
    sp.close()


def test_form_from_vertices_matches_reference(ctx, dev):
    """J-free path: geometry kept as vertices, D assembled by k_diffusion_setup_trilinear, q-point gradients
    rebuilt from vertices — against the reference's outputs (which used its own stored J)."""
    c = dev.c
    sp = dev.space(geometry="vertices")
    f = dev.form(sp, assemble=True)
    close(ctx.to_host(f.mult(dev["x"])), c["y"])
    close(ctx.to_host(f.assemble_diagonal()), c["diag"])
    close(ctx.to_host(sp.qphysgrad(dev["x"])), c["xq_physgrad"])
    close(ctx.to_host(sp.domain_lf(dev["lf_fq"])), c["lf_b"])
    import ctypes as C
    pd = b200pa.lib().b200pa_form_pa_diff(f.h)
    out = np.empty(6 * dev.NE * dev.Q ** 3)
    b200pa.check(b200pa.lib().b200pa_ctx_download(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(pd), C.c_size_t(out.nbytes)))
    close(out, c["pa_diff"])
    f.close()
    sp.close()


@pytest.mark.parametrize("p,dims", [(1, (9, 7, 5)), (2, (7, 5, 3)), (3, (5, 3, 3)), (4, (3, 3, 2)), (5, (3, 2, 2)), (6, (2, 2, 3))])
def test_midsize_against_oracle(ctx, p, dims):
    """product-side builder + device geometry + fused path vs the CPU oracle on meshes whose element count
    is not a multiple of the kernel's batch size (tail batches), skewed geometry, q-data coefficients,
    essential dofs on two faces: apply, constrained apply, diagonal, EliminateRHS, 6 PCG iterations."""
    m = b200pa.hex_build(*dims, p, 1.0, 0.8, 0.6, skew=True)
    b = b200pa.basis(p)
    D, Q, ne, nd = p + 1, p + 2, m["ne"], m["ndofs"]
    sp = b200pa.Space(ctx, D, Q, ne, nd, m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    rng = np.random.default_rng(100 + p)
    nq = ne * Q ** 3
    kq, mq = 0.5 + rng.random(nq), 1.0 + rng.random(nq)
    ess = b200pa.essential_dofs(m["bdr_attr"], [2, 3])
    f = b200pa.Form(sp)
    f.assemble_diffusion(kq)
    f.assemble_mass(mq)
    f.set_essential(ess)
    J, detJ = sp.J(), sp.detJ()
    pa_d = orc.diffusion_setup(Q, ne, b["W"], J, kq)
    pa_m = orc.mass_setup(Q, ne, b["W"], detJ, mq)
    op = orc.Operator(D, Q, ne, nd, m["gather_map"], b["B"], b["G"], pa_d, pa_m, ess)
    un = orc.Operator(D, Q, ne, nd, m["gather_map"], b["B"], b["G"], pa_d, pa_m, None)
    x, rhs, x0 = rng.random(nd), rng.random(nd), rng.random(nd)
    xd = ctx.to_dev(x)
    close(ctx.to_host(f.mult(xd)), un.mult(x))
    close(ctx.to_host(f.constrained_mult(xd)), op.constrained_mult(x))
    close(ctx.to_host(f.assemble_diagonal()), un.diag())
    bd = ctx.to_dev(rhs)
    f.eliminate_rhs(ctx.to_dev(x0), bd)
    close(ctx.to_host(bd), op.eliminate_rhs(x0, rhs))
    X = ctx.to_dev(x0)
    res, norms = f.pcg(f.jacobi(), bd, X, 0.0, 0.0, 6)
    xo, it, conv, fn, no = op.pcg(op.jacobi_dinv(), op.eliminate_rhs(x0, rhs), x0, 0.0, 0.0, 6)
    assert res.final_iter == it == 6
    close(ctx.to_host(X), xo, TOL_PCG)
    close(norms, no, 1e-9)
    f.close()
    sp.close()


def test_chebyshev_smoother_and_pcg(ctx, dev):
    """OperatorChebyshevSmoother (linalg/solvers.cpp:455-657) - coefficients, the power-method eigenvalue
    estimate, the smoother and PCG preconditioned with it against the reference's own outputs and the oracle"""
    c = dev.c
    sp = dev.space()
    f = dev.form(sp)
    dinv = f.jacobi()
    lam_ref = float(c["cheb_max_eig"][0])
    lam = f.power_method(dinv, ctx.to_dev(c["cheb_v0"]))
    assert abs(lam - lam_ref) <= 1e-12 * lam_ref
    assert abs(f.power_method(dinv, ctx.to_dev(b200pa.randomize(dev.nd, 12345))) - lam_ref) <= 1e-12 * lam_ref
    for order in range(1, 6):
        close(b200pa.chebyshev_coeffs(order, lam_ref), orc.chebyshev_coeffs(order, lam_ref), 1e-14)
        close(ctx.to_host(f.chebyshev_mult(dinv, order, lam_ref, dev["x"])), c[f"cheb_z{order}"])
    x = ctx.to_dev(c["X0"])
    res, norms = f.pcg_chebyshev(dinv, 3, lam_ref, dev["B_rhs"], x, 0.0, 0.0, 4)
    assert res.final_iter == 4
    close(ctx.to_host(x), c["X_cheb3_pcg4"], TOL_PCG)
    close(norms, c["cheb3_pcg_norms"], 1e-10)
    x = ctx.to_dev(c["X0"])
    res, _ = f.pcg_chebyshev(dinv, 3, lam_ref, dev["B_rhs"], x, 1e-8, 0.0, 5000)
    assert abs(res.final_iter - int(c["cheb3_tol_iters"][0])) <= 1 and bool(res.converged) == bool(c["cheb3_tol_converged"][0])
    close(ctx.to_host(x), c["X_cheb3_tol"], 1e-8)
    with pytest.raises(b200pa.B200paError, match="order"):
        f.chebyshev_mult(dinv, 6, lam_ref, dev["x"])
    f.close()
    sp.close()


@pytest.mark.parametrize("p,n", [(1, 104), (2, 52), (3, 35), (5, 21)])
def test_mult_host_pipelined_equals_device_apply(ctx, p, n):
    """b200pa_form_mult_host on a problem large enough for its pipelined route (x tiles up, element chunks, E->L reduction
    of the tiles a chunk completes, y tiles down - three streams): bit-identical to the device-resident apply, constrained
    and unconstrained, stored and factorised q-data, repeated calls, and equal to the serial route (B200PA_NO_PIPELINE)"""
    import torch
    m = b200pa.hex_build(n, n, n, p, sx=1.0, sy=0.8, sz=0.6)
    b = b200pa.basis(p)
    assert m["ndofs"] >= 1 << 20
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    rng = np.random.default_rng(p)
    nq = m["ne"] * (p + 2) ** 3
    kq, mq = 0.5 + rng.random(nq), 3.0 + rng.random(nq)
    xh = rng.random(m["ndofs"])
    xp = torch.from_numpy(xh).pin_memory()
    x = ctx.to_dev(xh)
    for fact in (False, True):
        f = b200pa.Form(sp)
        f.set_factorised(fact)
        f.assemble_diffusion(kq)
        f.assemble_mass(mq)
        f.set_essential(b200pa.essential_dofs(m["bdr_attr"], [1, 6]))
        for constrained in (False, True):
            yd = ctx.to_host(f.constrained_mult(x) if constrained else f.mult(x))
            for _ in range(2):
                yp = torch.full((m["ndofs"],), np.nan, dtype=torch.float64).pin_memory()
                f.mult_host(xp, yp, constrained=constrained)
                assert np.array_equal(yp.numpy(), yd)
            yh = np.full(m["ndofs"], np.nan)     # pageable host memory works too (copies just do not overlap)
            f.mult_host(xh, yh, constrained=constrained)
            assert np.array_equal(yh, yd)
            xn, yn = ctx.pinned(m["ndofs"], fill=xh), ctx.pinned(m["ndofs"], fill=np.nan)   # b200pa_host_alloc: NUMA-local, page-locked
            f.mult_host(xn, yn, constrained=constrained)
            assert np.array_equal(yn, yd) and ctx.host_node(xn) >= -1
            del xn, yn
        f.close()
    sp.close()


def test_host_alloc_rejects_foreign_pointers(ctx):
    """b200pa_host_free only takes what b200pa_host_alloc returned; blocks are usable as ordinary host memory"""
    a = ctx.pinned(1000, fill=3.0)
    assert a.sum() == 3000.0 and a.ctypes.data % 4096 == 0
    other = np.zeros(8)
    assert b200pa.lib().b200pa_host_free(ctx.h, b200pa._ptr(other)) != 0
    assert b"not returned by b200pa_host_alloc" in b200pa.lib().b200pa_last_error()
    assert b200pa.lib().b200pa_host_node(b200pa._ptr(other)) == -1
