"""GPU parity of the factorised diffusion q-data (affine meshes: w_q c_q per q-point + one tensor per element,
include/b200pa.h b200pa_form_set_factorised) against the golden fixtures (outputs of the unmodified reference,
which stores D at every q-point) and the CPU oracle; same tolerances as the stored form: 1e-12 per apply,
1e-10 after fixed PCG iterations, equal iteration counts.  A mesh with a non-affine element must be refused."""
import numpy as np
import pytest

import b200pa
import orc
from test_gpu_parity import TOL_APPLY, TOL_PCG, Dev, close

pytestmark = pytest.mark.gpu


@pytest.fixture
def dev(ctx, case):
    return Dev(ctx, case)


def fact_form(dev, sp, diff=True, mass=True, ess=True):
    f = b200pa.Form(sp)
    f.set_factorised(True)
    f.assemble_diffusion(dev.c["kq"] if diff else None)
    f.assemble_mass(dev.c["mq"] if mass else None)
    f.set_essential(dev.c["ess"] if ess else None)
    return f


def test_all_golden_meshes_are_affine(ctx, dev):
    sp = dev.space(geometry="vertices")
    assert sp.affine
    sp.close()


@pytest.mark.parametrize("geometry", ["vertices", "given"])
def test_factorised_mult_diag_match_reference(ctx, dev, geometry):
    """geometry from the vertices, or the reference's own Jacobians (what the MFEM binding passes)"""
    c = dev.c
    sp = dev.space(geometry=geometry)
    assert sp.affine
    f = fact_form(dev, sp)
    assert f.factorised
    close(ctx.to_host(f.mult(dev["x"])), c["y"])
    close(ctx.to_host(f.assemble_diagonal()), c["diag"])
    close(ctx.to_host(f.constrained_mult(dev["x"])), c["y_constrained"])
    f.close()
    # diffusion alone / switching one form between the two representations
    f = fact_form(dev, sp, mass=False)
    close(ctx.to_host(f.mult(dev["x"])), c["y_diff"])
    f.set_factorised(False)
    f.assemble_diffusion(c["kq"])
    assert not f.factorised
    close(ctx.to_host(f.mult(dev["x"])), c["y_diff"])
    f.close()
    sp.close()


def test_factorised_pcg_matches_oracle(ctx, dev):
    c = dev.c
    sp = dev.space(geometry="vertices")
    f = fact_form(dev, sp)
    op = orc.Operator(dev.D, dev.Q, dev.NE, dev.nd, c["gather_map"], c["B"], c["G"], c["pa_diff"], c["pa_mass"], c["ess"])
    rhs = np.random.default_rng(3).random(dev.nd)
    rhs[c["ess"]] = 0.0
    for rel_tol, iters in ((0.0, 6), (1e-8, 400)):
        xo, it, conv, fn, _ = op.pcg(op.jacobi_dinv(), rhs, np.zeros(dev.nd), rel_tol, 0.0, iters)
        x = ctx.zeros(dev.nd)
        res, _ = f.pcg(f.jacobi(), ctx.to_dev(rhs), x, rel_tol, 0.0, iters)
        assert abs(res.final_iter - it) <= (0 if rel_tol == 0.0 else 1)
        close(ctx.to_host(x), xo, TOL_PCG if rel_tol == 0.0 else 1e-6)
    f.close()
    sp.close()


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_factorised_equals_stored_on_slab(ctx, p):
    """anisotropic tissue slab with tail batches (NE not a multiple of any batch size), q-function coefficients:
    factorised == stored to 1e-13, bitwise reproducible run to run"""
    dims = (7, 5, 3)
    m = b200pa.hex_build(*dims, p, sx=2.0, sy=1.0, sz=0.25)
    b = b200pa.basis(p)
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    assert sp.affine
    rng = np.random.default_rng(p)
    nq = m["ne"] * (p + 2) ** 3
    kq, mq = 0.5 + rng.random(nq), 3.0 + rng.random(nq)
    x = ctx.to_dev(rng.random(m["ndofs"]))
    ys, ds = [], []
    for fact in (False, True, True):
        f = b200pa.Form(sp)
        f.set_factorised(fact)
        f.assemble_diffusion(kq)
        f.assemble_mass(mq)
        f.set_essential(None)
        ys.append(ctx.to_host(f.mult(x)))
        ds.append(ctx.to_host(f.assemble_diagonal()))
        f.close()
    close(ys[1], ys[0], 1e-13)
    close(ds[1], ds[0], 1e-13)
    assert np.array_equal(ys[1], ys[2]) and np.array_equal(ds[1], ds[2])
    sp.close()


def test_non_affine_mesh_is_refused(ctx):
    p = 2
    m = b200pa.hex_build(3, 3, 3, p)
    b = b200pa.basis(p)
    v = m["vertices"].copy().reshape(-1, 3)
    v[21] += [0.03, -0.02, 0.01]          # an interior vertex moved: its eight elements are no parallelepipeds
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], v.ravel(), m["elem_vertices"])
    assert not sp.affine
    f = b200pa.Form(sp)
    f.set_factorised(True)
    with pytest.raises(b200pa.B200paError, match="affine"):
        f.assemble_diffusion(np.array([1.0]))
    f.set_factorised(False)
    f.assemble_diffusion(np.array([1.0]))   # the stored form still works on this mesh
    f.set_essential(None)
    y = ctx.to_host(f.mult(ctx.to_dev(np.ones(m["ndofs"]))))
    assert np.max(np.abs(y)) < 1e-12        # constants are in the kernel of the diffusion operator
    f.close()
    sp.close()


# ---------------------------------------------------------------- fused q-data set-up + diagonal (one pass)
def _pa_diff_host(ctx, f, n):
    import ctypes as C
    out = np.empty(n)
    p = b200pa.lib().b200pa_form_pa_diff(f.h)
    b200pa.check(b200pa.lib().b200pa_ctx_download(ctx.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(p), C.c_size_t(out.nbytes)))
    return out


@pytest.mark.parametrize("fact", [False, True])
def test_fused_setup_diagonal_matches_reference(ctx, dev, fact):
    """b200pa_form_assemble_diffusion_with_diagonal on the golden meshes (all affine, geometry from the vertices):
    q-data and diagonal against the reference's pa_data / diagonal, and bit-identical to the two-pass route"""
    c = dev.c
    sp = dev.space(geometry="vertices")
    f = b200pa.Form(sp)
    f.set_factorised(fact)
    f.assemble_mass(c["mq"])
    f.set_essential(None)
    diag = ctx.to_host(f.assemble_diffusion_with_diagonal(c["kq"]))
    assert f.factorised == fact
    close(diag, c["diag"])
    close(ctx.to_host(f.mult(dev["x"])), c["y"])
    g = b200pa.Form(sp)
    g.set_factorised(fact)
    g.assemble_diffusion(c["kq"])
    g.assemble_mass(c["mq"])
    g.set_essential(None)
    nq = dev.NE * dev.Q ** 3
    n = nq if fact else 6 * nq
    assert np.array_equal(_pa_diff_host(ctx, f, n), _pa_diff_host(ctx, g, n))
    assert np.array_equal(diag, ctx.to_host(g.assemble_diagonal()))
    if not fact:
        close(_pa_diff_host(ctx, f, n), c["pa_diff"])
    f.close()
    g.close()
    sp.close()


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("const_c", [False, True])
def test_fused_setup_diagonal_on_slab(ctx, p, const_c):
    """tail batches and odd element counts (the scalar fields then start on odd 8-byte boundaries), q-function and constant
    coefficients, diffusion alone and with mass; fused == two passes, bit for bit, and reproducible"""
    dims = (7, 5, 3)
    m = b200pa.hex_build(*dims, p, sx=2.0, sy=1.0, sz=0.25)
    b = b200pa.basis(p)
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    rng = np.random.default_rng(10 + p)
    nq = m["ne"] * (p + 2) ** 3
    kq = np.array([0.7]) if const_c else 0.5 + rng.random(nq)
    mq = 3.0 + rng.random(nq)
    for mass in (True, False):
        for fact in (False, True):
            f, g = b200pa.Form(sp), b200pa.Form(sp)
            for h in (f, g):
                h.set_factorised(fact)
                if mass:
                    h.assemble_mass(mq)
                h.set_essential(None)
            d1 = ctx.to_host(f.assemble_diffusion_with_diagonal(kq))
            d1b = ctx.to_host(f.assemble_diffusion_with_diagonal(kq))
            g.assemble_diffusion(kq)
            d2 = ctx.to_host(g.assemble_diagonal())
            n = nq if fact else 6 * nq
            assert np.array_equal(_pa_diff_host(ctx, f, n), _pa_diff_host(ctx, g, n))
            assert np.array_equal(d1, d2) and np.array_equal(d1, d1b)
            x = ctx.to_dev(rng.random(m["ndofs"]))
            assert np.array_equal(ctx.to_host(f.mult(x)), ctx.to_host(g.mult(x)))
            f.close()
            g.close()
    sp.close()


def test_fused_setup_diagonal_falls_back_on_skewed_mesh(ctx):
    """a mesh with non-affine elements takes the two-pass route: same results as calling the two entry points"""
    p = 2
    m = b200pa.hex_build(4, 3, 3, p)
    b = b200pa.basis(p)
    v = m["vertices"].copy().reshape(-1, 3)
    v[21] += [0.03, -0.02, 0.01]
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], v.ravel(), m["elem_vertices"])
    assert not sp.affine
    rng = np.random.default_rng(5)
    kq = 0.5 + rng.random(m["ne"] * 64)
    f, g = b200pa.Form(sp), b200pa.Form(sp)
    for h in (f, g):
        h.assemble_mass(np.array([2.0]))
        h.set_essential(None)
    d1 = ctx.to_host(f.assemble_diffusion_with_diagonal(kq))
    g.assemble_diffusion(kq)
    assert np.array_equal(d1, ctx.to_host(g.assemble_diagonal()))
    f.close()
    g.close()
    sp.close()
