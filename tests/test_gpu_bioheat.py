"""GPU: the RF-ablation coupled step through the C ABI against the reference's
own run of it (tests/golden/bioheat_p2_n4.npz from oracle/_ref/ref_driver dump_bioheat), and
full-size property checks on the configs[1] mesh."""
import os

import numpy as np
import pytest

import b200pa
from conftest import GOLDEN
from test_gpu_parity import close

pytestmark = pytest.mark.gpu

P = dict(dt=0.5, rc=3.6e6, wbcb=4.0e4, Ta=37.0, k0=0.5, ak=0.02, s0=0.3, as_=0.015, V=30.0)


def t_init(lattice, gll, p, n):
    """GridFunction::ProjectCoefficient of 37 + 20 exp(-40 r^2): nodal values at the GLL points"""
    lat = lattice.reshape(-1, 3)
    xyz = (lat // p + gll[lat % p]) / n
    xyz[lat == p * n] = 1.0
    r2 = ((xyz - 0.5) ** 2).sum(1)
    return 37.0 + 20.0 * np.exp(-40.0 * r2)


def build(ctx, p, n):
    m = b200pa.hex_build(n, n, n, p)
    b = b200pa.basis(p)
    sp = b200pa.Space(ctx, p + 1, p + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
    sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
    return m, b, sp


def coupled_step(ctx, m, b, sp, T0, iters, rel_tol=0.0):
    """the product-side driver (b200pa/bioheat.py), plus the intermediate q-data the test inspects"""
    from b200pa.bioheat import CoupledStep
    n = round(m["ne"] ** (1 / 3))
    cs = CoupledStep(ctx, sp, m, (n, n, n))
    o = cs.step(T0, iters, iters, rel_tol)
    o.update(kq=cs.kq, sq=cs.sq, mq=cs.mq, ess=cs.ess, phi0=cs.phi_bc, fe=cs.fe, ft=cs.ft)
    cs.fm.close()
    return o


def test_coupled_step_matches_reference(ctx):
    g = dict(np.load(os.path.join(GOLDEN, "bioheat_p2_n4.npz")))
    p, n = 2, 4
    m, b, sp = build(ctx, p, n)
    T0h = t_init(m["lattice"], b["gll"], p, n)
    close(T0h, g["T0"], 1e-14)
    T0 = ctx.to_dev(g["T0"])
    close(ctx.to_host(sp.qvalues(T0)), g["Tq"])
    iters = int(g["iters"][0])
    o = coupled_step(ctx, m, b, sp, T0, iters)
    assert np.array_equal(o["ess"], g["ess"])
    close(ctx.to_host(o["kq"]), g["kq"])
    close(ctx.to_host(o["sq"]), g["sq"])
    close(ctx.to_host(o["mq"]), g["mq"], 0.0)
    close(o["phi0"], g["phi0"], 1e-15)
    close(ctx.to_host(o["Be"]), g["Be"])
    close(ctx.to_host(o["phi"]), g["phi"], 1e-10)
    gq = ctx.to_host(sp.qphysgrad(o["phi"]))
    close(gq, g["gradphi_q"], 1e-9)
    close(ctx.to_host(o["src"]), g["src_q"], 1e-9)
    close(ctx.to_host(o["rhs"]), g["rhs_T"], 1e-10)
    close(ctx.to_host(o["T1"]), g["T1"], 1e-10)
    # iteration counts to rel 1e-8 within +-1 (same systems as above, as the reference driver does)
    phi2 = ctx.to_dev(o["phi0"])
    res_e, _ = o["fe"].pcg(o["fe"].jacobi(), o["Be"], phi2, 1e-8, 0.0, 5000)
    assert abs(res_e.final_iter - int(g["iters_tol_phi"][0])) <= 1 and res_e.converged
    close(ctx.to_host(phi2), g["phi_tol"], 1e-6)
    T2 = T0.clone()
    res_t, _ = o["ft"].pcg(o["ft"].jacobi(), o["rhs"], T2, 1e-8, 0.0, 5000)
    assert abs(res_t.final_iter - int(g["iters_tol_T"][0])) <= 1 and res_t.converged
    close(ctx.to_host(T2), g["T1_tol"], 1e-8)
    o["fe"].close()
    o["ft"].close()
    sp.close()


@pytest.mark.parametrize("factorised", [False, True])
def test_three_coupled_steps_match_reference(ctx, factorised):
    """the time loop: T^{n+1} feeds k(T), sigma(T) of the next step; all state device-resident
    (tests/golden/bioheat_steps_p2_n4.npz: the reference run of the same three steps, 12 PCG its per solve);
    with the stored and with the factorised sigma(T), k(T) q-data"""
    from b200pa.bioheat import CoupledStep
    g = dict(np.load(os.path.join(GOLDEN, "bioheat_steps_p2_n4.npz")))
    p, n = 2, 4
    m, b, sp = build(ctx, p, n)
    cs = CoupledStep(ctx, sp, m, (n, n, n), factorised=factorised)
    iters, nsteps = int(g["iters"][0]), int(g["nsteps"][0])
    outs = cs.run(ctx.to_dev(g["T_step0"]), nsteps, iters, iters)
    for k, o in enumerate(outs, 1):
        close(ctx.to_host(o["T1"]), g[f"T_step{k}"], 1e-10)
        close(ctx.to_host(o["phi"]), g[f"phi_step{k}"], 1e-10)
    assert cs.fe.factorised == factorised and cs.ft.factorised == factorised
    cs.close()
    sp.close()


@pytest.mark.parametrize("p,n", [(2, 100), (1, 64), (3, 40), (4, 24)])
def test_full_size_properties(ctx, p, n):
    """BASELINE configs[1] size (p=2, N=100: 8,120,601 dofs) and sweep points: properties that do
    not need the CPU oracle — A 1 = M 1 for diffusion+mass (grad 1 = 0), 1^T M 1 = c * volume,
    symmetry (u, A v) = (v, A u), linearity, and agreement of the fused L->L path with the
    unfused kernel-level path (gather, E-apply, CSR scatter)."""
    import torch
    m, b, sp = build(ctx, p, n)
    nd = m["ndofs"]
    f = b200pa.Form(sp)
    f.assemble_diffusion(np.array([0.5]))
    f.assemble_mass(np.array([3.6]))
    f.set_essential(None)
    one = ctx.zeros(nd) + 1.0
    y1 = f.mult(one)
    vol = ctx.dot(one, y1)
    assert abs(vol - 3.6) <= 1e-11 * 3.6                      # 1^T (K + M) 1 = 3.6 * |Omega|
    with torch.cuda.stream(ctx.torch_stream):
        g = torch.Generator(device="cuda").manual_seed(1)
        u = torch.rand(nd, dtype=torch.float64, device="cuda", generator=g)
        v = torch.rand(nd, dtype=torch.float64, device="cuda", generator=g)
    Au, Av = f.mult(u), f.mult(v)
    s1, s2 = ctx.dot(v, Au), ctx.dot(u, Av)
    assert abs(s1 - s2) <= 1e-12 * abs(s1)
    w = ctx.add(u, 2.5, v)
    Aw = f.mult(w)
    ref = ctx.add(Au, 2.5, Av)
    ctx.sync()
    assert float((Aw - ref).abs().max() / ref.abs().max()) <= 1e-13
    # unfused path == fused path
    D, Q, ne = p + 1, p + 2, m["ne"]
    gm = ctx.to_dev(m["gather_map"])
    xE = ctx.restrict_mult(ne, D ** 3, gm, u)
    yE = ctx.zeros(ne * D ** 3)
    import ctypes as C
    pd = b200pa.lib().b200pa_form_pa_diff(f.h)
    pm = b200pa.lib().b200pa_form_pa_mass(f.h)
    ctx.diffusion_apply(ne, D, Q, b["B"], b["G"], pd, xE, yE)
    ctx.mass_apply(ne, D, Q, b["B"], pm, xE, yE)
    off = b200pa.lib().b200pa_space_offsets(sp.h)
    ind = b200pa.lib().b200pa_space_indices(sp.h)
    y2 = ctx.restrict_mult_transpose(nd, off, ind, yE)
    ctx.sync()
    assert float((y2 - Au).abs().max() / Au.abs().max()) <= 1e-13
    f.close()
    sp.close()


def test_config2_norm_matches_reference_probe(ctx):
    """p=2, N=100 (8,120,601 dofs), Diffusion(0.5)+Mass(3.6), x = Vector::Randomize(1) on the
    reference's own numbering: |Ax|_2 and |diag|_2 recorded from the reference CPU build in
    SURVEY.md Appendix A.2 / BASELINE.md (full-size parity, not just properties)."""
    import ctypes as C
    p, n = 2, 100
    m, b, sp = build(ctx, p, n)
    nd = m["ndofs"]
    assert nd == 8120601
    x = b200pa.randomize(nd, 1)   # Vector::Randomize(1)
    f = b200pa.Form(sp)
    f.assemble_diffusion(np.array([0.5]))
    f.assemble_mass(np.array([3.6]))
    f.set_essential(None)
    y = ctx.to_host(f.mult(ctx.to_dev(x)))
    d = ctx.to_host(f.assemble_diagonal())
    assert abs(np.linalg.norm(y) - 11.042738623592694) <= 1e-12 * 11.042738623592694
    assert abs(np.linalg.norm(d) - 36.909814340923262) <= 1e-12 * 36.909814340923262
    f.close()
    sp.close()


@pytest.mark.parametrize("p,n", [(1, 40), (2, 48), (3, 24), (4, 16), (6, 9)])
def test_bitwise_reproducible(ctx, p, n):
    """compute-sanitizer is closed on the GPU pool, so race-freedom of the shared-memory phases, the TMA staging
    and the atomic-free reductions is pinned the other way round: the apply, the diagonal and a 12-iteration PCG
    are bitwise identical run to run (deterministic reductions; a race shows up as a flipped bit)."""
    import torch
    m, b, sp = build(ctx, p, n)
    nd = m["ndofs"]
    f = b200pa.Form(sp)
    T0 = ctx.to_dev(37.0 + np.random.default_rng(3).random(nd))
    f.assemble_diffusion(sp.coeff_linear(0.5, 0.02, 37.0, T0))
    f.assemble_mass(np.array([3.6]))
    f.set_essential(b200pa.essential_dofs(m["bdr_attr"], [1, 6]))
    x = ctx.to_dev(np.random.default_rng(4).random(nd))
    dinv = f.jacobi()
    ref = None
    for rep in range(3):
        y = f.constrained_mult(x)
        d = f.assemble_diagonal()
        X = ctx.zeros(nd)
        res, norms = f.pcg(dinv, x, X, 0.0, 0.0, 12)
        ctx.sync()
        cur = (y.clone(), d.clone(), X.clone(), norms.copy())
        if ref is None:
            ref = cur
        else:
            assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and torch.equal(cur[2], ref[2])
            assert np.array_equal(cur[3], ref[3])
    f.close()
    sp.close()
