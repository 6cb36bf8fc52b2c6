"""GPU: object lifetimes behind the C ABI.  Every b200pa_*_create has a destroy that gives its device memory back: a
solver that re-creates spaces / forms / multigrid hierarchies (remeshing, p-adaptation) must not leak."""
import numpy as np
import pytest

import b200pa

pytestmark = pytest.mark.gpu


def one_cycle(ctx, p):
    """everything the library can allocate: space tables, geometry (stored J, then vertices), q-data in both forms, scratch,
    PCG / Chebyshev work vectors, marker masks, a host-buffer apply, pinned host blocks, transfers and a 2-level multigrid"""
    dims = (12, 10, 8)
    levels = []
    for pp in (1, p):
        m = b200pa.hex_build(*dims, pp, skew=True)
        b = b200pa.basis(pp)
        sp = b200pa.Space(ctx, pp + 1, pp + 2, m["ne"], m["ndofs"], m["gather_map"], b["B"], b["G"])
        sp.geometry_from_vertices(b["W"], m["vertices"], m["elem_vertices"])
        sp.set_attributes(1 + (np.arange(m["ne"]) % 3))
        nq = m["ne"] * (pp + 2) ** 3
        f = b200pa.Form(sp)
        f.set_markers(0, [1, 0, 1])      # diffusion acts on attributes 1 and 3 only; the mass term keeps every diagonal entry positive
        f.assemble_diffusion(0.5 + np.random.default_rng(1).random(nq))
        f.assemble_mass(np.array([3.6]))
        f.set_essential(b200pa.essential_dofs(m["bdr_attr"], [1, 6]))
        levels.append((m, sp, f))
    m, sp, f = levels[1]
    x = ctx.to_dev(np.random.default_rng(2).random(m["ndofs"]))
    y = f.constrained_mult(x)
    diag = ctx.empty(m["ndofs"])
    f.assemble_diffusion_with_diagonal(sp.coeff_linear(0.5, 0.02, 37.0, x), diag)
    X = ctx.zeros(m["ndofs"])
    f.pcg(f.jacobi(), y, X, 0.0, 0.0, 3)
    lam = f.power_method(f.jacobi(), ctx.to_dev(b200pa.randomize(m["ndofs"], 12345)))
    f.pcg_chebyshev(f.jacobi(), 2, lam, y, X, 0.0, 0.0, 2)
    g = b200pa.Form(sp)
    g.set_factorised(True)
    g.assemble_diffusion(sp.coeff_linear(0.5, 0.02, 37.0, x))
    g.set_essential(None)
    hx, hy = ctx.pinned(m["ndofs"], fill=1.0), ctx.pinned(m["ndofs"])
    g.mult_host(hx, hy)
    del hx, hy
    T = b200pa.Transfer(levels[0][2], f, b200pa.basis_transfer(1, p))
    mg = b200pa.Multigrid([levels[0][2], f], [T])
    mg.setup()
    mg.pcg(y, X, 0.0, 0.0, 2)
    mg.close()
    T.close()
    g.close()
    for _, s, ff in levels:
        ff.close()
        s.close()


@pytest.mark.parametrize("p", [2, 4])
def test_create_destroy_cycles_do_not_leak_device_memory(ctx, p):
    import gc

    import torch

    def free_bytes():
        gc.collect()
        ctx.sync()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        return torch.cuda.mem_get_info()[0]

    one_cycle(ctx, p)                 # first use: kernel images, per-device caches, the context's own scratch
    base = free_bytes()
    for _ in range(6):
        one_cycle(ctx, p)
    lost = base - free_bytes()
    assert lost <= 4 << 20, f"{lost / 2 ** 20:.1f} MiB of device memory not returned after 6 create/destroy cycles"
