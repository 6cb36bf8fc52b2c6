"""Element-attribute markers (BilinearForm::AddDomainIntegrator(bfi, elem_marker); multi-material domains):
PABilinearFormExtension::AddMultWithMarkers / AssembleDiagonal, fem/bilinearform_ext.cpp:370-454, 807-847 - the reference
pins them in tests/unit/fem/test_pa_kernels.cpp:696-750 (PA == full assembly).  Golden vectors: outputs of the unmodified
reference on three-material meshes for four marker combinations (tests/golden/make_golden.py `markers`).
CPU: the oracle's restatement of the orchestration against them; GPU: the product through the C ABI."""
import glob
import os

import numpy as np
import pytest

import orc
from conftest import GOLDEN

CASES = sorted(os.path.basename(f)[8:-4] for f in glob.glob(os.path.join(GOLDEN, "markers_*.npz")))


def load(tag):
    d = dict(np.load(os.path.join(GOLDEN, f"markers_{tag}.npz")))
    for k in ("p", "D1D", "Q1D", "NE", "ndofs"):
        d[k] = int(d[k][0])
    return d


def combos(c):
    for k in range(4):
        md, mm = c[f"marker_diff{k}"], c[f"marker_mass{k}"]
        yield k, (None if md[0] < 0 else md), (None if mm[0] < 0 else mm)


def rel(a, r):
    return float(np.max(np.abs(a - r)) / np.max(np.abs(r)))


@pytest.mark.parametrize("tag", CASES)
def test_oracle_markers_match_reference(tag):
    c = load(tag)
    pa_d = orc.diffusion_setup(c["Q1D"], c["NE"], c["W"], c["J"], c["kq"])
    pa_m = orc.mass_setup(c["Q1D"], c["NE"], c["W"], c["detJ"], c["mq"])
    op = orc.Operator(c["D1D"], c["Q1D"], c["NE"], c["ndofs"], c["gather_map"], c["B"], c["G"], pa_d, pa_m)
    for k, md, mm in combos(c):
        assert rel(orc.op_mult_markers(op, c["x"], c["elem_attr"], md, mm), c[f"y{k}"]) <= 1e-14
        assert rel(orc.op_diag_markers(op, c["elem_attr"], md, mm), c[f"diag{k}"]) <= 1e-14


def test_reference_diagonal_depends_on_integrator_order():
    """the order dependence the product reproduces is the reference's own: its PA diagonal differs from its full-assembly
    diagonal exactly when the LATER integrator (mass) carries a marker (recorded by `dump_markers`)"""
    c = load(CASES[0])
    for k, md, mm in combos(c):
        d = float(c[f"diag_fa_minus_pa{k}"][0])
        assert (d > 1e-3) == (mm is not None), (k, d)


@pytest.mark.gpu
@pytest.mark.parametrize("geometry", ["given", "vertices"])
@pytest.mark.parametrize("fact", [False, True])
@pytest.mark.parametrize("tag", CASES)
def test_gpu_markers_match_reference(ctx, tag, geometry, fact):
    import b200pa
    c = load(tag)
    sp = b200pa.Space(ctx, c["D1D"], c["Q1D"], c["NE"], c["ndofs"], c["gather_map"], c["B"], c["G"])
    if geometry == "given":
        sp.set_geometry(c["W"], c["J"], c["detJ"])
    else:
        sp.geometry_from_vertices(c["W"], c["vertices"], c["elem_vertices"])
    sp.set_attributes(c["elem_attr"])
    x = ctx.to_dev(c["x"])
    for k, md, mm in combos(c):
        f = b200pa.Form(sp)
        f.set_factorised(fact)
        # markers before assembly for the diffusion integrator, after it for the mass integrator: both orders must work
        f.set_markers(0, md)
        f.assemble_diffusion(c["kq"])
        f.assemble_mass(c["mq"])
        f.set_markers(1, mm)
        f.set_essential(None)
        assert rel(ctx.to_host(f.mult(x)), c[f"y{k}"]) <= 1e-12
        assert rel(ctx.to_host(f.assemble_diagonal()), c[f"diag{k}"]) <= 1e-12
        # the one-pass set-up + diagonal takes the same markers
        d2 = ctx.to_host(f.assemble_diffusion_with_diagonal(c["kq"]))
        assert rel(d2, c[f"diag{k}"]) <= 1e-12
        assert rel(ctx.to_host(f.mult(x)), c[f"y{k}"]) <= 1e-12
        f.close()
    sp.close()


@pytest.mark.gpu
def test_gpu_marker_errors(ctx):
    import b200pa
    c = load(CASES[0])
    sp = b200pa.Space(ctx, c["D1D"], c["Q1D"], c["NE"], c["ndofs"], c["gather_map"], c["B"], c["G"])
    sp.set_geometry(c["W"], c["J"], c["detJ"])
    f = b200pa.Form(sp)
    with pytest.raises(b200pa.B200paError, match="attributes"):
        f.set_markers(0, [1, 1, 1])
    sp.set_attributes(c["elem_attr"])
    with pytest.raises(b200pa.B200paError, match="exceeds"):
        f.set_markers(0, [1, 1])          # attribute 3 exists
    f.set_markers(1, [1, 0, 1])
    with pytest.raises(b200pa.B200paError, match="markers"):
        f.set_pa_data(ctx.zeros(6 * c["NE"] * c["Q1D"] ** 3), None)
    f.close()
    sp.close()
