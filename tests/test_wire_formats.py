"""CPU: results hand-off.  The product writes its mesh in the reference's "MFEM mesh v1.0" text format
and a field in GridFunction::Save's format; the UNMODIFIED reference (oracle/_ref/ref_driver load_check, test
infrastructure) loads both and must see the builder's numbering, boundary attributes and values."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


@pytest.mark.parametrize("p,dims,size,skew", [(1, (3, 2, 2), (1.0, 1.0, 1.0), False), (2, (4, 3, 2), (1.0, 0.7, 0.4), True),
                                              (3, (2, 3, 2), (2.0, 1.0, 0.25), False)])
def test_reference_loads_written_mesh_and_gridfunction(tmp_path, p, dims, size, skew):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver not built (needs the reference tree: make -C oracle ref)")
    import b200pa
    from make_golden import load_dump
    m = b200pa.hex_build(*dims, p, *size, skew=skew)
    b = b200pa.basis(p)
    lat = m["lattice"].reshape(-1, 3)
    xyz = (lat // p + b["gll"][lat % p]) / np.asarray(dims, dtype=np.float64)
    T = 37.0 + 20.0 * np.exp(-4.0 * ((xyz - 0.5) ** 2).sum(1))
    mesh_file, gf_file = tmp_path / "slab.mesh", tmp_path / "T.gf"
    b200pa.write_mesh(mesh_file, *dims, *size, skew=skew)
    b200pa.write_gridfunction(gf_file, p, T)
    out = tmp_path / "dump"
    subprocess.run([DRIVER, "load_check", str(out), str(mesh_file), str(gf_file)], check=True, stdout=subprocess.DEVNULL)
    d = load_dump(str(out))
    assert int(d["NE"][0]) == m["ne"] and int(d["NV"][0]) == m["nv"] and int(d["ndofs"][0]) == m["ndofs"] and int(d["order"][0]) == p
    assert np.array_equal(d["gather_map"], m["gather_map"])            # same H1 numbering through the file
    assert np.array_equal(d["values"], T)                               # 17 significant digits: lossless
    assert np.array_equal(d["vertices"], m["vertices"])
    assert np.array_equal(d["ess_z"], b200pa.essential_dofs(m["bdr_attr"], [1, 6]))
    nx, ny, nz = dims
    counts = np.bincount(d["bdr_attributes"], minlength=7)[1:]
    assert list(counts) == [nx * ny, nx * nz, ny * nz, nx * nz, ny * nz, nx * ny]
    assert d["l2_norm"][0] > 0.0
