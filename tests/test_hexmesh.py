"""CPU: the product-side synthetic-mesh builder (csrc/hexmesh.cpp) reproduces the reference's
numbering (tests/golden/numbering.npz, dumped from the unmodified reference) and its 1-D basis."""
import numpy as np
import pytest

import b200pa
from conftest import GOLDEN, golden_cases, load_case
import os

NUM = np.load(os.path.join(GOLDEN, "numbering.npz"))
TAGS = sorted(k[:-7] for k in NUM.files if k.endswith("_gather"))


def parse(tag):
    a = tag.split("_")
    return int(a[0][1:]), int(a[1]), int(a[2]), int(a[3][1:])


@pytest.mark.parametrize("tag", TAGS)
def test_numbering_matches_reference(tag):
    nx, ny, nz, p = parse(tag)
    m = b200pa.hex_build(nx, ny, nz, p)
    assert m["ndofs"] == int(NUM[tag + "_ndofs"][0])
    assert np.array_equal(m["elem_vertices"], NUM[tag + "_ev"])
    assert np.array_equal(m["gather_map"], NUM[tag + "_gather"])
    ess = b200pa.essential_dofs(m["bdr_attr"], [1, 2, 3, 4, 5, 6])
    assert np.array_equal(ess, NUM[tag + "_ess_all"])


@pytest.mark.parametrize("tag", golden_cases())
def test_case_mesh_and_basis(tag):
    c = load_case(tag)
    p = c["p"]
    # recover (nx,ny,nz,kind) from the tag
    kind = "skew" if "skew" in tag else "cart"
    dims = {"p1_skew3_func_z": (3, 3, 3), "p2_skew3_func_z": (3, 3, 3), "p2_cart432_const_all": (4, 3, 2),
            "p2_skew2_func_none": (2, 2, 2), "p3_skew2_func_z": (2, 2, 2), "p3_cart322_const_all": (3, 2, 2),
            "p4_skew2_func_z": (2, 2, 2), "p5_skew2_func_all": (2, 2, 2), "p6_skew2_func_z": (2, 1, 2)}[tag]
    size = (1.0, 0.7, 0.4) if tag == "p2_cart432_const_all" else (1.0, 1.0, 1.0)
    m = b200pa.hex_build(*dims, p, *size, skew=(kind == "skew"))
    assert np.array_equal(m["gather_map"], c["gather_map"])
    assert np.allclose(m["vertices"], c["vertices"], rtol=0, atol=1e-15)
    bc = tag.rsplit("_", 1)[1]
    attrs = {"all": [1, 2, 3, 4, 5, 6], "z": [1, 6], "none": []}[bc]
    assert np.array_equal(b200pa.essential_dofs(m["bdr_attr"], attrs), c["ess"])
    b = b200pa.basis(p)
    for k in ("B", "G", "W"):
        assert np.max(np.abs(b[k] - c[k])) <= 1e-13 * max(1.0, np.max(np.abs(c[k]))), k


def test_partition_lattice_consistency():
    """sub-box numbering: lattice coordinates of a part are those of the global mesh"""
    G = (4, 3, 2)
    p = 2
    glob = b200pa.hex_build(*G, p)
    key = {tuple(glob["lattice"][3 * i:3 * i + 3]): i for i in range(glob["ndofs"])}
    assert len(key) == glob["ndofs"]
    part = b200pa.hex_build(2, 3, 2, p, part=(4, 3, 2, 2, 0, 0))
    lat = part["lattice"].reshape(-1, 3)
    assert lat[:, 0].min() == 2 * p and lat[:, 0].max() == 4 * p
    assert all(tuple(r) in key for r in lat)
    # boundary attribute 5 (x = 0) is absent from the right half, 3 (x = max) present
    assert not np.any(part["bdr_attr"] & (1 << 4)) and np.any(part["bdr_attr"] & (1 << 2))


def test_rejects_bad_sizes():
    with pytest.raises(b200pa.B200paError):
        b200pa.hex_build(0, 1, 1, 1)
    with pytest.raises(b200pa.B200paError):
        b200pa.hex_sizes(1, 1, 1, 0)
