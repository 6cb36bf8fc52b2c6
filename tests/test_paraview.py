"""CPU: results hand-off in the reference's ParaView layout (ParaViewDataCollection, fem/datacollection.hpp:584).
Golden files: everything the UNMODIFIED reference's ParaViewDataCollection::Save wrote for a mesh and two fields handed to
it in its own text formats (tests/golden/make_golden.py `paraview`: ascii / binary / binary32, Lagrange hexahedra and
refined linear cells, levels of detail equal to and different from the order, one and two cycles).  b200pa_paraview_save
must produce the same directory tree, byte-identical .pvd / .pvtu files, and .vtu files with the same XML structure,
bit-equal integer arrays and float arrays equal to what the format's precision carries."""
import base64
import glob
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

CASES = sorted(os.path.basename(f)[9:-4] for f in glob.glob(os.path.join(GOLDEN, "paraview_*.npz")))


def decode(el, fmt):
    """(values, is_float) of one <DataArray>"""
    t = el.get("type")
    dt = {"Float64": np.float64, "Float32": np.float32, "Int32": np.int32, "UInt8": np.uint8}[t]
    txt = (el.text or "").strip()
    if fmt == "ascii":
        return np.array(txt.split(), dtype=np.float64 if t.startswith("Float") else np.int64), t.startswith("Float")
    assert txt[6:8] == "=="                       # 4 size bytes -> 8 characters, two of them padding
    n = int(np.frombuffer(base64.b64decode(txt[:8]), dtype=np.uint32)[0])
    raw = base64.b64decode(txt[8:])
    assert len(raw) == n
    return np.frombuffer(raw, dtype=dt), t.startswith("Float")


def compare_vtu(mine, ref, fmt):
    a, b = ET.fromstring(mine), ET.fromstring(ref)
    ea, eb = list(a.iter()), list(b.iter())
    assert [e.tag for e in ea] == [e.tag for e in eb]
    n_arrays = 0
    for x, y in zip(ea, eb):
        assert x.attrib == y.attrib, (x.tag, x.attrib, y.attrib)
        if x.tag != "DataArray":
            continue
        n_arrays += 1
        assert x.get("format") == ("ascii" if fmt == "ascii" else "binary")
        va, fl = decode(x, fmt)
        vb, _ = decode(y, fmt)
        assert va.shape == vb.shape, x.attrib
        if not fl:
            assert np.array_equal(va, vb), x.attrib
        else:
            # ascii carries 6 significant digits, Float32 24 bits, Float64 everything (the two writers evaluate the same
            # polynomial with different operation orders)
            rtol = {"ascii": 2e-6, "binary32": 2.4e-7, "binary": 1e-13}[fmt]
            scale = max(np.max(np.abs(vb)), 1e-300)
            assert np.max(np.abs(va.astype(np.float64) - vb.astype(np.float64))) <= rtol * scale, x.attrib
    return n_arrays


@pytest.mark.parametrize("tag", CASES)
def test_paraview_collection_matches_reference(tmp_path, tag):
    import b200pa
    from make_golden import PARAVIEW, paraview_fields
    p, dims, size, skew, fmt, ho, lod, cycles = PARAVIEW[tag]
    ref = {k.replace("|", "/"): bytes(v) for k, v in np.load(os.path.join(GOLDEN, f"paraview_{tag}.npz")).items()}
    m = b200pa.hex_build(*dims, p, *size, skew=skew)
    fields = paraview_fields(m, p, dims)
    out = str(tmp_path) + "/"
    for c in range(cycles):
        b200pa.paraview_save(out, "ablation", m, p, {k: (1.0 + c) * v for k, v in fields.items()}, cycle=c, time=0.25 * c,
                             levels_of_detail=lod, high_order=bool(ho), fmt=fmt)
    mine = {}
    for dp, _, fns in os.walk(out):
        for fn in fns:
            full = os.path.join(dp, fn)
            mine[os.path.relpath(full, out)] = open(full, "rb").read()
    assert sorted(mine) == sorted(ref)
    for path in sorted(ref):
        if path.endswith(".vtu"):
            assert compare_vtu(mine[path], ref[path], fmt) == 5 + len(fields)
            if fmt != "ascii":        # header and integer arrays are the same bytes; only float payloads may differ in the last bit
                la, lb = mine[path].split(b"\n"), ref[path].split(b"\n")
                assert len(la) == len(lb)
        else:
            assert mine[path] == ref[path], path   # .pvd, .pvtu: byte-identical


def test_paraview_pieces_of_a_partitioned_mesh(tmp_path):
    """two ranks write their own pieces, rank 0 the .pvtu naming both and the .pvd; attributes end up as cell data"""
    import b200pa
    from b200pa import partition
    GN, grid, p = (4, 2, 2), (2, 1, 1), 2
    out = str(tmp_path) + "/"
    for rank in range(2):
        m = partition.build_part(GN, grid, rank, p)
        v = np.arange(m["ndofs"], dtype=np.float64)
        b200pa.paraview_save(out, "parts", m, p, {"u": v}, cycle=3, time=1.5, rank=rank, nranks=2, fmt="ascii",
                             attributes=np.full(m["ne"], rank + 1, dtype=np.int32), append=False)
    pvtu = open(os.path.join(out, "parts", "Cycle000003", "data.pvtu")).read()
    assert '<Piece Source="proc000000.vtu"/>' in pvtu and '<Piece Source="proc000001.vtu"/>' in pvtu
    pvd = open(os.path.join(out, "parts", "parts.pvd")).read()
    assert 'timestep="1.5"' in pvd and 'file="Cycle000003/data.pvtu"' in pvd and pvd.rstrip().endswith("</VTKFile>")
    for rank in range(2):
        root = ET.parse(os.path.join(out, "parts", "Cycle000003", f"proc{rank:06d}.vtu")).getroot()
        piece = root.find("UnstructuredGrid/Piece")
        ne = 8
        assert int(piece.get("NumberOfCells")) == ne and int(piece.get("NumberOfPoints")) == ne * 27
        attr = piece.find("CellData/DataArray")
        assert [int(t) for t in attr.text.split()] == [rank + 1] * ne
        types = [e for e in piece.find("Cells") if e.get("Name") == "types"][0]
        assert set(types.text.split()) == {"72"}
    # values at the element's own GLL-lattice nodes are the nodal values when levels_of_detail == 2 and p == 2 (uniform == GLL)
    m = partition.build_part(GN, grid, 0, p)
    u = [e for e in ET.parse(os.path.join(out, "parts", "Cycle000003", "proc000000.vtu")).getroot().iter("DataArray") if e.get("Name") == "u"][0]
    vals = np.array(u.text.split(), dtype=np.float64)
    assert np.allclose(vals, m["gather_map"].astype(np.float64), rtol=1e-5)


def test_paraview_bad_arguments_fail_loudly(tmp_path):
    import b200pa
    m = b200pa.hex_build(1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="paraview_save"):
        b200pa.paraview_save(str(tmp_path) + "/", "", m, 1, {"u": np.zeros(m["ndofs"])})
    bad = dict(m, gather_map=m["gather_map"] + 100)
    with pytest.raises(RuntimeError, match="gather map entry out of range"):
        b200pa.paraview_save(str(tmp_path) + "/", "x", bad, 1, {"u": np.zeros(m["ndofs"])})
