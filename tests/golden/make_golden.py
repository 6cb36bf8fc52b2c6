#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs oracle/_ref/ref_driver, i.e. `make -C oracle ref`,
which compiles the reference sources under /root/reference).  The reference ships no stored
golden vectors for this path — its tests compare assembly levels — so the
vectors here are outputs of the reference itself, which is what pins oracle/pa_oracle.c.

    python tests/golden/make_golden.py

Fixtures (all float64 / int32, reference layouts):
  case_<tag>.npz      everything `ref_driver dump_case` writes for one small case
  numbering.npz       gather_map / ndofs for many (nx,ny,nz,p): pins the product-side hex builder
  bioheat_p2_n4.npz   the RF + bioheat coupled step of SURVEY §3.2/3.3 on a 4^3 mesh
  bioheat_steps_p2_n4.npz   three consecutive coupled steps (T^{n+1} feeds k(T), sigma(T) of the next one)
  markers_<tag>.npz   element-attribute markers: y = A x and the diagonal for four marker combinations (`dump_markers`)
  mg_<tag>.npz        p-multigrid: per-level tables and q-data, transfer matrices and P x / P^T x, one V-cycle, MG-PCG (`dump_mg`)
  paraview_<tag>.npz  every file ParaViewDataCollection::Save wrote (`paraview`) for a mesh + fields handed over in the
                      reference's text formats, as byte arrays keyed by relative path ('/' -> '|')
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def load_dump(d):
    out = {}
    for line in open(os.path.join(d, "manifest.txt")):
        name, ext, n = line.split()
        dt = np.float64 if ext == "f64" else np.int32
        a = np.fromfile(os.path.join(d, f"{name}.{ext}"), dtype=dt)
        assert a.size == int(n), (name, a.size, n)
        out[name] = a
    return out


def run(args):
    with tempfile.TemporaryDirectory() as t:
        d = os.path.join(t, "d")
        cmd = [DRIVER, args[0], d] + [str(a) for a in args[1:]]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
        return load_dump(d)


# tag: (p, kind, nx, ny, nz, sx, sy, sz, coef, bc, pcg_iters)
CASES = {
    "p1_skew3_func_z": (1, "skew", 3, 3, 3, 1, 1, 1, "func", "zfaces", 10),
    "p2_skew3_func_z": (2, "skew", 3, 3, 3, 1, 1, 1, "func", "zfaces", 10),
    "p2_cart432_const_all": (2, "cart", 4, 3, 2, 1.0, 0.7, 0.4, "const", "all", 10),
    "p2_skew2_func_none": (2, "skew", 2, 2, 2, 1, 1, 1, "func", "none", 5),
    "p3_skew2_func_z": (3, "skew", 2, 2, 2, 1, 1, 1, "func", "zfaces", 10),
    "p3_cart322_const_all": (3, "cart", 3, 2, 2, 1, 1, 1, "const", "all", 10),
    "p4_skew2_func_z": (4, "skew", 2, 2, 2, 1, 1, 1, "func", "zfaces", 10),
    "p5_skew2_func_all": (5, "skew", 2, 2, 2, 1, 1, 1, "func", "all", 10),
    "p6_skew2_func_z": (6, "skew", 2, 1, 2, 1, 1, 1, "func", "zfaces", 10),
}

NUMBERING = [(1, 1, 1), (2, 2, 2), (3, 3, 3), (4, 3, 2), (2, 5, 3), (5, 5, 5), (6, 4, 7), (3, 8, 5),
             (8, 8, 8), (7, 2, 9)]


def markers():
    """element-attribute markers (three-material meshes, four marker combinations): markers_<tag>.npz"""
    for tag, c in {"p2_skew332": (2, "skew", 3, 3, 2), "p3_cart232": (3, "cart", 2, 3, 2), "p1_skew433": (1, "skew", 4, 3, 3)}.items():
        d = run(["dump_markers"] + list(c))
        np.savez_compressed(os.path.join(HERE, f"markers_{tag}.npz"), **d)
        print("markers", tag, sum(v.nbytes for v in d.values()) // 1024, "KiB raw")


def multigrid():
    """p-multigrid (examples/ex26.cpp hierarchy, diffusion + mass): mg_<tag>.npz"""
    for tag, c in {"skew222_z_p124": ("skew", 2, 2, 2, "zfaces", 1, 2, 4), "cart322_none_p123": ("cart", 3, 2, 2, "none", 1, 2, 3),
                   "skew322_all_p13": ("skew", 3, 2, 2, "all", 1, 3)}.items():
        d = run(["dump_mg"] + list(c))
        np.savez_compressed(os.path.join(HERE, f"mg_{tag}.npz"), **d)
        print("mg", tag, sum(v.nbytes for v in d.values()) // 1024, "KiB raw")


# tag: (p, dims, size, skew, format, high_order, levels_of_detail, cycles)
PARAVIEW = {
    "p2_skew322_ascii_ho": (2, (3, 2, 2), (1.0, 0.7, 0.4), True, "ascii", 1, 2, 2),
    "p3_cart221_binary_ho": (3, (2, 2, 1), (1.0, 1.0, 0.5), False, "binary", 1, 3, 1),
    "p2_cart222_binary32_lo": (2, (2, 2, 2), (1.0, 1.0, 1.0), False, "binary32", 0, 2, 2),
    "p1_skew211_ascii_lo": (1, (2, 1, 1), (2.0, 1.0, 1.0), True, "ascii", 0, 1, 1),
    "p2_skew212_binary_lod4": (2, (2, 1, 2), (1.0, 1.0, 1.0), True, "binary", 1, 4, 1),
}


def paraview_fields(m, p, dims):
    """the nodal fields both writers are given: name -> values in the builder's L-dof numbering"""
    sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))
    import b200pa
    lat = m["lattice"].reshape(-1, 3)
    xyz = (lat // p + b200pa.basis(p)["gll"][lat % p]) / np.asarray(dims, dtype=np.float64)
    return {"T": 37.0 + 20.0 * np.exp(-4.0 * ((xyz - 0.5) ** 2).sum(1)),
            "phi": 30.0 * (1.0 - xyz[:, 2]) + np.sin(3.0 * xyz[:, 0]) * xyz[:, 1]}


def paraview():
    sys.path.insert(0, os.path.join(ROOT, "cardiac-ablation-ecm2_b200"))
    import b200pa
    for tag, (p, dims, size, skew, fmt, ho, lod, cycles) in PARAVIEW.items():
        m = b200pa.hex_build(*dims, p, *size, skew=skew)
        fields = paraview_fields(m, p, dims)
        with tempfile.TemporaryDirectory() as t:
            mesh_file = os.path.join(t, "slab.mesh")
            b200pa.write_mesh(mesh_file, *dims, *size, skew=skew)
            args = []
            for name, v in fields.items():
                b200pa.write_gridfunction(os.path.join(t, name + ".gf"), p, v)
                args.append(f"{name}={os.path.join(t, name + '.gf')}")
            out = os.path.join(t, "out") + "/"
            subprocess.run([DRIVER, "paraview", out, "ablation", mesh_file, fmt, str(ho), str(lod), str(cycles)] + args, check=True,
                           stdout=subprocess.DEVNULL)
            files = {}
            for dp, _, fns in os.walk(out):
                for fn in fns:
                    full = os.path.join(dp, fn)
                    files[os.path.relpath(full, out).replace("/", "|")] = np.frombuffer(open(full, "rb").read(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, f"paraview_{tag}.npz"), **files)
        print("paraview", tag, sorted(files), sum(v.nbytes for v in files.values()) // 1024, "KiB raw")


def main():
    if not os.path.exists(DRIVER):
        sys.exit("oracle/_ref/ref_driver missing: run `make -C oracle ref` in the build container")
    for tag, c in CASES.items():
        d = run(["dump_case"] + list(c))
        np.savez_compressed(os.path.join(HERE, f"case_{tag}.npz"), **d)
        print(tag, sum(v.nbytes for v in d.values()) // 1024, "KiB raw")
    num = {}
    for (nx, ny, nz) in NUMBERING:
        for p in (1, 2, 3, 4):
            if p >= 3 and nx * ny * nz > 200:
                continue
            d = run(["dump_case", p, "cart", nx, ny, nz, 1, 1, 1, "const", "all", 1])
            key = f"n{nx}_{ny}_{nz}_p{p}"
            num[key + "_gather"] = d["gather_map"]
            num[key + "_ndofs"] = d["ndofs"]
            num[key + "_ess_all"] = d["ess"]
            num[key + "_ev"] = d["elem_vertices"]
    np.savez_compressed(os.path.join(HERE, "numbering.npz"), **num)
    d = run(["dump_bioheat", 2, 4, 8])
    np.savez_compressed(os.path.join(HERE, "bioheat_p2_n4.npz"), **d)
    d = run(["dump_bioheat_steps", 2, 4, 12, 3])
    np.savez_compressed(os.path.join(HERE, "bioheat_steps_p2_n4.npz"), **d)
    markers()
    multigrid()
    paraview()
    print("done")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "markers":   # only the fixtures added in round 2
        markers()
    elif len(sys.argv) > 1 and sys.argv[1] == "multigrid":
        multigrid()
    elif len(sys.argv) > 1 and sys.argv[1] == "paraview":
        paraview()
    else:
        main()
