"""ctypes binding of libb200pa.so (include/b200pa.h) for the Python-side tests, bench.py and smoke().

The product is the C-ABI shared library; this module only loads it and passes raw pointers
(torch CUDA tensors provide device memory and streams — plumbing, not compute).  There is no
CPU fallback: loading fails loudly when the library has not been built, and every compute entry
point fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200PA_LIB") or os.path.join(os.path.dirname(_HERE), "libb200pa.so")  # env: tuning builds
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
vp = C.c_void_p


class B200paError(RuntimeError):
    pass


class PcgResult(C.Structure):
    _fields_ = [("final_iter", C.c_int), ("converged", C.c_int), ("final_norm", C.c_double),
                ("initial_norm", C.c_double)]


# every symbol include/b200pa.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = """
b200pa_version b200pa_last_error b200pa_launch_count
b200pa_ctx_create b200pa_ctx_destroy b200pa_ctx_sync b200pa_ctx_stream b200pa_ctx_upload b200pa_ctx_download b200pa_malloc b200pa_free b200pa_memset b200pa_copy b200pa_host_alloc b200pa_host_free b200pa_host_node b200pa_paraview_save
b200pa_restrict_mult b200pa_restrict_mult_transpose b200pa_diffusion_setup b200pa_mass_setup
b200pa_diffusion_apply b200pa_mass_apply b200pa_diffusion_diag b200pa_mass_diag b200pa_qvalues
b200pa_qphysgrad b200pa_domain_lf b200pa_dot b200pa_add b200pa_jacobi_setup b200pa_jacobi_mult
b200pa_coeff_eval
b200pa_space_create b200pa_space_destroy b200pa_space_set_geometry b200pa_space_geometry_from_vertices
b200pa_space_offsets b200pa_space_indices b200pa_space_gather_map b200pa_space_J b200pa_space_detJ b200pa_space_W
b200pa_space_is_affine b200pa_form_set_factorised b200pa_form_is_factorised
b200pa_space_qvalues b200pa_space_qphysgrad b200pa_space_coeff_linear b200pa_space_joule b200pa_space_domain_lf
b200pa_form_create b200pa_form_destroy b200pa_form_assemble_diffusion b200pa_form_assemble_mass
b200pa_form_set_pa_data b200pa_form_pa_diff b200pa_form_pa_mass b200pa_form_set_essential b200pa_form_mult
b200pa_form_constrained_mult b200pa_form_mult_phases b200pa_form_mult_host b200pa_form_assemble_diagonal b200pa_form_eliminate_rhs
b200pa_form_assemble_diffusion_with_diagonal b200pa_space_set_attributes b200pa_form_set_markers
b200pa_pcg_solve b200pa_pcg_solve_host
b200pa_chebyshev_coeffs b200pa_power_method b200pa_chebyshev_mult b200pa_pcg_solve_chebyshev
b200pa_comm_unique_id b200pa_comm_create b200pa_comm_destroy b200pa_comm_set_tables b200pa_comm_build_tables
b200pa_comm_owner_mask b200pa_comm_px_prepare b200pa_comm_px_connect b200pa_comm_px_error b200pa_comm_px_enabled b200pa_comm_px_disable b200pa_form_set_comm b200pa_comm_exchange_sum b200pa_comm_bcast b200pa_comm_allreduce_sum
b200pa_hex_sizes b200pa_hex_build b200pa_hex_build_part b200pa_hex_dof_lattice b200pa_basis b200pa_randomize
b200pa_hex_write_mesh b200pa_write_gridfunction b200pa_basis_transfer
b200pa_transfer_create b200pa_transfer_destroy b200pa_transfer_mult b200pa_transfer_mult_transpose
b200pa_mg_create b200pa_mg_destroy b200pa_mg_set_cycle b200pa_mg_set_coarse_solver b200pa_mg_setup b200pa_mg_max_eig b200pa_mg_coarse_iterations b200pa_mg_mult b200pa_pcg_solve_mg
""".split()


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200paError(f"{LIB_PATH} is missing: run `make -C cardiac-ablation-ecm2_b200` "
                              "(or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.b200pa_last_error.restype = C.c_char_p
        L.b200pa_launch_count.restype = C.c_longlong
        L.b200pa_ctx_stream.restype = vp
        for name in ("offsets", "indices", "gather_map", "J", "detJ", "W"):
            getattr(L, f"b200pa_space_{name}").restype = vp
        L.b200pa_form_pa_diff.restype = vp
        L.b200pa_form_pa_mass.restype = vp
        L.b200pa_comm_owner_mask.restype = vp
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise B200paError(lib().b200pa_last_error().decode())


def launch_count():
    return int(lib().b200pa_launch_count())


def _ptr(t):
    """device/host pointer of a torch tensor, numpy array, int or None"""
    if t is None:
        return vp(0)
    if isinstance(t, int):
        return vp(t)
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return vp(t.ctypes.data)
    assert t.is_contiguous()
    return vp(t.data_ptr())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ------------------------------------------------------------------ host-side builder (no GPU)
def hex_sizes(nx, ny, nz, p):
    ne, nv, nd = C.c_longlong(), C.c_longlong(), C.c_longlong()
    check(lib().b200pa_hex_sizes(nx, ny, nz, p, C.byref(ne), C.byref(nv), C.byref(nd)))
    return ne.value, nv.value, nd.value


def hex_build(nx, ny, nz, p, sx=1.0, sy=1.0, sz=1.0, skew=False, part=None, want=("gather_map", "elem_vertices",
              "vertices", "elem_ijk", "bdr_attr", "lattice")):
    """Mesh::MakeCartesian3D + H1 order-p numbering.  part = (GNX,GNY,GNZ,ox,oy,oz) for a sub-box."""
    ne, nv, nd = hex_sizes(nx, ny, nz, p)
    D3 = (p + 1) ** 3
    out = {"ne": ne, "nv": nv, "ndofs": nd, "p": p, "D1D": p + 1, "Q1D": p + 2}
    arr = {
        "gather_map": np.empty(ne * D3, np.int32), "elem_vertices": np.empty(8 * ne, np.int32),
        "vertices": np.empty(3 * nv, np.float64), "elem_ijk": np.empty(3 * ne, np.int32),
        "bdr_attr": np.empty(nd, np.uint8), "lattice": np.empty(3 * nd, np.int32)}
    for k in arr:
        out[k] = arr[k] if k in want else None
    G = part if part is not None else (nx, ny, nz, 0, 0, 0)
    check(lib().b200pa_hex_build_part(*[int(v) for v in G], nx, ny, nz, p, C.c_double(sx), C.c_double(sy), C.c_double(sz),
                                      int(bool(skew)), _ptr(out["gather_map"]), _ptr(out["elem_vertices"]),
                                      _ptr(out["vertices"]), _ptr(out["elem_ijk"]), _ptr(out["bdr_attr"]),
                                      _ptr(out["lattice"])))
    return out


def basis(p, q1d=None):
    q1d = p + 2 if q1d is None else q1d
    D = p + 1
    B, G = np.empty(q1d * D), np.empty(q1d * D)
    w1d, W, gll = np.empty(q1d), np.empty(q1d ** 3), np.empty(D)
    check(lib().b200pa_basis(p, q1d, _ptr(B), _ptr(G), _ptr(w1d), _ptr(W), _ptr(gll)))
    return {"B": B, "G": G, "w1d": w1d, "W": W, "gll": gll}


def basis_transfer(pc, pf):
    """[pf+1, pc+1] column-major: coarse GLL-nodal basis at the fine GLL nodes (order-refinement transfer)"""
    B = np.empty((pf + 1) * (pc + 1))
    check(lib().b200pa_basis_transfer(int(pc), int(pf), _ptr(B)))
    return B


def randomize(n, seed=1):
    """Vector::Randomize(seed): the reference's test / benchmark input vector"""
    out = np.empty(int(n))
    check(lib().b200pa_randomize(int(seed), C.c_longlong(int(n)), _ptr(out)))
    return out


def write_mesh(path, nx, ny, nz, sx=1.0, sy=1.0, sz=1.0, skew=False):
    """Mesh::Print ("MFEM mesh v1.0") of the mesh hex_build numbers"""
    check(lib().b200pa_hex_write_mesh(str(path).encode(), int(nx), int(ny), int(nz), C.c_double(sx), C.c_double(sy), C.c_double(sz),
                                      int(bool(skew))))


def write_gridfunction(path, p, values):
    """GridFunction::Save of a scalar H1 order-p field in the builder's L-dof numbering"""
    v = _f64(values)
    check(lib().b200pa_write_gridfunction(str(path).encode(), int(p), C.c_longlong(v.size), _ptr(v)))


def paraview_save(prefix_path, collection, mesh, p, fields, cycle=0, time=0.0, rank=0, nranks=1, levels_of_detail=None,
                  high_order=True, fmt="binary", attributes=None, append=None):
    """ParaViewDataCollection::Save for a mesh dict of hex_build / partition.build_part (gather_map, vertices,
    elem_vertices) and {name: host values in L-dof numbering}; fmt: ascii | binary | binary32"""
    names = sorted(fields)                     # the reference's field map iterates in name order
    vals = [_f64(fields[n]) for n in names]
    gm, vx, ev = _i32(mesh["gather_map"]), _f64(mesh["vertices"]), _i32(mesh["elem_vertices"])
    at = None if attributes is None else _i32(attributes)
    c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
    c_vals = (vp * len(names))(*[v.ctypes.data for v in vals])
    check(lib().b200pa_paraview_save(str(prefix_path).encode(), str(collection).encode(), int(cycle), C.c_double(time), int(rank), int(nranks),
                                     int(p), C.c_longlong(mesh["ne"]), C.c_longlong(mesh["ndofs"]), _ptr(gm), _ptr(vx), _ptr(ev), _ptr(at),
                                     len(names), c_names, c_vals, int(p if levels_of_detail is None else levels_of_detail), int(bool(high_order)),
                                     {"ascii": 0, "binary": 1, "binary32": 2}[fmt], int(cycle > 0 if append is None else append)))


def chebyshev_coeffs(order, max_eig):
    """OperatorChebyshevSmoother::Setup's polynomial coefficients (linalg/solvers.cpp:571-621)"""
    c = np.zeros(int(order))
    check(lib().b200pa_chebyshev_coeffs(int(order), C.c_double(max_eig), _ptr(c)))
    return c


def essential_dofs(bdr_attr, attrs):
    """GetEssentialTrueDofs for a list of boundary attributes (1..6): ascending dof ids."""
    mask = 0
    for a in attrs:
        mask |= 1 << (a - 1)
    return np.nonzero(bdr_attr & mask)[0].astype(np.int32)


def comm_build_tables(rank, ndofs, nbr_rank, shared_offsets, shared_ldofs):
    nbr_rank, shared_offsets, shared_ldofs = _i32(nbr_rank), _i32(shared_offsets), _i32(shared_ldofs)
    n_nbr = len(nbr_rank)
    ns = C.c_int(0)
    check(lib().b200pa_comm_build_tables(rank, ndofs, n_nbr, _ptr(nbr_rank), _ptr(shared_offsets), _ptr(shared_ldofs),
                                         C.byref(ns), None, None, None, None))
    n_send = int(shared_offsets[n_nbr]) if n_nbr else 0
    sh_ldof, sh_off = np.zeros(max(ns.value, 1), np.int32), np.zeros(ns.value + 1, np.int32)
    sh_src, mask = np.zeros(n_send + ns.value + 1, np.int32), np.zeros(max(ndofs, 1), np.uint8)
    check(lib().b200pa_comm_build_tables(rank, ndofs, n_nbr, _ptr(nbr_rank), _ptr(shared_offsets), _ptr(shared_ldofs),
                                         C.byref(ns), _ptr(sh_ldof), _ptr(sh_off), _ptr(sh_src), _ptr(mask)))
    return sh_ldof[:ns.value], sh_off, sh_src[:sh_off[ns.value]], mask[:ndofs]


# ------------------------------------------------------------------------------ GPU handles
class Context:
    """One per GPU.  By default kernels are enqueued on torch's current stream of that device
    (the legacy default stream is passed as cudaStreamLegacy), so torch allocations / copies and
    the library's kernels are ordered without extra synchronisation; own_stream=True lets the
    library create its own non-blocking stream instead."""

    def __init__(self, device=0, own_stream=False):
        import torch
        self.h = vp()
        self.device = torch.device("cuda", device)
        if own_stream:
            check(lib().b200pa_ctx_create(int(device), None, C.byref(self.h)))
            self.stream_ptr = lib().b200pa_ctx_stream(self.h)
            self.torch_stream = torch.cuda.ExternalStream(self.stream_ptr, device=self.device)
        else:
            if not torch.cuda.is_available():
                check(lib().b200pa_ctx_create(int(device), None, C.byref(self.h)))  # raises: no CUDA device
            self.torch_stream = torch.cuda.current_stream(self.device)
            sp = self.torch_stream.cuda_stream or 1  # 0 (default stream) -> cudaStreamLegacy
            check(lib().b200pa_ctx_create(int(device), vp(sp), C.byref(self.h)))
            self.stream_ptr = sp

    def sync(self):
        check(lib().b200pa_ctx_sync(self.h))

    def close(self):
        if self.h:
            lib().b200pa_ctx_destroy(self.h)
            self.h = vp()

    # device memory through torch, ordered on the context's stream
    def to_dev(self, a, dtype=None):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
        with torch.cuda.stream(self.torch_stream):
            d = t.to(self.device, non_blocking=False)
        return d

    def empty(self, n, dtype=None):
        import torch
        with torch.cuda.stream(self.torch_stream):
            return torch.empty(int(n), dtype=dtype or torch.float64, device=self.device)

    def zeros(self, n, dtype=None):
        import torch
        with torch.cuda.stream(self.torch_stream):
            return torch.zeros(int(n), dtype=dtype or torch.float64, device=self.device)

    def pinned(self, n, fill=None):
        """float64 numpy vector in page-locked host memory on the GPU's NUMA node (b200pa_host_alloc) for the *_host entry
        points; freed when the array (and every view of it) is gone.  `.numa_node` of the owner: -1 = kernel default"""
        import weakref
        p = vp()
        check(lib().b200pa_host_alloc(self.h, C.c_size_t(8 * int(n)), C.byref(p)))
        buf = (C.c_double * int(n)).from_address(p.value)
        a = np.ctypeslib.as_array(buf)
        addr = p.value
        weakref.finalize(buf, lambda: lib().b200pa_host_free(None, vp(addr)))   # may run after the context is closed
        if fill is not None:
            a[:] = fill
        return a

    def host_node(self, a):
        return int(lib().b200pa_host_node(_ptr(a)))

    def to_host(self, t):
        self.sync()
        return t.cpu().numpy()

    # ---- level-1 kernels (device tensors in, device tensors out)
    def restrict_mult(self, ne, nd, gmap, x):
        y = self.empty(ne * nd)
        check(lib().b200pa_restrict_mult(self.h, ne, nd, _ptr(gmap), _ptr(x), _ptr(y)))
        return y

    def restrict_mult_transpose(self, ndofs, offsets, indices, xE, abs_=False):
        y = self.empty(ndofs)
        check(lib().b200pa_restrict_mult_transpose(self.h, ndofs, _ptr(offsets), _ptr(indices), _ptr(xE), _ptr(y), int(abs_)))
        return y

    def diffusion_setup(self, q1d, ne, W, J, Cq):
        D = self.empty(6 * q1d ** 3 * ne)
        check(lib().b200pa_diffusion_setup(self.h, q1d, ne, _ptr(W), _ptr(J), _ptr(Cq), C.c_longlong(Cq.numel()), _ptr(D)))
        return D

    def mass_setup(self, q1d, ne, W, detJ, Cq):
        v = self.empty(q1d ** 3 * ne)
        check(lib().b200pa_mass_setup(self.h, q1d ** 3, ne, _ptr(W), _ptr(detJ), _ptr(Cq), C.c_longlong(Cq.numel()), _ptr(v)))
        return v

    def diffusion_apply(self, ne, d1d, q1d, B, G, D, xE, yE):
        check(lib().b200pa_diffusion_apply(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(_f64(G)), _ptr(D), _ptr(xE), _ptr(yE)))
        return yE

    def mass_apply(self, ne, d1d, q1d, B, v, xE, yE):
        check(lib().b200pa_mass_apply(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(v), _ptr(xE), _ptr(yE)))
        return yE

    def diffusion_diag(self, ne, d1d, q1d, B, G, D, dE):
        check(lib().b200pa_diffusion_diag(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(_f64(G)), _ptr(D), _ptr(dE)))
        return dE

    def mass_diag(self, ne, d1d, q1d, B, v, dE):
        check(lib().b200pa_mass_diag(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(v), _ptr(dE)))
        return dE

    def qvalues(self, ne, d1d, q1d, B, xE):
        y = self.empty(ne * q1d ** 3)
        check(lib().b200pa_qvalues(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(xE), _ptr(y)))
        return y

    def qphysgrad(self, ne, d1d, q1d, B, G, J, xE):
        g = self.empty(3 * ne * q1d ** 3)
        check(lib().b200pa_qphysgrad(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(_f64(G)), _ptr(J), _ptr(xE), _ptr(g)))
        return g

    def domain_lf(self, ne, d1d, q1d, B, detJ, W, f, bE):
        check(lib().b200pa_domain_lf(self.h, ne, d1d, q1d, _ptr(_f64(B)), _ptr(detJ), _ptr(W), _ptr(f),
                                     C.c_longlong(f.numel()), _ptr(bE)))
        return bE

    def dot(self, a, b):
        r = C.c_double(0)
        check(lib().b200pa_dot(self.h, C.c_longlong(a.numel()), _ptr(a), _ptr(b), C.byref(r)))
        return r.value

    def add(self, v1, alpha, v2, out=None):
        out = self.empty(v1.numel()) if out is None else out
        check(lib().b200pa_add(self.h, C.c_longlong(v1.numel()), _ptr(v1), C.c_double(alpha), _ptr(v2), _ptr(out)))
        return out

    def jacobi_setup(self, diag, ess, damping=1.0):
        dinv = self.empty(diag.numel())
        n_ess = 0 if ess is None else ess.numel()
        check(lib().b200pa_jacobi_setup(self.h, diag.numel(), _ptr(diag), n_ess, _ptr(ess) if n_ess else None,
                                        C.c_double(damping), _ptr(dinv)))
        return dinv

    def jacobi_mult(self, dinv, r):
        z = self.empty(r.numel())
        check(lib().b200pa_jacobi_mult(self.h, r.numel(), _ptr(dinv), _ptr(r), _ptr(z)))
        return z

    def coeff_eval(self, kind, n, a, b, T0, T=None, s=None, g=None):
        out = self.empty(n)
        check(lib().b200pa_coeff_eval(self.h, kind, C.c_longlong(n), C.c_double(a), C.c_double(b), C.c_double(T0),
                                      _ptr(T), _ptr(s), _ptr(g), _ptr(out)))
        return out


class Space:
    """ElementRestriction + DofToQuad + GeometricFactors of one H1 hex space."""

    def __init__(self, ctx, d1d, q1d, ne, ndofs, gather_map, B, G):
        self.ctx, self.d1d, self.q1d, self.ne, self.ndofs = ctx, int(d1d), int(q1d), int(ne), int(ndofs)
        self.nd, self.nq = self.d1d ** 3, self.q1d ** 3
        self.h = vp()
        self._keep = []
        gm = gather_map if not isinstance(gather_map, np.ndarray) else _i32(gather_map)
        check(lib().b200pa_space_create(ctx.h, self.d1d, self.q1d, self.ne, self.ndofs, _ptr(gm), _ptr(_f64(B)),
                                        _ptr(_f64(G)), C.byref(self.h)))

    def set_geometry(self, W, J, detJ):
        """W host or device; J/detJ device tensors are referenced (kept alive here), host arrays copied."""
        self._keep += [J, detJ]
        W = _f64(W) if isinstance(W, np.ndarray) else W
        J = _f64(J) if isinstance(J, np.ndarray) else J
        detJ = _f64(detJ) if isinstance(detJ, np.ndarray) else detJ
        check(lib().b200pa_space_set_geometry(self.h, _ptr(W), _ptr(J), _ptr(detJ)))

    def set_attributes(self, attr):
        """Mesh::GetAttribute(e) for every element (needed by integrator markers)"""
        a = _i32(attr)
        assert len(a) == self.ne
        check(lib().b200pa_space_set_attributes(self.h, _ptr(a)))

    @property
    def affine(self):
        return lib().b200pa_space_is_affine(self.h) == 1

    def geometry_from_vertices(self, W, vertices, elem_vertices):
        v, ev = _f64(vertices), _i32(elem_vertices)
        check(lib().b200pa_space_geometry_from_vertices(self.h, _ptr(_f64(W)), len(v) // 3, _ptr(v), _ptr(ev)))

    def _view(self, name, n, dtype):
        """host copy of one of the space's device arrays"""
        p = getattr(lib(), f"b200pa_space_{name}")(self.h)
        out = np.empty(n, np.float64 if dtype == "f64" else np.int32)
        check(lib().b200pa_ctx_download(self.ctx.h, _ptr(out), vp(p), C.c_size_t(out.nbytes)))
        return out

    def offsets(self):
        return self._view("offsets", self.ndofs + 1, "i32")

    def indices(self):
        return self._view("indices", self.ne * self.nd, "i32")

    def J(self):
        return self._view("J", 9 * self.ne * self.nq, "f64")

    def detJ(self):
        return self._view("detJ", self.ne * self.nq, "f64")

    def qvalues(self, xL):
        y = self.ctx.empty(self.ne * self.nq)
        check(lib().b200pa_space_qvalues(self.h, _ptr(xL), _ptr(y)))
        return y

    def qphysgrad(self, xL):
        g = self.ctx.empty(3 * self.ne * self.nq)
        check(lib().b200pa_space_qphysgrad(self.h, _ptr(xL), _ptr(g)))
        return g

    def coeff_linear(self, a, b, T0, TL, out=None):
        out = self.ctx.empty(self.ne * self.nq) if out is None else out
        check(lib().b200pa_space_coeff_linear(self.h, C.c_double(a), C.c_double(b), C.c_double(T0), _ptr(TL), _ptr(out)))
        return out

    def joule(self, phiL, sigma_q, add, out=None):
        out = self.ctx.empty(self.ne * self.nq) if out is None else out
        check(lib().b200pa_space_joule(self.h, _ptr(phiL), _ptr(sigma_q), C.c_double(add), _ptr(out)))
        return out

    def domain_lf(self, f, out=None):
        out = self.ctx.empty(self.ndofs) if out is None else out
        check(lib().b200pa_space_domain_lf(self.h, _ptr(f), C.c_longlong(f.numel()), _ptr(out)))
        return out

    def close(self):
        if self.h:
            lib().b200pa_space_destroy(self.h)
            self.h = vp()


class Form:
    """PABilinearFormExtension (+ ConstrainedOperator) with diffusion and/or mass on a Space."""

    def __init__(self, space):
        self.sp, self.ctx = space, space.ctx
        self.h = vp()
        self._keep = []
        check(lib().b200pa_form_create(space.h, C.byref(self.h)))

    def set_factorised(self, on=True):
        """factorised diffusion q-data (affine meshes): w_q c_q per q-point + one tensor per element"""
        check(lib().b200pa_form_set_factorised(self.h, 1 if on else 0))

    @property
    def factorised(self):
        return bool(lib().b200pa_form_is_factorised(self.h))

    def assemble_diffusion(self, Cq):
        if Cq is None:
            check(lib().b200pa_form_assemble_diffusion(self.h, None, C.c_longlong(0)))
            return
        Cq = _f64(Cq) if isinstance(Cq, np.ndarray) else Cq
        n = Cq.size if isinstance(Cq, np.ndarray) else Cq.numel()
        check(lib().b200pa_form_assemble_diffusion(self.h, _ptr(Cq), C.c_longlong(n)))

    def assemble_mass(self, Cq):
        if Cq is None:
            check(lib().b200pa_form_assemble_mass(self.h, None, C.c_longlong(0)))
            return
        Cq = _f64(Cq) if isinstance(Cq, np.ndarray) else Cq
        n = Cq.size if isinstance(Cq, np.ndarray) else Cq.numel()
        check(lib().b200pa_form_assemble_mass(self.h, _ptr(Cq), C.c_longlong(n)))

    def set_pa_data(self, pa_diff, pa_mass):
        self._keep = [pa_diff, pa_mass]
        check(lib().b200pa_form_set_pa_data(self.h, _ptr(pa_diff), _ptr(pa_mass)))

    def set_markers(self, which, marker):
        """AddDomainIntegrator(bfi, elem_marker): which = 0 diffusion / 1 mass; marker[a-1] != 0 <=> acts on attribute a"""
        if marker is None:
            check(lib().b200pa_form_set_markers(self.h, int(which), 0, None))
            return
        mk = _i32(marker)
        check(lib().b200pa_form_set_markers(self.h, int(which), len(mk), _ptr(mk)))

    def set_essential(self, ess):
        ess = _i32(ess if ess is not None else np.zeros(0, np.int32))
        self.n_ess = len(ess)
        self.ess_dev = self.ctx.to_dev(ess) if len(ess) else None
        check(lib().b200pa_form_set_essential(self.h, len(ess), _ptr(ess) if len(ess) else None))

    def set_comm(self, comm):
        self._comm = comm
        check(lib().b200pa_form_set_comm(self.h, comm.h if comm is not None else None))

    def mult(self, x, y=None):
        y = self.ctx.empty(self.sp.ndofs) if y is None else y
        check(lib().b200pa_form_mult(self.h, _ptr(x), _ptr(y)))
        return y

    def mult_phases(self, x, y, phases):
        check(lib().b200pa_form_mult_phases(self.h, _ptr(x), _ptr(y), int(phases)))
        return y

    def constrained_mult(self, x, y=None):
        y = self.ctx.empty(self.sp.ndofs) if y is None else y
        check(lib().b200pa_form_constrained_mult(self.h, _ptr(x), _ptr(y)))
        return y

    def mult_host(self, x_host, y_host, constrained=False):
        check(lib().b200pa_form_mult_host(self.h, int(constrained), _ptr(x_host), _ptr(y_host)))
        return y_host

    def assemble_diagonal(self, diag=None):
        diag = self.ctx.empty(self.sp.ndofs) if diag is None else diag
        check(lib().b200pa_form_assemble_diagonal(self.h, _ptr(diag)))
        return diag

    def assemble_diffusion_with_diagonal(self, Cq, diag=None):
        """AssemblePA of the diffusion integrator + the form's diagonal in one pass over the q-points"""
        diag = self.ctx.empty(self.sp.ndofs) if diag is None else diag
        Cq = _f64(Cq) if isinstance(Cq, np.ndarray) else Cq
        n = Cq.size if isinstance(Cq, np.ndarray) else Cq.numel()
        check(lib().b200pa_form_assemble_diffusion_with_diagonal(self.h, _ptr(Cq), C.c_longlong(n), _ptr(diag)))
        return diag

    def jacobi_from(self, diag, damping=1.0):
        """OperatorJacobiSmoother from an already assembled diagonal"""
        return self.ctx.jacobi_setup(diag, self.ess_dev, damping)

    def eliminate_rhs(self, x, b):
        check(lib().b200pa_form_eliminate_rhs(self.h, _ptr(x), _ptr(b)))
        return b

    def jacobi(self, damping=1.0):
        """OperatorJacobiSmoother(a, ess): dinv on the device."""
        return self.ctx.jacobi_setup(self.assemble_diagonal(), self.ess_dev, damping)

    def pcg(self, dinv, b, x, rel_tol=0.0, abs_tol=0.0, max_iter=100, want_norms=True, host=False):
        res = PcgResult()
        norms = np.zeros(max_iter + 2) if want_norms else None
        fn = lib().b200pa_pcg_solve_host if host else lib().b200pa_pcg_solve
        check(fn(self.h, _ptr(dinv), _ptr(b), _ptr(x), C.c_double(rel_tol), C.c_double(abs_tol), int(max_iter),
                 C.byref(res), _ptr(norms) if want_norms else None))
        if getattr(self, "_comm", None) is not None:
            self._comm.check_p2p()
        return res, (norms[:res.final_iter + 1] if want_norms else None)

    # ---- OperatorChebyshevSmoother (linalg/solvers.cpp:455-657)
    def power_method(self, dinv, v0, num_steps=10, tol=1e-8):
        """largest eigenvalue of Dinv*A; v0 (device) = start vector, overwritten (reference: Vector::Randomize(12345))"""
        lam = C.c_double(0.0)
        check(lib().b200pa_power_method(self.h, _ptr(dinv), _ptr(v0), int(num_steps), C.c_double(tol), C.byref(lam)))
        return lam.value

    def chebyshev_mult(self, dinv, order, max_eig, x, y=None):
        y = self.ctx.empty(self.sp.ndofs) if y is None else y
        check(lib().b200pa_chebyshev_mult(self.h, _ptr(dinv), int(order), C.c_double(max_eig), _ptr(x), _ptr(y)))
        return y

    def pcg_chebyshev(self, dinv, order, max_eig, b, x, rel_tol=0.0, abs_tol=0.0, max_iter=100, want_norms=True):
        res = PcgResult()
        norms = np.zeros(max_iter + 2) if want_norms else None
        check(lib().b200pa_pcg_solve_chebyshev(self.h, _ptr(dinv), int(order), C.c_double(max_eig), _ptr(b), _ptr(x), C.c_double(rel_tol),
                                               C.c_double(abs_tol), int(max_iter), C.byref(res), _ptr(norms) if want_norms else None))
        if getattr(self, "_comm", None) is not None:
            self._comm.check_p2p()
        return res, (norms[:res.final_iter + 1] if want_norms else None)

    def close(self):
        if self.h:
            lib().b200pa_form_destroy(self.h)
            self.h = vp()


class Transfer:
    """TensorProductPRefinementTransferOperator between two forms on the same mesh (coarse, fine)"""

    def __init__(self, coarse, fine, B):
        self.fc, self.ff, self.ctx = coarse, fine, coarse.ctx
        self.h = vp()
        check(lib().b200pa_transfer_create(coarse.h, fine.h, _ptr(_f64(B)), C.byref(self.h)))

    def mult(self, xc, yf=None):
        yf = self.ctx.empty(self.ff.sp.ndofs) if yf is None else yf
        check(lib().b200pa_transfer_mult(self.h, _ptr(xc), _ptr(yf)))
        return yf

    def mult_transpose(self, xf, yc=None):
        yc = self.ctx.empty(self.fc.sp.ndofs) if yc is None else yc
        check(lib().b200pa_transfer_mult_transpose(self.h, _ptr(xf), _ptr(yc)))
        return yc

    def close(self):
        if self.h:
            lib().b200pa_transfer_destroy(self.h)
            self.h = vp()


class Multigrid:
    """Multigrid (fem/multigrid.hpp) over forms[0] (coarsest) .. forms[-1], Chebyshev smoothers, CG coarse solve"""

    def __init__(self, forms, transfers):
        self.forms, self.transfers, self.ctx = forms, transfers, forms[0].ctx
        self.h = vp()
        fa = (vp * len(forms))(*[f.h for f in forms])
        ta = (vp * max(len(transfers), 1))(*[t.h for t in transfers])
        check(lib().b200pa_mg_create(len(forms), fa, ta, C.byref(self.h)))

    def set_cycle(self, wcycle=False, pre=1, post=1):
        check(lib().b200pa_mg_set_cycle(self.h, int(wcycle), int(pre), int(post)))

    def set_coarse_solver(self, rel_tol, abs_tol, max_iter, jacobi=False):
        check(lib().b200pa_mg_set_coarse_solver(self.h, C.c_double(rel_tol), C.c_double(abs_tol), int(max_iter), int(jacobi)))

    def setup(self, order=None, max_eig=None):
        n = len(self.forms)
        o = None if order is None else _i32(order)
        e = None if max_eig is None else _f64(max_eig)
        assert (o is None or len(o) == n) and (e is None or len(e) == n)
        check(lib().b200pa_mg_setup(self.h, _ptr(o), _ptr(e)))

    def max_eig(self, level):
        lib().b200pa_mg_max_eig.restype = C.c_double
        return lib().b200pa_mg_max_eig(self.h, int(level))

    def coarse_iterations(self):
        return int(lib().b200pa_mg_coarse_iterations(self.h))

    def mult(self, x, y=None):
        y = self.ctx.empty(self.forms[-1].sp.ndofs) if y is None else y
        check(lib().b200pa_mg_mult(self.h, _ptr(x), _ptr(y)))
        return y

    def pcg(self, b, x, rel_tol=0.0, abs_tol=0.0, max_iter=100, want_norms=True):
        res = PcgResult()
        norms = np.zeros(max_iter + 2) if want_norms else None
        check(lib().b200pa_pcg_solve_mg(self.h, _ptr(b), _ptr(x), C.c_double(rel_tol), C.c_double(abs_tol), int(max_iter),
                                        C.byref(res), _ptr(norms) if want_norms else None))
        return res, (norms[:res.final_iter + 1] if want_norms else None)

    def close(self):
        if self.h:
            lib().b200pa_mg_destroy(self.h)
            self.h = vp()


class Comm:
    """Shared-dof exchange + all-reduce over NCCL (one rank per GPU)."""

    def __init__(self, ctx, nccl_id, rank, nranks):
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        self.h = vp()
        # nccl_id=None: peer-memory transport only (ranks that share one GPU; NCCL refuses those)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(bytes(nccl_id)) if nccl_id is not None else None
        check(lib().b200pa_comm_create(ctx.h, idbuf, rank, nranks, C.byref(self.h)))

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        check(lib().b200pa_comm_unique_id(buf))
        return bytes(buf)

    def set_tables(self, ndofs, nbr_rank, shared_offsets, shared_ldofs, p2p=None):
        """neighbour tables; p2p=None enables the peer-memory path when torch.distributed is initialised
        (B200PA_NO_P2P=1 keeps NCCL send/recv + all-reduce)"""
        nbr_rank, shared_offsets, shared_ldofs = _i32(nbr_rank), _i32(shared_offsets), _i32(shared_ldofs)
        check(lib().b200pa_comm_set_tables(self.h, ndofs, len(nbr_rank), _ptr(nbr_rank), _ptr(shared_offsets),
                                           _ptr(shared_ldofs)))
        self._nbr_rank, self._offs = [int(v) for v in nbr_rank], [int(v) for v in shared_offsets]
        if p2p is None:
            p2p = os.environ.get("B200PA_NO_P2P", "0") != "1"
        if p2p and self.nranks > 1:
            self.enable_p2p()

    def enable_p2p(self):
        """map every rank's mailbox into every other rank (CUDA IPC) - collective over torch.distributed"""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or self.nranks > 8:
            return False
        hbuf = (C.c_ubyte * 64)()
        check(lib().b200pa_comm_px_prepare(self.h, hbuf))
        info = [None] * self.nranks
        dist.all_gather_object(info, (bytes(hbuf), self._nbr_rank, self._offs))
        handles = (C.c_ubyte * (64 * self.nranks)).from_buffer_copy(b"".join(i[0] for i in info))
        n = len(self._nbr_rank)
        roff, rns = (C.c_longlong * max(n, 1))(), (C.c_longlong * max(n, 1))()
        for k, q in enumerate(self._nbr_rank):
            qn, qo = info[q][1], info[q][2]
            idx = qn.index(self.rank)
            roff[k], rns[k] = qo[idx], qo[len(qn)]
        ok = lib().b200pa_comm_px_connect(self.h, handles, roff, rns) == 0
        oks = [None] * self.nranks
        dist.all_gather_object(oks, ok)          # all ranks take the same decision
        if not all(oks):
            lib().b200pa_comm_px_disable(self.h)  # NCCL send/recv + all-reduce stay in place
            return False
        dist.barrier()
        return True

    def p2p_enabled(self):
        return bool(lib().b200pa_comm_px_enabled(self.h))

    def check_p2p(self):
        if lib().b200pa_comm_px_error(self.h):
            raise B200paError("b200pa: a peer-memory wait timed out (a rank did not take part in the exchange / all-reduce)")

    def exchange_sum(self, y):
        check(lib().b200pa_comm_exchange_sum(self.h, _ptr(y)))

    def bcast(self, x):
        check(lib().b200pa_comm_bcast(self.h, _ptr(x)))

    def allreduce_sum(self, vals):
        check(lib().b200pa_comm_allreduce_sum(self.h, _ptr(vals), vals.numel()))

    def close(self):
        if self.h:
            lib().b200pa_comm_destroy(self.h)
            self.h = vp()
