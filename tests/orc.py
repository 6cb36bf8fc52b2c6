"""ctypes binding of oracle/liboracle.so (the CPU restatement) — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the
product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class OrcOperator(C.Structure):
    _fields_ = [("NE", C.c_int), ("D1D", C.c_int), ("Q1D", C.c_int), ("ndofs", C.c_int),
                ("gather_map", c_ip), ("offsets", c_ip), ("indices", c_ip),
                ("B", c_dp), ("G", c_dp), ("pa_diff", c_dp), ("pa_mass", c_dp),
                ("n_ess", C.c_int), ("ess", c_ip)]


def build():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("pa_oracle.c", "pa_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_dot.restype = C.c_double
        _lib.orc_pcg.restype = C.c_int
        _lib.orc_pcg_prec.restype = C.c_int
        _lib.orc_chebyshev_coeffs.restype = C.c_int
        _lib.orc_power_method.restype = C.c_double
        _lib.orc_jacobi_setup.restype = C.c_int
    return _lib


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Operator:
    """Keeps the numpy arrays alive behind an orc_operator struct."""

    def __init__(self, D1D, Q1D, NE, ndofs, gather_map, B, G, pa_diff=None, pa_mass=None, ess=None,
                 offsets=None, indices=None):
        self.D1D, self.Q1D, self.NE, self.ndofs = int(D1D), int(Q1D), int(NE), int(ndofs)
        self.nd = self.D1D ** 3
        self.gather_map = i32(gather_map)
        if offsets is None:
            offsets = np.zeros(self.ndofs + 1, np.int32)
            indices = np.zeros(self.NE * self.nd, np.int32)
            lib().orc_restriction_tables(self.NE, self.nd, self.ndofs, ip(self.gather_map), ip(offsets), ip(indices))
        self.offsets, self.indices = i32(offsets), i32(indices)
        self.B, self.G = f64(B), f64(G)
        self.pa_diff = None if pa_diff is None else f64(pa_diff)
        self.pa_mass = None if pa_mass is None else f64(pa_mass)
        self.ess = i32(ess if ess is not None else np.zeros(0, np.int32))
        self.s = OrcOperator(self.NE, self.D1D, self.Q1D, self.ndofs, ip(self.gather_map), ip(self.offsets),
                             ip(self.indices), dp(self.B), dp(self.G), dp(self.pa_diff), dp(self.pa_mass),
                             len(self.ess), ip(self.ess))
        self.workE = np.zeros(2 * self.NE * self.nd)
        self.work = np.zeros(2 * self.ndofs)

    def mult(self, x):
        y = np.zeros(self.ndofs)
        lib().orc_op_mult(C.byref(self.s), dp(f64(x)), dp(y), dp(self.workE))
        return y

    def constrained_mult(self, x):
        y = np.zeros(self.ndofs)
        lib().orc_constrained_mult(C.byref(self.s), dp(f64(x)), dp(y), dp(self.work), dp(self.workE))
        return y

    def diag(self):
        d = np.zeros(self.ndofs)
        lib().orc_op_diag(C.byref(self.s), dp(d), dp(self.workE))
        return d

    def eliminate_rhs(self, x, b):
        b = f64(b).copy()
        lib().orc_eliminate_rhs(C.byref(self.s), dp(f64(x)), dp(b), dp(self.work), dp(self.workE))
        return b

    def jacobi_dinv(self, damping=1.0):
        dinv = np.zeros(self.ndofs)
        rc = lib().orc_jacobi_setup(self.ndofs, dp(self.diag()), len(self.ess), ip(self.ess), C.c_double(damping), dp(dinv))
        assert rc == 0, "zero diagonal"
        return dinv

    def pcg(self, dinv, b, x0, rel_tol, abs_tol, max_iter):
        x = f64(x0).copy()
        conv = C.c_int(0)
        fn = C.c_double(0)
        norms = np.zeros(max_iter + 2)
        it = lib().orc_pcg(C.byref(self.s), dp(f64(dinv)), dp(f64(b)), dp(x), C.c_double(rel_tol), C.c_double(abs_tol),
                           int(max_iter), C.byref(conv), C.byref(fn), dp(norms))
        return x, it, bool(conv.value), fn.value, norms[:it + 1]

    def power_method(self, dinv, v0, num_steps=10, tol=1e-8):
        """largest eigenvalue of Dinv*A from the start vector v0 (the reference: Vector::Randomize(12345))"""
        v = f64(v0).copy()
        return lib().orc_power_method(C.byref(self.s), dp(f64(dinv)), dp(v), int(num_steps), C.c_double(tol))

    def chebyshev_mult(self, dinv, order, max_eig, x):
        c = chebyshev_coeffs(order, max_eig)
        y = np.zeros(self.ndofs)
        lib().orc_chebyshev_mult(C.byref(self.s), dp(f64(dinv)), int(order), dp(c), dp(f64(x)), dp(y), dp(self.work), dp(self.workE))
        return y

    def pcg_chebyshev(self, dinv, order, max_eig, b, x0, rel_tol, abs_tol, max_iter):
        x = f64(x0).copy()
        conv = C.c_int(0)
        fn = C.c_double(0)
        norms = np.zeros(max_iter + 2)
        it = lib().orc_pcg_prec(C.byref(self.s), dp(f64(dinv)), int(order), C.c_double(max_eig), dp(f64(b)), dp(x), C.c_double(rel_tol),
                                C.c_double(abs_tol), int(max_iter), C.byref(conv), C.byref(fn), dp(norms))
        return x, it, bool(conv.value), fn.value, norms[:it + 1]


def chebyshev_coeffs(order, max_eig):
    c = np.zeros(order)
    assert lib().orc_chebyshev_coeffs(int(order), C.c_double(max_eig), dp(c)) == 0, "order outside 1..5"
    return c


def restrict_mult(NE, nd, gather_map, x):
    y = np.zeros(NE * nd)
    lib().orc_restrict_mult(NE, nd, ip(i32(gather_map)), dp(f64(x)), dp(y))
    return y


def restrict_mult_transpose(ndofs, offsets, indices, xE, abs_=False):
    y = np.zeros(ndofs)
    lib().orc_restrict_mult_transpose(ndofs, ip(i32(offsets)), ip(i32(indices)), dp(f64(xE)), dp(y), int(abs_))
    return y


def diffusion_setup(Q1D, NE, W, J, Cq):
    D = np.zeros(6 * Q1D ** 3 * NE)
    Cq = f64(Cq)
    lib().orc_diffusion_setup(Q1D, NE, dp(f64(W)), dp(f64(J)), dp(Cq), C.c_long(Cq.size), dp(D))
    return D


def mass_setup(Q1D, NE, W, detJ, Cq):
    v = np.zeros(Q1D ** 3 * NE)
    Cq = f64(Cq)
    lib().orc_mass_setup(Q1D ** 3, NE, dp(f64(W)), dp(f64(detJ)), dp(Cq), C.c_long(Cq.size), dp(v))
    return v


def diffusion_apply(NE, D1D, Q1D, B, G, D, xE, yE=None):
    y = np.zeros(NE * D1D ** 3) if yE is None else f64(yE).copy()
    lib().orc_diffusion_apply(NE, D1D, Q1D, dp(f64(B)), dp(f64(G)), dp(f64(D)), dp(f64(xE)), dp(y))
    return y


def mass_apply(NE, D1D, Q1D, B, v, xE, yE=None):
    y = np.zeros(NE * D1D ** 3) if yE is None else f64(yE).copy()
    lib().orc_mass_apply(NE, D1D, Q1D, dp(f64(B)), dp(f64(v)), dp(f64(xE)), dp(y))
    return y


def diffusion_diag(NE, D1D, Q1D, B, G, D):
    y = np.zeros(NE * D1D ** 3)
    lib().orc_diffusion_diag(NE, D1D, Q1D, dp(f64(B)), dp(f64(G)), dp(f64(D)), dp(y))
    return y


def mass_diag(NE, D1D, Q1D, B, v):
    y = np.zeros(NE * D1D ** 3)
    lib().orc_mass_diag(NE, D1D, Q1D, dp(f64(B)), dp(f64(v)), dp(y))
    return y


def jacobi_mult(dinv, r):
    z = np.zeros(len(r))
    lib().orc_jacobi_mult(len(r), dp(f64(dinv)), dp(f64(r)), dp(z))
    return z


def dot(a, b):
    return lib().orc_dot(C.c_long(len(a)), dp(f64(a)), dp(f64(b)))


def qvalues(NE, D1D, Q1D, B, xE):
    y = np.zeros(NE * Q1D ** 3)
    lib().orc_qvalues(NE, D1D, Q1D, dp(f64(B)), dp(f64(xE)), dp(y))
    return y


def qphysgrad(NE, D1D, Q1D, B, G, J, xE):
    g = np.zeros(3 * NE * Q1D ** 3)
    lib().orc_qphysgrad(NE, D1D, Q1D, dp(f64(B)), dp(f64(G)), dp(f64(J)), dp(f64(xE)), dp(g))
    return g


def domain_lf(NE, D1D, Q1D, B, detJ, W, f):
    b = np.zeros(NE * D1D ** 3)
    f = f64(f)
    lib().orc_domain_lf(NE, D1D, Q1D, dp(f64(B)), dp(f64(detJ)), dp(f64(W)), dp(f), C.c_long(f.size), dp(b))
    return b


# ---- element-attribute markers: the reference's orchestration restated over the element-level functions above
def _marked(attr, marker):
    attr = np.asarray(attr)
    mk = np.asarray(marker)
    return (attr > 0) & (mk[np.maximum(attr, 1) - 1] != 0)


def op_mult_markers(op, x, attr, marker_diff, marker_mass):
    """PABilinearFormExtension::MultInternal with AddMultWithMarkers (fem/bilinearform_ext.cpp:528-556, 807-847):
    every integrator's E-vector result is added for the elements its marker includes (None = all)"""
    nd = op.nd
    xE = restrict_mult(op.NE, nd, op.gather_map, x)
    yE = np.zeros(op.NE * nd)
    for pa, marker, fn in ((op.pa_diff, marker_diff, lambda v: diffusion_apply(op.NE, op.D1D, op.Q1D, op.B, op.G, op.pa_diff, v)),
                           (op.pa_mass, marker_mass, lambda v: mass_apply(op.NE, op.D1D, op.Q1D, op.B, op.pa_mass, v))):
        if pa is None:
            continue
        tmp = fn(xE)
        if marker is not None:
            tmp = tmp.reshape(op.NE, nd) * _marked(attr, marker)[:, None]
        yE += tmp.ravel()
    return restrict_mult_transpose(op.ndofs, op.offsets, op.indices, yE)


def op_diag_markers(op, attr, marker_diff, marker_mass):
    """PABilinearFormExtension::AssembleDiagonal (fem/bilinearform_ext.cpp:370-423): after every integrator the
    accumulated element diagonal of the elements its marker excludes is set to zero (earlier integrators included)"""
    nd = op.nd
    dE = np.zeros(op.NE * nd)
    for pa, marker, fn in ((op.pa_diff, marker_diff, lambda: diffusion_diag(op.NE, op.D1D, op.Q1D, op.B, op.G, op.pa_diff)),
                           (op.pa_mass, marker_mass, lambda: mass_diag(op.NE, op.D1D, op.Q1D, op.B, op.pa_mass))):
        if pa is None:
            continue
        dE += fn()
        if marker is not None:
            dE = (dE.reshape(op.NE, nd) * _marked(attr, marker)[:, None]).ravel()
    return restrict_mult_transpose(op.ndofs, op.offsets, op.indices, dE, abs_=True)
