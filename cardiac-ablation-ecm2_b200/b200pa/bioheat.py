"""The RF-ablation coupled step of SURVEY.md §3.2/§3.3 composed from the C-ABI entry points
(host-side driver; the reference composes the same step from its own PA API - the tests and the bench
compare against that composition run by the unmodified reference).

  (1) electrostatics  div sigma(T) grad phi = 0, phi = V on z=0, 0 on z=1      -> PCG + Jacobi
  (2) Joule source    q = sigma |grad phi|^2 + w_b rho_b c_b T_a               -> q-point kernel
  (3) bioheat         (rho c/dt + w_b rho_b c_b) M T1 + K(k(T0)) T1 = (rho c/dt) M T0 + b(q)   -> PCG + Jacobi
"""
import numpy as np

from . import Form, essential_dofs

PHYS = dict(dt=0.5, rc=3.6e6, wbcb=4.0e4, Ta=37.0, k0=0.5, ak=0.02, s0=0.3, as_=0.015, V=30.0)


def initial_temperature(lattice, gll, p, GN):
    """GridFunction::ProjectCoefficient of 37 + 20 exp(-40 r^2): nodal values at the GLL points"""
    lat = lattice.reshape(-1, 3)
    xyz = (lat // p + gll[lat % p]) / np.asarray(GN, dtype=np.float64)
    return 37.0 + 20.0 * np.exp(-40.0 * ((xyz - 0.5) ** 2).sum(1))


class CoupledStep:
    """Keeps the three forms and the q-data buffers alive across time steps (device resident)."""

    def __init__(self, ctx, sp, mesh, GN, P=PHYS, comm=None, factorised=False):
        self.ctx, self.sp, self.m, self.P, self.comm = ctx, sp, mesh, P, comm
        nq = sp.ne * sp.nq
        self.nq = nq
        self.kq, self.sq, self.src = ctx.empty(nq), ctx.empty(nq), ctx.empty(nq)
        self.mq = ctx.coeff_eval(1, nq, P["rc"] / P["dt"] + P["wbcb"], 0.0, 0.0)
        self.ess = essential_dofs(mesh["bdr_attr"], [1, 6])
        lat = mesh["lattice"].reshape(-1, 3)
        self.phi_bc = np.zeros(mesh["ndofs"])
        self.phi_bc[self.ess] = P["V"] * (1.0 - lat[self.ess, 2] / (mesh["p"] * GN[2]))
        self.phi_bc_dev = ctx.to_dev(self.phi_bc)       # uploaded once: no host copy inside the time loop
        self.diag = ctx.empty(mesh["ndofs"])
        self.fe, self.ft, self.fm = Form(sp), Form(sp), Form(sp)
        if factorised:  # affine mesh: sigma(T), k(T) q-data as one scalar per q-point (b200pa_form_set_factorised)
            self.fe.set_factorised(True)
            self.ft.set_factorised(True)
        self.fe.set_essential(self.ess)
        self.ft.set_essential(None)
        self.fm.assemble_mass(np.array([P["rc"] / P["dt"]]))
        self.fm.set_essential(None)
        for f in (self.fe, self.ft, self.fm):
            if comm is not None:
                f.set_comm(comm)

    def step(self, T0, iters_e, iters_t, rel_tol=0.0):
        ctx, sp, P = self.ctx, self.sp, self.P
        sp.coeff_linear(P["k0"], P["ak"], 37.0, T0, out=self.kq)
        sp.coeff_linear(P["s0"], P["as_"], 37.0, T0, out=self.sq)
        # (1)
        dinv_e = self.fe.jacobi_from(self.fe.assemble_diffusion_with_diagonal(self.sq, self.diag))
        phi = self.phi_bc_dev.clone()
        Be = ctx.zeros(self.m["ndofs"])
        self.fe.eliminate_rhs(phi, Be)
        res_e, _ = self.fe.pcg(dinv_e, Be, phi, rel_tol, 0.0, iters_e, want_norms=False)
        # (2)
        sp.joule(phi, self.sq, P["wbcb"] * P["Ta"], out=self.src)
        # (3)
        self.ft.assemble_mass(self.mq)
        dinv_t = self.ft.jacobi_from(self.ft.assemble_diffusion_with_diagonal(self.kq, self.diag))
        rhs = sp.domain_lf(self.src)
        if self.comm is not None:
            self.comm.exchange_sum(rhs)
        rhs = ctx.add(rhs, 1.0, self.fm.mult(T0))
        T1 = T0.clone()
        res_t, _ = self.ft.pcg(dinv_t, rhs, T1, rel_tol, 0.0, iters_t, want_norms=False)
        return dict(phi=phi, Be=Be, src=self.src, rhs=rhs, T1=T1, res_e=res_e, res_t=res_t)

    def run(self, T0, nsteps, iters_e, iters_t, rel_tol=0.0):
        """time loop (≙ BackwardEulerSolver::Step over ImplicitSolve, linalg/ode.cpp:682-696): T, phi and all
        q-data stay on the device between steps; returns the list of temperatures [T^1 .. T^nsteps]"""
        out, T = [], T0
        for _ in range(nsteps):
            o = self.step(T, iters_e, iters_t, rel_tol)
            T = o["T1"]
            out.append(o)
        return out

    def close(self):
        for f in (self.fe, self.ft, self.fm):
            f.close()
