"""GPU: the UNMODIFIED reference (stock MFEM, CPU build) driving the B200 through the drop-in binding
cardiac-ablation-ecm2_b200/host/mfem_b200pa.hpp, compared in-process with the reference's own CPU
partial-assembly path (oracle/shim_check.cpp -> oracle/_ref/shim_check, built in the build container
by `make -C oracle ref`; it travels to the GPU box with the snapshot)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "shim_check")


def run(args, timeout=900):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/shim_check not built (needs the reference tree: make -C oracle ref)")
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    out = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout, env=env)
    recs = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return recs


@pytest.mark.parametrize("p,dims", [(1, (5, 4, 3)), (2, (6, 5, 4)), (3, (4, 3, 3)), (4, (3, 3, 2)), (5, (2, 3, 2)), (6, (2, 2, 2))])
def test_mfem_drop_in(p, dims):
    recs = run(["apply", p, *dims])
    assert len(recs) == 2
    for r in recs:
        assert r["ok"]
        assert max(r["integrator_level"].values()) <= 1e-12
        f = r["fused"]
        assert max(f["apply"], f["diag"], f["constrained"], f["rhs"]) <= 1e-12 and f["pcg10"] <= 1e-10
        assert abs(f["iters_ref"] - f["iters_gpu"]) <= 1
        assert max(r["factorised"].values()) <= 1e-12
        ch = r["chebyshev3"]
        assert abs(ch["iters_ref"] - ch["iters_gpu"]) <= 1 and ch["iters_gpu"] < f["iters_gpu"] and ch["solution"] <= 1e-6


def test_ex1_config0():
    """BASELINE configs[0]: ex1 -pa -o 3, inline-hex with 3 refinements (912,673 dofs): the reference
    needs 197 Jacobi-PCG iterations; the drop-in must agree within +-1."""
    r = run(["ex1", 3, 3, "omp"][:3])[0]
    assert r["ok"] and r["ndofs"] == 912673
    assert abs(r["iters_ref"] - 197) <= 1 and abs(r["iters_gpu"] - r["iters_ref"]) <= 1


@pytest.mark.parametrize("p,n", [(2, 4), (3, 3)])
def test_bioheat_time_stepping_through_mfem_ode_solver(p, n):
    """mfem::BackwardEulerSolver stepping b200::BioheatOperator (TimeDependentOperator::ImplicitSolve
    on the GPU) against the same operator written with the reference's PA forms + CGSolver + OperatorJacobiSmoother:
    3 steps, temperature equal to 1e-10 at a fixed iteration count, iteration counts to 1e-8 within +-1 per step;
    stored and factorised q-data"""
    recs = run(["bioheat", p, n, 3])
    assert len(recs) == 2 and {bool(r["factorised"]) for r in recs} == {False, True}
    for r in recs:
        assert r["ok"] and r["T_rel_diff_fixed_iters"] <= 1e-10
        assert abs(r["iters_ref"] - r["iters_gpu"]) <= 3


@pytest.mark.parametrize("p,n", [(2, 4), (3, 3), (1, 5)])
def test_iterative_solver_surface_and_markers(p, n):
    """b200::PCGSolver handed to code that only knows mfem::IterativeSolver (SetPreconditioner / SetOperator / SetMonitor /
    PrintLevel): same iteration counts, recorded (B r, r) history, printed lines and norms as mfem::CGSolver with
    OperatorJacobiSmoother, and as plain CG without a preconditioner; element-attribute markers through
    b200::PAOperator against BilinearForm::AddDomainIntegrator(bfi, marker) on a three-material mesh"""
    r = run(["surface", p, n])[0]
    assert r["ok"]
    assert max(r["markers"].values()) <= 1e-12
    s = r["iterative_solver"]
    assert abs(s["iters_ref"] - s["iters_gpu"]) <= 1 and s["monitor_norms"] <= 1e-9
    assert abs(r["plain_cg"]["iters_ref"] - r["plain_cg"]["iters_gpu"]) <= 1


@pytest.mark.parametrize("p,n", [(2, 4), (3, 3)])
def test_rf_coupled_operator_through_mfem_ode_solver(p, n):
    """BASELINE configs[2] behind the C++ surface: b200::RFCoupledOperator (electrostatics with sigma(T) + Joule heat +
    bioheat stage, device-resident) stepped by mfem::BackwardEulerSolver against the same composition written with the
    reference's own PA forms, solvers and q-point interpolators; stored and factorised q-data"""
    recs = run(["rf", p, n, 2])
    assert len(recs) == 2 and {bool(r["factorised"]) for r in recs} == {False, True}
    for r in recs:
        assert r["ok"] and r["T_rel_diff_fixed_iters"] <= 1e-9 and r["phi_rel_diff"] <= 1e-8
        assert abs(r["iters_T_ref"] - r["iters_T_gpu"]) <= 2 and abs(r["iters_phi_ref"] - r["iters_phi_gpu"]) <= 2


@pytest.mark.parametrize("n,pmax", [(3, 2), (2, 4)])
def test_p_multigrid_through_mfem(n, pmax):
    """b200::PMultigrid (order-refined hierarchy, Chebyshev smoothers, CG coarse solve, V-cycle on the GPU) as
    the preconditioner of b200::PCGSolver against mfem::GeometricMultigrid + CGSolver composed as examples/ex26.cpp"""
    r = run(["mg", n, pmax])[0]
    assert r["ok"] and r["vcycle"] <= 1e-8 and abs(r["iters_ref"] - r["iters_gpu"]) <= 1
    assert r["iters_gpu"] < r["iters_gpu_jacobi"]


def test_example_application_runs(tmp_path):
    """examples/rf_ablation.cpp - a stock MFEM application (mesh, spaces, BackwardEulerSolver, ParaViewDataCollection are the
    reference's) whose operator / solver classes come from the binding - linked against the unmodified reference: it heats
    the slab, its PCG solves converge, and the reference's own ParaView writer saves the fields the GPU produced"""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "rf_ablation")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/rf_ablation not built (needs the reference tree: make -C oracle ref)")
    out = subprocess.run([exe, "-n", "12", "-o", "2", "-dt", "0.5", "-tf", "3", "-vs", "2", "-pv"], capture_output=True, text=True, timeout=600,
                         cwd=str(tmp_path), env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    steps = re.findall(r"step (\d+), t = ([\d.]+) s: max T = ([\d.eE+-]+) C, PCG iterations phi / T: (\d+) / (\d+)", out.stdout)
    assert len(steps) >= 3
    temps = [float(s[2]) for s in steps]
    assert temps[0] > 37.0 and all(b > a for a, b in zip(temps, temps[1:]))            # Joule heating
    assert all(0 < int(s[3]) < 500 and 0 < int(s[4]) < 500 for s in steps)             # both solves converged within the cap
    assert "Iteration :" in out.stdout and "Average reduction factor" in out.stdout    # IterativeSolver print level 3, replayed
    assert os.path.exists(tmp_path / "rf_ablation" / "rf_ablation.pvd")
    assert len(list((tmp_path / "rf_ablation").glob("Cycle*/proc000000.vtu"))) >= 3
