// TEST INFRASTRUCTURE ONLY — built in the build container (needs the reference headers and
// oracle/_ref/libmfem_ref.a), travels to the GPU box as oracle/_ref/shim_check and is executed only
// by tests/test_gpu_mfem_shim.py.
//
// One process, two hosts of the same path: the UNMODIFIED reference (CPU) and the reference driving
// the B200 through cardiac-ablation-ecm2_b200/host/mfem_b200pa.hpp (the drop-in binding).  Same
// Mesh / FiniteElementSpace / Coefficient objects, same input vectors; prints one JSON line with the
// relative differences and exits non-zero when a north-star tolerance is missed:
//   1e-12 per operator apply, 1e-10 on the solution after a fixed number of PCG iterations,
//   iteration counts to a tolerance within +-1.
#include "mfem.hpp"
#include "../cardiac-ablation-ecm2_b200/host/mfem_b200pa.hpp"

#include <algorithm>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <cmath>
#include <iostream>
#include <sstream>
using namespace mfem;
using namespace std;

static double kfun(const Vector &x) { return sin(8.0 * M_PI * x[0]) * cos(6.0 * M_PI * x[1]) * sin(4.0 * M_PI * x[2]) + 2.0; }
static double mfun(const Vector &x) { return 3.0 + x[0] * x[1] + 0.5 * cos(3.0 * x[2]); }

static double rel(const Vector &a, const Vector &r)
{
   Vector d(a); d -= r;
   return d.Normlinf() / r.Normlinf();
}

static int apply_case(int p, int nx, int ny, int nz, bool bc)
{
   Mesh mesh = Mesh::MakeCartesian3D(nx, ny, nz, Element::HEXAHEDRON, 1.0, 0.8, 0.6);
   for (int i = 0; i < mesh.GetNV(); ++i) { real_t *v = mesh.GetVertex(i); v[1] += 0.2 * v[0]; v[2] += 0.3 * v[0]; }
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   FunctionCoefficient kc(kfun), mc(mfun);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0;
   if (bc) { ess_bdr[0] = 1; ess_bdr[5] = 1; }
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   const int n = fes.GetNDofs();
   Vector x(n); x.Randomize(1);

   // (0) the reference, CPU
   BilinearForm a(&fes); a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a.AddDomainIntegrator(new mfem::DiffusionIntegrator(kc));
   a.AddDomainIntegrator(new mfem::MassIntegrator(mc));
   a.Assemble();
   Vector y0(n), d0(n); a.Mult(x, y0); a.AssembleDiagonal(d0);

   // (1) the reference's BilinearForm/PABilinearFormExtension with the GPU integrators plugged in
   BilinearForm a1(&fes); a1.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a1.AddDomainIntegrator(new b200::DiffusionIntegrator(kc));
   a1.AddDomainIntegrator(new b200::MassIntegrator(mc));
   a1.Assemble();
   Vector y1(n), d1(n); a1.Mult(x, y1); a1.AssembleDiagonal(d1);

   // (2) the fused operator + device PCG
   b200::PAOperator A2(fes, &kc, &mc, ess);
   Vector y2(n), d2(n); A2.MultUnconstrained(x, y2); A2.AssembleDiagonal(d2);

   // (3) the same with the factorised diffusion q-data (the sheared Cartesian mesh is affine)
   b200::PAOperator A3(fes, &kc, &mc, ess, true);
   Vector y3(n), d3(n); A3.MultUnconstrained(x, y3); A3.AssembleDiagonal(d3);
   const double e_apply3 = rel(y3, y0), e_diag3 = rel(d3, d0);
   const bool ok3 = A3.Factorised() && e_apply3 <= 1e-12 && e_diag3 <= 1e-12;

   // linear system: same FormLinearSystem on the reference side; EliminateRHS on ours
   GridFunction xg(&fes); xg = 0.0;
   FunctionCoefficient bcf([](const Vector &X) { return 30.0 * (1.0 - X(2)) + X(0); });
   if (ess.Size()) { xg.ProjectBdrCoefficient(bcf, ess_bdr); }
   LinearForm b(&fes); ConstantCoefficient one(1.0);
   b.AddDomainIntegrator(new DomainLFIntegrator(one)); b.Assemble();
   Vector b_copy(b);   // FormLinearSystem eliminates in place when P is the identity (X,B alias x,b)
   OperatorPtr A0; Vector X0, B0;
   a.FormLinearSystem(ess, xg, b, A0, X0, B0);
   Vector yc0(n), yc2(n); A0->Mult(x, yc0); A2.Mult(x, yc2);
   Vector B2(b_copy); A2.EliminateRHS(xg, B2);
   OperatorJacobiSmoother M0(a, ess);

   auto solve0 = [&](double rtol, int maxit, Vector &X, int &its, bool &conv)
   {
      CGSolver cg; cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(maxit); cg.SetPrintLevel(-1);
      cg.SetOperator(*A0); cg.SetPreconditioner(M0); cg.iterative_mode = true;
      X = X0; cg.Mult(B0, X); its = cg.GetNumIterations(); conv = cg.GetConverged();
   };
   auto solve2 = [&](double rtol, int maxit, Vector &X, int &its, bool &conv)
   {
      // the reference's own sequence of calls: preconditioner object, SetOperator, SetPreconditioner
      b200::JacobiSmoother M2(A2);
      b200::PCGSolver cg; cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(maxit); cg.SetPrintLevel(-1);
      cg.SetOperator(A2); cg.SetPreconditioner(M2); cg.iterative_mode = true;
      X = xg; cg.Mult(B2, X); its = cg.GetNumIterations(); conv = cg.GetConverged();
   };
   Vector Xa(n), Xb(n), Xc(n), Xd(n);
   int ia, ib, ic, id; bool ca, cb, cc, cd;
   solve0(0.0, 10, Xa, ia, ca); solve2(0.0, 10, Xb, ib, cb);
   solve0(1e-8, 5000, Xc, ic, cc); solve2(1e-8, 5000, Xd, id, cd);

   // Chebyshev-preconditioned CG: the reference's OperatorChebyshevSmoother (order 3, its own power-method estimate)
   // against b200::PCGSolver::SetChebyshev(3)
   int ie = 0, ig = 0; double e_cheb = 0.0, lam_gpu = 0.0;
   {
      Vector dg(n); a.AssembleDiagonal(dg);
      OperatorChebyshevSmoother C3(*A0, dg, ess, 3);
      CGSolver cg; cg.SetRelTol(1e-8); cg.SetAbsTol(0.0); cg.SetMaxIter(5000); cg.SetPrintLevel(-1);
      cg.SetOperator(*A0); cg.SetPreconditioner(C3); cg.iterative_mode = true;
      Vector Xe(n); Xe = X0; cg.Mult(B0, Xe); ie = cg.GetNumIterations();
      b200::PCGSolver cg2; cg2.SetRelTol(1e-8); cg2.SetAbsTol(0.0); cg2.SetMaxIter(5000); cg2.SetChebyshev(3);
      cg2.SetOperator(A2); cg2.iterative_mode = true;
      Vector Xg(n); Xg = xg; cg2.Mult(B2, Xg); ig = cg2.GetNumIterations(); lam_gpu = cg2.GetMaxEigEstimate();
      e_cheb = rel(Xg, Xe);
   }
   const bool ok_cheb = abs(ie - ig) <= 1 && e_cheb <= 1e-6 && ig < id;

   const double e_apply1 = rel(y1, y0), e_diag1 = rel(d1, d0), e_apply2 = rel(y2, y0), e_diag2 = rel(d2, d0);
   const double e_con = rel(yc2, yc0), e_rhs = rel(B2, B0), e_pcg = rel(Xb, Xa), e_tol = rel(Xd, Xc);
   const bool ok = ok3 && ok_cheb && e_apply1 <= 1e-12 && e_diag1 <= 1e-12 && e_apply2 <= 1e-12 && e_diag2 <= 1e-12 && e_con <= 1e-12 &&
                   e_rhs <= 1e-12 && e_pcg <= 1e-10 && ia == 10 && ib == 10 && abs(ic - id) <= 1 && cc == cd;
   cout << "{\"kind\":\"shim_apply\",\"p\":" << p << ",\"ne\":" << mesh.GetNE() << ",\"ndofs\":" << n << ",\"bc\":" << bc
        << ",\"integrator_level\":{\"apply\":" << e_apply1 << ",\"diag\":" << e_diag1 << "}"
        << ",\"fused\":{\"apply\":" << e_apply2 << ",\"diag\":" << e_diag2 << ",\"constrained\":" << e_con << ",\"rhs\":" << e_rhs
        << ",\"pcg10\":" << e_pcg << ",\"pcg_tol\":" << e_tol << ",\"iters_ref\":" << ic << ",\"iters_gpu\":" << id << "}"
        << ",\"chebyshev3\":{\"iters_ref\":" << ie << ",\"iters_gpu\":" << ig << ",\"solution\":" << e_cheb << ",\"max_eig\":" << lam_gpu << "}"
        << ",\"factorised\":{\"apply\":" << e_apply3 << ",\"diag\":" << e_diag3 << "}"
        << ",\"ok\":" << (ok ? "true" : "false") << "}" << endl;
   return ok ? 0 : 1;
}

// configs[0]: examples/ex1.cpp -pa -o 3 on the inline 4^3 hex mesh with 3 uniform refinements
static int ex1_case(int order, int ref)
{
   Mesh mesh = Mesh::MakeCartesian3D(4, 4, 4, Element::HEXAHEDRON, 1.0, 1.0, 1.0);
   for (int l = 0; l < ref; l++) { mesh.UniformRefinement(); }
   H1_FECollection fec(order, 3);
   FiniteElementSpace fes(&mesh, &fec);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   LinearForm b(&fes); ConstantCoefficient one(1.0);
   b.AddDomainIntegrator(new DomainLFIntegrator(one)); b.Assemble();
   GridFunction x(&fes); x = 0.0;
   Vector b_copy(b);
   // reference
   BilinearForm a(&fes); a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a.AddDomainIntegrator(new mfem::DiffusionIntegrator(one)); a.Assemble();
   OperatorPtr A; Vector B, X;
   a.FormLinearSystem(ess, x, b, A, X, B);
   OperatorJacobiSmoother M(a, ess);
   CGSolver cg; cg.SetRelTol(sqrt(1e-12)); cg.SetAbsTol(0.0); cg.SetMaxIter(400); cg.SetPrintLevel(-1);
   cg.SetOperator(*A); cg.SetPreconditioner(M);
   tic_toc.Clear(); tic_toc.Start(); cg.Mult(B, X); tic_toc.Stop();
   const double t_ref = tic_toc.RealTime();
   // drop-in
   b200::PAOperator A2(fes, &one, nullptr, ess);
   Vector B2(b_copy), X2(fes.GetNDofs()); X2 = 0.0;   // x itself now holds the reference solution (X aliases x)
   { Vector x0(fes.GetNDofs()); x0 = 0.0; A2.EliminateRHS(x0, B2); }
   b200::PCGSolver cg2; cg2.SetRelTol(sqrt(1e-12)); cg2.SetAbsTol(0.0); cg2.SetMaxIter(400);
   b200::JacobiSmoother M2;
   cg2.SetPreconditioner(M2); cg2.SetOperator(A2);   // SetOperator hands the operator to the preconditioner, as in the reference
   cg2.iterative_mode = false;
   tic_toc.Clear(); tic_toc.Start(); cg2.Mult(B2, X2); tic_toc.Stop();
   const double t_gpu = tic_toc.RealTime();
   const double e = rel(X2, X);
   const std::vector<double> &h = cg2.GetResidualHistory();
   cerr << "gpu (Br,r): it0 " << h[0] << " it1 " << h[1] << " last " << h[cg2.GetNumIterations()] << " rhs diff " << rel(B2, B) << endl;
   const bool ok = abs(cg.GetNumIterations() - cg2.GetNumIterations()) <= 1 && cg.GetConverged() == cg2.GetConverged() && e <= 1e-6;
   cout << "{\"kind\":\"shim_ex1\",\"order\":" << order << ",\"ref\":" << ref << ",\"ndofs\":" << fes.GetNDofs()
        << ",\"iters_ref\":" << cg.GetNumIterations() << ",\"iters_gpu\":" << cg2.GetNumIterations()
        << ",\"final_norm_ref\":" << cg.GetFinalNorm() << ",\"final_norm_gpu\":" << cg2.GetFinalNorm()
        << ",\"solution_rel_diff\":" << e << ",\"t_pcg_ref_s\":" << t_ref << ",\"t_pcg_gpu_s\":" << t_gpu
        << ",\"ok\":" << (ok ? "true" : "false") << "}" << endl;
   return ok ? 0 : 1;
}

// (f)2 of SURVEY.md 8: the Pennes bioheat equation stepped by the reference's BackwardEulerSolver, once over a
// TimeDependentOperator written with the reference's own PA forms / CGSolver / OperatorJacobiSmoother (pattern:
// examples/ex16.cpp:326-379), once over b200::BioheatOperator (the same stage equation solved on the GPU).
class RefBioheat : public TimeDependentOperator
{
   FiniteElementSpace &fes;
   b200::BioheatOperator::Physics ph;
   double rtol; int maxit;
public:
   mutable int total_iters = 0;
   RefBioheat(FiniteElementSpace &f, const b200::BioheatOperator::Physics &p, double rt, int mi)
      : TimeDependentOperator(f.GetVSize(), 0.0, IMPLICIT), fes(f), ph(p), rtol(rt), maxit(mi) {}
   void Mult(const Vector &, Vector &) const override { MFEM_ABORT("implicit only"); }
   void ImplicitSolve(const real_t dt, const Vector &T, Vector &k) override
   {
      const int n = fes.GetNDofs();
      GridFunction kg(&fes), kdt(&fes);
      for (int i = 0; i < n; i++) { kg[i] = ph.k0 * (1.0 + ph.ak * (T[i] - ph.Tref)); kdt[i] = dt * ph.k0 * (1.0 + ph.ak * (T[i] - ph.Tref)); }
      GridFunctionCoefficient kc(&kg), kdtc(&kdt);
      ConstantCoefficient wc(ph.w), cmc(ph.rc + dt * ph.w), srcc(ph.w * ph.Ta + ph.q);
      BilinearForm K(&fes), A(&fes);
      K.SetAssemblyLevel(AssemblyLevel::PARTIAL); A.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      K.AddDomainIntegrator(new mfem::DiffusionIntegrator(kc)); K.AddDomainIntegrator(new mfem::MassIntegrator(wc));
      A.AddDomainIntegrator(new mfem::DiffusionIntegrator(kdtc)); A.AddDomainIntegrator(new mfem::MassIntegrator(cmc));
      K.Assemble(); A.Assemble();
      LinearForm lf(&fes); lf.AddDomainIntegrator(new DomainLFIntegrator(srcc)); lf.Assemble();
      Vector z(n), rhs(lf); K.Mult(T, z); rhs -= z;
      Array<int> none;
      OperatorJacobiSmoother M(A, none);
      CGSolver cg; cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(maxit); cg.SetPrintLevel(-1);
      cg.SetOperator(A); cg.SetPreconditioner(M);
      k = 0.0; cg.Mult(rhs, k);
      total_iters += cg.GetNumIterations();
   }
};

static int bioheat_case(int p, int nx, int nsteps, bool factorised)
{
   Mesh mesh = Mesh::MakeCartesian3D(nx, nx, nx, Element::HEXAHEDRON, 1.0, 1.0, 0.5);
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   b200::BioheatOperator::Physics ph; ph.q = 2.0e5;
   FunctionCoefficient T0c([](const Vector &X) { const double r2 = pow(X(0) - 0.5, 2) + pow(X(1) - 0.5, 2) + pow(X(2) - 0.25, 2); return 37.0 + 20.0 * exp(-40.0 * r2); });
   const double dt = 0.5;
   auto run = [&](TimeDependentOperator &op, Vector &T)
   {
      GridFunction g(&fes); g.ProjectCoefficient(T0c); T = g;
      BackwardEulerSolver ode; ode.Init(op);
      real_t t = 0.0;
      for (int s = 0; s < nsteps; s++) { real_t h = dt; ode.Step(T, t, h); }
   };
   const int n = fes.GetNDofs();
   // fixed iteration count: the north star's 1e-10 on the solution
   Vector Ta(n), Tb(n), Tc(n), Td(n);
   RefBioheat r1(fes, ph, 0.0, 12); run(r1, Ta);
   b200::BioheatOperator g1(fes, ph, factorised); g1.SetSolverOptions(0.0, 0.0, 12); run(g1, Tb);
   // to a tolerance: iteration counts within +-1 per step
   RefBioheat r2(fes, ph, 1e-8, 500); run(r2, Tc);
   b200::BioheatOperator g2(fes, ph, factorised); g2.SetSolverOptions(1e-8, 0.0, 500); run(g2, Td);
   const double e_fixed = rel(Tb, Ta), e_tol = rel(Td, Tc);
   const bool ok = e_fixed <= 1e-10 && e_tol <= 1e-7 && abs(r2.total_iters - g2.TotalIterations()) <= nsteps && g2.LastConverged() &&
                   g1.Factorised() == factorised;
   cout << "{\"kind\":\"shim_bioheat\",\"p\":" << p << ",\"ndofs\":" << n << ",\"steps\":" << nsteps << ",\"factorised\":" << factorised
        << ",\"T_rel_diff_fixed_iters\":" << e_fixed << ",\"T_rel_diff_tol\":" << e_tol << ",\"iters_ref\":" << r2.total_iters
        << ",\"iters_gpu\":" << g2.TotalIterations() << ",\"ok\":" << (ok ? "true" : "false") << "}" << endl;
   return ok ? 0 : 1;
}

// b200::PCGSolver as an mfem::IterativeSolver: handed to code that only knows the base class, with a monitor and the
// reference's print levels; plain CG (no preconditioner); element-attribute markers through b200::PAOperator
static int solve_through_base(IterativeSolver &cg, const Operator &A, Solver *M, const Vector &B, Vector &X, IterativeSolverMonitor &mon)
{
   cg.SetRelTol(1e-10); cg.SetAbsTol(0.0); cg.SetMaxIter(400);
   cg.SetPrintLevel(IterativeSolver::PrintLevel().Iterations().Summary());
   cg.SetMonitor(mon);
   if (M) { cg.SetPreconditioner(*M); }
   cg.SetOperator(A);
   cg.iterative_mode = false;
   cg.Mult(B, X);
   return cg.GetNumIterations();
}

struct Recorder : public IterativeSolverMonitor
{
   vector<double> norms; int finals = 0; double x_norm = 0.0, r_norm = 0.0;
   void MonitorResidual(int it, real_t norm, const Vector &r, bool final) override
   {
      if (!final) { if ((int)norms.size() <= it) { norms.resize(it + 1); } norms[it] = norm; }
      else { finals++; r_norm = r.Norml2(); }
   }
   void MonitorSolution(int, real_t, const Vector &x, bool final) override { if (final) { x_norm = x.Norml2(); } }
};

static int surface_case(int p, int nx)
{
   Mesh mesh = Mesh::MakeCartesian3D(nx, nx + 1, nx, Element::HEXAHEDRON, 1.0, 0.8, 0.6);
   for (int e = 0; e < mesh.GetNE(); e++) { mesh.SetAttribute(e, 1 + e % 3); }   // tissue / blood / electrode
   mesh.SetAttributes();
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   FunctionCoefficient kc(kfun), mc(mfun);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   const int n = fes.GetNDofs();
   Vector x(n); x.Randomize(1);
   bool ok = true;
   std::ostringstream js;

   // ---- markers: diffusion on attributes {1,2}, mass on {1,3}; and diffusion everywhere + mass on {2} (the reference's
   // diagonal then drops the diffusion part outside attribute 2, fem/bilinearform_ext.cpp:374-399)
   const int combos[2][2][3] = {{{1, 1, 0}, {1, 0, 1}}, {{-1, -1, -1}, {0, 1, 0}}};
   double e_mk_apply = 0.0, e_mk_diag = 0.0;
   for (int k = 0; k < 2; k++)
   {
      Array<int> md(3), mm(3);
      for (int i = 0; i < 3; i++) { md[i] = combos[k][0][i]; mm[i] = combos[k][1][i]; }
      const bool has_d = md[0] >= 0;
      BilinearForm a(&fes); a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      if (has_d) { a.AddDomainIntegrator(new mfem::DiffusionIntegrator(kc), md); } else { a.AddDomainIntegrator(new mfem::DiffusionIntegrator(kc)); }
      a.AddDomainIntegrator(new mfem::MassIntegrator(mc), mm);
      a.Assemble();
      Vector y0(n), d0(n); a.Mult(x, y0); a.AssembleDiagonal(d0);
      Array<int> none;
      b200::PAOperator A2(fes, &kc, &mc, none, false, has_d ? &md : nullptr, &mm);
      Vector y2(n), d2(n); A2.MultUnconstrained(x, y2); A2.AssembleDiagonal(d2);
      e_mk_apply = max(e_mk_apply, rel(y2, y0)); e_mk_diag = max(e_mk_diag, rel(d2, d0));
   }
   ok = ok && e_mk_apply <= 1e-12 && e_mk_diag <= 1e-12;

   // ---- IterativeSolver surface
   BilinearForm a(&fes); a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a.AddDomainIntegrator(new mfem::DiffusionIntegrator(kc)); a.AddDomainIntegrator(new mfem::MassIntegrator(mc));
   a.Assemble();
   GridFunction xg(&fes); xg = 0.0;
   LinearForm b(&fes); ConstantCoefficient one(1.0);
   b.AddDomainIntegrator(new DomainLFIntegrator(one)); b.Assemble();
   Vector b_copy(b);
   OperatorPtr A0; Vector X0, B0;
   a.FormLinearSystem(ess, xg, b, A0, X0, B0);
   b200::PAOperator A2(fes, &kc, &mc, ess);
   Vector B2(b_copy); A2.EliminateRHS(xg, B2);
   double e_it = 0.0, e_norms = 0.0, e_plain = 0.0;
   int it_ref = 0, it_gpu = 0, it_ref_plain = 0, it_gpu_plain = 0, lines_ref = 0, lines_gpu = 0;
   for (int with_prec = 1; with_prec >= 0; with_prec--)
   {
      OperatorJacobiSmoother M0(a, ess);
      b200::JacobiSmoother M2;
      CGSolver cg0; b200::PCGSolver cg2;
      Recorder r0, r2;
      std::ostringstream o0, o2;
      Vector Xa(n), Xb(n);
      mfem::out.SetStream(o0);
      const int i0 = solve_through_base(cg0, *A0, with_prec ? (Solver *)&M0 : nullptr, B0, Xa, r0);
      mfem::out.SetStream(o2);
      const int i2 = solve_through_base(cg2, A2, with_prec ? (Solver *)&M2 : nullptr, B2, Xb, r2);
      mfem::out.SetStream(std::cout);
      double en = 0.0;
      for (size_t i = 0; i < min(r0.norms.size(), r2.norms.size()); i++) { en = max(en, fabs(r0.norms[i] - r2.norms[i]) / r0.norms[0]); }
      const string s0 = o0.str(), s2 = o2.str();
      const int l0 = (int)count(s0.begin(), s0.end(), '\n'), l2 = (int)count(s2.begin(), s2.end(), '\n');
      ok = ok && abs(i0 - i2) <= 1 && en <= 1e-9 && rel(Xb, Xa) <= 1e-7 && r2.finals == 1 && fabs(r2.x_norm - Xb.Norml2()) <= 1e-12 * Xb.Norml2() &&
           abs(l0 - l2) <= 1 && cg2.GetConverged() == cg0.GetConverged() && fabs(cg2.GetFinalNorm() - cg0.GetFinalNorm()) <= 1e-6 * cg0.GetInitialNorm() &&
           fabs(cg2.GetInitialNorm() - cg0.GetInitialNorm()) <= 1e-10 * cg0.GetInitialNorm();
      if (with_prec) { it_ref = i0; it_gpu = i2; e_it = rel(Xb, Xa); e_norms = en; lines_ref = l0; lines_gpu = l2; }
      else { it_ref_plain = i0; it_gpu_plain = i2; e_plain = rel(Xb, Xa); }
   }
   cout << "{\"kind\":\"shim_surface\",\"p\":" << p << ",\"ndofs\":" << n << ",\"markers\":{\"apply\":" << e_mk_apply << ",\"diag\":" << e_mk_diag << "}"
        << ",\"iterative_solver\":{\"iters_ref\":" << it_ref << ",\"iters_gpu\":" << it_gpu << ",\"solution\":" << e_it << ",\"monitor_norms\":" << e_norms
        << ",\"printed_lines_ref\":" << lines_ref << ",\"printed_lines_gpu\":" << lines_gpu << "}"
        << ",\"plain_cg\":{\"iters_ref\":" << it_ref_plain << ",\"iters_gpu\":" << it_gpu_plain << ",\"solution\":" << e_plain << "}"
        << ",\"ok\":" << (ok ? "true" : "false") << "}" << endl;
   return ok ? 0 : 1;
}

// configs[2]: the RF-ablation coupled problem stepped by the reference's BackwardEulerSolver - the reference composition
// (PA forms + CGSolver + OperatorJacobiSmoother + QuadratureInterpolator::PhysDerivatives + DomainLFIntegrator on a
// QuadratureFunctionCoefficient; Joule heat as miniapps/electromagnetics/joule_solver.cpp:898-906) against b200::RFCoupledOperator
class RefRF : public TimeDependentOperator
{
   FiniteElementSpace &fes;
   b200::BioheatOperator::Physics ph;
   b200::RFCoupledOperator::RF rfp;
   Array<int> ess_bdr, ess;
   GridFunction phi_bc;
   double rtol; int maxit;
public:
   mutable int total_iters = 0, total_iters_e = 0;
   GridFunction phi;
   RefRF(FiniteElementSpace &f, const b200::BioheatOperator::Physics &p, const b200::RFCoupledOperator::RF &r, const Array<int> &eb,
         const GridFunction &bc, double rt, int mi)
      : TimeDependentOperator(f.GetVSize(), 0.0, IMPLICIT), fes(f), ph(p), rfp(r), ess_bdr(eb), phi_bc(bc), rtol(rt), maxit(mi), phi(&f)
   { fes.GetEssentialTrueDofs(ess_bdr, ess); }
   void Mult(const Vector &, Vector &) const override { MFEM_ABORT("implicit only"); }
   void ImplicitSolve(const real_t dt, const Vector &T, Vector &k) override
   {
      const int n = fes.GetNDofs();
      Mesh &mesh = *fes.GetMesh();
      const FiniteElement &el = *fes.GetTypicalFE();
      const IntegrationRule &ir = mfem::DiffusionIntegrator::GetRule(el, el);
      QuadratureSpace qs(mesh, ir);
      GridFunction kg(&fes), kdt(&fes), sg(&fes);
      for (int i = 0; i < n; i++)
      {
         kg[i] = ph.k0 * (1.0 + ph.ak * (T[i] - ph.Tref)); kdt[i] = dt * ph.k0 * (1.0 + ph.ak * (T[i] - ph.Tref));
         sg[i] = rfp.s0 * (1.0 + rfp.as * (T[i] - ph.Tref));
      }
      GridFunctionCoefficient kc(&kg), kdtc(&kdt), sc(&sg);
      // (1) electrostatics
      BilinearForm ae(&fes); ae.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      ae.AddDomainIntegrator(new mfem::DiffusionIntegrator(sc)); ae.Assemble();
      phi = phi_bc;
      LinearForm be(&fes); be.Assemble();
      OperatorPtr Ae; Vector Xe, Be; ae.FormLinearSystem(ess, phi, be, Ae, Xe, Be);
      OperatorJacobiSmoother Me(ae, ess);
      CGSolver cge; cge.SetRelTol(rfp.rel_tol); cge.SetAbsTol(0.0); cge.SetMaxIter(rfp.max_iter); cge.SetPrintLevel(-1);
      cge.SetOperator(*Ae); cge.SetPreconditioner(Me); cge.iterative_mode = true;
      cge.Mult(Be, Xe);
      ae.RecoverFEMSolution(Xe, be, phi);
      total_iters_e += cge.GetNumIterations();
      // (2) Joule heat at the q-points
      const Operator *R = fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
      Vector ephi(R->Height()), eT(R->Height()); R->Mult(phi, ephi); R->Mult(T, eT);
      const QuadratureInterpolator *qi = fes.GetQuadratureInterpolator(qs);
      qi->SetOutputLayout(QVectorLayout::byVDIM);
      Vector gq(3 * qs.GetSize()), Tq(qs.GetSize()); qi->PhysDerivatives(ephi, gq); qi->Values(eT, Tq);
      QuadratureFunction rq(qs);
      for (int i = 0; i < qs.GetSize(); i++)
      {
         const double s = rfp.s0 * (1.0 + rfp.as * (Tq[i] - ph.Tref));
         rq[i] = s * (gq[3 * i] * gq[3 * i] + gq[3 * i + 1] * gq[3 * i + 1] + gq[3 * i + 2] * gq[3 * i + 2]) + ph.w * ph.Ta + ph.q;
      }
      QuadratureFunctionCoefficient rcf(rq);
      // (3) bioheat stage
      ConstantCoefficient wc(ph.w), cmc(ph.rc + dt * ph.w);
      BilinearForm K(&fes), A(&fes);
      K.SetAssemblyLevel(AssemblyLevel::PARTIAL); A.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      K.AddDomainIntegrator(new mfem::DiffusionIntegrator(kc)); K.AddDomainIntegrator(new mfem::MassIntegrator(wc));
      A.AddDomainIntegrator(new mfem::DiffusionIntegrator(kdtc)); A.AddDomainIntegrator(new mfem::MassIntegrator(cmc));
      K.Assemble(); A.Assemble();
      LinearForm lf(&fes); lf.AddDomainIntegrator(new DomainLFIntegrator(rcf, &ir)); lf.UseFastAssembly(true); lf.Assemble();
      Vector z(n), rhs(lf); K.Mult(T, z); rhs -= z;
      Array<int> none;
      OperatorJacobiSmoother M(A, none);
      CGSolver cg; cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(maxit); cg.SetPrintLevel(-1);
      cg.SetOperator(A); cg.SetPreconditioner(M);
      k = 0.0; cg.Mult(rhs, k);
      total_iters += cg.GetNumIterations();
   }
};

static int rf_case(int p, int nx, int nsteps, bool factorised)
{
   Mesh mesh = Mesh::MakeCartesian3D(nx, nx, nx, Element::HEXAHEDRON, 1.0, 1.0, 0.5);
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   b200::BioheatOperator::Physics ph;
   b200::RFCoupledOperator::RF rfp; rfp.rel_tol = 1e-10; rfp.max_iter = 2000;
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   GridFunction bc(&fes); bc = 0.0;
   FunctionCoefficient phibc([](const Vector &X) { return 30.0 * (1.0 - X(2) / 0.5); });
   bc.ProjectBdrCoefficient(phibc, ess_bdr);
   FunctionCoefficient T0c([](const Vector &X) { const double r2 = pow(X(0) - 0.5, 2) + pow(X(1) - 0.5, 2) + pow(X(2) - 0.25, 2); return 37.0 + 20.0 * exp(-40.0 * r2); });
   const double dt = 0.5;
   auto run = [&](TimeDependentOperator &op, Vector &T)
   {
      GridFunction g(&fes); g.ProjectCoefficient(T0c); T = g;
      BackwardEulerSolver ode; ode.Init(op);
      real_t t = 0.0;
      for (int s = 0; s < nsteps; s++) { real_t h = dt; ode.Step(T, t, h); }
   };
   const int n = fes.GetNDofs();
   Vector Ta(n), Tb(n), Tc(n), Td(n), phib(n);
   RefRF r1(fes, ph, rfp, ess_bdr, bc, 0.0, 12); run(r1, Ta);
   b200::RFCoupledOperator g1(fes, ph, rfp, ess, bc, factorised); g1.SetSolverOptions(0.0, 0.0, 12); run(g1, Tb);
   g1.GetPotential(phib);
   RefRF r2(fes, ph, rfp, ess_bdr, bc, 1e-8, 500); run(r2, Tc);
   b200::RFCoupledOperator g2(fes, ph, rfp, ess, bc, factorised); g2.SetSolverOptions(1e-8, 0.0, 500); run(g2, Td);
   const double e_fixed = rel(Tb, Ta), e_tol = rel(Td, Tc), e_phi = rel(phib, r1.phi);
   const bool ok = e_fixed <= 1e-9 && e_tol <= 1e-7 && e_phi <= 1e-8 && abs(r2.total_iters - g2.TotalIterations()) <= nsteps &&
                   abs(r2.total_iters_e - g2.TotalPotentialIterations()) <= nsteps && g2.LastConverged() && g1.Factorised() == factorised;
   cout << "{\"kind\":\"shim_rf\",\"p\":" << p << ",\"ndofs\":" << n << ",\"steps\":" << nsteps << ",\"factorised\":" << factorised
        << ",\"T_rel_diff_fixed_iters\":" << e_fixed << ",\"T_rel_diff_tol\":" << e_tol << ",\"phi_rel_diff\":" << e_phi
        << ",\"iters_T_ref\":" << r2.total_iters << ",\"iters_T_gpu\":" << g2.TotalIterations()
        << ",\"iters_phi_ref\":" << r2.total_iters_e << ",\"iters_phi_gpu\":" << g2.TotalPotentialIterations()
        << ",\"ok\":" << (ok ? "true" : "false") << "}" << endl;
   return ok ? 0 : 1;
}

// (f)4 of SURVEY.md 8: p-multigrid.  The reference: GeometricMultigrid over an order-refined hierarchy with the smoothers and
// coarse solver of examples/ex26.cpp (for diffusion + mass), as the preconditioner of CGSolver; against b200::PMultigrid + PCGSolver.
struct RefMG : public GeometricMultigrid
{
   RefMG(FiniteElementSpaceHierarchy &h, Array<int> &ess_bdr, Coefficient &kc, Coefficient &mc) : GeometricMultigrid(h, ess_bdr)
   {
      for (int l = 0; l < h.GetNumLevels(); ++l)
      {
         FiniteElementSpace &fes = h.GetFESpaceAtLevel(l);
         BilinearForm *form = new BilinearForm(&fes);
         form->SetAssemblyLevel(AssemblyLevel::PARTIAL);
         form->AddDomainIntegrator(new mfem::DiffusionIntegrator(kc));
         form->AddDomainIntegrator(new mfem::MassIntegrator(mc));
         form->Assemble();
         bfs.Append(form);
         OperatorPtr opr; opr.SetType(Operator::ANY_TYPE);
         bfs[l]->FormSystemMatrix(*essentialTrueDofs[l], opr);
         opr.SetOperatorOwner(false);
         if (l == 0)
         {
            CGSolver *pcg = new CGSolver();
            pcg->SetPrintLevel(-1); pcg->SetMaxIter(200); pcg->SetRelTol(sqrt(1e-4)); pcg->SetAbsTol(0.0);
            pcg->SetOperator(*opr.Ptr());
            AddLevel(opr.Ptr(), pcg, true, true);
         }
         else
         {
            Vector diag(fes.GetTrueVSize());
            bfs[l]->AssembleDiagonal(diag);
            AddLevel(opr.Ptr(), new OperatorChebyshevSmoother(*opr, diag, *essentialTrueDofs[l], 2), true, true);
         }
      }
   }
};

static int mg_case(int nx, int pmax)
{
   Mesh *mesh = new Mesh(Mesh::MakeCartesian3D(nx, nx, nx + 1, Element::HEXAHEDRON, 1.0, 0.8, 0.6));
   for (int i = 0; i < mesh->GetNV(); ++i) { real_t *v = mesh->GetVertex(i); v[1] += 0.2 * v[0]; v[2] += 0.3 * v[0]; }
   vector<FiniteElementCollection *> fecs;
   fecs.push_back(new H1_FECollection(1, 3));
   FiniteElementSpaceHierarchy h(mesh, new FiniteElementSpace(mesh, fecs[0]), true, true);
   for (int p = 2; p <= pmax; p *= 2) { fecs.push_back(new H1_FECollection(p, 3)); h.AddOrderRefinedLevel(fecs.back()); }
   FunctionCoefficient kc(kfun), mc(mfun);
   Array<int> ess_bdr(mesh->bdr_attributes.Max()); ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   FiniteElementSpace &ff = h.GetFinestFESpace();
   const int n = ff.GetNDofs();
   RefMG M0(h, ess_bdr, kc, mc);
   M0.SetCycleType(Multigrid::CycleType::VCYCLE, 1, 1);
   b200::PMultigrid M2(h, &kc, &mc, ess_bdr);
   Array<int> ess; ff.GetEssentialTrueDofs(ess_bdr, ess);
   // one cycle
   Vector x(n); x.Randomize(1);
   for (int i = 0; i < ess.Size(); i++) { x[ess[i]] = 0.0; }
   Vector y0(n), y2(n); y0 = 0.0;
   M0.Mult(x, y0); M2.Mult(x, y2);
   const double e_cycle = rel(y2, y0);
   // preconditioned solve
   GridFunction xg(&ff); xg = 0.0;
   LinearForm b(&ff); ConstantCoefficient one(1.0);
   b.AddDomainIntegrator(new DomainLFIntegrator(one)); b.Assemble();
   Vector b_copy(b);
   OperatorHandle A0; Vector X0, B0;
   M0.FormFineLinearSystem(xg, b, A0, X0, B0);
   Vector B2(b_copy); M2.FineOperator().EliminateRHS(xg, B2);
   CGSolver cg0; cg0.SetRelTol(1e-8); cg0.SetAbsTol(0.0); cg0.SetMaxIter(500); cg0.SetPrintLevel(-1);
   cg0.SetOperator(*A0); cg0.SetPreconditioner(M0);
   Vector Xa(n); Xa = 0.0; cg0.Mult(B0, Xa);
   b200::PCGSolver cg2; cg2.SetRelTol(1e-8); cg2.SetAbsTol(0.0); cg2.SetMaxIter(500); cg2.SetPrintLevel(-1);
   cg2.SetPreconditioner(M2); cg2.SetOperator(M2.FineOperator());
   Vector Xb(n); Xb = 0.0; cg2.Mult(B2, Xb);
   // Jacobi-PCG on the same system for the iteration-count comparison
   b200::JacobiSmoother J2; b200::PCGSolver cgj; cgj.SetRelTol(1e-8); cgj.SetAbsTol(0.0); cgj.SetMaxIter(5000); cgj.SetPrintLevel(-1);
   cgj.SetPreconditioner(J2); cgj.SetOperator(M2.FineOperator());
   Vector Xj(n); Xj = 0.0; cgj.Mult(B2, Xj);
   const double e_sol = rel(Xb, Xa);
   const bool ok = e_cycle <= 1e-8 && abs(cg0.GetNumIterations() - cg2.GetNumIterations()) <= 1 && cg2.GetConverged() && e_sol <= 1e-6 &&
                   cg2.GetNumIterations() < cgj.GetNumIterations();
   cout << "{\"kind\":\"shim_mg\",\"levels\":" << h.GetNumLevels() << ",\"ndofs\":" << n << ",\"vcycle\":" << e_cycle << ",\"iters_ref\":" << cg0.GetNumIterations()
        << ",\"iters_gpu\":" << cg2.GetNumIterations() << ",\"iters_gpu_jacobi\":" << cgj.GetNumIterations() << ",\"solution\":" << e_sol
        << ",\"ok\":" << (ok ? "true" : "false") << "}" << endl;
   return ok ? 0 : 1;
}

static void on_segv(int sig)
{
   void *bt[64];
   const int n = backtrace(bt, 64);
   const char msg[] = "shim_check: fatal signal, backtrace:\n";
   if (write(2, msg, sizeof(msg) - 1) < 0) {}
   backtrace_symbols_fd(bt, n, 2);
   _exit(128 + sig);
}

int main(int argc, char **argv)
{
   signal(SIGSEGV, on_segv);
   const string cmd = argc > 1 ? argv[1] : "";
   cout.precision(6);
   Device device((cmd == "apply" && argc > 6) ? argv[6] : "cpu");
   if (cmd == "apply" && argc >= 6) { return apply_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), true) | apply_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), false); }
   if (cmd == "ex1") { return ex1_case(argc > 2 ? atoi(argv[2]) : 3, argc > 3 ? atoi(argv[3]) : 3); }
   if (cmd == "bioheat" && argc >= 5) { return bioheat_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), false) | bioheat_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), true); }
   if (cmd == "mg" && argc >= 4) { return mg_case(atoi(argv[2]), atoi(argv[3])); }
   if (cmd == "surface" && argc >= 4) { return surface_case(atoi(argv[2]), atoi(argv[3])); }
   if (cmd == "rf" && argc >= 5) { return rf_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), false) | rf_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), true); }
   cerr << "usage: shim_check apply p nx ny nz | ex1 [order refinements] | bioheat p nx steps | surface p nx | rf p nx steps\n";
   return 2;
}
