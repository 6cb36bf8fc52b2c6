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

#include <cmath>
#include <iostream>
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
      b200::PCGSolver cg; cg.SetRelTol(rtol); cg.SetAbsTol(0.0); cg.SetMaxIter(maxit);
      cg.SetOperator(A2); cg.iterative_mode = true;
      X = xg; cg.Mult(B2, X); its = cg.GetNumIterations(); conv = cg.GetConverged();
   };
   Vector Xa(n), Xb(n), Xc(n), Xd(n);
   int ia, ib, ic, id; bool ca, cb, cc, cd;
   solve0(0.0, 10, Xa, ia, ca); solve2(0.0, 10, Xb, ib, cb);
   solve0(1e-8, 5000, Xc, ic, cc); solve2(1e-8, 5000, Xd, id, cd);

   const double e_apply1 = rel(y1, y0), e_diag1 = rel(d1, d0), e_apply2 = rel(y2, y0), e_diag2 = rel(d2, d0);
   const double e_con = rel(yc2, yc0), e_rhs = rel(B2, B0), e_pcg = rel(Xb, Xa), e_tol = rel(Xd, Xc);
   const bool ok = ok3 && e_apply1 <= 1e-12 && e_diag1 <= 1e-12 && e_apply2 <= 1e-12 && e_diag2 <= 1e-12 && e_con <= 1e-12 &&
                   e_rhs <= 1e-12 && e_pcg <= 1e-10 && ia == 10 && ib == 10 && abs(ic - id) <= 1 && cc == cd;
   cout << "{\"kind\":\"shim_apply\",\"p\":" << p << ",\"ne\":" << mesh.GetNE() << ",\"ndofs\":" << n << ",\"bc\":" << bc
        << ",\"integrator_level\":{\"apply\":" << e_apply1 << ",\"diag\":" << e_diag1 << "}"
        << ",\"fused\":{\"apply\":" << e_apply2 << ",\"diag\":" << e_diag2 << ",\"constrained\":" << e_con << ",\"rhs\":" << e_rhs
        << ",\"pcg10\":" << e_pcg << ",\"pcg_tol\":" << e_tol << ",\"iters_ref\":" << ic << ",\"iters_gpu\":" << id << "}"
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
   cg2.SetOperator(A2);
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

int main(int argc, char **argv)
{
   const string cmd = argc > 1 ? argv[1] : "";
   cout.precision(6);
   Device device(argc > 6 ? argv[6] : "cpu");
   if (cmd == "apply" && argc >= 6) { return apply_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), true) | apply_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), false); }
   if (cmd == "ex1") { return ex1_case(argc > 2 ? atoi(argv[2]) : 3, argc > 3 ? atoi(argv[3]) : 3); }
   cerr << "usage: shim_check apply p nx ny nz | ex1 [order refinements]\n";
   return 2;
}
