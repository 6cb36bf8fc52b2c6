// TEST INFRASTRUCTURE ONLY — never linked into, or called from, the product path.
//
// Driver that links the UNMODIFIED reference (stock MFEM 4.9.1-dev sources under
// /root/reference, compiled by oracle/Makefile into oracle/_ref/libmfem_ref.a) and
//   * dumps golden vectors for every row of SURVEY.md §8(a): the arrays the hot path
//     consumes (B, G, W, gather_map/offsets/indices, J, detJ, q-data) and what the
//     reference computes from them (pa_data, E- and L-vector applies, diagonals,
//     Jacobi, PCG iterates, q-point values/gradients, RHS, the bioheat/RF step);
//   * times the reference CPU path (serial "cpu" or OpenMP "omp" device) for
//     bench.py's cpu_baseline / --impl reference arm.
// It reads nothing from /root/reference at run time (meshes are generated with
// Mesh::MakeCartesian3D; config 1's data/inline-hex.mesh is the INLINE 4x4x4 hex
// mesh, i.e. MakeCartesian3D(4,4,4), see `--check-inline`).
//
// Output format of `dump_*`: one raw little-endian file per array in <outdir>
// (<name>.f64 / <name>.i32) plus <outdir>/manifest.txt with "name dtype count".
#include "mfem.hpp"
#include "fem/integ/bilininteg_diffusion_kernels.hpp"
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <sys/stat.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace mfem;
using namespace std;

static double now()
{
   return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- dump helpers
struct Dumper
{
   string dir;
   ofstream man;
   explicit Dumper(const string &d) : dir(d)
   {
      mkdir(dir.c_str(), 0755);
      man.open(dir + "/manifest.txt");
   }
   void raw(const string &name, const char *ext, const void *p, size_t bytes, size_t n)
   {
      ofstream f(dir + "/" + name + "." + ext, ios::binary);
      f.write((const char *)p, bytes);
      man << name << " " << ext << " " << n << "\n";
   }
   void f64(const string &name, const double *p, size_t n) { raw(name, "f64", p, 8 * n, n); }
   void i32(const string &name, const int *p, size_t n) { raw(name, "i32", p, 4 * n, n); }
   void vec(const string &name, const Vector &v) { f64(name, v.HostRead(), v.Size()); }
   void arr(const string &name, const Array<int> &a) { i32(name, a.HostRead(), a.Size()); }
   void arr(const string &name, const Array<double> &a) { f64(name, a.HostRead(), a.Size()); }
   void scalar(const string &name, double v) { f64(name, &v, 1); }
   void iscalar(const string &name, int v) { i32(name, &v, 1); }
};

struct NormRecorder : public IterativeSolverMonitor
{
   vector<double> norms;
   void MonitorResidual(int it, real_t norm, const Vector &, bool final) override
   {
      if (!final) { if ((int)norms.size() <= it) { norms.resize(it + 1); } norms[it] = norm; }
   }
};

// Mesh kinds: "cart" = MakeCartesian3D(nx,ny,nz,HEX,sx,sy,sz) (SFC element order);
//             "skew" = the same + the vertex remap of tests/unit/fem/test_pa_coeff.cpp:33-39;
//             "inline3" = config 1: MakeCartesian3D(4,4,4) + 3 uniform refinements.
static Mesh make_mesh(const string &kind, int nx, int ny, int nz, double sx, double sy, double sz)
{
   if (kind == "inline3")
   {
      Mesh m = Mesh::MakeCartesian3D(4, 4, 4, Element::HEXAHEDRON, 1.0, 1.0, 1.0);
      for (int l = 0; l < nx; l++) { m.UniformRefinement(); }
      return m;
   }
   Mesh m = Mesh::MakeCartesian3D(nx, ny, nz, Element::HEXAHEDRON, sx, sy, sz);
   if (kind == "skew")
   {
      for (int i = 0; i < m.GetNV(); ++i)
      {
         real_t *v = m.GetVertex(i);
         v[1] += 0.2 * v[0];
         v[2] += 0.3 * v[0];
      }
   }
   return m;
}

// Coefficient functions used for the "func" coefficient kind (q-data sampled by the
// reference's own Coefficient::Project).  Same shape as test_pa_coeff.cpp:45-58.
static double kfun(const Vector &x)
{
   return sin(8.0 * M_PI * x[0]) * cos(6.0 * M_PI * x[1]) * sin(4.0 * M_PI * x[2]) + 2.0;
}
static double mfun(const Vector &x)
{
   return 3.0 + x[0] * x[1] + 0.5 * cos(3.0 * x[2]);
}

struct MassPA : public MassIntegrator
{
   using MassIntegrator::MassIntegrator;
   const Vector &PAData() const { return pa_data; }
};

// ------------------------------------------------------------------ dump_case
// Everything §8(a) a1–a16, a20, a21 for one (mesh, order, coefficient kind, BC) case.
static int dump_case(int argc, char **argv)
{
   if (argc < 12)
   {
      cerr << "dump_case OUT p kind nx ny nz sx sy sz coef(const|func) bc(none|all|zfaces) [pcg_iters]\n";
      return 2;
   }
   Dumper D(argv[2]);
   const int p = atoi(argv[3]);
   const string kind = argv[4];
   const int nx = atoi(argv[5]), ny = atoi(argv[6]), nz = atoi(argv[7]);
   const double sx = atof(argv[8]), sy = atof(argv[9]), sz = atof(argv[10]);
   const string coef = argv[11];
   const string bc = argc > 12 ? argv[12] : "none";
   const int pcg_iters = argc > 13 ? atoi(argv[13]) : 10;

   Device device("cpu");
   Mesh mesh = make_mesh(kind, nx, ny, nz, sx, sy, sz);
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   const FiniteElement &el = *fes.GetTypicalFE();
   const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
   {
      // the mass rule must be the same rule; verify, do not assume
      ElementTransformation &T0 = *mesh.GetTypicalElementTransformation();
      const IntegrationRule &irm = MassIntegrator::GetRule(el, el, T0);
      MFEM_VERIFY(irm.GetNPoints() == ir.GetNPoints(), "mass/diffusion rules differ");
   }
   const DofToQuad &maps = el.GetDofToQuad(ir, DofToQuad::TENSOR);
   const int D1D = maps.ndof, Q1D = maps.nqpt, NE = mesh.GetNE(), ND = fes.GetNDofs();
   const int NQ = ir.GetNPoints();
   D.iscalar("p", p); D.iscalar("D1D", D1D); D.iscalar("Q1D", Q1D); D.iscalar("NE", NE);
   D.iscalar("ndofs", ND);
   D.arr("B", maps.B); D.arr("G", maps.G); D.arr("Bt", maps.Bt); D.arr("Gt", maps.Gt);
   D.arr("W", ir.GetWeights());

   // a1: restriction tables
   const ElementRestriction *R = dynamic_cast<const ElementRestriction *>(
                                    fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   MFEM_VERIFY(R, "no ElementRestriction");
   D.arr("gather_map", R->GatherMap()); D.arr("offsets", R->Offsets()); D.arr("indices", R->Indices());

   // a20: geometric factors; also vertices (inputs of the product-side builder)
   const GeometricFactors *geom = mesh.GetGeometricFactors(
                                     ir, GeometricFactors::JACOBIANS | GeometricFactors::DETERMINANTS |
                                     GeometricFactors::COORDINATES);
   D.vec("J", geom->J); D.vec("detJ", geom->detJ); D.vec("Xq", geom->X);
   {
      Vector vx(3 * mesh.GetNV());
      for (int i = 0; i < mesh.GetNV(); i++) { for (int d = 0; d < 3; d++) { vx(3 * i + d) = mesh.GetVertex(i)[d]; } }
      D.vec("vertices", vx);
      Array<int> ev(8 * NE);
      for (int e = 0; e < NE; e++) { const int *v = mesh.GetElement(e)->GetVertices(); for (int j = 0; j < 8; j++) { ev[8 * e + j] = v[j]; } }
      D.arr("elem_vertices", ev);
   }

   // coefficients → q-data through the reference's own projection (a19)
   QuadratureSpace qs(mesh, ir);
   ConstantCoefficient kc_const(0.5), mc_const(3.6);
   FunctionCoefficient kc_fun(kfun), mc_fun(mfun);
   Coefficient &kc = (coef == "const") ? (Coefficient &)kc_const : (Coefficient &)kc_fun;
   Coefficient &mc = (coef == "const") ? (Coefficient &)mc_const : (Coefficient &)mc_fun;
   CoefficientVector kq(kc, qs, CoefficientStorage::COMPRESSED);
   CoefficientVector mq(mc, qs, CoefficientStorage::COMPRESSED);
   D.vec("kq", kq); D.vec("mq", mq);   // size 1 (constant) or NQ*NE

   // a5: diffusion pa_data via the reference setup kernel itself
   Vector pa_diff(6 * NQ * NE);
   internal::PADiffusionSetup(3, 3, D1D, Q1D, 1, NE, ir.GetWeights(), geom->J, kq, pa_diff);
   D.vec("pa_diff", pa_diff);

   // integrators (a6–a10) at E-vector level
   DiffusionIntegrator *di = new DiffusionIntegrator(kc);
   MassPA *mi = new MassPA(mc);
   BilinearForm a(&fes);
   a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a.AddDomainIntegrator(di);
   a.AddDomainIntegrator(mi);
   a.Assemble();
   D.vec("pa_mass", mi->PAData());

   Vector x(ND); x.Randomize(1);
   D.vec("x", x);
   Vector xE(R->Height()); R->Mult(x, xE); D.vec("xE", xE);             // a2
   Vector yE(R->Height());
   yE = 0.0; di->AddMultPA(xE, yE); D.vec("yE_diff", yE);                 // a6
   { Vector yL(ND); R->MultTranspose(yE, yL); D.vec("y_diff", yL); }      // a3
   yE = 0.0; mi->AddMultPA(xE, yE); D.vec("yE_mass", yE);                 // a9
   { Vector yL(ND); R->MultTranspose(yE, yL); D.vec("y_mass", yL); }
   yE = 0.0; di->AddMultPA(xE, yE); mi->AddMultPA(xE, yE); D.vec("yE", yE);
   Vector y(ND); a.Mult(x, y); D.vec("y", y);                             // a11 L→L
   yE = 0.0; di->AssembleDiagonalPA(yE); D.vec("dE_diff", yE);            // a7
   yE = 0.0; mi->AssembleDiagonalPA(yE); D.vec("dE_mass", yE);            // a10
   Vector diag(ND); a.AssembleDiagonal(diag); D.vec("diag", diag);        // a4+a7+a10

   // a12: essential dofs + constrained operator
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0;
   if (bc == "all") { ess_bdr = 1; }
   else if (bc == "zfaces") { ess_bdr[0] = 1; ess_bdr[5] = 1; }   // z=0 is attr 1, z=sz is attr 6
   Array<int> ess;
   fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.arr("ess", ess);
   {
      GridFunction xg(&fes); xg = 0.0;
      FunctionCoefficient bcf([&](const Vector &X) { return 30.0 * (1.0 - X(2) / sz) + X(0); });
      if (ess.Size()) { xg.ProjectBdrCoefficient(bcf, ess_bdr); }
      LinearForm b(&fes);
      ConstantCoefficient one(1.0);
      b.AddDomainIntegrator(new DomainLFIntegrator(one));
      b.Assemble();
      D.vec("b_L", b); D.vec("x0_L", xg);
      OperatorPtr A; Vector X, B;
      a.FormLinearSystem(ess, xg, b, A, X, B);
      D.vec("B_rhs", B); D.vec("X0", X);
      Vector yc(ND); A->Mult(x, yc); D.vec("y_constrained", yc);
      OperatorJacobiSmoother M(a, ess);                                   // a13
      { Vector r(ND), z(ND); r = x; M.Mult(r, z); D.vec("jacobi_z", z); }
      // a14: PCG — fixed iteration counts, then to tolerance
      Array<int> its;
      its.Append(1); its.Append(2); its.Append(pcg_iters);
      for (int k = 0; k < its.Size(); k++)
      {
         CGSolver cg; NormRecorder rec;
         cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(its[k]); cg.SetPrintLevel(-1);
         cg.SetOperator(*A); cg.SetPreconditioner(M); cg.SetMonitor(rec);
         cg.iterative_mode = true;
         Vector Xk(X);
         cg.Mult(B, Xk);
         D.vec("X_pcg" + to_string(its[k]), Xk);
         if (k == its.Size() - 1) { D.f64("pcg_norms", rec.norms.data(), rec.norms.size()); }
      }
      {
         CGSolver cg; NormRecorder rec;
         cg.SetRelTol(1e-8); cg.SetAbsTol(0.0); cg.SetMaxIter(5000); cg.SetPrintLevel(-1);
         cg.SetOperator(*A); cg.SetPreconditioner(M); cg.SetMonitor(rec);
         cg.iterative_mode = true;
         Vector Xk(X);
         cg.Mult(B, Xk);
         D.vec("X_pcg_tol", Xk);
         D.iscalar("pcg_tol_iters", cg.GetNumIterations());
         D.iscalar("pcg_tol_converged", cg.GetConverged());
         D.scalar("pcg_tol_final_norm", cg.GetFinalNorm());
         D.f64("pcg_tol_norms", rec.norms.data(), rec.norms.size());
      }
   }
   // (f)4 of SURVEY 8: OperatorChebyshevSmoother (linalg/solvers.cpp:455-657) on the constrained operator, with the
   // largest eigenvalue of D^-1 A from the reference's power method exactly as its second constructor runs it
   // (solvers.cpp:497-511: ProductOperator(OperatorJacobiSmoother(diag, ess, 1.0), A), 10 steps, 1e-8, seed 12345)
   {
      GridFunction xg(&fes); xg = 0.0;
      FunctionCoefficient bcf([&](const Vector &X) { return 30.0 * (1.0 - X(2) / sz) + X(0); });
      if (ess.Size()) { xg.ProjectBdrCoefficient(bcf, ess_bdr); }
      LinearForm b(&fes);
      ConstantCoefficient one(1.0);
      b.AddDomainIntegrator(new DomainLFIntegrator(one));
      b.Assemble();
      OperatorPtr A; Vector X, B;
      a.FormLinearSystem(ess, xg, b, A, X, B);
      OperatorJacobiSmoother invD(diag, ess, 1.0);
      ProductOperator dp(&invD, A.Ptr(), false, false);
      PowerMethod pm;
      Vector ev(ND);
      const double lam = pm.EstimateLargestEigenvalue(dp, ev, 10, 1e-8, 12345);
      D.scalar("cheb_max_eig", lam);
      { Vector v0(ND); v0.Randomize(12345); D.vec("cheb_v0", v0); }
      for (int order = 1; order <= 5; order++)
      {
         OperatorChebyshevSmoother C(*A, diag, ess, order, lam);
         Vector z(ND); C.Mult(x, z);
         D.vec("cheb_z" + to_string(order), z);
      }
      OperatorChebyshevSmoother C3(*A, diag, ess, 3, lam);
      {
         CGSolver cg; NormRecorder rec;
         cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(4); cg.SetPrintLevel(-1);
         cg.SetOperator(*A); cg.SetPreconditioner(C3); cg.SetMonitor(rec);
         cg.iterative_mode = true;
         Vector Xk(X);
         cg.Mult(B, Xk);
         D.vec("X_cheb3_pcg4", Xk);
         D.f64("cheb3_pcg_norms", rec.norms.data(), rec.norms.size());
      }
      {
         CGSolver cg;
         cg.SetRelTol(1e-8); cg.SetAbsTol(0.0); cg.SetMaxIter(5000); cg.SetPrintLevel(-1);
         cg.SetOperator(*A); cg.SetPreconditioner(C3);
         cg.iterative_mode = true;
         Vector Xk(X);
         cg.Mult(B, Xk);
         D.vec("X_cheb3_tol", Xk);
         D.iscalar("cheb3_tol_iters", cg.GetNumIterations());
         D.iscalar("cheb3_tol_converged", cg.GetConverged());
      }
   }
   // a17/a18: q-point values and physical gradients of x; K15 RHS with q-data source
   {
      const QuadratureInterpolator *qi = fes.GetQuadratureInterpolator(qs);
      qi->SetOutputLayout(QVectorLayout::byVDIM);
      Vector vq(NQ * NE), gq(3 * NQ * NE);
      qi->Values(xE, vq); D.vec("xq_values", vq);
      qi->PhysDerivatives(xE, gq); D.vec("xq_physgrad", gq);
      QuadratureFunction fq(qs);
      for (int i = 0; i < fq.Size(); i++) { fq(i) = 1.0 + vq(i) * vq(i); }
      QuadratureFunctionCoefficient fqc(fq);
      LinearForm lf(&fes);
      lf.AddDomainIntegrator(new DomainLFIntegrator(fqc, &ir));
      lf.UseFastAssembly(true);
      lf.Assemble();
      D.vec("lf_fq", fq); D.vec("lf_b", lf);
   }
   cout << "dump_case ok: p=" << p << " NE=" << NE << " ndofs=" << ND << " Q1D=" << Q1D
        << setprecision(17) << " |y|=" << y.Norml2() << " |diag|=" << diag.Norml2() << endl;
   return 0;
}

// --------------------------------------------------------------- bioheat step
// One RF-ablation coupled step expressed with the reference's own PA API
// (parameters §8d).  Used both for dumps and for timing.
struct BioheatParams
{
   double dt = 0.5, rc = 3.6e6, wbcb = 4.0e4, Ta = 37.0, k0 = 0.5, ak = 0.02, s0 = 0.3, as = 0.015, V = 30.0;
};

static int bioheat(int argc, char **argv, bool dump)
{
   // dump_bioheat OUT p N iters [dev]    |   time_bioheat p N iters dev
   int ai = 2;
   Dumper *D = nullptr;
   if (dump) { D = new Dumper(argv[ai++]); }
   if (argc < ai + 3) { cerr << "usage: [dump_bioheat OUT|time_bioheat] p N iters [dev]\n"; return 2; }
   const int p = atoi(argv[ai++]), N = atoi(argv[ai++]), iters = atoi(argv[ai++]);
   const char *dev = argc > ai ? argv[ai] : "cpu";
   Device device(dev);
   const BioheatParams P;
   double t0 = now();
   Mesh mesh = Mesh::MakeCartesian3D(N, N, N, Element::HEXAHEDRON, 1.0, 1.0, 1.0);
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   const FiniteElement &el = *fes.GetTypicalFE();
   const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
   QuadratureSpace qs(mesh, ir);
   const double t_mesh = now() - t0;
   GridFunction T0(&fes), T1(&fes), phi(&fes);
   FunctionCoefficient Tinit([](const Vector &x)
   {
      const double r2 = (x(0) - .5) * (x(0) - .5) + (x(1) - .5) * (x(1) - .5) + (x(2) - .5) * (x(2) - .5);
      return 37.0 + 20.0 * exp(-40.0 * r2);
   });
   T0.ProjectCoefficient(Tinit);
   // T at q-points → k(T), sigma(T), mass coefficient
   t0 = now();
   QuadratureFunction Tq(qs), kq(qs), sq(qs), mq(qs);
   Tq.ProjectGridFunction(T0);
   {
      const int n = Tq.Size();
      auto t = Tq.Read(); auto k = kq.Write(); auto s = sq.Write(); auto m = mq.Write();
      const double k0 = P.k0, ak = P.ak, s0 = P.s0, as = P.as, mval = P.rc / P.dt + P.wbcb;
      mfem::forall(n, [=] MFEM_HOST_DEVICE (int i)
      {
         k[i] = k0 * (1.0 + ak * (t[i] - 37.0));
         s[i] = s0 * (1.0 + as * (t[i] - 37.0));
         m[i] = mval;
      });
   }
   const double t_coef = now() - t0;
   QuadratureFunctionCoefficient kc(kq), sc(sq), mc(mq);
   // (1) electrostatics
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   const double V = P.V;
   FunctionCoefficient phibc([=](const Vector &x) { return V * (1.0 - x(2)); });
   phi = 0.0; phi.ProjectBdrCoefficient(phibc, ess_bdr);
   BilinearForm ae(&fes); ae.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   ae.AddDomainIntegrator(new DiffusionIntegrator(sc));
   t0 = now(); ae.Assemble(); const double t_asm_e = now() - t0;
   LinearForm be(&fes); be.Assemble();
   OperatorPtr Ae; Vector Xe, Be; ae.FormLinearSystem(ess, phi, be, Ae, Xe, Be);
   OperatorJacobiSmoother Me(ae, ess);
   CGSolver cge; cge.SetRelTol(0.0); cge.SetAbsTol(0.0); cge.SetMaxIter(iters); cge.SetPrintLevel(-1);
   cge.SetOperator(*Ae); cge.SetPreconditioner(Me);
   if (D) { D->vec("T0", T0); D->vec("Tq", Tq); D->vec("kq", kq); D->vec("sq", sq); D->vec("mq", mq);
            D->arr("ess", ess); D->vec("phi0", phi); D->vec("Be", Be); }
   t0 = now(); cge.Mult(Be, Xe); const double t_cg_e = now() - t0;
   ae.RecoverFEMSolution(Xe, be, phi);
   // (2) Joule source at q-points
   t0 = now();
   const Operator *R = fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
   Vector ephi(R->Height()); R->Mult(phi, ephi);
   const QuadratureInterpolator *qi = fes.GetQuadratureInterpolator(qs);
   qi->SetOutputLayout(QVectorLayout::byVDIM);
   Vector gq(3 * qs.GetSize()); qi->PhysDerivatives(ephi, gq);
   QuadratureFunction rq(qs);
   {
      const int n = qs.GetSize();
      auto g = gq.Read(); auto s = sq.Read(); auto r = rq.Write();
      const double src0 = P.wbcb * P.Ta;
      mfem::forall(n, [=] MFEM_HOST_DEVICE (int i)
      {
         const double gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
         r[i] = s[i] * (gx * gx + gy * gy + gz * gz) + src0;
      });
   }
   const double t_joule = now() - t0;
   QuadratureFunctionCoefficient rcf(rq);
   // (3) bioheat backward-Euler step
   BilinearForm at(&fes); at.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   at.AddDomainIntegrator(new DiffusionIntegrator(kc));
   at.AddDomainIntegrator(new MassIntegrator(mc));
   t0 = now(); at.Assemble(); const double t_asm_t = now() - t0;
   ConstantCoefficient rcdt(P.rc / P.dt);
   BilinearForm mrc(&fes); mrc.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   mrc.AddDomainIntegrator(new MassIntegrator(rcdt)); mrc.Assemble();
   LinearForm bt(&fes); bt.AddDomainIntegrator(new DomainLFIntegrator(rcf, &ir)); bt.UseFastAssembly(true);
   t0 = now(); bt.Assemble(); Vector mT(fes.GetVSize()); mrc.Mult(T0, mT); bt += mT;
   const double t_rhs = now() - t0;
   Array<int> noess; T1 = T0; OperatorPtr At; Vector Xt, Bt; at.FormLinearSystem(noess, T1, bt, At, Xt, Bt, 1);
   OperatorJacobiSmoother Mt(at, noess);
   CGSolver cgt; cgt.SetRelTol(0.0); cgt.SetAbsTol(0.0); cgt.SetMaxIter(iters); cgt.SetPrintLevel(-1);
   cgt.SetOperator(*At); cgt.SetPreconditioner(Mt); cgt.iterative_mode = true;
   t0 = now(); cgt.Mult(Bt, Xt); const double t_cg_t = now() - t0;
   at.RecoverFEMSolution(Xt, bt, T1);
   if (D)
   {
      D->vec("phi", phi); D->vec("gradphi_q", gq); D->vec("src_q", rq); D->vec("rhs_T", bt); D->vec("T1", T1);
      D->iscalar("iters", iters);
      // iterations to rel 1e-8 for both solves (iteration-count parity, ±1)
      CGSolver c2; c2.SetRelTol(1e-8); c2.SetAbsTol(0.0); c2.SetMaxIter(5000); c2.SetPrintLevel(-1);
      c2.SetOperator(*At); c2.SetPreconditioner(Mt); c2.iterative_mode = true;
      Vector X2(T0); c2.Mult(Bt, X2);
      D->iscalar("iters_tol_T", c2.GetNumIterations()); D->vec("T1_tol", X2);
      CGSolver c3; c3.SetRelTol(1e-8); c3.SetAbsTol(0.0); c3.SetMaxIter(5000); c3.SetPrintLevel(-1);
      c3.SetOperator(*Ae); c3.SetPreconditioner(Me); c3.iterative_mode = true;
      GridFunction p3(&fes); p3 = 0.0; p3.ProjectBdrCoefficient(phibc, ess_bdr);
      Vector X3(p3); c3.Mult(Be, X3);
      D->iscalar("iters_tol_phi", c3.GetNumIterations()); D->vec("phi_tol", X3);
   }
   int nthreads = 1;
#ifdef _OPENMP
   if (string(dev) == "omp") { nthreads = omp_get_max_threads(); }
#endif
   cout << setprecision(17)
        << "{\"kind\":\"bioheat_step\",\"p\":" << p << ",\"N\":" << N << ",\"ndofs\":" << fes.GetNDofs()
        << ",\"iters\":" << iters << ",\"device\":\"" << dev << "\",\"threads\":" << nthreads
        << ",\"phi_norm\":" << phi.Norml2() << ",\"src_sum\":" << rq.Sum() << ",\"T1_norm\":" << T1.Norml2()
        << ",\"T1_max\":" << T1.Max()
        << ",\"t_mesh\":" << t_mesh << ",\"t_coef\":" << t_coef << ",\"t_asm_e\":" << t_asm_e
        << ",\"t_cg_e\":" << t_cg_e << ",\"t_joule\":" << t_joule << ",\"t_asm_t\":" << t_asm_t
        << ",\"t_rhs\":" << t_rhs << ",\"t_cg_t\":" << t_cg_t << "}" << endl;
   delete D;
   return 0;
}


// ------------------------------------------------------- several coupled steps
// dump_bioheat_steps OUT p N iters nsteps: the coupled RF + bioheat step repeated, T^{n+1} feeding k(T), sigma(T)
// of the next step (the time loop a BackwardEulerSolver::Step / ImplicitSolve driver runs, linalg/ode.cpp:682-696).
static int bioheat_steps(int argc, char **argv)
{
   if (argc < 7) { cerr << "usage: dump_bioheat_steps OUT p N iters nsteps\n"; return 2; }
   Dumper D(argv[2]);
   const int p = atoi(argv[3]), N = atoi(argv[4]), iters = atoi(argv[5]), nsteps = atoi(argv[6]);
   Device device("cpu");
   const BioheatParams P;
   Mesh mesh = Mesh::MakeCartesian3D(N, N, N, Element::HEXAHEDRON, 1.0, 1.0, 1.0);
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   const FiniteElement &el = *fes.GetTypicalFE();
   const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
   QuadratureSpace qs(mesh, ir);
   GridFunction T(&fes), phi(&fes);
   FunctionCoefficient Tinit([](const Vector &x)
   {
      const double r2 = (x(0) - .5) * (x(0) - .5) + (x(1) - .5) * (x(1) - .5) + (x(2) - .5) * (x(2) - .5);
      return 37.0 + 20.0 * exp(-40.0 * r2);
   });
   T.ProjectCoefficient(Tinit);
   D.vec("T_step0", T);
   D.iscalar("iters", iters); D.iscalar("nsteps", nsteps);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   const double V = P.V;
   FunctionCoefficient phibc([=](const Vector &x) { return V * (1.0 - x(2)); });
   const Operator *R = fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC);
   const QuadratureInterpolator *qi = fes.GetQuadratureInterpolator(qs);
   qi->SetOutputLayout(QVectorLayout::byVDIM);
   for (int step = 1; step <= nsteps; step++)
   {
      QuadratureFunction Tq(qs), kq(qs), sq(qs), mq(qs), rq(qs);
      Tq.ProjectGridFunction(T);
      for (int i = 0; i < Tq.Size(); i++)
      {
         kq(i) = P.k0 * (1.0 + P.ak * (Tq(i) - 37.0));
         sq(i) = P.s0 * (1.0 + P.as * (Tq(i) - 37.0));
         mq(i) = P.rc / P.dt + P.wbcb;
      }
      QuadratureFunctionCoefficient kc(kq), sc(sq), mc(mq);
      phi = 0.0; phi.ProjectBdrCoefficient(phibc, ess_bdr);
      BilinearForm ae(&fes); ae.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      ae.AddDomainIntegrator(new DiffusionIntegrator(sc)); ae.Assemble();
      LinearForm be(&fes); be.Assemble();
      OperatorPtr Ae; Vector Xe, Be; ae.FormLinearSystem(ess, phi, be, Ae, Xe, Be);
      OperatorJacobiSmoother Me(ae, ess);
      CGSolver cge; cge.SetRelTol(0.0); cge.SetAbsTol(0.0); cge.SetMaxIter(iters); cge.SetPrintLevel(-1);
      cge.SetOperator(*Ae); cge.SetPreconditioner(Me);
      cge.Mult(Be, Xe);
      ae.RecoverFEMSolution(Xe, be, phi);
      Vector ephi(R->Height()); R->Mult(phi, ephi);
      Vector gq(3 * qs.GetSize()); qi->PhysDerivatives(ephi, gq);
      for (int i = 0; i < qs.GetSize(); i++)
      {
         const double gx = gq(3 * i), gy = gq(3 * i + 1), gz = gq(3 * i + 2);
         rq(i) = sq(i) * (gx * gx + gy * gy + gz * gz) + P.wbcb * P.Ta;
      }
      QuadratureFunctionCoefficient rcf(rq);
      BilinearForm at(&fes); at.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      at.AddDomainIntegrator(new DiffusionIntegrator(kc));
      at.AddDomainIntegrator(new MassIntegrator(mc)); at.Assemble();
      ConstantCoefficient rcdt(P.rc / P.dt);
      BilinearForm mrc(&fes); mrc.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      mrc.AddDomainIntegrator(new MassIntegrator(rcdt)); mrc.Assemble();
      LinearForm bt(&fes); bt.AddDomainIntegrator(new DomainLFIntegrator(rcf, &ir)); bt.UseFastAssembly(true);
      bt.Assemble(); Vector mT(fes.GetVSize()); mrc.Mult(T, mT); bt += mT;
      GridFunction T1(&fes); T1 = T;
      Array<int> noess; OperatorPtr At; Vector Xt, Bt; at.FormLinearSystem(noess, T1, bt, At, Xt, Bt, 1);
      OperatorJacobiSmoother Mt(at, noess);
      CGSolver cgt; cgt.SetRelTol(0.0); cgt.SetAbsTol(0.0); cgt.SetMaxIter(iters); cgt.SetPrintLevel(-1);
      cgt.SetOperator(*At); cgt.SetPreconditioner(Mt); cgt.iterative_mode = true;
      cgt.Mult(Bt, Xt);
      at.RecoverFEMSolution(Xt, bt, T1);
      T = T1;
      D.vec("T_step" + to_string(step), T);
      D.vec("phi_step" + to_string(step), phi);
   }
   cout << setprecision(17) << "dump_bioheat_steps ok: |T|=" << T.Norml2() << " max T=" << T.Max() << endl;
   return 0;
}

// ---------------------------------------------------------------- time_apply
// CPU reference timing of the PA diffusion+mass apply (L→L, A.Mult) and of a fixed
// number of Jacobi-PCG iterations, on MakeCartesian3D(N^3), order p, q-data
// coefficients (bioheat-like k(T), mass) — the same operator bench.py times on the GPU.
static int time_apply(int argc, char **argv)
{
   if (argc < 7) { cerr << "time_apply p N reps warmup dev [pcg_iters]\n"; return 2; }
   const int p = atoi(argv[2]), N = atoi(argv[3]), reps = atoi(argv[4]), warm = atoi(argv[5]);
   const char *dev = argv[6];
   const int pcg_iters = argc > 7 ? atoi(argv[7]) : 0;
   Device device(dev);
   double t0 = now();
   Mesh mesh = Mesh::MakeCartesian3D(N, N, N, Element::HEXAHEDRON, 1.0, 1.0, 1.0);
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   const FiniteElement &el = *fes.GetTypicalFE();
   const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
   QuadratureSpace qs(mesh, ir);
   QuadratureFunction kq(qs), mq(qs);
   {
      const BioheatParams P;
      GridFunction T0(&fes);
      FunctionCoefficient Tinit([](const Vector &x)
      {
         const double r2 = (x(0) - .5) * (x(0) - .5) + (x(1) - .5) * (x(1) - .5) + (x(2) - .5) * (x(2) - .5);
         return 37.0 + 20.0 * exp(-40.0 * r2);
      });
      T0.ProjectCoefficient(Tinit);
      QuadratureFunction Tq(qs); Tq.ProjectGridFunction(T0);
      Tq.HostRead(); kq.HostWrite(); mq.HostWrite();   // under Device("cuda") the projection ran on the device
      for (int i = 0; i < Tq.Size(); i++) { kq(i) = P.k0 * (1.0 + P.ak * (Tq(i) - 37.0)); mq(i) = P.rc / P.dt + P.wbcb; }
   }
   QuadratureFunctionCoefficient kc(kq), mc(mq);
   BilinearForm a(&fes); a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a.AddDomainIntegrator(new DiffusionIntegrator(kc));
   a.AddDomainIntegrator(new MassIntegrator(mc));
   a.Assemble();
   const double t_setup = now() - t0;
   const int ND = fes.GetNDofs();
   Vector x(ND), y(ND); x.Randomize(1); y = 0.0;
   x.UseDevice(true); y.UseDevice(true);
   for (int i = 0; i < warm; i++) { a.Mult(x, y); }
   MFEM_DEVICE_SYNC;   // no-op on the CPU devices; with the reference's CUDA backend (make refcuda) the kernels are asynchronous
   vector<double> ts(reps);
   for (int i = 0; i < reps; i++) { t0 = now(); a.Mult(x, y); MFEM_DEVICE_SYNC; ts[i] = now() - t0; }
   double tsum = 0, tmin = 1e300, tmax = 0;
   for (double t : ts) { tsum += t; tmin = min(tmin, t); tmax = max(tmax, t); }
   double t_pcg = 0.0;
   if (pcg_iters > 0)
   {
      Array<int> noess;
      OperatorJacobiSmoother M(a, noess);
      CGSolver cg; cg.SetRelTol(0.0); cg.SetAbsTol(0.0); cg.SetMaxIter(pcg_iters); cg.SetPrintLevel(-1);
      cg.SetOperator(a); cg.SetPreconditioner(M);
      Vector X(ND); X = 0.0;
      MFEM_DEVICE_SYNC;
      t0 = now(); cg.Mult(x, X); MFEM_DEVICE_SYNC; t_pcg = now() - t0;
   }
   int nthreads = 1;
#ifdef _OPENMP
   if (string(dev) == "omp") { nthreads = omp_get_max_threads(); }
#endif
   cout << setprecision(17)
        << "{\"kind\":\"time_apply\",\"p\":" << p << ",\"N\":" << N << ",\"NE\":" << mesh.GetNE() << ",\"ndofs\":" << ND
        << ",\"device\":\"" << dev << "\",\"threads\":" << nthreads << ",\"reps\":" << reps << ",\"warmup\":" << warm
        << ",\"t_setup\":" << t_setup << ",\"t_apply_mean\":" << tsum / reps << ",\"t_apply_min\":" << tmin
        << ",\"t_apply_max\":" << tmax << ",\"y_norm\":" << y.Norml2()
        << ",\"pcg_iters\":" << pcg_iters << ",\"t_pcg\":" << t_pcg << "}" << endl;
   return 0;
}

// ------------------------------------------------------------------------ ex1
// Config 1: examples/ex1.cpp -pa -o 3 on the INLINE 4^3 hex mesh with 3 refinements:
// Jacobi-PCG iteration count to (Br,r) <= 1e-12 (Br,r)_0 (197).
static int ex1(int argc, char **argv)
{
   // ex1 [OUT|-] order refinements dev
   const string out = argc > 2 ? argv[2] : "-";
   const int order = argc > 3 ? atoi(argv[3]) : 3;
   const int ref = argc > 4 ? atoi(argv[4]) : 3;
   const char *dev = argc > 5 ? argv[5] : "cpu";
   Device device(dev);
   Mesh mesh = make_mesh("inline3", ref, 0, 0, 1, 1, 1);
   H1_FECollection fec(order, 3);
   FiniteElementSpace fes(&mesh, &fec);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   LinearForm b(&fes); ConstantCoefficient one(1.0);
   b.AddDomainIntegrator(new DomainLFIntegrator(one)); b.Assemble();
   GridFunction x(&fes); x = 0.0;
   BilinearForm a(&fes); a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
   a.AddDomainIntegrator(new DiffusionIntegrator(one));
   a.Assemble();
   OperatorPtr A; Vector B, X;
   a.FormLinearSystem(ess, x, b, A, X, B);
   OperatorJacobiSmoother M(a, ess);
   CGSolver cg; NormRecorder rec;
   cg.SetRelTol(sqrt(1e-12)); cg.SetAbsTol(0.0); cg.SetMaxIter(400); cg.SetPrintLevel(-1);
   cg.SetOperator(*A); cg.SetPreconditioner(M); cg.SetMonitor(rec);
   double t0 = now(); cg.Mult(B, X); const double t_pcg = now() - t0;
   if (out != "-")
   {
      Dumper D(out);
      const FiniteElement &el = *fes.GetTypicalFE();
      const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
      const DofToQuad &maps = el.GetDofToQuad(ir, DofToQuad::TENSOR);
      const ElementRestriction *R = dynamic_cast<const ElementRestriction *>(
                                       fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
      const GeometricFactors *geom = mesh.GetGeometricFactors(ir, GeometricFactors::JACOBIANS);
      D.iscalar("p", order); D.iscalar("D1D", maps.ndof); D.iscalar("Q1D", maps.nqpt);
      D.iscalar("NE", mesh.GetNE()); D.iscalar("ndofs", fes.GetNDofs());
      D.arr("B", maps.B); D.arr("G", maps.G); D.arr("W", ir.GetWeights());
      D.arr("gather_map", R->GatherMap()); D.arr("offsets", R->Offsets()); D.arr("indices", R->Indices());
      D.vec("J", geom->J); D.arr("ess", ess); D.vec("B_rhs", B); D.vec("X", X);
      D.f64("pcg_norms", rec.norms.data(), rec.norms.size());
      D.iscalar("pcg_iters", cg.GetNumIterations());
   }
   cout << setprecision(17) << "{\"kind\":\"ex1\",\"order\":" << order << ",\"ref\":" << ref << ",\"ndofs\":" << fes.GetNDofs()
        << ",\"iters\":" << cg.GetNumIterations() << ",\"converged\":" << cg.GetConverged()
        << ",\"final_norm\":" << cg.GetFinalNorm() << ",\"t_pcg\":" << t_pcg << ",\"X_norm\":" << X.Norml2() << "}" << endl;
   return 0;
}

// Verify (in the build container only) that data/inline-hex.mesh is MakeCartesian3D(4,4,4).
static int check_inline(const char *path)
{
   Mesh a(path, 1, 1), b = Mesh::MakeCartesian3D(4, 4, 4, Element::HEXAHEDRON, 1.0, 1.0, 1.0);
   bool same = a.GetNE() == b.GetNE() && a.GetNV() == b.GetNV() && a.GetNBE() == b.GetNBE();
   for (int e = 0; same && e < a.GetNE(); e++)
   {
      const int *u = a.GetElement(e)->GetVertices(), *v = b.GetElement(e)->GetVertices();
      for (int j = 0; j < 8; j++) { same = same && u[j] == v[j]; }
   }
   for (int i = 0; same && i < a.GetNV(); i++) { for (int d = 0; d < 3; d++) { same = same && a.GetVertex(i)[d] == b.GetVertex(i)[d]; } }
   cout << "inline-hex == MakeCartesian3D(4,4,4): " << (same ? "yes" : "NO") << endl;
   return same ? 0 : 1;
}

// --------------------------------------------------------------- dump_markers
// dump_markers OUT p kind nx ny nz: element-attribute markers (BilinearForm::AddDomainIntegrator(bfi, elem_marker),
// PABilinearFormExtension::AddMultWithMarkers / AssembleDiagonal, fem/bilinearform_ext.cpp:370-454, 807-847) on a
// three-material mesh (attribute 1 + e % 3).  Marker combinations k = 0..3, each with y = A x and the diagonal.
static int dump_markers(int argc, char **argv)
{
   if (argc < 8) { cerr << "dump_markers OUT p kind nx ny nz\n"; return 2; }
   Dumper D(argv[2]);
   const int p = atoi(argv[3]);
   const string kind = argv[4];
   const int nx = atoi(argv[5]), ny = atoi(argv[6]), nz = atoi(argv[7]);
   Device device("cpu");
   Mesh mesh = make_mesh(kind, nx, ny, nz, 1.0, 1.0, 1.0);
   for (int e = 0; e < mesh.GetNE(); e++) { mesh.SetAttribute(e, 1 + e % 3); }
   mesh.SetAttributes();
   H1_FECollection fec(p, 3);
   FiniteElementSpace fes(&mesh, &fec);
   const FiniteElement &el = *fes.GetTypicalFE();
   const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
   const DofToQuad &maps = el.GetDofToQuad(ir, DofToQuad::TENSOR);
   const int NE = mesh.GetNE(), ND = fes.GetNDofs();
   D.iscalar("p", p); D.iscalar("D1D", maps.ndof); D.iscalar("Q1D", maps.nqpt); D.iscalar("NE", NE); D.iscalar("ndofs", ND);
   D.arr("B", maps.B); D.arr("G", maps.G); D.arr("W", ir.GetWeights());
   const ElementRestriction *R = dynamic_cast<const ElementRestriction *>(fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   D.arr("gather_map", R->GatherMap());
   const GeometricFactors *geom = mesh.GetGeometricFactors(ir, GeometricFactors::JACOBIANS | GeometricFactors::DETERMINANTS);
   D.vec("J", geom->J); D.vec("detJ", geom->detJ);
   {
      Vector vx(3 * mesh.GetNV());
      for (int i = 0; i < mesh.GetNV(); i++) { for (int d = 0; d < 3; d++) { vx(3 * i + d) = mesh.GetVertex(i)[d]; } }
      D.vec("vertices", vx);
      Array<int> ev(8 * NE), at(NE);
      for (int e = 0; e < NE; e++)
      {
         const int *v = mesh.GetElement(e)->GetVertices();
         for (int j = 0; j < 8; j++) { ev[8 * e + j] = v[j]; }
         at[e] = mesh.GetAttribute(e);
      }
      D.arr("elem_vertices", ev); D.arr("elem_attr", at);
   }
   QuadratureSpace qs(mesh, ir);
   FunctionCoefficient kc(kfun), mc(mfun);
   CoefficientVector kq(kc, qs, CoefficientStorage::COMPRESSED), mq(mc, qs, CoefficientStorage::COMPRESSED);
   D.vec("kq", kq); D.vec("mq", mq);
   Vector x(ND); x.Randomize(1);
   D.vec("x", x);
   // marker combinations: -1 = the integrator has no marker
   const int combos[4][2][3] = {{{1, 1, 0}, {1, 0, 1}}, {{-1, -1, -1}, {0, 1, 0}}, {{1, 0, 1}, {-1, -1, -1}}, {{0, 1, 1}, {1, 1, 0}}};
   for (int k = 0; k < 4; k++)
   {
      Array<int> md(3), mm(3);
      for (int i = 0; i < 3; i++) { md[i] = combos[k][0][i]; mm[i] = combos[k][1][i]; }
      const bool has_d = md[0] >= 0, has_m = mm[0] >= 0;
      BilinearForm a(&fes);
      a.SetAssemblyLevel(AssemblyLevel::PARTIAL);
      if (has_d) { a.AddDomainIntegrator(new DiffusionIntegrator(kc), md); } else { a.AddDomainIntegrator(new DiffusionIntegrator(kc)); }
      if (has_m) { a.AddDomainIntegrator(new MassIntegrator(mc), mm); } else { a.AddDomainIntegrator(new MassIntegrator(mc)); }
      a.Assemble();
      Vector y(ND), diag(ND);
      a.Mult(x, y);
      a.AssembleDiagonal(diag);
      const string tag = to_string(k);
      D.arr("marker_diff" + tag, md); D.arr("marker_mass" + tag, mm);
      D.vec("y" + tag, y); D.vec("diag" + tag, diag);
      // the same against the reference's own full assembly (its test: tests/unit/fem/test_pa_kernels.cpp:696-750)
      BilinearForm fa(&fes);
      if (has_d) { fa.AddDomainIntegrator(new DiffusionIntegrator(kc), md); } else { fa.AddDomainIntegrator(new DiffusionIntegrator(kc)); }
      if (has_m) { fa.AddDomainIntegrator(new MassIntegrator(mc), mm); } else { fa.AddDomainIntegrator(new MassIntegrator(mc)); }
      fa.Assemble(); fa.Finalize();
      Vector yf(ND), df(ND);
      fa.Mult(x, yf);
      fa.SpMat().GetDiag(df);
      yf -= y; df -= diag;
      cout << "markers " << k << ": |y_fa - y_pa| = " << yf.Normlinf() << "  |diag_fa - diag_pa| = " << df.Normlinf()
           << "  min diag_pa = " << diag.Min() << endl;
      D.scalar("diag_fa_minus_pa" + tag, df.Normlinf());
   }
   cout << "dump_markers ok: p=" << p << " NE=" << NE << " ndofs=" << ND << endl;
   return 0;
}

// -------------------------------------------------------------------- dump_mg
// dump_mg OUT kind nx ny nz bc orders...: the p-multigrid of examples/ex26.cpp (GeometricMultigrid over an order-refined
// FiniteElementSpaceHierarchy, OperatorChebyshevSmoother of order 2 on the upper levels, unpreconditioned CG on the
// coarsest) for the diffusion + mass operator with function coefficients.  Per level: the space tables and q-data; per
// transfer: its 1-D matrix, P x and P^T x (raw and with the essential-dof constraints); one V-cycle; PCG preconditioned by it.
struct RefMG : public GeometricMultigrid
{
   FunctionCoefficient kc, mc;
   Array<double> eigs;
   RefMG(FiniteElementSpaceHierarchy &h, Array<int> &ess_bdr) : GeometricMultigrid(h, ess_bdr), kc(kfun), mc(mfun)
   {
      for (int l = 0; l < h.GetNumLevels(); ++l)
      {
         FiniteElementSpace &fes = h.GetFESpaceAtLevel(l);
         BilinearForm *form = new BilinearForm(&fes);
         form->SetAssemblyLevel(AssemblyLevel::PARTIAL);
         form->AddDomainIntegrator(new DiffusionIntegrator(kc));
         form->AddDomainIntegrator(new MassIntegrator(mc));
         form->Assemble();
         bfs.Append(form);
         OperatorPtr opr; opr.SetType(Operator::ANY_TYPE);
         Array<int> none;
         const Array<int> &ess = essentialTrueDofs.Size() ? *essentialTrueDofs[l] : none;
         bfs[l]->FormSystemMatrix(ess, opr);
         opr.SetOperatorOwner(false);
         if (l == 0)
         {
            CGSolver *pcg = new CGSolver();
            pcg->SetPrintLevel(-1); pcg->SetMaxIter(200); pcg->SetRelTol(sqrt(1e-4)); pcg->SetAbsTol(0.0);
            pcg->SetOperator(*opr.Ptr());
            AddLevel(opr.Ptr(), pcg, true, true);
            eigs.Append(0.0);
         }
         else
         {
            Vector diag(fes.GetTrueVSize());
            bfs[l]->AssembleDiagonal(diag);
            // the estimate OperatorChebyshevSmoother's power-method constructor makes (linalg/solvers.cpp:497-511), recorded
            OperatorJacobiSmoother invD(diag, ess, 1.0);
            ProductOperator dp(&invD, opr.Ptr(), false, false);
            PowerMethod pm;
            Vector ev(fes.GetTrueVSize());
            const double lam = pm.EstimateLargestEigenvalue(dp, ev, 10, 1e-8, 12345);
            eigs.Append(lam);
            AddLevel(opr.Ptr(), new OperatorChebyshevSmoother(*opr, diag, ess, 2, lam), true, true);
         }
      }
   }
   const Operator *ConstrainedP(int l) const { return prolongations[l]; }
   const Array<int> &Ess(int l) const { static Array<int> none; return essentialTrueDofs.Size() ? *essentialTrueDofs[l] : none; }
   BilinearForm *Form(int l) { return bfs[l]; }
};

static int dump_mg(int argc, char **argv)
{
   if (argc < 9) { cerr << "dump_mg OUT kind nx ny nz bc(none|zfaces|all) order0 order1 ...\n"; return 2; }
   Dumper D(argv[2]);
   const string kind = argv[3];
   const int nx = atoi(argv[4]), ny = atoi(argv[5]), nz = atoi(argv[6]);
   const string bc = argv[7];
   vector<int> orders;
   for (int i = 8; i < argc; i++) { orders.push_back(atoi(argv[i])); }
   Device device("cpu");
   Mesh *mesh = new Mesh(make_mesh(kind, nx, ny, nz, 1.0, 1.0, 1.0));
   vector<FiniteElementCollection *> fecs;
   fecs.push_back(new H1_FECollection(orders[0], 3));
   FiniteElementSpace *coarse = new FiniteElementSpace(mesh, fecs[0]);
   FiniteElementSpaceHierarchy h(mesh, coarse, true, true);
   for (size_t l = 1; l < orders.size(); l++)
   {
      fecs.push_back(new H1_FECollection(orders[l], 3));
      h.AddOrderRefinedLevel(fecs.back());
   }
   Array<int> ess_bdr(mesh->bdr_attributes.Max()); ess_bdr = 0;
   if (bc == "all") { ess_bdr = 1; }
   else if (bc == "zfaces") { ess_bdr[0] = 1; ess_bdr[5] = 1; }
   RefMG mg(h, ess_bdr);
   mg.SetCycleType(Multigrid::CycleType::VCYCLE, 1, 1);
   const int NL = h.GetNumLevels(), NE = mesh->GetNE();
   D.iscalar("nlevels", NL); D.iscalar("NE", NE);
   {
      Vector vx(3 * mesh->GetNV());
      for (int i = 0; i < mesh->GetNV(); i++) { for (int d = 0; d < 3; d++) { vx(3 * i + d) = mesh->GetVertex(i)[d]; } }
      D.vec("vertices", vx);
      Array<int> ev(8 * NE);
      for (int e = 0; e < NE; e++) { const int *v = mesh->GetElement(e)->GetVertices(); for (int j = 0; j < 8; j++) { ev[8 * e + j] = v[j]; } }
      D.arr("elem_vertices", ev);
   }
   for (int l = 0; l < NL; l++)
   {
      FiniteElementSpace &fes = h.GetFESpaceAtLevel(l);
      const FiniteElement &el = *fes.GetTypicalFE();
      const IntegrationRule &ir = DiffusionIntegrator::GetRule(el, el);
      const DofToQuad &maps = el.GetDofToQuad(ir, DofToQuad::TENSOR);
      const string t = to_string(l);
      D.iscalar("order" + t, orders[l]); D.iscalar("ndofs" + t, fes.GetNDofs());
      D.arr("B" + t, maps.B); D.arr("G" + t, maps.G); D.arr("W" + t, ir.GetWeights());
      const ElementRestriction *R = dynamic_cast<const ElementRestriction *>(fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
      D.arr("gather_map" + t, R->GatherMap());
      QuadratureSpace qs(*mesh, ir);
      CoefficientVector kq(mg.kc, qs, CoefficientStorage::COMPRESSED), mq(mg.mc, qs, CoefficientStorage::COMPRESSED);
      D.vec("kq" + t, kq); D.vec("mq" + t, mq);
      D.arr("ess" + t, mg.Ess(l));
      D.scalar("max_eig" + t, mg.eigs[l]);
   }
   for (int l = 0; l + 1 < NL; l++)
   {
      FiniteElementSpace &lf = h.GetFESpaceAtLevel(l), &hf = h.GetFESpaceAtLevel(l + 1);
      // the 1-D matrix as TensorProductPRefinementTransferOperator's constructor forms it (fem/transfer.cpp:2240-2262)
      const FiniteElement &el = *lf.GetTypicalFE();
      const TensorBasisElement *htel = dynamic_cast<const TensorBasisElement *>(hf.GetTypicalFE());
      const Array<int> &hdofmap = htel->GetDofMap();
      const IntegrationRule &irn = hf.GetTypicalFE()->GetNodes();
      IntegrationRule irLex = irn;
      for (int i = 0; i < irn.GetNPoints(); ++i) { const int j = hdofmap[i] >= 0 ? hdofmap[i] : -1 - hdofmap[i]; irLex.IntPoint(i) = irn.IntPoint(j); }
      const DofToQuad &maps = el.GetDofToQuad(irLex, DofToQuad::TENSOR);
      const string t = to_string(l);
      D.arr("PB" + t, maps.B);
      Vector xc(lf.GetNDofs()), xf(hf.GetNDofs()), yf(hf.GetNDofs()), yc(lf.GetNDofs());
      xc.Randomize(11 + l); xf.Randomize(23 + l);
      D.vec("xc" + t, xc); D.vec("xf" + t, xf);
      h.GetProlongationAtLevel(l)->Mult(xc, yf); D.vec("P_xc" + t, yf);
      h.GetProlongationAtLevel(l)->MultTranspose(xf, yc); D.vec("Pt_xf" + t, yc);
      mg.ConstrainedP(l)->Mult(xc, yf); D.vec("Pc_xc" + t, yf);
      mg.ConstrainedP(l)->MultTranspose(xf, yc); D.vec("Pct_xf" + t, yc);
   }
   // one V-cycle and the preconditioned solve on the finest level
   FiniteElementSpace &ff = h.GetFinestFESpace();
   const int n = ff.GetNDofs();
   Vector x(n); x.Randomize(1);
   {
      const Array<int> &ess = mg.Ess(NL - 1);
      for (int i = 0; i < ess.Size(); i++) { x[ess[i]] = 0.0; }   // a residual of the constrained system vanishes there
   }
   D.vec("x", x);
   Vector y(n); y = 0.0;
   mg.Mult(x, y); D.vec("vcycle_x", y);
   {
      GridFunction xg(&ff); xg = 0.0;
      LinearForm b(&ff); ConstantCoefficient one(1.0);
      b.AddDomainIntegrator(new DomainLFIntegrator(one)); b.Assemble();
      OperatorPtr A; Vector X, B;
      // (GeometricMultigrid::FormFineLinearSystem dereferences essentialTrueDofs.Last(), which does not exist without
      //  essential boundaries: go through the form directly - the same call with the same list)
      Array<int> ess_fine(mg.Ess(NL - 1));
      mg.Form(NL - 1)->FormLinearSystem(ess_fine, xg, b, A, X, B);
      D.vec("B_rhs", B);
      for (int pass = 0; pass < 2; pass++)
      {
         CGSolver cg; NormRecorder rec;
         cg.SetRelTol(pass == 0 ? 0.0 : 1e-8); cg.SetAbsTol(0.0); cg.SetMaxIter(pass == 0 ? 3 : 500); cg.SetPrintLevel(-1);
         cg.SetOperator(*A); cg.SetPreconditioner(mg); cg.SetMonitor(rec);
         Vector Xk(X); Xk = 0.0;
         cg.Mult(B, Xk);
         if (pass == 0) { D.vec("X_mgpcg3", Xk); D.f64("mgpcg3_norms", rec.norms.data(), rec.norms.size()); }
         else { D.vec("X_mgpcg_tol", Xk); D.iscalar("mgpcg_tol_iters", cg.GetNumIterations()); D.iscalar("mgpcg_tol_converged", cg.GetConverged()); }
      }
      // Jacobi-PCG iteration count on the same system, for the record
      OperatorJacobiSmoother M(*mg.Form(NL - 1), mg.Ess(NL - 1));
      CGSolver cg; cg.SetRelTol(1e-8); cg.SetAbsTol(0.0); cg.SetMaxIter(5000); cg.SetPrintLevel(-1);
      cg.SetOperator(*A); cg.SetPreconditioner(M);
      Vector Xj(X); Xj = 0.0; cg.Mult(B, Xj);
      D.iscalar("jacobi_pcg_tol_iters", cg.GetNumIterations());
   }
   cout << "dump_mg ok: levels=" << NL << " NE=" << NE << " fine ndofs=" << n << setprecision(17) << " |M x|=" << y.Norml2() << endl;
   for (auto f : fecs) { (void)f; }
   return 0;
}

// ------------------------------------------------------------------ load_check
// The reference loading the product's wire formats: Mesh(file) + GridFunction(mesh, file); dumps what it sees so that
// tests/test_wire_formats.py can compare with the builder (numbering, boundary attributes, values, an L2 norm).
static int load_check(int argc, char **argv)
{
   if (argc < 5) { cerr << "load_check OUT mesh_file gridfunction_file\n"; return 2; }
   Dumper D(argv[2]);
   Device device("cpu");
   Mesh mesh(argv[3], 1, 1);
   ifstream gin(argv[4]);
   MFEM_VERIFY(gin.good(), "cannot open the GridFunction file");
   GridFunction gf(&mesh, gin);
   FiniteElementSpace &fes = *gf.FESpace();
   const ElementRestriction *R = dynamic_cast<const ElementRestriction *>(fes.GetElementRestriction(ElementDofOrdering::LEXICOGRAPHIC));
   MFEM_VERIFY(R, "no ElementRestriction");
   D.iscalar("NE", mesh.GetNE()); D.iscalar("NV", mesh.GetNV()); D.iscalar("NBE", mesh.GetNBE());
   D.iscalar("ndofs", fes.GetNDofs()); D.iscalar("order", fes.GetMaxElementOrder());
   D.arr("gather_map", R->GatherMap());
   D.vec("values", gf);
   Array<int> ess_bdr(mesh.bdr_attributes.Max()); ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   Array<int> ess; fes.GetEssentialTrueDofs(ess_bdr, ess);
   D.arr("ess_z", ess);
   Array<int> battr(mesh.GetNBE());
   for (int i = 0; i < mesh.GetNBE(); i++) { battr[i] = mesh.GetBdrAttribute(i); }
   D.arr("bdr_attributes", battr);
   ConstantCoefficient zero(0.0);
   D.scalar("l2_norm", gf.ComputeL2Error(zero));
   {
      Vector vx(3 * mesh.GetNV());
      for (int i = 0; i < mesh.GetNV(); i++) { for (int d = 0; d < 3; d++) { vx(3 * i + d) = mesh.GetVertex(i)[d]; } }
      D.vec("vertices", vx);
   }
   cout << "load_check ok: NE=" << mesh.GetNE() << " ndofs=" << fes.GetNDofs() << endl;
   return 0;
}

// ------------------------------------------------------------------ paraview
// ParaViewDataCollection::Save by the unmodified reference for a mesh + fields the product wrote in the reference's own
// text formats (so both writers see the same numbering and values): `cycles` saves at cycle c, time 0.25 c, field c
// scaled by (1 + c).   paraview OUTDIR NAME mesh_file format(ascii|binary|binary32) high_order lod cycles name=gf_file ...
static int paraview(int argc, char **argv)
{
   if (argc < 10) { cerr << "paraview OUTDIR NAME mesh_file ascii|binary|binary32 high_order lod cycles name=gf_file ...\n"; return 2; }
   Device device("cpu");
   Mesh mesh(argv[4], 1, 1);
   const string fmt = argv[5];
   const bool high_order = atoi(argv[6]) != 0;
   const int lod = atoi(argv[7]), cycles = atoi(argv[8]);
   vector<unique_ptr<GridFunction>> gfs;
   vector<string> names;
   for (int i = 9; i < argc; i++)
   {
      const string a = argv[i];
      const size_t eq = a.find('=');
      MFEM_VERIFY(eq != string::npos, "name=gf_file expected");
      ifstream gin(a.substr(eq + 1));
      MFEM_VERIFY(gin.good(), "cannot open the GridFunction file");
      gfs.emplace_back(new GridFunction(&mesh, gin));
      names.push_back(a.substr(0, eq));
   }
   ParaViewDataCollection pv(argv[3], &mesh);
   pv.SetPrefixPath(argv[2]);
   pv.SetLevelsOfDetail(lod);
   pv.SetHighOrderOutput(high_order);
   pv.SetDataFormat(fmt == "ascii" ? VTKFormat::ASCII : fmt == "binary32" ? VTKFormat::BINARY32 : VTKFormat::BINARY);
   pv.SetCompressionLevel(0);
   for (size_t i = 0; i < gfs.size(); i++) { pv.RegisterField(names[i], gfs[i].get()); }
   vector<Vector> base;
   for (auto &g : gfs) { base.emplace_back(*g); }
   for (int c = 0; c < cycles; c++)
   {
      for (size_t i = 0; i < gfs.size(); i++) { gfs[i]->Set(1.0 + c, base[i]); }
      pv.SetCycle(c);
      pv.SetTime(0.25 * c);
      pv.Save();
   }
   cout << "paraview ok: NE=" << mesh.GetNE() << " fields=" << gfs.size() << endl;
   return 0;
}

int main(int argc, char **argv)
{
   const string cmd = argc > 1 ? argv[1] : "";
   if (cmd == "dump_case") { return dump_case(argc, argv); }
   if (cmd == "dump_bioheat") { return bioheat(argc, argv, true); }
   if (cmd == "time_bioheat") { return bioheat(argc, argv, false); }
   if (cmd == "dump_bioheat_steps") { return bioheat_steps(argc, argv); }
   if (cmd == "time_apply") { return time_apply(argc, argv); }
   if (cmd == "load_check") { return load_check(argc, argv); }
   if (cmd == "dump_markers") { return dump_markers(argc, argv); }
   if (cmd == "dump_mg") { return dump_mg(argc, argv); }
   if (cmd == "paraview") { return paraview(argc, argv); }
   if (cmd == "ex1") { return ex1(argc, argv); }
   if (cmd == "--check-inline" && argc > 2) { return check_inline(argv[2]); }
   cerr << "usage: ref_driver dump_case|dump_bioheat|time_bioheat|time_apply|ex1|--check-inline ...\n";
   return 2;
}
