// mfem_b200pa.hpp — the reference-side binding: MFEM classes whose PA virtuals run on a B200 through
// the C ABI of libb200pa.so (include/b200pa.h).  Header-only; include it from an MFEM application
// (a CPU build of MFEM is enough: the host library never sees a device pointer) and link -lb200pa.
//
//   drop-in level 1  b200::DiffusionIntegrator / b200::MassIntegrator
//        subclasses of mfem::DiffusionIntegrator / mfem::MassIntegrator that override
//        AssemblePA / AddMultPA / AddMultTransposePA / AssembleDiagonalPA (fem/bilininteg.hpp:52-97).
//        Hand them to an ordinary BilinearForm with AssemblyLevel::PARTIAL; MFEM's own
//        PABilinearFormExtension (fem/bilinearform_ext.cpp:332-564) keeps orchestrating gather,
//        integrator calls and scatter, E-vectors cross PCIe per call.  Bit-for-tolerance, not fast.
//   drop-in level 2  b200::PAOperator (an mfem::Operator) + b200::PCGSolver (an mfem::Solver)
//        the whole L->L apply (ElementRestriction::Mult, both AddMultPA, MultTranspose) as one fused
//        launch pair, ConstrainedOperator semantics, OperatorJacobiSmoother and CGSolver::Mult
//        (linalg/solvers.cpp:869-1050) resident on the GPU; vectors cross PCIe once per solve.
//
// Error convention: a nonzero status from the C ABI becomes mfem::mfem_error(b200pa_last_error()),
// i.e. MFEM_ABORT semantics (general/error.hpp:26-64).  Unsupported spaces (non-hex, vdim > 1,
// orders > 6, non-default rules, matrix coefficients) abort: there is no CPU fallback.
#ifndef MFEM_B200PA_HPP
#define MFEM_B200PA_HPP

#include "mfem.hpp"

#include <cmath>
#include <iomanip>
#include <memory>
#include <vector>

#include "../../include/b200pa.h"

namespace b200
{

inline void Check(int rc)
{
   if (rc) { mfem::mfem_error(b200pa_last_error()); }
}

/// One context per process and GPU (≙ mfem::Device for this path).
inline b200pa_ctx Ctx(int device = 0)
{
   static b200pa_ctx ctx = nullptr;
   if (!ctx) { Check(b200pa_ctx_create(device, nullptr, &ctx)); }
   return ctx;
}

/// RAII device buffer owned by the library side.
class DeviceBuffer
{
   void *p = nullptr;
   size_t bytes = 0;
public:
   DeviceBuffer() = default;
   DeviceBuffer(const DeviceBuffer &) = delete;
   DeviceBuffer &operator=(const DeviceBuffer &) = delete;
   ~DeviceBuffer() { if (p) { b200pa_free(Ctx(), p); } }
   void Resize(size_t n)
   {
      if (n <= bytes && p) { return; }
      if (p) { Check(b200pa_free(Ctx(), p)); p = nullptr; }
      Check(b200pa_malloc(Ctx(), n, &p));
      bytes = n;
   }
   void Upload(const void *src, size_t n) { Resize(n); Check(b200pa_ctx_upload(Ctx(), p, src, n)); }
   void Download(void *dst, size_t n) const { Check(b200pa_ctx_download(Ctx(), dst, p, n)); }
   double *D() const { return static_cast<double *>(p); }
   int *I() const { return static_cast<int *>(p); }
};

namespace internal
{
/// The quadrature rule, tensor maps and q-data a PA integrator of the reference sets up.
struct PASetup
{
   int ne = 0, d1d = 0, q1d = 0, nq = 0;
   const mfem::IntegrationRule *ir = nullptr;
   const mfem::DofToQuad *maps = nullptr;
   void Init(const mfem::FiniteElementSpace &fes, const mfem::IntegrationRule *rule)
   {
      mfem::Mesh *mesh = fes.GetMesh();
      MFEM_VERIFY(mesh->Dimension() == 3 && mesh->GetNumGeometries(3) == 1 &&
                  mesh->GetElementBaseGeometry(0) == mfem::Geometry::CUBE,
                  "b200pa: 3-D hexahedral meshes only");
      MFEM_VERIFY(fes.GetVDim() == 1 && !fes.IsVariableOrder() && fes.Conforming(), "b200pa: scalar conforming H1 spaces only");
      const mfem::FiniteElement &el = *fes.GetTypicalFE();
      ir = rule;
      maps = &el.GetDofToQuad(*ir, mfem::DofToQuad::TENSOR);
      ne = fes.GetNE(); d1d = maps->ndof; q1d = maps->nqpt; nq = ir->GetNPoints();
   }
};
} // namespace internal

/// (Q grad u, grad v), scalar Q: mfem::DiffusionIntegrator with its PA virtuals on the GPU.
class DiffusionIntegrator : public mfem::DiffusionIntegrator
{
   internal::PASetup s;
   DeviceBuffer d_pa;
   mutable DeviceBuffer d_x, d_y;
public:
   using mfem::DiffusionIntegrator::DiffusionIntegrator;

   /// ≙ DiffusionIntegrator::AssemblePA, fem/integ/bilininteg_diffusion_pa.cpp:89-142
   void AssemblePA(const mfem::FiniteElementSpace &fes) override
   {
      MFEM_VERIFY(!VQ && !MQ, "b200pa: scalar diffusion coefficients only");
      const mfem::FiniteElement &el = *fes.GetTypicalFE();
      s.Init(fes, IntRule ? IntRule : &GetRule(el, el));
      mfem::Mesh *mesh = fes.GetMesh();
      const mfem::GeometricFactors *geom = mesh->GetGeometricFactors(*s.ir, mfem::GeometricFactors::JACOBIANS);
      mfem::QuadratureSpace qs(*mesh, *s.ir);
      mfem::CoefficientVector coeff(qs, mfem::CoefficientStorage::COMPRESSED);
      if (Q) { coeff.Project(*Q); } else { coeff.SetConstant(1.0); }
      DeviceBuffer dW, dJ, dC;
      dW.Upload(s.ir->GetWeights().HostRead(), sizeof(double) * s.nq);
      dJ.Upload(geom->J.HostRead(), sizeof(double) * geom->J.Size());
      dC.Upload(coeff.HostRead(), sizeof(double) * coeff.Size());
      d_pa.Resize(sizeof(double) * 6 * (size_t)s.nq * s.ne);
      Check(b200pa_diffusion_setup(Ctx(), s.q1d, s.ne, dW.D(), dJ.D(), dC.D(), coeff.Size(), d_pa.D()));
      Check(b200pa_ctx_sync(Ctx()));
   }

   /// ≙ DiffusionIntegrator::AddMultPA, fem/integ/bilininteg_diffusion_pa.cpp:40-74 (y += A_E x)
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override
   {
      const size_t b = sizeof(double) * x.Size();
      d_x.Upload(x.HostRead(), b);
      d_y.Upload(y.HostRead(), b);
      Check(b200pa_diffusion_apply(Ctx(), s.ne, s.d1d, s.q1d, s.maps->B.HostRead(), s.maps->G.HostRead(), d_pa.D(), d_x.D(), d_y.D()));
      d_y.Download(y.HostReadWrite(), b);
   }
   void AddMultTransposePA(const mfem::Vector &x, mfem::Vector &y) const override { AddMultPA(x, y); } // symmetric

   /// ≙ DiffusionIntegrator::AssembleDiagonalPA, fem/integ/bilininteg_diffusion_pa.cpp:22-37
   void AssembleDiagonalPA(mfem::Vector &diag) override
   {
      const size_t b = sizeof(double) * diag.Size();
      d_y.Upload(diag.HostRead(), b);
      Check(b200pa_diffusion_diag(Ctx(), s.ne, s.d1d, s.q1d, s.maps->B.HostRead(), s.maps->G.HostRead(), d_pa.D(), d_y.D()));
      d_y.Download(diag.HostReadWrite(), b);
   }
   const double *DevicePAData() const { return d_pa.D(); }
};

/// (Q u, v): mfem::MassIntegrator with its PA virtuals on the GPU.
class MassIntegrator : public mfem::MassIntegrator
{
   internal::PASetup s;
   DeviceBuffer d_pa;
   mutable DeviceBuffer d_x, d_y;
public:
   using mfem::MassIntegrator::MassIntegrator;

   /// ≙ MassIntegrator::AssemblePA, fem/integ/bilininteg_mass_pa.cpp:24-79
   void AssemblePA(const mfem::FiniteElementSpace &fes) override
   {
      const mfem::FiniteElement &el = *fes.GetTypicalFE();
      mfem::Mesh *mesh = fes.GetMesh();
      mfem::ElementTransformation *T0 = mesh->GetTypicalElementTransformation();
      s.Init(fes, IntRule ? IntRule : &GetRule(el, el, *T0));
      MFEM_VERIFY(el.GetMapType() == mfem::FiniteElement::VALUE, "b200pa: VALUE map type only");
      const mfem::GeometricFactors *geom = mesh->GetGeometricFactors(*s.ir, mfem::GeometricFactors::DETERMINANTS);
      mfem::QuadratureSpace qs(*mesh, *s.ir);
      mfem::CoefficientVector coeff(Q, qs, mfem::CoefficientStorage::COMPRESSED);
      DeviceBuffer dW, dD, dC;
      dW.Upload(s.ir->GetWeights().HostRead(), sizeof(double) * s.nq);
      dD.Upload(geom->detJ.HostRead(), sizeof(double) * geom->detJ.Size());
      dC.Upload(coeff.HostRead(), sizeof(double) * coeff.Size());
      d_pa.Resize(sizeof(double) * (size_t)s.nq * s.ne);
      Check(b200pa_mass_setup(Ctx(), s.nq, s.ne, dW.D(), dD.D(), dC.D(), coeff.Size(), d_pa.D()));
      Check(b200pa_ctx_sync(Ctx()));
   }

   /// ≙ MassIntegrator::AddMultPA, fem/integ/bilininteg_mass_pa.cpp:140-169
   void AddMultPA(const mfem::Vector &x, mfem::Vector &y) const override
   {
      const size_t b = sizeof(double) * x.Size();
      d_x.Upload(x.HostRead(), b);
      d_y.Upload(y.HostRead(), b);
      Check(b200pa_mass_apply(Ctx(), s.ne, s.d1d, s.q1d, s.maps->B.HostRead(), d_pa.D(), d_x.D(), d_y.D()));
      d_y.Download(y.HostReadWrite(), b);
   }
   void AddMultTransposePA(const mfem::Vector &x, mfem::Vector &y) const override { AddMultPA(x, y); }

   /// ≙ MassIntegrator::AssembleDiagonalPA, fem/integ/bilininteg_mass_pa.cpp:127-138
   void AssembleDiagonalPA(mfem::Vector &diag) override
   {
      const size_t b = sizeof(double) * diag.Size();
      d_y.Upload(diag.HostRead(), b);
      Check(b200pa_mass_diag(Ctx(), s.ne, s.d1d, s.q1d, s.maps->B.HostRead(), d_pa.D(), d_y.D()));
      d_y.Download(diag.HostReadWrite(), b);
   }
};

/// The fused operator: PABilinearFormExtension::Mult + ConstrainedOperator for a form made of a
/// DiffusionIntegrator and/or a MassIntegrator with scalar coefficients (either may be null), added in that order.
/// diff_marker / mass_marker: the element-attribute markers of BilinearForm::AddDomainIntegrator(bfi, elem_marker)
/// (multi-material domains; semantics of fem/bilinearform_ext.cpp:370-454, 807-847, see include/b200pa.h).
class PAOperator : public mfem::Operator
{
   const mfem::FiniteElementSpace &fes;
   b200pa_space sp = nullptr;
   b200pa_form form = nullptr;
   mfem::Array<int> ess;
   const mfem::IntegrationRule *ir = nullptr;
   friend class PCGSolver;
   friend class JacobiSmoother;
   friend class ChebyshevSmoother;
   friend class PMultigrid;

   void Project(mfem::Coefficient *c, std::vector<double> &out)
   {
      mfem::QuadratureSpace qs(*fes.GetMesh(), *ir);
      mfem::CoefficientVector cv(c, qs, mfem::CoefficientStorage::COMPRESSED);
      out.assign(cv.HostRead(), cv.HostRead() + cv.Size());
   }
public:
   /// factorised = true: on a mesh whose elements are all affine (b200pa_space_is_affine) the diffusion q-data is kept
   /// as w_q c_q per q-point + one tensor per element (b200pa_form_set_factorised); other meshes keep the stored form
   PAOperator(const mfem::FiniteElementSpace &fes_, mfem::Coefficient *kdiff, mfem::Coefficient *cmass,
              const mfem::Array<int> &ess_tdof_list, bool factorised = false,
              const mfem::Array<int> *diff_marker = nullptr, const mfem::Array<int> *mass_marker = nullptr)
      : mfem::Operator(fes_.GetVSize()), fes(fes_)
   {
      const mfem::FiniteElement &el = *fes.GetTypicalFE();
      internal::PASetup s;
      ir = &mfem::DiffusionIntegrator::GetRule(el, el);
      s.Init(fes, ir);
      // ElementRestriction tables (fem/restriction.cpp:26-107) and tensor maps (fem/fe/fe_base.cpp:2619-2662)
      const mfem::ElementRestriction *R = dynamic_cast<const mfem::ElementRestriction *>(
                                             fes.GetElementRestriction(mfem::ElementDofOrdering::LEXICOGRAPHIC));
      MFEM_VERIFY(R, "b200pa: ElementRestriction expected");
      Check(b200pa_space_create(Ctx(), s.d1d, s.q1d, s.ne, fes.GetNDofs(), R->GatherMap().HostRead(), s.maps->B.HostRead(),
                                s.maps->G.HostRead(), &sp));
      const mfem::GeometricFactors *geom = fes.GetMesh()->GetGeometricFactors(
                                              *ir, mfem::GeometricFactors::JACOBIANS | mfem::GeometricFactors::DETERMINANTS);
      Check(b200pa_space_set_geometry(sp, ir->GetWeights().HostRead(), geom->J.HostRead(), geom->detJ.HostRead()));
      Check(b200pa_form_create(sp, &form));
      if (factorised && b200pa_space_is_affine(sp) == 1) { Check(b200pa_form_set_factorised(form, 1)); }
      if (diff_marker || mass_marker)
      {
         // elem_attributes of PABilinearFormExtension (fem/bilinearform_ext.cpp: SetupRestrictionOperators)
         mfem::Array<int> attr(s.ne);
         for (int e = 0; e < s.ne; e++) { attr[e] = fes.GetMesh()->GetAttribute(e); }
         Check(b200pa_space_set_attributes(sp, attr.HostRead()));
         if (diff_marker) { Check(b200pa_form_set_markers(form, 0, diff_marker->Size(), diff_marker->HostRead())); }
         if (mass_marker) { Check(b200pa_form_set_markers(form, 1, mass_marker->Size(), mass_marker->HostRead())); }
      }
      std::vector<double> q;
      if (kdiff) { Project(kdiff, q); Check(b200pa_form_assemble_diffusion(form, q.data(), (long long)q.size())); }
      if (cmass) { Project(cmass, q); Check(b200pa_form_assemble_mass(form, q.data(), (long long)q.size())); }
      ess_tdof_list.Copy(ess);
      Check(b200pa_form_set_essential(form, ess.Size(), ess.HostRead()));
   }
   ~PAOperator() { b200pa_form_destroy(form); b200pa_space_destroy(sp); }
   bool Factorised() const { return b200pa_form_is_factorised(form) == 1; }

   /// ≙ ConstrainedOperator::Mult (linalg/operator.cpp:710-714) of the PA form, host vectors
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      Check(b200pa_form_mult_host(form, 1, x.HostRead(), y.HostWrite()));
   }
   /// ≙ PABilinearFormExtension::Mult (fem/bilinearform_ext.cpp:487-564): unconstrained A x
   void MultUnconstrained(const mfem::Vector &x, mfem::Vector &y) const
   {
      Check(b200pa_form_mult_host(form, 0, x.HostRead(), y.HostWrite()));
   }
   /// ≙ PABilinearFormExtension::AssembleDiagonal (fem/bilinearform_ext.cpp:370-454)
   void AssembleDiagonal(mfem::Vector &diag) const override
   {
      DeviceBuffer d;
      d.Resize(sizeof(double) * height);
      Check(b200pa_form_assemble_diagonal(form, d.D()));
      d.Download(diag.HostWrite(), sizeof(double) * height);
   }
   /// ≙ ConstrainedOperator::EliminateRHS (linalg/operator.cpp:559-584)
   void EliminateRHS(const mfem::Vector &x, mfem::Vector &b) const
   {
      DeviceBuffer dx, db;
      dx.Upload(x.HostRead(), sizeof(double) * height);
      db.Upload(b.HostRead(), sizeof(double) * height);
      Check(b200pa_form_eliminate_rhs(form, dx.D(), db.D()));
      db.Download(b.HostReadWrite(), sizeof(double) * height);
   }
   const mfem::Array<int> &GetEssentialTrueDofs() const { return ess; }
};

/// ≙ OperatorJacobiSmoother(a, ess_tdof_list, damping) (linalg/solvers.cpp:331-453) for a b200::PAOperator: the diagonal is
/// assembled and inverted on the GPU.  A Solver of its own (Mult works on host vectors, so mfem::CGSolver can use it) and
/// the preconditioner b200::PCGSolver fuses into its device-resident loop.
class JacobiSmoother : public mfem::Solver
{
   const PAOperator *op = nullptr;
   DeviceBuffer dinv;
   mutable DeviceBuffer d_r, d_z;
   double damping;
public:
   explicit JacobiSmoother(double damping_ = 1.0) : mfem::Solver(0, false), damping(damping_) {}
   JacobiSmoother(const PAOperator &A, double damping_ = 1.0) : mfem::Solver(0, false), damping(damping_) { SetOperator(A); }
   void SetOperator(const mfem::Operator &o) override
   {
      op = dynamic_cast<const PAOperator *>(&o);
      MFEM_VERIFY(op, "b200::JacobiSmoother works on a b200::PAOperator");
      height = width = op->Height();
      DeviceBuffer diag, dess;
      diag.Resize(sizeof(double) * height);
      dinv.Resize(sizeof(double) * height);
      Check(b200pa_form_assemble_diagonal(op->form, diag.D()));
      dess.Resize(sizeof(int) * std::max(op->ess.Size(), 1));
      if (op->ess.Size()) { dess.Upload(op->ess.HostRead(), sizeof(int) * op->ess.Size()); }
      Check(b200pa_jacobi_setup(Ctx(), height, diag.D(), op->ess.Size(), dess.I(), damping, dinv.D()));
   }
   /// z = dinv .* r  (OperatorJacobiSmoother::Mult with iterative_mode = false, linalg/solvers.cpp:427-453)
   void Mult(const mfem::Vector &r, mfem::Vector &z) const override
   {
      MFEM_VERIFY(op, "b200::JacobiSmoother: SetOperator first");
      MFEM_VERIFY(!iterative_mode, "b200::JacobiSmoother: iterative_mode is not supported");
      d_r.Upload(r.HostRead(), sizeof(double) * height);
      d_z.Resize(sizeof(double) * height);
      Check(b200pa_jacobi_mult(Ctx(), height, dinv.D(), d_r.D(), d_z.D()));
      d_z.Download(z.HostWrite(), sizeof(double) * height);
   }
   const PAOperator *GetOperator() const { return op; }
   const double *DeviceDinv() const { return dinv.D(); }
};

/// ≙ OperatorChebyshevSmoother(A, diag, ess, order[, max_eig_estimate]) (linalg/solvers.cpp:455-657); without an estimate the
/// reference's power method (10 steps, 1e-8, seed 12345, linalg/solvers.cpp:497-511) runs on the GPU at SetOperator time.
class ChebyshevSmoother : public mfem::Solver
{
   JacobiSmoother jac;
   int order;
   double max_eig;
   mutable DeviceBuffer d_x, d_y;
public:
   explicit ChebyshevSmoother(int order_, double max_eig_estimate = 0.0) : mfem::Solver(0, false), jac(1.0), order(order_), max_eig(max_eig_estimate) {}
   void SetOperator(const mfem::Operator &o) override
   {
      jac.SetOperator(o);
      height = width = jac.Height();
      if (max_eig <= 0.0)
      {
         mfem::Vector v0(height);
         v0.Randomize(12345);
         DeviceBuffer dv;
         dv.Upload(v0.HostRead(), sizeof(double) * height);
         Check(b200pa_power_method(jac.GetOperator()->form, jac.DeviceDinv(), dv.D(), 10, 1e-8, &max_eig));
      }
   }
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      MFEM_VERIFY(jac.GetOperator(), "b200::ChebyshevSmoother: SetOperator first");
      d_x.Upload(x.HostRead(), sizeof(double) * height);
      d_y.Resize(sizeof(double) * height);
      Check(b200pa_chebyshev_mult(jac.GetOperator()->form, jac.DeviceDinv(), order, max_eig, d_x.D(), d_y.D()));
      d_y.Download(y.HostWrite(), sizeof(double) * height);
   }
   int Order() const { return order; }
   double GetMaxEigEstimate() const { return max_eig; }
   const JacobiSmoother &Jacobi() const { return jac; }
};

/// p-multigrid on the GPU: mfem::GeometricMultigrid over an order-refined FiniteElementSpaceHierarchy (fem/multigrid.hpp) with the
/// smoothers and coarse solver of examples/ex26.cpp - OperatorChebyshevSmoother (order 2, power-method estimate) on the upper
/// levels, unpreconditioned CG (rel. tol. 1e-2, 200 iterations) on the coarsest - for the diffusion (+ mass) operator.  The
/// order-refinement transfers use the 1-D matrices TensorProductPRefinementTransferOperator builds (fem/transfer.cpp:2240-2262).
/// A Solver: Mult is one V-cycle from a zero guess (MultigridBase::Mult); b200::PCGSolver fuses it as its preconditioner.
class PMultigrid : public mfem::Solver
{
   std::vector<std::unique_ptr<PAOperator>> ops;
   std::vector<b200pa_transfer> transfers;
   b200pa_mg mg = nullptr;
   mutable DeviceBuffer d_x, d_y;
   friend class PCGSolver;
public:
   PMultigrid(const mfem::FiniteElementSpaceHierarchy &h, mfem::Coefficient *kdiff, mfem::Coefficient *cmass, const mfem::Array<int> &ess_bdr,
              int cheb_order = 2, bool factorised = false)
      : mfem::Solver(h.GetFinestFESpace().GetVSize(), false)
   {
      const int nl = h.GetNumLevels();
      bool have_ess = false;
      for (int i = 0; i < ess_bdr.Size(); i++) { have_ess = have_ess || ess_bdr[i]; }
      std::vector<b200pa_form> forms;
      for (int l = 0; l < nl; l++)
      {
         const mfem::FiniteElementSpace &fes = h.GetFESpaceAtLevel(l);
         mfem::Array<int> ess;
         if (have_ess) { fes.GetEssentialTrueDofs(ess_bdr, ess); }
         ops.emplace_back(new PAOperator(fes, kdiff, cmass, ess, factorised));
         forms.push_back(ops.back()->form);
      }
      for (int l = 0; l + 1 < nl; l++)
      {
         const mfem::FiniteElementSpace &lf = h.GetFESpaceAtLevel(l), &hf = h.GetFESpaceAtLevel(l + 1);
         MFEM_VERIFY(lf.GetMesh() == hf.GetMesh(), "b200::PMultigrid: order refinement only (the levels share one mesh)");
         const mfem::TensorBasisElement *htel = dynamic_cast<const mfem::TensorBasisElement *>(hf.GetTypicalFE());
         MFEM_VERIFY(htel, "b200::PMultigrid: tensor-product elements expected");
         const mfem::Array<int> &hdofmap = htel->GetDofMap();
         const mfem::IntegrationRule &irn = hf.GetTypicalFE()->GetNodes();
         mfem::IntegrationRule irLex = irn;
         for (int i = 0; i < irn.GetNPoints(); ++i) { const int j = hdofmap[i] >= 0 ? hdofmap[i] : -1 - hdofmap[i]; irLex.IntPoint(i) = irn.IntPoint(j); }
         const mfem::DofToQuad &maps = lf.GetTypicalFE()->GetDofToQuad(irLex, mfem::DofToQuad::TENSOR);
         b200pa_transfer t = nullptr;
         Check(b200pa_transfer_create(forms[l], forms[l + 1], maps.B.HostRead(), &t));
         transfers.push_back(t);
      }
      Check(b200pa_mg_create(nl, forms.data(), transfers.data(), &mg));
      Check(b200pa_mg_set_cycle(mg, 0, 1, 1));
      Check(b200pa_mg_set_coarse_solver(mg, 1e-2, 0.0, 200, 0));
      std::vector<int> order(nl, cheb_order);
      Check(b200pa_mg_setup(mg, order.data(), nullptr));
   }
   ~PMultigrid()
   {
      b200pa_mg_destroy(mg);
      for (b200pa_transfer t : transfers) { b200pa_transfer_destroy(t); }
   }
   void SetCycleType(bool wcycle, int pre, int post) { Check(b200pa_mg_set_cycle(mg, wcycle ? 1 : 0, pre, post)); }
   void SetCoarseSolver(double rel_tol, double abs_tol, int max_iter, bool jacobi) { Check(b200pa_mg_set_coarse_solver(mg, rel_tol, abs_tol, max_iter, jacobi ? 1 : 0)); Check(b200pa_mg_setup(mg, nullptr, nullptr)); }
   void SetOperator(const mfem::Operator &) override {}
   const PAOperator &FineOperator() const { return *ops.back(); }
   double GetMaxEigEstimate(int level) const { return b200pa_mg_max_eig(mg, level); }
   /// one cycle: y = M x (MultigridBase::Mult, fem/multigrid.cpp:107-134)
   void Mult(const mfem::Vector &x, mfem::Vector &y) const override
   {
      d_x.Upload(x.HostRead(), sizeof(double) * height);
      d_y.Resize(sizeof(double) * height);
      Check(b200pa_mg_mult(mg, d_x.D(), d_y.D()));
      d_y.Download(y.HostWrite(), sizeof(double) * height);
   }
};

/// mfem::CGSolver on the GPU (linalg/solvers.cpp:869-1050): an mfem::IterativeSolver - it can be handed wherever the
/// reference takes an IterativeSolver& - whose Mult runs the whole loop device-resident on a b200::PAOperator.
///   SetPreconditioner   b200::JacobiSmoother, b200::ChebyshevSmoother or b200::PMultigrid (fused into the loop); none = plain
///                       CG, as in the reference; any other Solver aborts - there is no CPU fallback.
///   SetPrintLevel       the reference's PrintLevel flags print the reference's lines from the recorded (B r, r) history.
///   SetMonitor          MonitorResidual / MonitorSolution are called for every iteration AFTER the solve with the recorded
///                       norms; the vectors they receive are the final residual and solution (the loop keeps its iterates
///                       on the device), and a controller cannot stop the iteration (HasConverged is ignored, a controller
///                       that RequiresUpdatedSolution is refused).
/// Same meaning of iterative_mode, rel/abs tolerance, GetNumIterations, GetConverged, GetInitialNorm, GetFinalNorm.
class PCGSolver : public mfem::IterativeSolver
{
   const PAOperator *op = nullptr;
   mutable DeviceBuffer ones;
   mutable std::vector<double> norms;
   // shortcuts that own their smoother: UseJacobi(damping), SetChebyshev(order)
   std::unique_ptr<ChebyshevSmoother> own_cheb;
   std::unique_ptr<JacobiSmoother> own_jac;
public:
   PCGSolver() : mfem::IterativeSolver() {}
   void UseJacobi(double damping = 1.0) { own_jac.reset(new JacobiSmoother(damping)); prec = own_jac.get(); }
   void SetChebyshev(int order, double max_eig = 0.0) { own_cheb.reset(new ChebyshevSmoother(order, max_eig)); prec = own_cheb.get(); }
   double GetMaxEigEstimate() const { return own_cheb ? own_cheb->GetMaxEigEstimate() : 0.0; }

   void SetPreconditioner(mfem::Solver &pr) override
   {
      MFEM_VERIFY(dynamic_cast<JacobiSmoother *>(&pr) || dynamic_cast<ChebyshevSmoother *>(&pr) || dynamic_cast<PMultigrid *>(&pr),
                  "b200::PCGSolver: the preconditioner must be a b200::JacobiSmoother, ChebyshevSmoother or PMultigrid (no CPU fallback)");
      mfem::IterativeSolver::SetPreconditioner(pr);
   }
   void SetOperator(const mfem::Operator &o) override
   {
      op = dynamic_cast<const PAOperator *>(&o);
      MFEM_VERIFY(op, "b200::PCGSolver works on a b200::PAOperator");
      mfem::IterativeSolver::SetOperator(o); // sets height/width and hands the operator to the preconditioner
   }
   void Mult(const mfem::Vector &b, mfem::Vector &x) const override
   {
      MFEM_VERIFY(op, "b200::PCGSolver: SetOperator first");
      MFEM_VERIFY(!ControllerRequiresUpdate(), "b200::PCGSolver: controllers that need the updated solution every iteration are not supported");
      if (!iterative_mode) { x = 0.0; }
      norms.assign(max_iter + 2, 0.0);
      b200pa_pcg_result res{};
      const JacobiSmoother *jac = dynamic_cast<const JacobiSmoother *>(prec);
      const ChebyshevSmoother *cheb = dynamic_cast<const ChebyshevSmoother *>(prec);
      if (jac && !jac->GetOperator()) { const_cast<JacobiSmoother *>(jac)->SetOperator(*op); }
      if (cheb && !cheb->Jacobi().GetOperator()) { const_cast<ChebyshevSmoother *>(cheb)->SetOperator(*op); }
      const PMultigrid *pmg = dynamic_cast<const PMultigrid *>(prec);
      if (pmg)
      {
         MFEM_VERIFY(&pmg->FineOperator() == op, "b200::PCGSolver: the operator must be the multigrid's finest-level operator (PMultigrid::FineOperator)");
         DeviceBuffer db, dx;
         db.Upload(b.HostRead(), sizeof(double) * height);
         dx.Upload(x.HostRead(), sizeof(double) * height);
         Check(b200pa_pcg_solve_mg(pmg->mg, db.D(), dx.D(), rel_tol, abs_tol, max_iter, &res, norms.data()));
         dx.Download(x.HostReadWrite(), sizeof(double) * height);
      }
      else if (cheb)
      {
         DeviceBuffer db, dx;
         db.Upload(b.HostRead(), sizeof(double) * height);
         dx.Upload(x.HostRead(), sizeof(double) * height);
         Check(b200pa_pcg_solve_chebyshev(op->form, cheb->Jacobi().DeviceDinv(), cheb->Order(), cheb->GetMaxEigEstimate(), db.D(), dx.D(),
                                          rel_tol, abs_tol, max_iter, &res, norms.data()));
         dx.Download(x.HostReadWrite(), sizeof(double) * height);
      }
      else
      {
         const double *dinv = nullptr;
         if (jac) { dinv = jac->DeviceDinv(); }
         else
         {
            // no preconditioner: d = r (linalg/solvers.cpp:889-892) == Jacobi with dinv = 1
            std::vector<double> one(height, 1.0);
            ones.Upload(one.data(), sizeof(double) * height);
            dinv = ones.D();
         }
         Check(b200pa_pcg_solve_host(op->form, dinv, b.HostRead(), x.HostReadWrite(), rel_tol, abs_tol, max_iter, &res, norms.data()));
      }
      final_iter = res.final_iter;
      converged = res.converged != 0;
      initial_norm = res.initial_norm;
      final_norm = res.final_norm;
      Report(b, x);
   }
   const std::vector<double> &GetResidualHistory() const { return norms; }
private:
   // the output and monitor calls of CGSolver::Mult (linalg/solvers.cpp:897-1047), from the recorded history
   void Report(const mfem::Vector &b, const mfem::Vector &x) const
   {
      using std::setw;
      const PrintLevel &po = print_options;
      const double nom0 = norms[0];
      if (po.iterations || po.first_and_last)
      {
         mfem::out << "   Iteration : " << setw(3) << 0 << "  (B r, r) = " << nom0 << (po.first_and_last ? " ...\n" : "\n");
      }
      if (nom0 < 0.0)
      {
         if (po.warnings) { mfem::out << "PCG: The preconditioner is not positive definite. (Br, r) = " << nom0 << '\n'; }
      }
      else if (final_iter > 0 || !converged)
      {
         const double betanom = norms[final_iter];
         if (po.iterations) { for (int i = 1; i <= final_iter; i++) { mfem::out << "   Iteration : " << setw(3) << i << "  (B r, r) = " << norms[i] << std::endl; } }
         if (betanom < 0.0 && po.warnings) { mfem::out << "PCG: The preconditioner is not positive definite. (Br, r) = " << betanom << '\n'; }
         if (po.first_and_last && !po.iterations) { mfem::out << "   Iteration : " << setw(3) << final_iter << "  (B r, r) = " << betanom << '\n'; }
         if (po.summary || (po.warnings && !converged)) { mfem::out << "PCG: Number of iterations: " << final_iter << '\n'; }
         if ((po.summary || po.iterations || po.first_and_last) && final_iter > 0)
         {
            mfem::out << "Average reduction factor = " << pow(betanom / nom0, 0.5 / final_iter) << '\n';
         }
         if (po.warnings && !converged) { mfem::out << "PCG: No convergence!" << '\n'; }
      }
      if (controller)
      {
         // final residual for the monitor: r = b - A x (one more apply; only when somebody is watching)
         mfem::Vector r(height);
         op->Mult(x, r);
         subtract(b, r, r);
         for (int i = 0; i <= final_iter; i++) { Monitor(i, norms[i], r, x, false); }
         Monitor(final_iter, final_iter == 0 ? norms[0] : final_norm, r, x, true);
      }
   }
};

/// The Pennes bioheat equation as an mfem::TimeDependentOperator whose implicit solve runs on the GPU (pattern:
/// ConductionOperator::ImplicitSolve, examples/ex16.cpp:326-379; stepped by BackwardEulerSolver::Step,
/// linalg/ode.cpp:682-696 - or any SDIRK solver, they all call ImplicitSolve):
///     rho c dT/dt = div k(T) grad T - w (T - Ta) + q,    k(T) = k0 (1 + ak (T - Tref)),   natural BCs
///     ImplicitSolve(dt, T, dT):  [M(rho c + dt w) + dt K(k(T))] dT = -[K(k(T)) + M(w)] T + (w Ta + q, v)
/// T, the q-data and all solver vectors stay on the device; only T comes up and dT goes back per call.
/// With EnableRF (or through b200::RFCoupledOperator) q is the Joule heat of the RF field, re-solved at every stage.
class BioheatOperator : public mfem::TimeDependentOperator
{
public:
   struct Physics { double rc = 3.6e6, w = 4.0e4, Ta = 37.0, q = 0.0, k0 = 0.5, ak = 0.02, Tref = 37.0; };
protected:
   const mfem::FiniteElementSpace &fes;
   Physics ph;
   b200pa_space sp = nullptr;
   b200pa_form fA = nullptr, fK = nullptr, fE = nullptr;
   mutable DeviceBuffer dT, dk, drhs, dz, dlf, dkq, dkq_dt, dsrc, dcm, diag, dinv, dess;
   // RF part: sigma(T) q-data, potential, its Dirichlet data (uploaded once), eliminated RHS, Joule source q-data
   mutable DeviceBuffer dsq, dphi, dphi_bc, dBe, dsrcq, dinvE, dessE;
   bool rf = false;
   double s0 = 0.3, as = 0.015;
   int n_essE = 0;
   long long nq = 0;
   double rel_tol = 1e-8, abs_tol = 0.0, rel_tol_e = 1e-8;
   int max_iter = 500, max_iter_e = 500;
   mutable b200pa_pcg_result res{}, resE{};
   mutable int total_iters = 0, total_iters_e = 0;
public:
   BioheatOperator(const mfem::FiniteElementSpace &fes_, const Physics &p, bool factorised = false)
      : mfem::TimeDependentOperator(fes_.GetVSize(), 0.0, mfem::TimeDependentOperator::IMPLICIT), fes(fes_), ph(p)
   {
      const mfem::FiniteElement &el = *fes.GetTypicalFE();
      internal::PASetup s;
      const mfem::IntegrationRule *ir = &mfem::DiffusionIntegrator::GetRule(el, el);
      s.Init(fes, ir);
      const mfem::ElementRestriction *R = dynamic_cast<const mfem::ElementRestriction *>(
                                             fes.GetElementRestriction(mfem::ElementDofOrdering::LEXICOGRAPHIC));
      MFEM_VERIFY(R, "b200pa: ElementRestriction expected");
      Check(b200pa_space_create(Ctx(), s.d1d, s.q1d, s.ne, fes.GetNDofs(), R->GatherMap().HostRead(), s.maps->B.HostRead(),
                                s.maps->G.HostRead(), &sp));
      const mfem::GeometricFactors *geom = fes.GetMesh()->GetGeometricFactors(
                                              *ir, mfem::GeometricFactors::JACOBIANS | mfem::GeometricFactors::DETERMINANTS);
      Check(b200pa_space_set_geometry(sp, ir->GetWeights().HostRead(), geom->J.HostRead(), geom->detJ.HostRead()));
      Check(b200pa_form_create(sp, &fA));
      Check(b200pa_form_create(sp, &fK));
      Check(b200pa_form_create(sp, &fE));
      if (factorised && b200pa_space_is_affine(sp) == 1)
      {
         Check(b200pa_form_set_factorised(fA, 1));
         Check(b200pa_form_set_factorised(fK, 1));
         Check(b200pa_form_set_factorised(fE, 1));
      }
      Check(b200pa_form_set_essential(fA, 0, nullptr));
      Check(b200pa_form_set_essential(fK, 0, nullptr));
      nq = (long long)s.nq * s.ne;
      const size_t nb = sizeof(double) * height;
      for (DeviceBuffer *b : {&dT, &dk, &drhs, &dz, &dlf, &diag, &dinv}) { b->Resize(nb); }
      dkq.Resize(sizeof(double) * nq); dkq_dt.Resize(sizeof(double) * nq);
      dsrc.Resize(sizeof(double)); dcm.Resize(sizeof(double)); dess.Resize(sizeof(int));
      // constant parts: M(w) of the explicit operator and the load vector (w Ta + q, v)
      const double wv = ph.w, src = ph.w * ph.Ta + ph.q;
      Check(b200pa_form_assemble_mass(fK, &wv, 1));
      dsrc.Upload(&src, sizeof(double));
      Check(b200pa_space_domain_lf(sp, dsrc.D(), 1, dlf.D()));
   }
   ~BioheatOperator() { b200pa_form_destroy(fA); b200pa_form_destroy(fK); b200pa_form_destroy(fE); b200pa_space_destroy(sp); }
   void SetSolverOptions(double rtol, double atol, int maxit) { rel_tol = rtol; abs_tol = atol; max_iter = maxit; }

   /// The RF field that heats the tissue: at every stage  div sigma(T) grad phi = 0,  phi = phi_bc on the essential dofs
   /// (the electrodes), sigma(T) = s0 (1 + as (T - Tref)), and the heat source becomes q + sigma(T) |grad phi|^2 (Joule
   /// heating: miniapps/electromagnetics/joule_solver.cpp:898-906).  phi_bc: an L-vector holding the Dirichlet values on
   /// ess_tdofs (anything elsewhere); it is uploaded here, once - nothing crosses PCIe for phi inside the time loop.
   void EnableRF(const mfem::Array<int> &ess_tdofs, const mfem::Vector &phi_bc, double sigma0, double a_sigma, double rtol_e = 1e-8,
                 int maxit_e = 500)
   {
      rf = true; s0 = sigma0; as = a_sigma; rel_tol_e = rtol_e; max_iter_e = maxit_e;
      n_essE = ess_tdofs.Size();
      Check(b200pa_form_set_essential(fE, n_essE, ess_tdofs.HostRead()));
      dessE.Resize(sizeof(int) * std::max(n_essE, 1));
      if (n_essE) { dessE.Upload(ess_tdofs.HostRead(), sizeof(int) * n_essE); }
      const size_t nb = sizeof(double) * height;
      mfem::Vector bc(height);
      bc = 0.0;
      for (int i = 0; i < n_essE; i++) { bc[ess_tdofs[i]] = phi_bc[ess_tdofs[i]]; }
      dphi_bc.Upload(bc.HostRead(), nb);
      for (DeviceBuffer *b : {&dphi, &dBe, &dinvE}) { b->Resize(nb); }
      dsq.Resize(sizeof(double) * nq); dsrcq.Resize(sizeof(double) * nq);
   }
   void SetRFSolverOptions(double rtol_e, int maxit_e) { rel_tol_e = rtol_e; max_iter_e = maxit_e; }

   /// dT = k solving the backward-Euler stage equation at T (TimeDependentOperator::ImplicitSolve, linalg/operator.hpp:343)
   void ImplicitSolve(const mfem::real_t dt, const mfem::Vector &T, mfem::Vector &dT_dt) override
   {
      const size_t nb = sizeof(double) * height;
      dT.Upload(T.HostRead(), nb);
      if (rf)
      {
         // (1) electrostatics with sigma(T): q-data + Jacobi diagonal in one pass, EliminateRHS, PCG from the Dirichlet lift
         Check(b200pa_space_coeff_linear(sp, s0, as, ph.Tref, dT.D(), dsq.D()));
         Check(b200pa_form_assemble_diffusion_with_diagonal(fE, dsq.D(), nq, diag.D()));
         Check(b200pa_jacobi_setup(Ctx(), height, diag.D(), n_essE, dessE.I(), 1.0, dinvE.D()));
         Check(b200pa_copy(Ctx(), height, dphi_bc.D(), dphi.D()));
         Check(b200pa_memset(Ctx(), dBe.D(), 0, nb));
         Check(b200pa_form_eliminate_rhs(fE, dphi.D(), dBe.D()));
         Check(b200pa_pcg_solve(fE, dinvE.D(), dBe.D(), dphi.D(), rel_tol_e, 0.0, max_iter_e, &resE, nullptr));
         total_iters_e += resE.final_iter;
         // (2) Joule source at the q-points (grad phi never stored) and its load vector
         Check(b200pa_space_joule(sp, dphi.D(), dsq.D(), ph.w * ph.Ta + ph.q, dsrcq.D()));
         Check(b200pa_space_domain_lf(sp, dsrcq.D(), nq, dlf.D()));
      }
      // k(T) at the quadrature points, once for K and once scaled by dt for the system operator
      Check(b200pa_space_coeff_linear(sp, ph.k0, ph.ak, ph.Tref, dT.D(), dkq.D()));
      Check(b200pa_space_coeff_linear(sp, dt * ph.k0, ph.ak, ph.Tref, dT.D(), dkq_dt.D()));
      Check(b200pa_form_assemble_diffusion(fK, dkq.D(), nq));
      const double cm = ph.rc + dt * ph.w;
      Check(b200pa_form_assemble_mass(fA, &cm, 1));
      // system operator: q-data and Jacobi diagonal in one pass over the q-points
      Check(b200pa_form_assemble_diffusion_with_diagonal(fA, dkq_dt.D(), nq, diag.D()));
      // rhs = (w Ta + q, v) - [K + M(w)] T
      Check(b200pa_form_mult(fK, dT.D(), dz.D()));
      Check(b200pa_add(Ctx(), height, dlf.D(), -1.0, dz.D(), drhs.D()));
      // Jacobi-PCG from a zero initial guess
      Check(b200pa_jacobi_setup(Ctx(), height, diag.D(), 0, dess.I(), 1.0, dinv.D()));
      Check(b200pa_memset(Ctx(), dk.D(), 0, nb));
      Check(b200pa_pcg_solve(fA, dinv.D(), drhs.D(), dk.D(), rel_tol, abs_tol, max_iter, &res, nullptr));
      total_iters += res.final_iter;
      dk.Download(dT_dt.HostWrite(), nb);
   }
   /// explicit form, for completeness (rho c M)^-1 of the right-hand side is not provided: implicit solvers only
   void Mult(const mfem::Vector &, mfem::Vector &) const override { MFEM_ABORT("b200::BioheatOperator is IMPLICIT: use an implicit ODE solver"); }
   int LastIterations() const { return res.final_iter; }
   int TotalIterations() const { return total_iters; }
   bool LastConverged() const { return res.converged != 0; }
   bool Factorised() const { return b200pa_form_is_factorised(fA) == 1; }
   /// RF: the potential of the last stage (downloaded on request), iteration counts of the potential solves
   void GetPotential(mfem::Vector &phi) const { phi.SetSize(height); dphi.Download(phi.HostWrite(), sizeof(double) * height); }
   int LastPotentialIterations() const { return resE.final_iter; }
   int TotalPotentialIterations() const { return total_iters_e; }
};

/// The RF-ablation coupled problem (BASELINE configs[2]) as ONE TimeDependentOperator: every implicit stage solves the
/// electrostatic problem with sigma(T), evaluates the Joule heat sigma |grad phi|^2 at the quadrature points and solves the
/// bioheat stage with k(T) - all device-resident, stepped by the reference's ODE solvers.
class RFCoupledOperator : public BioheatOperator
{
public:
   struct RF { double s0 = 0.3, as = 0.015, rel_tol = 1e-8; int max_iter = 500; };
   RFCoupledOperator(const mfem::FiniteElementSpace &fes_, const Physics &p, const RF &r, const mfem::Array<int> &ess_phi_tdofs,
                     const mfem::Vector &phi_bc, bool factorised = false)
      : BioheatOperator(fes_, p, factorised)
   {
      EnableRF(ess_phi_tdofs, phi_bc, r.s0, r.as, r.rel_tol, r.max_iter);
   }
};

} // namespace b200

#endif
