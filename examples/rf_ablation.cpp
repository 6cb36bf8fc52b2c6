// RF ablation of a tissue slab with the reference's own driver classes and the B200 hot path behind them.
//
// What an MFEM application changes to run its Pennes-bioheat / electrostatics solves on a B200: the operator and solver
// classes come from cardiac-ablation-ecm2_b200/host/mfem_b200pa.hpp (real subclasses of mfem::TimeDependentOperator,
// mfem::Operator, mfem::IterativeSolver); the mesh, the spaces, the ODE solver, the boundary data and the output stay the
// reference's (cf. examples/ex16.cpp for the time loop, miniapps/electromagnetics/joule for the coupling).  A CPU build
// of MFEM is enough - device memory lives behind the C ABI (include/b200pa.h).
//
//   g++ -O2 -std=c++17 -I$MFEM_DIR examples/rf_ablation.cpp -o rf_ablation -L$MFEM_DIR -lmfem \
//       -I. -Lcardiac-ablation-ecm2_b200 -lb200pa -Wl,-rpath,$PWD/cardiac-ablation-ecm2_b200
//   ./rf_ablation -n 32 -o 2 -dt 0.5 -tf 10 -V 30 -pv
#include "mfem.hpp"
#include "cardiac-ablation-ecm2_b200/host/mfem_b200pa.hpp"

#include <iostream>

using namespace mfem;

int main(int argc, char *argv[])
{
   int n = 16, order = 2, vis_steps = 5;
   double dt = 0.5, t_final = 5.0, V = 30.0;
   bool paraview = false, factorised = true;
   OptionsParser args(argc, argv);
   args.AddOption(&n, "-n", "--elements", "Elements per direction of the slab.");
   args.AddOption(&order, "-o", "--order", "H1 order (1-6).");
   args.AddOption(&dt, "-dt", "--time-step", "Time step [s].");
   args.AddOption(&t_final, "-tf", "--t-final", "Final time [s].");
   args.AddOption(&V, "-V", "--voltage", "Electrode potential on the face z = 0 [V] (z = 1 is grounded).");
   args.AddOption(&vis_steps, "-vs", "--visualization-steps", "Save every n-th step.");
   args.AddOption(&paraview, "-pv", "--paraview", "-no-pv", "--no-paraview", "ParaView data collection output.");
   args.AddOption(&factorised, "-fq", "--factorised-qdata", "-sq", "--stored-qdata",
                  "Keep the diffusion q-data factorised (affine meshes) or as the reference stores it.");
   args.Parse();
   if (!args.Good()) { args.PrintUsage(std::cout); return 1; }
   args.PrintOptions(std::cout);

   // tissue slab 4 x 4 x 2 cm, hexahedra
   Mesh mesh = Mesh::MakeCartesian3D(n, n, n, Element::HEXAHEDRON, 0.04, 0.04, 0.02);
   H1_FECollection fec(order, 3);
   FiniteElementSpace fes(&mesh, &fec);
   std::cout << "unknowns: " << fes.GetTrueVSize() << std::endl;

   // electrodes: attributes 1 (z = 0) and 6 (z = top) of MakeCartesian3D carry the Dirichlet data of the potential
   Array<int> ess_bdr(mesh.bdr_attributes.Max());
   ess_bdr = 0; ess_bdr[0] = 1; ess_bdr[5] = 1;
   Array<int> ess_phi;
   fes.GetEssentialTrueDofs(ess_bdr, ess_phi);
   GridFunction phi(&fes);
   FunctionCoefficient phi_bc([V](const Vector &x) { return V * (1.0 - x(2) / 0.02); });
   phi.ProjectCoefficient(phi_bc);

   GridFunction T(&fes);
   ConstantCoefficient T_body(37.0);
   T.ProjectCoefficient(T_body);

   // Pennes bioheat: rho c dT/dt = div k(T) grad T - w (T - Ta) + sigma(T) |grad phi|^2,  div sigma(T) grad phi = 0
   b200::BioheatOperator::Physics tissue;            // rho c, perfusion w, arterial Ta, k(T) = k0 (1 + ak (T - Tref))
   b200::RFCoupledOperator::RF rf;                    // sigma(T) = s0 (1 + as (T - Tref)), tolerance of the potential solve
   b200::RFCoupledOperator oper(fes, tissue, rf, ess_phi, phi, factorised);
   oper.SetSolverOptions(1e-8, 0.0, 500);

   BackwardEulerSolver ode;                          // or SDIRK23Solver, SDIRK33Solver ... (linalg/ode.hpp)
   ode.Init(oper);

   ParaViewDataCollection pv("rf_ablation", &mesh);
   if (paraview)
   {
      pv.SetLevelsOfDetail(order);
      pv.SetHighOrderOutput(true);
      pv.SetDataFormat(VTKFormat::BINARY);
      pv.RegisterField("temperature", &T);
      pv.RegisterField("potential", &phi);
   }

   double t = 0.0;
   for (int step = 1; t < t_final - 1e-12 * dt; step++)
   {
      double dt_real = std::min(dt, t_final - t);
      ode.Step(T, t, dt_real);                        // one electrostatic + one bioheat solve on the GPU
      if (step % vis_steps == 0 || t >= t_final - 1e-12 * dt)
      {
         std::cout << "step " << step << ", t = " << t << " s: max T = " << T.Max() << " C, PCG iterations phi / T: "
                   << oper.LastPotentialIterations() << " / " << oper.LastIterations() << std::endl;
         if (paraview)
         {
            oper.GetPotential(phi);
            pv.SetCycle(step); pv.SetTime(t); pv.Save();
         }
      }
   }

   // a stationary solve through the Operator / IterativeSolver surface: the potential for the final temperature field
   {
      GridFunctionCoefficient Tc(&T);
      TransformedCoefficient sigma(&Tc, [](double Tq) { return 0.3 * (1.0 + 0.015 * (Tq - 37.0)); });
      b200::PAOperator A(fes, &sigma, nullptr, ess_phi);
      Vector x(fes.GetVSize()), B(fes.GetVSize());
      phi.ProjectCoefficient(phi_bc);
      x = 0.0; B = 0.0;
      for (int i = 0; i < ess_phi.Size(); i++) { x(ess_phi[i]) = phi(ess_phi[i]); }   // Dirichlet lift, zero elsewhere
      A.EliminateRHS(x, B);
      b200::ChebyshevSmoother prec(3);
      b200::PCGSolver cg;
      cg.SetRelTol(1e-10); cg.SetMaxIter(2000); cg.SetPrintLevel(3);
      cg.SetPreconditioner(prec);
      cg.SetOperator(A);
      cg.Mult(B, x);
      phi = x;
   }
   return 0;
}
