// The reference's usage pattern (tests_cpp/eigen_2d_euler_riemann_explicit/main.cc) against the B200 engine through
// include/pda_b200_eigen.hpp: load a mesh directory, create a problem, initial condition, Jacobian with the fixed
// pattern; evaluate when a GPU is present.  argv[1] = mesh directory written by create_full_mesh.py / pda_mesh_write.
#include <cmath>
#include <cstdio>
#include <filesystem>
#include <string>

#include "pda_b200_eigen.hpp"

// the two boundary-condition functors of the reference's tests_cpp/eigen_2d_swe_custom_bcs/main.cc (6-58), argument
// for argument: ghost fill and Jacobian-factor call operators
struct Dirichlet {
  template <class ConnecRowType, class StateT, class T>
  void operator()(const int, ConnecRowType const&, const double, const double, const StateT&, int numDofPerCell,
                  const double, T& ghostValues) const {
    if (numDofPerCell != 3) return;
    ghostValues(0) = 0.00001; ghostValues(1) = 0.004; ghostValues(2) = 0.001;
  }
  template <class ConnecRowType, class FactorsType>
  void operator()(ConnecRowType const&, const double, const double, int, FactorsType& factorsForBCJac) const {
    factorsForBCJac = {0., 0., 0.};
  }
};
struct HomogNeumann {
  template <class ConnecRowType, class StateT, class T>
  void operator()(const int, ConnecRowType const& connectivityRow, const double, const double, const StateT& currentState,
                  int numDofPerCell, const double, T& ghostValues) const {
    const int cellGID = connectivityRow[0];
    const auto uIndex = cellGID * numDofPerCell;
    ghostValues[0] = currentState(uIndex); ghostValues[1] = currentState(uIndex + 1); ghostValues[2] = currentState(uIndex + 2);
  }
  template <class ConnecRowType, class FactorsType>
  void operator()(ConnecRowType const&, const double, const double, int, FactorsType& factorsForBCJac) const {
    factorsForBCJac = {1., 1., 1.};
  }
};

int main(int argc, char** argv) {
  namespace pda = pressiodemoapps_b200;
  if (argc < 2) return 2;
  const auto meshObj = pda::load_cellcentered_uniform_mesh_eigen(argv[1]);
  auto appObj = pda::create_problem_eigen(meshObj, pda::Euler2d::Riemann, pda::InviscidFluxReconstruction::Weno5, 2);
  using app_t = decltype(appObj);
  typename app_t::state_type state = appObj.initialCondition();
  auto V = appObj.createRightHandSide();
  auto J = appObj.createJacobian();
  std::printf("dofs %d nnz %lld gamma %.2f dx %.4f\n", (int)state.size(), (long long)J.nonZeros(), appObj.gamma(), meshObj.dx());
  if ((int)state.size() != appObj.totalDofStencilMesh() || J.rows() != V.size()) return 3;
  if (pda_device_count() > 0) {
    appObj.rightHandSideAndJacobian(state, 0.0, V, J);
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> B =
        Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>::Random(state.size(), 3);
    auto R = appObj.createApplyJacobianResult(B);
    appObj.applyJacobian(state, B, 0.0, R);
    const double err = (J * B - R).cwiseAbs().maxCoeff();
    std::printf("|V|max %.6e  |J*B - applyJacobian|max %.3e\n", V.cwiseAbs().maxCoeff(), err);
    if (!(err < 1e-9)) return 4;
    // reference-order mode (the reference's formulas and operation order): a second evaluation through the same
    // surface; the values may differ from the fast kernels only at the conditioning level of the WENO gradients
    appObj.setOption("order", "reference");
    if (appObj.getOption("jacobian_order") != "reference") return 7;
    auto V2 = appObj.createRightHandSide();
    auto J2 = appObj.createJacobian();
    appObj.rightHandSideAndJacobian(state, 0.0, V2, J2);
    const double dv = (V2 - V).cwiseAbs().maxCoeff();
    std::printf("reference-order vs fast: |dV|max %.3e\n", dv);
    if (!(dv < 1e-9 * (1.0 + V.cwiseAbs().maxCoeff()))) return 8;
    appObj.setOption("order", "fast");
  } else {
    try { appObj.rightHandSide(state, 0.0, V); return 5; }   // must refuse: no CPU fallback
    catch (const std::runtime_error& e) { std::printf("no device: %s\n", e.what()); }
  }
  // boundary-face gradients, written like tests_cpp/gradients/main.cc:24-97
  {
    const auto x = meshObj.viewX();
    const auto y = meshObj.viewY();
    Eigen::VectorXd f(meshObj.stencilMeshSize());
    for (int i = 0; i < x.size(); ++i) f(i) = std::sin(M_PI * x(i) * y(i));
    pda::GradientEvaluator<pda::Mesh> grads(meshObj);
    const auto G = meshObj.graph();
    int count = 0;
    double err = 0.0;
    if (pda_device_count() > 0) {
      grads(f);
      for (auto rowInd : meshObj.graphRowsOfCellsStrictlyOnBd()) {
        if (!meshObj.cellHasLeftFaceOnBoundary2d(rowInd)) continue;
        const auto& face = grads.queryFace(G(rowInd, 0), pda::FacePosition::Left);
        const double gold = face.centerCoordinates[1] * M_PI * std::cos(M_PI * face.centerCoordinates[0] * face.centerCoordinates[1]);
        err += (face.normalGradient - gold) * (face.normalGradient - gold);
        ++count;
      }
      std::printf("left-wall faces %d, rmse of the one-sided normal gradient %.3e\n", count, std::sqrt(err / std::max(count, 1)));
      if (count == 0 || !(std::sqrt(err / count) < 0.5)) return 6;
    } else {
      try { grads(f); return 7; }
      catch (const std::runtime_error& e) { std::printf("gradients, no device: %s\n", e.what()); }
    }
  }
  // custom boundary conditions through host functors (create_problem_eigen with four functors, swe2d.hpp:187-281)
  // against the same rules as device tables: identical velocity and Jacobian
  {
    const auto fo = pda::InviscidFluxReconstruction::FirstOrder;
    auto sweF = pda::create_problem_eigen(meshObj, pda::Swe2d::CustomBCs, fo, Dirichlet(), Dirichlet(), HomogNeumann(), HomogNeumann());
    auto sweD = pda::create_problem_eigen(meshObj, pda::Swe2d::CustomBCs, fo);
    const double dirich[3] = {0.00001, 0.004, 0.001};
    for (int side = 0; side < 4; ++side)
      if (pda_problem_set_bc(sweD.handle(), side, side < 2 ? PDA_BC_DIRICHLET : PDA_BC_HOMOG_NEUMANN, dirich) != PDA_OK) return 8;
    Eigen::VectorXd u = sweF.initialCondition();
    for (int i = 0; i < u.size(); ++i) u(i) += 0.5;   // like verifyJacobian, main.cc:66-69
    auto vF = sweF.createRightHandSide(), vD = sweD.createRightHandSide();
    auto jF = sweF.createJacobian(), jD = sweD.createJacobian();
    if (pda_device_count() > 0) {
      sweF.rightHandSideAndJacobian(u, 0.0, vF, jF);
      sweD.rightHandSideAndJacobian(u, 0.0, vD, jD);
      const double dv = (vF - vD).cwiseAbs().maxCoeff();
      const double dj = (Eigen::VectorXd::Map(jF.valuePtr(), jF.nonZeros()) - Eigen::VectorXd::Map(jD.valuePtr(), jD.nonZeros())).cwiseAbs().maxCoeff();
      // central finite-difference check of J*a, like the second-order check of the reference test (main.cc:97-104)
      Eigen::VectorXd a = Eigen::VectorXd::Random(u.size());
      const double eps = 1e-6;
      Eigen::VectorXd u2 = u + eps * a, u3 = u - eps * a, v2 = vF, v3 = vF;
      sweF.rightHandSide(u2, 0.0, v2);
      sweF.rightHandSide(u3, 0.0, v3);
      const double fd = ((v2 - v3) / (2.0 * eps) - jF * a).cwiseAbs().maxCoeff();
      std::printf("custom BC functors vs device rules: |dV|max %.1e |dJ|max %.1e ; |J a - FD|max %.2e\n", dv, dj, fd);
      if (dv != 0.0 || dj != 0.0 || !(fd < 1e-5)) return 9;
    } else {
      try { sweF.rightHandSide(u, 0.0, vF); return 10; }
      catch (const std::runtime_error& e) { std::printf("custom BC functors, no device: %s\n", e.what()); }
    }
  }
  // the named factories and remaining overloads of the reference's public headers (host side only)
  {
    const auto w3 = pda::InviscidFluxReconstruction::Weno3;
    const auto visc = pda::ViscousFluxReconstruction::FirstOrder;
    // Gray-Scott wants a periodic 3-point-stencil mesh (diffusion_reaction_2d_prob_class.hpp): written next to argv[1]
    const std::string dir3 = std::string(argv[1]) + "_s3";
    {
      const int32_t n3[3] = {16, 16, 1}, per3[3] = {1, 1, 0};
      const double b3[6] = {-1.25, 1.25, -1.25, 1.25, 0.0, 0.0};
      pda_mesh h3 = nullptr;
      std::filesystem::create_directories(dir3);
      if (pda_mesh_make_lattice(2, n3, b3, per3, 3, &h3) != PDA_OK || pda_mesh_write(h3, dir3.c_str()) != PDA_OK) return 15;
      pda_mesh_free(h3);
    }
    const auto mesh3 = pda::load_cellcentered_uniform_mesh_eigen(dir3);
    auto gs = pda::create_gray_scott_2d_problem_eigen(mesh3, visc, 2e-4, 5e-5, 0.042, 0.062);   // the defaults
    auto gs0 = pda::create_problem_eigen(mesh3, pda::DiffusionReaction2d::GrayScott);
    auto bur = pda::create_problem_eigen(meshObj, pda::AdvectionDiffusion2d::BurgersOutflow, w3, visc);
    auto slip = pda::create_slip_wall_swe_2d_problem_eigen(meshObj, w3, 9.8, -3.0, 0.125);
    auto cs = pda::create_cross_shock_problem_eigen(meshObj, w3, 0.2, 9.0, 1.5);
    auto cs0 = pda::create_cross_shock_problem_eigen(meshObj, w3);
    if (gs.numDofPerCell() != 2 || gs0.numDofPerCell() != 2 || bur.numDofPerCell() != 2) return 11;
    if ((gs.initialCondition() - gs0.initialCondition()).cwiseAbs().maxCoeff() != 0.0) return 12;
    if (slip.gravity() != 9.8 || slip.coriolis() != -3.0) return 13;
    if (cs.queryParameter("crossShockDensity") != 0.2 || cs0.queryParameter("crossShockDensity") != 0.1) return 14;
    // a 1D mesh written next to the 2D one
    const std::string dir1 = std::string(argv[1]) + "_1d";
    const int32_t n1[3] = {50, 1, 1}, per1[3] = {1, 0, 0};
    const double b1[6] = {-1.0, 1.0, 0.0, 0.0, 0.0, 0.0};
    pda_mesh h1 = nullptr;
    std::filesystem::create_directories(dir1);
    if (pda_mesh_make_lattice(1, n1, b1, per1, 3, &h1) != PDA_OK || pda_mesh_write(h1, dir1.c_str()) != PDA_OK) return 15;
    pda_mesh_free(h1);
    const auto mesh1 = pda::load_cellcentered_uniform_mesh_eigen(dir1);
    const auto fo1 = pda::InviscidFluxReconstruction::FirstOrder;   // 3-point-stencil mesh (DiffusionReaction1d needs it)
    auto adv = pda::create_linear_advection_1d_problem_eigen(mesh1, fo1, 2.5);
    auto adv2 = pda::create_linear_advection_1d_problem_eigen(mesh1, fo1, pda::InviscidFluxScheme::Rusanov, 2.5);
    auto dr1 = pda::create_problem_eigen(mesh1, pda::DiffusionReaction1d::ProblemA);
    if (adv.totalDofStencilMesh() != 50 || adv2.totalDofStencilMesh() != 50 || dr1.numDofPerCell() != 1) return 16;
    std::printf("named factories ok\n");
  }
  std::printf("cpp_shim_demo ok\n");
  return 0;
}
