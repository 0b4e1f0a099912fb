// The reference's usage pattern (tests_cpp/eigen_2d_euler_riemann_explicit/main.cc) against the B200 engine through
// include/pda_b200_eigen.hpp: load a mesh directory, create a problem, initial condition, Jacobian with the fixed
// pattern; evaluate when a GPU is present.  argv[1] = mesh directory written by create_full_mesh.py / pda_mesh_write.
#include <cmath>
#include <cstdio>

#include "pda_b200_eigen.hpp"

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
  std::printf("cpp_shim_demo ok\n");
  return 0;
}
