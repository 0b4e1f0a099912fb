// The reference's usage pattern (tests_cpp/eigen_2d_euler_riemann_explicit/main.cc) against the B200 engine through
// include/pda_b200_eigen.hpp: load a mesh directory, create a problem, initial condition, Jacobian with the fixed
// pattern; evaluate when a GPU is present.  argv[1] = mesh directory written by create_full_mesh.py / pda_mesh_write.
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
  std::printf("cpp_shim_demo ok\n");
  return 0;
}
