// pda_b200_eigen.hpp -- the C++ shim of INTEGRATION.md as a real header: the public surface of the reference's
// PublicProblemEigenMixinCpp (include/pressiodemoapps/adapter_cpp.hpp:59-264) and of load_cellcentered_uniform_mesh_eigen
// (mesh.hpp:87-91) on top of the C-ABI include/pda_b200.h.  Needs Eigen 3.4 on the include path (the reference
// vendors it under tpls/eigen3).  A user of the reference switches namespace `pressiodemoapps` -> `pressiodemoapps_b200`.
#pragma once
#include <Eigen/Core>
#include <Eigen/Sparse>

#include <algorithm>
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "pda_b200.h"

namespace pressiodemoapps_b200 {

// enumerator order = the reference's (euler1d.hpp:64-69, euler2d.hpp:66-76, euler3d.hpp:65-68, swe2d.hpp:65-68, ...)
enum class InviscidFluxReconstruction { FirstOrder = 0, Weno3 = 1, Weno5 = 2 };
enum class InviscidFluxScheme { Rusanov = 0 };            // schemes_info.hpp
enum class ViscousFluxReconstruction { FirstOrder = 0 };
enum class ViscousFluxScheme { Central = 0 };
enum class Euler1d { PeriodicSmooth = 0, Sod, Lax, ShuOsher };
enum class Euler2d { PeriodicSmooth = 0, KelvinHelmholtz, SedovFull, SedovSymmetry, Riemann, NormalShock,
                     DoubleMachReflection, CrossShock, testingonlyneumann };
enum class Euler3d { PeriodicSmooth = 0, SedovSymmetry };
enum class Swe2d { SlipWall = 0, CustomBCs };
enum class DiffusionReaction2d { ProblemA = 0, GrayScott };
enum class AdvectionDiffusion2d { BurgersPeriodic = 0, BurgersOutflow };
enum class AdvectionDiffusionReaction2d { ProblemA = 0 };
enum class Advection1d { PeriodicLinear = 0 };
enum class DiffusionReaction1d { ProblemA = 0 };
enum class FacePosition { Left = 0, Front, Right, Back, Bottom, Top };   // schemes_info.hpp:112-114

namespace impl {
inline void check(pda_status s) { if (s != PDA_OK) throw std::runtime_error(pda_last_error()); }
template <class E> struct family_of;
template <> struct family_of<Euler1d> { static constexpr int value = PDA_FAMILY_EULER1D; };
template <> struct family_of<Euler2d> { static constexpr int value = PDA_FAMILY_EULER2D; };
template <> struct family_of<Euler3d> { static constexpr int value = PDA_FAMILY_EULER3D; };
template <> struct family_of<Swe2d> { static constexpr int value = PDA_FAMILY_SWE2D; };
template <> struct family_of<DiffusionReaction2d> { static constexpr int value = PDA_FAMILY_DIFFUSION_REACTION2D; };
template <> struct family_of<AdvectionDiffusion2d> { static constexpr int value = PDA_FAMILY_ADVECTION_DIFFUSION2D; };
template <> struct family_of<AdvectionDiffusionReaction2d> { static constexpr int value = PDA_FAMILY_ADVECTION_DIFFUSION_REACTION2D; };
template <> struct family_of<Advection1d> { static constexpr int value = PDA_FAMILY_ADVECTION1D; };
template <> struct family_of<DiffusionReaction1d> { static constexpr int value = PDA_FAMILY_DIFFUSION_REACTION1D; };
// dofs per cell of the 2D families that accept custom boundary-condition functors (compile-time, like the reference's
// std::array<scalar_type, numDofPerCell> Jacobian factors: swe_2d_prob_class.hpp:1054, euler_2d_prob_class.hpp)
template <class E> struct custom_bc_ndpc;
template <> struct custom_bc_ndpc<Swe2d> { static constexpr int value = 3; };
template <> struct custom_bc_ndpc<Euler2d> { static constexpr int value = 4; };
template <> struct custom_bc_ndpc<AdvectionDiffusion2d> { static constexpr int value = 2; };

// A user functor with the reference's two call operators (custom_bcs_functions.hpp:60-164; e.g. the Dirichlet /
// HomogNeumann structs of tests_cpp/eigen_2d_swe_custom_bcs/main.cc:6-58) behind the C callbacks of
// pda_problem_set_bc_callback: Eigen maps stand in for the reference's graph row, state and ghost-row arguments.
struct BcFunctorBase {
  virtual ~BcFunctorBase() = default;
  int ncols = 0, ghostLen = 0;
  int32_t nState = 0;
};
template <class F, int NDPC>
struct BcFunctorHolder final : BcFunctorBase {
  F f;
  explicit BcFunctorHolder(F ff) : f(std::move(ff)) {}
  using conn_t = Eigen::Map<const Eigen::Matrix<int32_t, 1, Eigen::Dynamic>>;
  static void ghost(void* user, int32_t nearBdRow, const int32_t* graphRow, double x, double y, const double* U, int ndpc,
                    double cellWidth, double* ghostValues) {
    auto* self = static_cast<BcFunctorHolder*>(user);
    conn_t conn(graphRow, self->ncols);
    Eigen::Map<const Eigen::VectorXd> state(U, self->nState);
    Eigen::Map<Eigen::Matrix<double, 1, Eigen::Dynamic>> gv(ghostValues, self->ghostLen);
    self->f(static_cast<int>(nearBdRow), conn, x, y, state, ndpc, cellWidth, gv);
  }
  static void factors(void* user, const int32_t* graphRow, double x, double y, int ndpc, double* out) {
    auto* self = static_cast<BcFunctorHolder*>(user);
    conn_t conn(graphRow, self->ncols);
    std::array<double, NDPC> fac;
    for (int d = 0; d < NDPC; ++d) fac[(size_t)d] = out[d];
    self->f(conn, x, y, ndpc, fac);
    for (int d = 0; d < NDPC; ++d) out[d] = fac[(size_t)d];
  }
};
}  // namespace impl

// CellCenteredUniformMesh look-alike (impl/mesh_ccu.hpp:115-159)
class Mesh {
 public:
  using scalar_type = double;
  using index_t = int32_t;
  using graph_t = Eigen::Matrix<int32_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;
  explicit Mesh(const std::string& dir) { impl::check(pda_mesh_load(dir.c_str(), &h_)); }
  // slab window of a full lattice for rank `rank` of `nranks` (multi-GPU sharding, pda_mesh_make_slab_window): owned
  // planes plus halo planes where a neighbour rank exists; an ordinary mesh for every problem factory
  static Mesh slabWindow(const Mesh& full, int rank, int nranks) {
    Mesh w;
    impl::check(pda_mesh_make_slab_window(full.handle(), rank, nranks, &w.h_));
    return w;
  }
  // {plane_cells, k0, k1, halo_planes_below, halo_planes_above, rank, nranks, dim}
  std::array<int64_t, 8> slabWindowInfo() const {
    std::array<int64_t, 8> info{};
    impl::check(pda_mesh_slab_window_info(h_, info.data()));
    return info;
  }
  Mesh(const Mesh&) = delete;
  Mesh& operator=(const Mesh&) = delete;
  Mesh(Mesh&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  ~Mesh() { if (h_) pda_mesh_free(h_); }
  int dimensionality() const { return pda_mesh_dimensionality(h_); }
  int stencilSize() const { return pda_mesh_stencil_size(h_); }
  index_t stencilMeshSize() const { return pda_mesh_stencil_mesh_size(h_); }
  index_t sampleMeshSize() const { return pda_mesh_sample_mesh_size(h_); }
  index_t numCellsInner() const { return pda_mesh_num_cells_inner(h_); }
  index_t numCellsNearBd() const { return pda_mesh_num_cells_near_bd(h_); }
  bool isFullyPeriodic() const { return pda_mesh_is_fully_periodic(h_) != 0; }
  double dx() const { return deltas_(0); }
  double dy() const { return deltas_(1); }
  double dz() const { return deltas_(2); }
  double dxInv() const { return deltas_(3); }
  double dyInv() const { return deltas_(4); }
  double dzInv() const { return deltas_(5); }
  graph_t graph() const {
    graph_t g(sampleMeshSize(), pda_mesh_graph_cols(h_));
    impl::check(pda_mesh_graph(h_, g.data()));
    return g;
  }
  Eigen::VectorXd viewX() const { return coord_(0); }
  Eigen::VectorXd viewY() const { return coord_(1); }
  Eigen::VectorXd viewZ() const { return coord_(2); }
  std::vector<index_t> graphRowsOfCellsAwayFromBd() const {
    std::vector<index_t> r((size_t)numCellsInner());
    if (!r.empty()) impl::check(pda_mesh_rows_inner(h_, r.data()));
    return r;
  }
  std::vector<index_t> graphRowsOfCellsNearBd() const {
    std::vector<index_t> r((size_t)numCellsNearBd());
    if (!r.empty()) impl::check(pda_mesh_rows_near_bd(h_, r.data()));
    return r;
  }
  // impl/mesh_ccu.hpp:153-155 and 298-326 (2D)
  std::vector<index_t> graphRowsOfCellsStrictlyOnBd() const {
    std::vector<index_t> r((size_t)std::max<index_t>(0, pda_mesh_num_cells_strictly_on_bd(h_)));
    if (!r.empty()) impl::check(pda_mesh_rows_strictly_on_bd(h_, r.data()));
    return r;
  }
  bool cellHasLeftFaceOnBoundary2d(index_t rowInd) const { return firstLayer_(rowInd, 1) == -1; }
  bool cellHasFrontFaceOnBoundary2d(index_t rowInd) const { return firstLayer_(rowInd, 2) == -1; }
  bool cellHasRightFaceOnBoundary2d(index_t rowInd) const { return firstLayer_(rowInd, 3) == -1; }
  bool cellHasBackFaceOnBoundary2d(index_t rowInd) const { return firstLayer_(rowInd, 4) == -1; }
  pda_mesh handle() const { return h_; }

 private:
  Mesh() = default;   // slabWindow()
  index_t firstLayer_(index_t rowInd, int col) const {
    if (graphCache_.size() == 0) graphCache_ = graph();
    return graphCache_(rowInd, col);
  }
  mutable graph_t graphCache_;
  double deltas_(int i) const {
    double d[3], di[3];
    impl::check(pda_mesh_deltas(h_, d, di));
    return i < 3 ? d[i] : di[i - 3];
  }
  Eigen::VectorXd coord_(int axis) const {
    const auto n = stencilMeshSize();
    Eigen::VectorXd x(n), y(n), z(n);
    impl::check(pda_mesh_coordinates(h_, x.data(), y.data(), z.data()));
    return axis == 0 ? x : (axis == 1 ? y : z);
  }
  pda_mesh h_ = nullptr;
};

inline Mesh load_cellcentered_uniform_mesh_eigen(const std::string& dir) { return Mesh(dir); }

// GradientEvaluator<MeshType, MaxNumDofPerCell> look-alike (gradient.hpp:61-121): normal gradients at the faces on the
// domain boundary of a 2D mesh, computed on the GPU by pda_gradient_compute_host.
template <class MeshType = Mesh, std::size_t MaxNumDofPerCell = 1>
class GradientEvaluator {
 public:
  struct Face {   // the exposition-only struct of gradient.hpp:95-113
    std::array<double, 3> centerCoordinates = {};
    std::conditional_t<MaxNumDofPerCell == 1, double, std::array<double, MaxNumDofPerCell>> normalGradient = {};
    int normalDirection = {};   // 1 = x, 2 = y
  };
  explicit GradientEvaluator(const MeshType& mesh) {
    impl::check(pda_gradient_create(mesh.handle(), (int)MaxNumDofPerCell, &h_));
    const auto n = (size_t)pda_gradient_num_faces(h_);
    faces_.resize(n);
    std::vector<int32_t> dir(n);
    std::vector<double> cen(3 * n);
    impl::check(pda_gradient_faces(h_, nullptr, nullptr, nullptr, dir.data(), cen.data()));
    for (size_t k = 0; k < n; ++k) {
      faces_[k].centerCoordinates = {cen[3 * k], cen[3 * k + 1], cen[3 * k + 2]};
      faces_[k].normalDirection = dir[k];
    }
  }
  GradientEvaluator(const GradientEvaluator&) = delete;
  GradientEvaluator& operator=(const GradientEvaluator&) = delete;
  ~GradientEvaluator() { if (h_) pda_gradient_free(h_); }
  template <class FieldType> void operator()(const FieldType& field) { compute_(field, 1); }
  template <class FieldType> void operator()(const FieldType& field, int numDofPerCell) { compute_(field, numDofPerCell); }
  const Face& queryFace(int32_t cellGID, FacePosition fp) const {
    int32_t k = -1;
    impl::check(pda_gradient_query_face(h_, cellGID, static_cast<int>(fp), &k));
    return faces_[(size_t)k];
  }

 private:
  template <class FieldType> void compute_(const FieldType& field, int nd) {
    std::vector<double> g(faces_.size() * (size_t)std::max(nd, 1));
    impl::check(pda_gradient_compute_host(h_, field.data(), nd, g.data()));   // refuses nd > MaxNumDofPerCell
    for (size_t k = 0; k < faces_.size(); ++k) {
      if constexpr (MaxNumDofPerCell == 1) faces_[k].normalGradient = g[k];
      else for (int j = 0; j < nd; ++j) faces_[k].normalGradient[(size_t)j] = g[k * (size_t)nd + (size_t)j];
    }
  }
  pda_gradient h_ = nullptr;
  std::vector<Face> faces_;
};

// PublicProblemEigenMixinCpp look-alike (adapter_cpp.hpp:59-264).  The mesh must outlive the problem, like in the
// reference (std::reference_wrapper, euler_2d_prob_class.hpp:1277).
class Problem {
 public:
  using scalar_type = double;
  using independent_variable_type = double;
  using state_type = Eigen::VectorXd;
  using rhs_type = Eigen::VectorXd;
  using right_hand_side_type = Eigen::VectorXd;
  using jacobian_type = Eigen::SparseMatrix<double, Eigen::RowMajor, int32_t>;

  Problem(const Mesh& m, int family, int id, int recon, int icFlag,
          const std::unordered_map<std::string, double>& params, int device) {
    std::vector<const char*> names;
    std::vector<double> vals;
    for (auto& kv : params) { names.push_back(kv.first.c_str()); vals.push_back(kv.second); }
    impl::check(pda_problem_create(m.handle(), family, id, recon, icFlag, (int)names.size(),
                                   names.empty() ? nullptr : names.data(), vals.empty() ? nullptr : vals.data(), device, &h_));
  }
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;
  Problem(Problem&& o) noexcept : h_(o.h_) { o.h_ = nullptr; for (int i = 0; i < 4; ++i) bc_[i] = std::move(o.bc_[i]); }
  ~Problem() { if (h_) pda_problem_free(h_); }

  int numDofPerCell() const { return pda_problem_num_dof_per_cell(h_); }
  int32_t totalDofSampleMesh() const { return pda_problem_total_dof_sample_mesh(h_); }
  int32_t totalDofStencilMesh() const { return pda_problem_total_dof_stencil_mesh(h_); }
  double queryParameter(const std::string& name) const {
    double v;
    impl::check(pda_problem_query_parameter(h_, name.c_str(), &v));
    return v;
  }
  // evaluation mode (an extension of the reference's surface, include/pda_b200.h pda_problem_set_option):
  // setOption("order", "reference") -> the reference's formulas, operation and accumulation order (values bit-identical
  // to the CPU reference); "fast" (default) -> the structured kernels
  void setOption(const std::string& name, const std::string& value) { impl::check(pda_problem_set_option(h_, name.c_str(), value.c_str())); }
  std::string getOption(const std::string& name) const {
    char buf[64] = {0};
    impl::check(pda_problem_get_option(h_, name.c_str(), buf, (int)sizeof(buf)));
    return std::string(buf);
  }
  double gamma() const { return queryParameter("gamma"); }
  double gravity() const { return queryParameter("gravity"); }
  double coriolis() const { return queryParameter("coriolis"); }

  state_type initialCondition() const {
    state_type u(totalDofStencilMesh());
    impl::check(pda_problem_initial_condition(h_, u.data()));
    return u;
  }
  state_type createState() const { return state_type::Zero(totalDofStencilMesh()); }
  rhs_type createRightHandSide() const { return rhs_type::Zero(totalDofSampleMesh()); }
  rhs_type createRhs() const { return createRightHandSide(); }
  jacobian_type createJacobian() const {   // fixed pattern, explicit zeros kept, like the reference
    int64_t nnz = 0;
    impl::check(pda_problem_jacobian_nnz(h_, &nnz));
    jacobian_type J(totalDofSampleMesh(), totalDofStencilMesh());
    J.resizeNonZeros(nnz);
    impl::check(pda_problem_jacobian_pattern(h_, J.outerIndexPtr(), J.innerIndexPtr()));
    std::fill(J.valuePtr(), J.valuePtr() + nnz, 0.0);
    return J;
  }
  template <class Operand>
  Operand createApplyJacobianResult(const Operand& B) const { return Operand::Zero(totalDofSampleMesh(), B.cols()); }
  template <class Operand>
  Operand createResultOfJacobianActionOn(const Operand& B) const { return createApplyJacobianResult(B); }

  void rightHandSide(const state_type& U, double t, rhs_type& V) const {
    impl::check(pda_problem_velocity_host(h_, U.data(), t, V.data()));
  }
  void rhs(const state_type& U, double t, rhs_type& V) const { rightHandSide(U, t, V); }
  void operator()(const state_type& U, double t, rhs_type& V) const { rightHandSide(U, t, V); }
  void rightHandSideAndJacobian(const state_type& U, double t, rhs_type& V, jacobian_type& J) const {
    impl::check(pda_problem_velocity_and_jacobian_host(h_, U.data(), t, V.data(), J.valuePtr()));
  }
  void operator()(const state_type& U, double t, rhs_type& V, jacobian_type& J, bool computeJac) const {
    if (computeJac) rightHandSideAndJacobian(U, t, V, J); else rightHandSide(U, t, V);
  }
  void jacobian(const state_type& U, double t, jacobian_type& J) const {
    impl::check(pda_problem_velocity_and_jacobian_host(h_, U.data(), t, nullptr, J.valuePtr()));
  }
  template <class Operand>   // Eigen vector or matrix (col- or row-major)
  void applyJacobian(const state_type& U, const Operand& B, double t, Operand& R) const {
    impl::check(pda_problem_apply_jacobian_host(h_, U.data(), B.data(), (int)B.cols(),
                                                Operand::IsRowMajor ? PDA_LAYOUT_ROW_MAJOR : PDA_LAYOUT_COL_MAJOR, t, R.data()));
  }
  // explicit time stepping with the state resident in HBM (no reference counterpart: its tests borrow pressio's steppers)
  void advance(int stepper, state_type& U, double t0, double dt, int32_t nsteps) const {
    impl::check(pda_problem_advance_host(h_, stepper, U.data(), t0, dt, nsteps));
  }
  pda_problem handle() const { return h_; }

  // custom boundary conditions: one functor per side (0 Left, 1 Front, 2 Right, 3 Back), kept alive by the problem
  template <int NDPC, class F>
  void installBcFunctor(int side, const Mesh& m, int recon, F f) {
    auto holder = std::make_shared<impl::BcFunctorHolder<F, NDPC>>(std::move(f));
    holder->ncols = (m.stencilSize() - 1) * m.dimensionality() + 1;
    holder->nState = totalDofStencilMesh();
    holder->ghostLen = (recon + 1) * NDPC;   // ghost layers of the scheme x dofs per cell
    impl::check(pda_problem_set_bc_callback(h_, side, &impl::BcFunctorHolder<F, NDPC>::ghost,
                                            &impl::BcFunctorHolder<F, NDPC>::factors, holder.get()));
    bc_[side] = std::move(holder);
  }

 private:
  pda_problem h_ = nullptr;
  std::shared_ptr<impl::BcFunctorBase> bc_[4];
};

// create_problem_eigen(mesh, <enum>, recon[, icFlag][, {name: value}])  (euler1d.hpp:82-99, euler2d.hpp:88-186, ...)
template <class ProbEnum>
Problem create_problem_eigen(const Mesh& m, ProbEnum e, InviscidFluxReconstruction r, int icFlag = 1,
                             const std::unordered_map<std::string, double>& params = {}, int device = 0) {
  return Problem(m, impl::family_of<ProbEnum>::value, static_cast<int>(e), static_cast<int>(r), icFlag, params, device);
}
template <class ProbEnum>
Problem create_problem_eigen(const Mesh& m, ProbEnum e, InviscidFluxReconstruction r,
                             const std::unordered_map<std::string, double>& params, int device = 0) {
  return Problem(m, impl::family_of<ProbEnum>::value, static_cast<int>(e), static_cast<int>(r), 1, params, device);
}

// ---- the remaining overloads and named factories of the reference's public headers (argument for argument)
// DiffusionReaction1d / DiffusionReaction2d: (mesh, enum[, ViscousFluxReconstruction])
// (diffusion_reaction1d.hpp:112-126, diffusion_reaction2d.hpp:126-180)
inline Problem create_problem_eigen(const Mesh& m, DiffusionReaction1d e, int device = 0) {
  return Problem(m, PDA_FAMILY_DIFFUSION_REACTION1D, static_cast<int>(e), 0, 1, {}, device);
}
inline Problem create_problem_eigen(const Mesh& m, DiffusionReaction2d e,
                                    ViscousFluxReconstruction = ViscousFluxReconstruction::FirstOrder, int device = 0) {
  return Problem(m, PDA_FAMILY_DIFFUSION_REACTION2D, static_cast<int>(e), 0, 1, {}, device);
}
// create_gray_scott_2d_problem_eigen(mesh, viscRecon, Du, Dv, F, k)  (diffusion_reaction2d.hpp:259-285)
inline Problem create_gray_scott_2d_problem_eigen(const Mesh& m, ViscousFluxReconstruction, double diffusion_u,
                                                  double diffusion_v, double feedRate, double killRate, int device = 0) {
  return Problem(m, PDA_FAMILY_DIFFUSION_REACTION2D, static_cast<int>(DiffusionReaction2d::GrayScott), 0, 1,
                 {{"Du", diffusion_u}, {"Dv", diffusion_v}, {"F", feedRate}, {"k", killRate}}, device);
}
// AdvectionDiffusion2d (Burgers): (mesh, enum, recon, viscRecon[, {name: value}])  (advection_diffusion2d.hpp:79-152)
inline Problem create_problem_eigen(const Mesh& m, AdvectionDiffusion2d e, InviscidFluxReconstruction r,
                                    ViscousFluxReconstruction, const std::unordered_map<std::string, double>& params = {},
                                    int device = 0) {
  return Problem(m, PDA_FAMILY_ADVECTION_DIFFUSION2D, static_cast<int>(e), static_cast<int>(r), 1, params, device);
}
// create_linear_advection_1d_problem_eigen(mesh, recon, [InviscidFluxScheme,] velocity[, ic])  (advection1d.hpp:107-152)
inline Problem create_linear_advection_1d_problem_eigen(const Mesh& m, InviscidFluxReconstruction r, double velocity,
                                                        int ic = 1, int device = 0) {
  return Problem(m, PDA_FAMILY_ADVECTION1D, static_cast<int>(Advection1d::PeriodicLinear), static_cast<int>(r), ic,
                 {{"velocity", velocity}}, device);
}
inline Problem create_linear_advection_1d_problem_eigen(const Mesh& m, InviscidFluxReconstruction r, InviscidFluxScheme,
                                                        double velocity, int device = 0) {
  return create_linear_advection_1d_problem_eigen(m, r, velocity, 1, device);
}
// create_slip_wall_swe_2d_problem_eigen(mesh, recon, gravity, coriolis, pulseMagnitude)  (swe2d.hpp:283-309, legacy)
inline Problem create_slip_wall_swe_2d_problem_eigen(const Mesh& m, InviscidFluxReconstruction r, double gravity,
                                                     double coriolis, double pulseMagnitude, int device = 0) {
  return Problem(m, PDA_FAMILY_SWE2D, static_cast<int>(Swe2d::SlipWall), static_cast<int>(r), 1,
                 {{"gravity", gravity}, {"coriolis", coriolis}, {"pulseMagnitude", pulseMagnitude}}, device);
}
// create_cross_shock_problem_eigen(mesh, recon[, density, inletXVel, bottomYVel])  (euler2d.hpp:252-298)
inline Problem create_cross_shock_problem_eigen(const Mesh& m, InviscidFluxReconstruction r, int device = 0) {
  return Problem(m, PDA_FAMILY_EULER2D, static_cast<int>(Euler2d::CrossShock), static_cast<int>(r), 1, {}, device);
}
inline Problem create_cross_shock_problem_eigen(const Mesh& m, InviscidFluxReconstruction r, double density,
                                                double inletXVel, double bottomYVel, int device = 0) {
  return Problem(m, PDA_FAMILY_EULER2D, static_cast<int>(Euler2d::CrossShock), static_cast<int>(r), 1,
                 {{"crossShockDensity", density}, {"crossShockInletXVel", inletXVel}, {"crossShockBottomYVel", bottomYVel}},
                 device);
}

// create_problem_eigen(mesh, <enum>, recon, BCsLeft, BCsFront, BCsRight, BCsBack[, icFlag])  -- the custom-BC overloads
// (swe2d.hpp:187-281, euler2d.hpp:188-250, advection_diffusion2d.hpp): arbitrary host functors with the reference's
// two call operators; evaluated on the host for the boundary cells at every evaluation (the slow, fully general path --
// device-expressible rules go through pda_problem_set_bc and never leave HBM)
template <class ProbEnum, class FL, class FF, class FR, class FB,
          class = std::enable_if_t<!std::is_arithmetic<std::decay_t<FL>>::value &&
                                   !std::is_same<std::decay_t<FL>, std::unordered_map<std::string, double>>::value>>
Problem create_problem_eigen(const Mesh& m, ProbEnum e, InviscidFluxReconstruction r, FL&& bcLeft, FF&& bcFront,
                             FR&& bcRight, FB&& bcBack, int icFlag = 1, int device = 0) {
  constexpr int nd = impl::custom_bc_ndpc<ProbEnum>::value;
  Problem p(m, impl::family_of<ProbEnum>::value, static_cast<int>(e), static_cast<int>(r), icFlag, {}, device);
  p.template installBcFunctor<nd>(0, m, static_cast<int>(r), std::decay_t<FL>(std::forward<FL>(bcLeft)));
  p.template installBcFunctor<nd>(1, m, static_cast<int>(r), std::decay_t<FF>(std::forward<FF>(bcFront)));
  p.template installBcFunctor<nd>(2, m, static_cast<int>(r), std::decay_t<FR>(std::forward<FR>(bcRight)));
  p.template installBcFunctor<nd>(3, m, static_cast<int>(r), std::decay_t<FB>(std::forward<FB>(bcBack)));
  return p;
}

}  // namespace pressiodemoapps_b200
