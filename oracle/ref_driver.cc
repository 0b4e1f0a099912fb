// oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C-ABI shim around the UNMODIFIED reference (header-only pressio-demoapps under
// /root/reference/include + its vendored Eigen under /root/reference/tpls/eigen3).  It is compiled
// by oracle/Makefile into oracle/_ref/libpda_ref.so (serial, parity) and oracle/_ref/libpda_ref_omp.so
// (OpenMP, CPU baseline timing).  No reference source is copied: everything is #included from where
// it lies.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load these libraries.
//
// Reference entry points exercised:
//   load_cellcentered_uniform_mesh_eigen           include/pressiodemoapps/mesh.hpp:87-91
//   create_problem_eigen (Euler1d/2d/3d, Swe2d)    euler1d.hpp:82-99 euler2d.hpp:88-186 euler3d.hpp:80-122 swe2d.hpp:134-185
//   create_gray_scott_2d_problem_eigen             diffusion_reaction2d.hpp:259-285
//   create_diffusion_reaction_{1d,2d}_problem_A_eigen  diffusion_reaction1d.hpp:128-212 diffusion_reaction2d.hpp:186-257
//   create_problem_eigen (AdvectionDiffusion2d)    advection_diffusion2d.hpp:79-152
//   create_advdiffreac_2d_problem_A_eigen          advection_diffusion_reaction2d.hpp:140-160
//   create_linear_advection_1d_problem_eigen       advection1d.hpp:133-152
//   rightHandSide / rightHandSideAndJacobian / applyJacobian   adapter_cpp.hpp:162-259
#define PRESSIODEMOAPPS_ENABLE_TESTS 1   // exposes viewGhost*() (euler_2d_prob_class.hpp:205-210)
#include "pressiodemoapps/euler1d.hpp"
#include "pressiodemoapps/euler2d.hpp"
#include "pressiodemoapps/euler3d.hpp"
#include "pressiodemoapps/swe2d.hpp"
#include "pressiodemoapps/diffusion_reaction2d.hpp"
#include "pressiodemoapps/advection_diffusion2d.hpp"
// the reference's ADR class does not compile with OpenMP (`#pragma omp for` followed by a declaration,
// advection_diffusion_reaction_2d_prob_class.hpp:490-493): the OpenMP timing build leaves it out
#ifndef PDA_REFDRV_NO_ADR2D
#include "pressiodemoapps/advection_diffusion_reaction2d.hpp"
#endif
#include "pressiodemoapps/advection1d.hpp"
#include "pressiodemoapps/diffusion_reaction1d.hpp"

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace pda = pressiodemoapps;
using mesh_t = pda::cellcentered_uniform_mesh_eigen_type;
using vec_t  = Eigen::VectorXd;
using jac_t  = Eigen::SparseMatrix<double, Eigen::RowMajor, int32_t>;
using cmap_t = Eigen::Map<const vec_t>;
using mmap_t = Eigen::Map<vec_t>;

namespace {

struct ProblemBase {
  virtual ~ProblemBase() = default;
  virtual int ndpc() const = 0;
  virtual int nDofStencil() const = 0;
  virtual int nDofSample() const = 0;
  virtual vec_t ic() const = 0;
  virtual void velocity(const vec_t& U, double t, vec_t& V) const = 0;
  virtual void velocityAndJacobian(const vec_t& U, double t, vec_t& V, jac_t& J) const = 0;
  virtual jac_t createJacobian() const = 0;
  virtual void applyJacobianVec(const vec_t& U, const vec_t& B, double t, vec_t& R) const = 0;
  virtual int ghosts(int /*side*/, double* /*out*/) const { return -1; }
};

template <class P, bool HasGhostView>
struct ProblemHolder final : ProblemBase {
  P prob;
  template <class... A> explicit ProblemHolder(A&&... a) : prob(std::forward<A>(a)...) {}
  int ndpc() const override { return prob.numDofPerCell(); }
  int nDofStencil() const override { return prob.totalDofStencilMesh(); }
  int nDofSample() const override { return prob.totalDofSampleMesh(); }
  vec_t ic() const override { return prob.initialCondition(); }
  void velocity(const vec_t& U, double t, vec_t& V) const override { prob.rightHandSide(U, t, V); }
  void velocityAndJacobian(const vec_t& U, double t, vec_t& V, jac_t& J) const override {
    prob.rightHandSideAndJacobian(U, t, V, J);
  }
  jac_t createJacobian() const override { return prob.createJacobian(); }
  void applyJacobianVec(const vec_t& U, const vec_t& B, double t, vec_t& R) const override {
    prob.applyJacobian(U, B, t, R);
  }
  int ghosts(int side, double* out) const override {
    if constexpr (HasGhostView) {
      const auto& g = (side == 0) ? prob.viewGhostLeft()
                    : (side == 1) ? prob.viewGhostFront()
                    : (side == 2) ? prob.viewGhostRight() : prob.viewGhostBack();
      if (out) std::memcpy(out, g.data(), sizeof(double) * g.rows() * g.cols());
      return int(g.rows() * g.cols());
    } else {
      (void)side; (void)out;
      return -1;
    }
  }
};

struct Handle {
  std::unique_ptr<mesh_t> mesh;
  std::unique_ptr<ProblemBase> prob;
  jac_t J;          // scratch with the fixed pattern
  std::string err;
};

thread_local std::string g_err;

template <class P, bool G = false, class... A>
std::unique_ptr<ProblemBase> hold(P&& p) {
  return std::unique_ptr<ProblemBase>(new ProblemHolder<std::decay_t<P>, G>(std::move(p)));
}

pda::InviscidFluxReconstruction recon(int r) {
  switch (r) {
    case 0: return pda::InviscidFluxReconstruction::FirstOrder;
    case 1: return pda::InviscidFluxReconstruction::Weno3;
    case 2: return pda::InviscidFluxReconstruction::Weno5;
  }
  throw std::runtime_error("ref_driver: bad reconstruction enum");
}

}  // namespace

extern "C" {

const char* pdaref_last_error() { return g_err.c_str(); }

int pdaref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// family: 1 Euler1d, 2 Euler2d, 3 Euler3d, 4 Swe2d, 5 DiffusionReaction2d (0 ProblemA, 1 GrayScott),
//         6 AdvectionDiffusion2d (Burgers), 7 AdvectionDiffusionReaction2d, 8 Advection1d, 9 DiffusionReaction1d
void* pdaref_create(const char* meshDir, int family, int probEnum, int reconEnum, int icFlag,
                    int nParams, const char* const* names, const double* values) {
  try {
    auto h = std::make_unique<Handle>();
    h->mesh = std::make_unique<mesh_t>(pda::load_cellcentered_uniform_mesh_eigen(std::string(meshDir)));
    const mesh_t& m = *h->mesh;
    std::unordered_map<std::string, double> up;
    for (int i = 0; i < nParams; ++i) up[names[i]] = values[i];

    switch (family) {
      case 1:
        h->prob = hold(pda::create_problem_eigen(m, static_cast<pda::Euler1d>(probEnum), recon(reconEnum)));
        break;
      case 2: {
        auto pe = static_cast<pda::Euler2d>(probEnum);
        if (pe == pda::Euler2d::CrossShock) {
          h->prob = hold<decltype(pda::create_cross_shock_problem_eigen(m, recon(reconEnum))), true>(
              pda::create_cross_shock_problem_eigen(m, recon(reconEnum)));
        } else if (nParams > 0) {
          h->prob = hold<decltype(pda::create_problem_eigen(m, pe, recon(reconEnum), icFlag, up)), true>(
              pda::create_problem_eigen(m, pe, recon(reconEnum), icFlag, up));
        } else {
          h->prob = hold<decltype(pda::create_problem_eigen(m, pe, recon(reconEnum), icFlag)), true>(
              pda::create_problem_eigen(m, pe, recon(reconEnum), icFlag));
        }
        break;
      }
      case 3:
        h->prob = hold(pda::create_problem_eigen(m, static_cast<pda::Euler3d>(probEnum), recon(reconEnum)));
        break;
      case 4: {
        auto pe = static_cast<pda::Swe2d>(probEnum);
        h->prob = hold<decltype(pda::create_problem_eigen(m, pe, recon(reconEnum), icFlag, up)), true>(
            pda::create_problem_eigen(m, pe, recon(reconEnum), icFlag, up));
        break;
      }
      case 5: {
        if (probEnum == 0) {
          const double D = up.count("diffusion") ? up["diffusion"] : 0.01;
          const double k = up.count("reaction") ? up["reaction"] : 0.01;
          const auto visc = pda::ViscousFluxReconstruction::FirstOrder;
          if (up.count("testSource")) {   // time-dependent analytic source (parity of pda_problem_set_source)
            auto f = [](const double& x, const double& y, const double& t, double& v) { v = std::cos(x * y + t); };
            h->prob = hold(pda::create_diffusion_reaction_2d_problem_A_eigen(m, visc, f, D, k));
          } else {
            h->prob = hold(pda::create_diffusion_reaction_2d_problem_A_eigen(m, visc, D, k));
          }
          break;
        }
        double Du = 0.0002, Dv = 0.00005, F = 0.042, k = 0.062;
        if (up.count("Du")) Du = up["Du"];
        if (up.count("Dv")) Dv = up["Dv"];
        if (up.count("F")) F = up["F"];
        if (up.count("k")) k = up["k"];
        h->prob = hold(pda::create_gray_scott_2d_problem_eigen(m, pda::ViscousFluxReconstruction::FirstOrder, Du, Dv, F, k));
        break;
      }
      case 6: {
        auto pe = static_cast<pda::AdvectionDiffusion2d>(probEnum);
        const auto visc = pda::ViscousFluxReconstruction::FirstOrder;
        if (nParams > 0) h->prob = hold(pda::create_problem_eigen(m, pe, recon(reconEnum), visc, up));
        else h->prob = hold(pda::create_problem_eigen(m, pe, recon(reconEnum), visc));
        break;
      }
#ifndef PDA_REFDRV_NO_ADR2D
      case 7: {
        if (nParams > 0) {
          const double ux = up.count("ux") ? up["ux"] : 0.5 * std::cos(M_PI / 3);
          const double uy = up.count("uy") ? up["uy"] : 0.5 * std::sin(M_PI / 3);
          const double D = up.count("diffusion") ? up["diffusion"] : 0.001;
          const double sg = up.count("sigma") ? up["sigma"] : 1.0;
          h->prob = hold(pda::create_advdiffreac_2d_problem_A_eigen(m, recon(reconEnum), ux, uy, D, sg));
        } else {
          h->prob = hold(pda::create_problem_eigen(m, pda::AdvectionDiffusionReaction2d::ProblemA, recon(reconEnum)));
        }
        break;
      }
#endif
      case 8: {
        const double vel = up.count("velocity") ? up["velocity"] : 1.0;
        h->prob = hold(pda::create_linear_advection_1d_problem_eigen(m, recon(reconEnum), vel, icFlag));
        break;
      }
      case 9: {
        const double D = up.count("diffusion") ? up["diffusion"] : 0.01;
        const double k = up.count("reaction") ? up["reaction"] : 0.01;
        if (up.count("testSource")) {
          auto f = [](const double& x, const double& t, double& v) { v = std::sin(x + t); };
          h->prob = hold(pda::create_diffusion_reaction_1d_problem_A_eigen(m, f, D, k));
        } else {
          h->prob = hold(pda::create_diffusion_reaction_1d_problem_A_eigen(m, D, k));
        }
        break;
      }
      default:
        throw std::runtime_error("ref_driver: unknown family");
    }
    h->J = h->prob->createJacobian();
    return h.release();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void pdaref_destroy(void* hv) { delete static_cast<Handle*>(hv); }

// what: 0 dim, 1 stencilSize, 2 sampleMeshSize, 3 stencilMeshSize, 4 graph cols, 5 numInner, 6 numNearBd,
//       7 isFullyPeriodic, 8 ndpc, 9 nDofStencil, 10 nDofSample, 11 nnz
long long pdaref_query(void* hv, int what) {
  auto* h = static_cast<Handle*>(hv);
  const mesh_t& m = *h->mesh;
  switch (what) {
    case 0: return m.dimensionality();
    case 1: return m.stencilSize();
    case 2: return m.sampleMeshSize();
    case 3: return m.stencilMeshSize();
    case 4: return m.graph().cols();
    case 5: return (long long)m.numCellsInner();
    case 6: return (long long)m.numCellsNearBd();
    case 7: return m.isFullyPeriodic() ? 1 : 0;
    case 8: return h->prob->ndpc();
    case 9: return h->prob->nDofStencil();
    case 10: return h->prob->nDofSample();
    case 11: return h->J.nonZeros();
  }
  return -1;
}

void pdaref_mesh_arrays(void* hv, int32_t* graph, double* x, double* y, double* z, int32_t* rowsInner,
                        int32_t* rowsNearBd, double* dxyz) {
  auto* h = static_cast<Handle*>(hv);
  const mesh_t& m = *h->mesh;
  if (graph) std::memcpy(graph, m.graph().data(), sizeof(int32_t) * m.graph().rows() * m.graph().cols());
  const auto n = m.stencilMeshSize();
  if (x) std::memcpy(x, m.viewX().data(), sizeof(double) * n);
  if (y) std::memcpy(y, m.viewY().data(), sizeof(double) * n);
  if (z) std::memcpy(z, m.viewZ().data(), sizeof(double) * n);
  if (rowsInner) std::memcpy(rowsInner, m.graphRowsOfCellsAwayFromBd().data(), sizeof(int32_t) * m.numCellsInner());
  if (rowsNearBd) std::memcpy(rowsNearBd, m.graphRowsOfCellsNearBd().data(), sizeof(int32_t) * m.numCellsNearBd());
  if (dxyz) {
    dxyz[0] = m.dx(); dxyz[1] = m.dy(); dxyz[2] = m.dz();
    dxyz[3] = m.dxInv(); dxyz[4] = m.dyInv(); dxyz[5] = m.dzInv();
  }
}

void pdaref_ic(void* hv, double* U) {
  auto* h = static_cast<Handle*>(hv);
  vec_t ic = h->prob->ic();
  std::memcpy(U, ic.data(), sizeof(double) * ic.size());
}

int pdaref_velocity(void* hv, const double* U, double t, double* V) {
  auto* h = static_cast<Handle*>(hv);
  try {
    vec_t u = cmap_t(U, h->prob->nDofStencil());
    vec_t v(h->prob->nDofSample());
    h->prob->velocity(u, t, v);
    std::memcpy(V, v.data(), sizeof(double) * v.size());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

int pdaref_velocity_and_jacobian(void* hv, const double* U, double t, double* V, double* vals) {
  auto* h = static_cast<Handle*>(hv);
  try {
    vec_t u = cmap_t(U, h->prob->nDofStencil());
    vec_t v(h->prob->nDofSample());
    h->prob->velocityAndJacobian(u, t, v, h->J);
    if (V) std::memcpy(V, v.data(), sizeof(double) * v.size());
    if (vals) std::memcpy(vals, h->J.valuePtr(), sizeof(double) * h->J.nonZeros());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

void pdaref_pattern(void* hv, int32_t* rowptr, int32_t* colidx) {
  auto* h = static_cast<Handle*>(hv);
  std::memcpy(rowptr, h->J.outerIndexPtr(), sizeof(int32_t) * (h->J.rows() + 1));
  std::memcpy(colidx, h->J.innerIndexPtr(), sizeof(int32_t) * h->J.nonZeros());
}

int pdaref_apply_jacobian(void* hv, const double* U, const double* B, double t, double* R) {
  auto* h = static_cast<Handle*>(hv);
  try {
    vec_t u = cmap_t(U, h->prob->nDofStencil());
    vec_t b = cmap_t(B, h->prob->nDofStencil());
    vec_t r(h->prob->nDofSample());
    h->prob->applyJacobianVec(u, b, t, r);
    std::memcpy(R, r.data(), sizeof(double) * r.size());
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// side: 0 left, 1 front, 2 right, 3 back.  Returns number of doubles (or -1 if the family has no view).
int pdaref_ghosts(void* hv, int side, double* out) {
  return static_cast<Handle*>(hv)->prob->ghosts(side, out);
}

// Timing legs (tests_perf/main.cc:37-62 pattern: warm-up call(s), then N timed calls).  Returns seconds per call.
double pdaref_time_velocity(void* hv, const double* U, double t, int warmup, int reps) {
  auto* h = static_cast<Handle*>(hv);
  vec_t u = cmap_t(U, h->prob->nDofStencil());
  vec_t v(h->prob->nDofSample());
  for (int i = 0; i < warmup; ++i) h->prob->velocity(u, t, v);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) h->prob->velocity(u, t, v);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count() / reps;
}

double pdaref_time_jacobian(void* hv, const double* U, double t, int warmup, int reps) {
  auto* h = static_cast<Handle*>(hv);
  vec_t u = cmap_t(U, h->prob->nDofStencil());
  vec_t v(h->prob->nDofSample());
  for (int i = 0; i < warmup; ++i) h->prob->velocityAndJacobian(u, t, v, h->J);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) h->prob->velocityAndJacobian(u, t, v, h->J);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count() / reps;
}

}  // extern "C"
