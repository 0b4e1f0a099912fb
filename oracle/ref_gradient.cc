// oracle/ref_gradient.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// C-ABI shim around the UNMODIFIED reference's GradientEvaluator (include/pressiodemoapps/gradient.hpp:61-121,
// impl/gradient_2d.hpp:62-293) and the mesh's graphRowsOfCellsStrictlyOnBd() (impl/mesh_ccu.hpp:153-155, 441-447).
// Compiled by oracle/Makefile into oracle/_ref/libpda_ref_grad.so; nothing of the reference is copied, its headers
// are #included from where they lie.  Faces are listed the way the reference's own test walks them
// (tests_cpp/gradients/main.cc:48-97): rows of graphRowsOfCellsStrictlyOnBd() in order, Left, Front, Right, Back.
#include "pressiodemoapps/mesh.hpp"
#include "pressiodemoapps/gradient.hpp"

#include <cstdint>
#include <cstring>
#include <string>

namespace pda = pressiodemoapps;
using mesh_t = pda::cellcentered_uniform_mesh_eigen_type;
using vec_t = Eigen::VectorXd;

namespace {
thread_local std::string g_err;

template <class Fn> void forEachBoundaryFace(const mesh_t& mesh, Fn&& fn) {
  const auto& G = mesh.graph();
  for (auto rowInd : mesh.graphRowsOfCellsStrictlyOnBd()) {
    const int cellGID = G(rowInd, 0);
    if (mesh.cellHasLeftFaceOnBoundary2d(rowInd)) fn(rowInd, cellGID, pda::FacePosition::Left);
    if (mesh.cellHasFrontFaceOnBoundary2d(rowInd)) fn(rowInd, cellGID, pda::FacePosition::Front);
    if (mesh.cellHasRightFaceOnBoundary2d(rowInd)) fn(rowInd, cellGID, pda::FacePosition::Right);
    if (mesh.cellHasBackFaceOnBoundary2d(rowInd)) fn(rowInd, cellGID, pda::FacePosition::Back);
  }
}

template <std::size_t MAXN>
void evalMany(const mesh_t& mesh, const vec_t& f, int ndpc, double* grad, double* centers, int32_t* normalDir) {
  pda::GradientEvaluator<mesh_t, MAXN> grads(mesh);
  grads(f, ndpc);
  int64_t k = 0;
  forEachBoundaryFace(mesh, [&](int, int cellGID, pda::FacePosition fp) {
    const auto& face = grads.queryFace(cellGID, fp);
    for (int j = 0; j < ndpc; ++j) grad[k * ndpc + j] = face.normalGradient[j];
    for (int c = 0; c < 3; ++c) centers[3 * k + c] = face.centerCoordinates[c];
    normalDir[k] = face.normalDirection;
    ++k;
  });
}
}  // namespace

extern "C" {

const char* pdaref_grad_last_error() { return g_err.c_str(); }

// rows of graphRowsOfCellsStrictlyOnBd(): returns the count, fills `rows` when non-null
int64_t pdaref_rows_strictly_on_bd(const char* meshDir, int32_t* rows) {
  try {
    auto mesh = pda::load_cellcentered_uniform_mesh_eigen(std::string(meshDir));
    const auto& r = mesh.graphRowsOfCellsStrictlyOnBd();
    if (rows)
      for (size_t i = 0; i < r.size(); ++i) rows[i] = (int32_t)r[i];
    return (int64_t)r.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// boundary faces: returns the count; fills (when non-null) cell gid, FacePosition (0 Left, 1 Front, 2 Right, 3 Back),
// parent graph row
int64_t pdaref_grad_faces(const char* meshDir, int32_t* cellGid, int32_t* position, int32_t* parentRow) {
  try {
    auto mesh = pda::load_cellcentered_uniform_mesh_eigen(std::string(meshDir));
    int64_t k = 0;
    forEachBoundaryFace(mesh, [&](int rowInd, int cellGID, pda::FacePosition fp) {
      if (cellGid) cellGid[k] = cellGID;
      if (position) position[k] = (int32_t)fp;
      if (parentRow) parentRow[k] = rowInd;
      ++k;
    });
    return k;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// normal gradients of `field` ([stencilMeshSize][ndpc]) at every boundary face, in the order of pdaref_grad_faces.
// ndpc == 1 and useScalarApi != 0 goes through operator()(field) of GradientEvaluator<Mesh,1> (scalar normalGradient),
// otherwise through operator()(field, ndpc) of GradientEvaluator<Mesh,5>.
int pdaref_grad_eval(const char* meshDir, int ndpc, int useScalarApi, const double* field, double* grad /*[n][ndpc]*/,
                     double* centers /*[n][3]*/, int32_t* normalDir) {
  try {
    auto mesh = pda::load_cellcentered_uniform_mesh_eigen(std::string(meshDir));
    vec_t f = Eigen::Map<const vec_t>(field, (Eigen::Index)mesh.stencilMeshSize() * ndpc);
    if (ndpc == 1 && useScalarApi) {
      pda::GradientEvaluator<mesh_t> grads(mesh);
      grads(f);
      int64_t k = 0;
      forEachBoundaryFace(mesh, [&](int, int cellGID, pda::FacePosition fp) {
        const auto& face = grads.queryFace(cellGID, fp);
        grad[k] = face.normalGradient;
        for (int c = 0; c < 3; ++c) centers[3 * k + c] = face.centerCoordinates[c];
        normalDir[k] = face.normalDirection;
        ++k;
      });
    } else {
      evalMany<5>(mesh, f, ndpc, grad, centers, normalDir);
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}  // extern "C"
