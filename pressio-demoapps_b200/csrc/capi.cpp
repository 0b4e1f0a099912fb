// capi.cpp -- extern "C" boundary declared in include/pda_b200.h.  Catches every C++ exception and turns it into
// a status code + thread-local message (the reference throws std::runtime_error or calls exit(); a drop-in
// library must do neither across an FFI).
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <string>

#include "../../include/pda_b200.h"
#include "common.hpp"
#include "engine.hpp"
#include "gradient.hpp"
#include "mesh.hpp"

struct pda_mesh_s { pda::Mesh m; };
struct pda_problem_s { pda::Problem* p; };
struct pda_gradient_s { pda::GradientEvaluator* g; };

namespace {
thread_local std::string g_lastError;

template <class F>
pda_status guarded(F&& f) {
  try {
    f();
    return PDA_OK;
  } catch (const pda::Error& e) {
    g_lastError = e.what();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_lastError = "out of host memory";
    return PDA_ERR_TOO_LARGE;
  } catch (const std::exception& e) {
    g_lastError = e.what();
    return PDA_ERR_INVALID;
  } catch (...) {
    g_lastError = "unknown error";
    return PDA_ERR_INVALID;
  }
}

pda::Mesh& M(pda_mesh m) {
  if (!m) throw pda::Error(pda::kInvalid, "null mesh handle");
  return m->m;
}
pda::Problem& P(pda_problem p) {
  if (!p || !p->p) throw pda::Error(pda::kInvalid, "null problem handle");
  return *p->p;
}
pda::GradientEvaluator& GE(pda_gradient g) {
  if (!g || !g->g) throw pda::Error(pda::kInvalid, "null gradient evaluator handle");
  return *g->g;
}
}  // namespace

extern "C" {

const char* pda_last_error(void) { return g_lastError.c_str(); }
const char* pda_version(void) { return "pda_b200 0.1 (sm_100a)"; }

int pda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------ mesh
pda_status pda_mesh_load(const char* dir, pda_mesh* out) {
  return guarded([&] {
    if (!dir || !out) throw pda::Error(pda::kInvalid, "mesh_load: null argument");
    *out = new pda_mesh_s{pda::Mesh::load(dir)};
  });
}

pda_status pda_mesh_make_lattice(int dim, const int32_t n[3], const double bounds[6], const int32_t periodic[3],
                                 int stencil, pda_mesh* out) {
  return guarded([&] {
    if (!n || !bounds || !periodic || !out) throw pda::Error(pda::kInvalid, "mesh_make_lattice: null argument");
    *out = new pda_mesh_s{pda::Mesh::makeLattice(dim, n, bounds, periodic, stencil)};
  });
}

pda_status pda_mesh_make_sample(pda_mesh full, const int32_t* gids, int64_t ngids, pda_mesh* out) {
  return guarded([&] {
    if (!gids || !out) throw pda::Error(pda::kInvalid, "mesh_make_sample: null argument");
    *out = new pda_mesh_s{pda::Mesh::makeSample(M(full), gids, ngids)};
  });
}

pda_status pda_mesh_make_slab_window(pda_mesh full, int rank, int nranks, pda_mesh* out) {
  return guarded([&] {
    if (!out) throw pda::Error(pda::kInvalid, "mesh_make_slab_window: null output");
    *out = new pda_mesh_s{pda::Mesh::makeWindow(M(full), rank, nranks)};
  });
}

pda_status pda_mesh_slab_window_info(pda_mesh m, int64_t info[8]) {
  return guarded([&] {
    if (!info) throw pda::Error(pda::kInvalid, "mesh_slab_window_info: null output");
    const pda::Mesh& w = M(m);
    if (!w.window) throw pda::Error(pda::kInvalid, "mesh_slab_window_info: not a slab window mesh");
    info[0] = w.winPlaneCells; info[1] = w.winK0; info[2] = w.winK1; info[3] = w.winHLo; info[4] = w.winHHi;
    info[5] = w.winRank; info[6] = w.winRanks;
    info[7] = w.dim;
  });
}

pda_status pda_mesh_from_arrays(int dim, int stencil, int32_t nSample, int32_t nStencil, const double dxyz[3],
                                const double* x, const double* y, const double* z, const int32_t* graph,
                                pda_mesh* out) {
  return guarded([&] {
    if (!dxyz || !x || !graph || !out) throw pda::Error(pda::kInvalid, "mesh_from_arrays: null argument");
    *out = new pda_mesh_s{pda::Mesh::fromArrays(dim, stencil, nSample, nStencil, dxyz, x, y, z, graph)};
  });
}

pda_status pda_mesh_write(pda_mesh m, const char* dir) {
  return guarded([&] {
    if (!dir) throw pda::Error(pda::kInvalid, "mesh_write: null directory");
    M(m).write(dir);
  });
}

pda_status pda_mesh_free(pda_mesh m) {
  delete m;
  return PDA_OK;
}

int pda_mesh_dimensionality(pda_mesh m) { return m ? m->m.dim : -1; }
int pda_mesh_stencil_size(pda_mesh m) { return m ? m->m.stencil : -1; }
int32_t pda_mesh_stencil_mesh_size(pda_mesh m) { return m ? m->m.nStencil : -1; }
int32_t pda_mesh_sample_mesh_size(pda_mesh m) { return m ? m->m.nSample : -1; }
int pda_mesh_graph_cols(pda_mesh m) { return m ? m->m.ncols() : -1; }
int pda_mesh_is_fully_periodic(pda_mesh m) { return m ? (m->m.fullyPeriodic ? 1 : 0) : -1; }
int pda_mesh_is_lattice(pda_mesh m) { return m ? (m->m.lattice ? 1 : 0) : -1; }
int32_t pda_mesh_num_cells_near_bd(pda_mesh m) { return m ? (int32_t)m->m.countNearBd() : -1; }
int32_t pda_mesh_num_cells_inner(pda_mesh m) { return m ? (int32_t)(m->m.nSample - m->m.countNearBd()) : -1; }

pda_status pda_mesh_deltas(pda_mesh m, double dxyz[3], double dxyz_inv[3]) {
  return guarded([&] {
    pda::Mesh& mm = M(m);
    for (int a = 0; a < 3; ++a) {
      if (dxyz) dxyz[a] = mm.d[a];
      if (dxyz_inv) dxyz_inv[a] = mm.dInv[a];
    }
  });
}

pda_status pda_mesh_coordinates(pda_mesh m, double* x, double* y, double* z) {
  return guarded([&] {
    pda::Mesh& mm = M(m);
    mm.ensureCoords();
    const size_t nb = sizeof(double) * (size_t)mm.nStencil;
    if (x) std::memcpy(x, mm.x.data(), nb);
    if (y) std::memcpy(y, mm.y.data(), nb);
    if (z) std::memcpy(z, mm.z.data(), nb);
  });
}

pda_status pda_mesh_graph(pda_mesh m, int32_t* graph) {
  return guarded([&] {
    pda::Mesh& mm = M(m);
    if (!graph) throw pda::Error(pda::kInvalid, "mesh_graph: null output");
    mm.ensureGraph();
    std::memcpy(graph, mm.graph.data(), sizeof(int32_t) * mm.graph.size());
  });
}

pda_status pda_mesh_rows_inner(pda_mesh m, int32_t* rows) {
  return guarded([&] {
    pda::Mesh& mm = M(m);
    mm.ensureRows();
    if (rows && !mm.rowsInner.empty()) std::memcpy(rows, mm.rowsInner.data(), sizeof(int32_t) * mm.rowsInner.size());
  });
}

pda_status pda_mesh_rows_near_bd(pda_mesh m, int32_t* rows) {
  return guarded([&] {
    pda::Mesh& mm = M(m);
    mm.ensureRows();
    if (rows && !mm.rowsNearBd.empty()) std::memcpy(rows, mm.rowsNearBd.data(), sizeof(int32_t) * mm.rowsNearBd.size());
  });
}

int32_t pda_mesh_num_cells_strictly_on_bd(pda_mesh m) {
  if (!m) return -1;
  int32_t n = -1;
  guarded([&] {
    std::vector<int32_t> rows;
    m->m.strictlyOnBdRows(rows);
    n = (int32_t)rows.size();
  });
  return n;
}

pda_status pda_mesh_rows_strictly_on_bd(pda_mesh m, int32_t* rows) {
  return guarded([&] {
    std::vector<int32_t> r;
    M(m).strictlyOnBdRows(r);
    if (rows && !r.empty()) std::memcpy(rows, r.data(), sizeof(int32_t) * r.size());
  });
}

pda_status pda_mesh_stencil_gids(pda_mesh m, int32_t* gids) {
  return guarded([&] {
    pda::Mesh& mm = M(m);
    if (mm.stencilGids.empty()) throw pda::Error(pda::kInvalid, "mesh_stencil_gids: not a sample mesh built by this library");
    std::memcpy(gids, mm.stencilGids.data(), sizeof(int32_t) * mm.stencilGids.size());
  });
}

// ------------------------------------------------------------------ problem
pda_status pda_problem_create(pda_mesh mesh, int family, int problem_id, int recon, int ic_flag, int nparams,
                              const char* const* names, const double* values, int device, pda_problem* out) {
  return guarded([&] {
    if (!out) throw pda::Error(pda::kInvalid, "problem_create: null output");
    if (nparams > 0 && (!names || !values)) throw pda::Error(pda::kInvalid, "problem_create: null parameter arrays");
    *out = new pda_problem_s{new pda::Problem(&M(mesh), family, problem_id, recon, ic_flag, nparams, names, values, device)};
  });
}

pda_status pda_problem_set_source(pda_problem p, const double* values) {
  return guarded([&] { P(p).setSource(values); });
}

pda_status pda_problem_set_option(pda_problem p, const char* name, const char* value) {
  return guarded([&] {
    if (!name || !value) throw pda::Error(pda::kInvalid, "set_option: null argument");
    P(p).setOption(name, value);
  });
}

pda_status pda_problem_get_option(pda_problem p, const char* name, char* value, int capacity) {
  return guarded([&] {
    if (!name || !value || capacity < 1) throw pda::Error(pda::kInvalid, "get_option: null argument");
    const std::string v = P(p).getOption(name);
    if ((int)v.size() + 1 > capacity) throw pda::Error(pda::kInvalid, "get_option: buffer too small");
    std::memcpy(value, v.c_str(), v.size() + 1);
  });
}

pda_status pda_problem_set_bc(pda_problem p, int side, int kind, const double* values) {
  return guarded([&] { P(p).setBc(side, kind, values); });
}

pda_status pda_problem_set_bc_callback(pda_problem p, int side, pda_bc_ghost_fn ghost, pda_bc_factor_fn factors, void* user) {
  return guarded([&] { P(p).setBcCallback(side, ghost, factors, user); });
}

pda_status pda_problem_set_bc_pointer(pda_problem p, int side, void* user) {
  return guarded([&] { P(p).setBcPointer(side, user); });
}

pda_status pda_problem_free(pda_problem p) {
  if (p) { delete p->p; delete p; }
  return PDA_OK;
}

int pda_problem_num_dof_per_cell(pda_problem p) { return (p && p->p) ? p->p->ndpc() : -1; }
int32_t pda_problem_total_dof_sample_mesh(pda_problem p) { return (p && p->p) ? p->p->nDofSample() : -1; }
int32_t pda_problem_total_dof_stencil_mesh(pda_problem p) { return (p && p->p) ? p->p->nDofStencil() : -1; }

pda_status pda_problem_query_parameter(pda_problem p, const char* name, double* value) {
  return guarded([&] {
    if (!name || !value) throw pda::Error(pda::kInvalid, "query_parameter: null argument");
    *value = P(p).queryParameter(name);
  });
}

pda_status pda_problem_initial_condition(pda_problem p, double* U) {
  return guarded([&] {
    if (!U) throw pda::Error(pda::kInvalid, "initial_condition: null output");
    P(p).initialCondition(U);
  });
}

pda_status pda_problem_jacobian_nnz(pda_problem p, int64_t* nnz) {
  return guarded([&] {
    if (!nnz) throw pda::Error(pda::kInvalid, "jacobian_nnz: null output");
    *nnz = P(p).jacobianNnz();
  });
}

pda_status pda_problem_jacobian_pattern(pda_problem p, int32_t* rowptr, int32_t* colidx) {
  return guarded([&] { P(p).jacobianPattern(rowptr, colidx); });
}

pda_status pda_problem_velocity_host(pda_problem p, const double* U, double t, double* V) {
  return guarded([&] { P(p).velocityHost(U, t, V); });
}

pda_status pda_problem_velocity_and_jacobian_host(pda_problem p, const double* U, double t, double* V, double* jv) {
  return guarded([&] { P(p).velocityAndJacobianHost(U, t, V, jv); });
}

pda_status pda_problem_apply_jacobian_host(pda_problem p, const double* U, const double* B, int ncols, int layout,
                                           double t, double* R) {
  return guarded([&] { P(p).applyJacobianHost(U, B, ncols, layout, t, R); });
}

pda_status pda_problem_velocity_dev(pda_problem p, const double* dU, double t, double* dV, void* stream) {
  return guarded([&] { P(p).velocityDev(dU, t, dV, stream); });
}

pda_status pda_problem_velocity_and_jacobian_dev(pda_problem p, const double* dU, double t, double* dV, double* dJ,
                                                 void* stream) {
  return guarded([&] { P(p).velocityAndJacobianDev(dU, t, dV, dJ, stream); });
}

pda_status pda_problem_apply_jacobian_dev(pda_problem p, const double* dU, const double* dB, int ncols, int layout,
                                          double t, double* dR, void* stream) {
  return guarded([&] { P(p).applyJacobianDev(dU, dB, ncols, layout, t, dR, stream); });
}

pda_status pda_problem_advance_dev(pda_problem p, int stepper, double* dU, double t0, double dt, int32_t nsteps, void* stream) {
  return guarded([&] { P(p).advanceDev(stepper, dU, t0, dt, nsteps, stream); });
}

pda_status pda_problem_advance_host(pda_problem p, int stepper, double* U, double t0, double dt, int32_t nsteps) {
  return guarded([&] { P(p).advanceHost(stepper, U, t0, dt, nsteps); });
}

pda_status pda_problem_ghosts(pda_problem p, int side, double* out) {
  return guarded([&] { P(p).ghosts(side, out); });
}

int64_t pda_problem_launch_count(pda_problem p) { return (p && p->p) ? p->p->launchCount() : -1; }

// ------------------------------------------------------------------ slabs
pda_status pda_problem_create_slab(pda_mesh lattice, int family, int problem_id, int recon, int rank, int nranks,
                                   int device, pda_problem* out) {
  return guarded([&] {
    if (!out) throw pda::Error(pda::kInvalid, "problem_create_slab: null output");
    auto* prob = new pda::Problem(&M(lattice), family, problem_id, recon, 1, 0, nullptr, nullptr, device);
    try { prob->makeSlab(rank, nranks); } catch (...) { delete prob; throw; }
    *out = new pda_problem_s{prob};
  });
}

pda_status pda_slab_extent(pda_problem p, int32_t* k0, int32_t* k1, int32_t* halo, int64_t* plane_dofs) {
  return guarded([&] { P(p).slabExtent(k0, k1, halo, plane_dofs); });
}

pda_status pda_slab_initial_condition(pda_problem p, double* U_owned) {
  return guarded([&] { P(p).slabInitialCondition(U_owned); });
}

pda_status pda_slab_velocity_interior_dev(pda_problem p, const double* dU, double t, double* dV, void* stream) {
  return guarded([&] { P(p).slabVelocityDev(dU, t, dV, stream, false); });
}

pda_status pda_slab_velocity_boundary_dev(pda_problem p, const double* dU, double t, double* dV, void* stream) {
  return guarded([&] { P(p).slabVelocityDev(dU, t, dV, stream, true); });
}

pda_status pda_slab_peer_handle(pda_problem p, unsigned char handle[64]) {
  return guarded([&] {
    if (!handle) throw pda::Error(pda::kInvalid, "slab_peer_handle: null output");
    P(p).slabPeerHandle(handle);
  });
}

pda_status pda_slab_peer_connect(pda_problem p, const unsigned char* handles) {
  return guarded([&] { P(p).slabPeerConnect(handles); });
}

pda_status pda_slab_peer_connect_local(pda_problem p, pda_problem lower, pda_problem upper) {
  return guarded([&] { P(p).slabPeerConnectLocal(&P(lower), &P(upper)); });
}

pda_status pda_slab_velocity_peer_dev(pda_problem p, const double* dU, double t, double* dV, void* stream) {
  return guarded([&] { P(p).slabVelocityPeerDev(dU, t, dV, stream); });
}

pda_status pda_slab_velocity_peer_host(pda_problem p, const double* U, double t, double* V) {
  return guarded([&] { P(p).slabVelocityPeerHost(U, t, V); });
}

// ------------------------------------------------------------------ boundary-face gradients
pda_status pda_gradient_create(pda_mesh mesh, int max_num_dof_per_cell, pda_gradient* out) {
  return guarded([&] {
    if (!out) throw pda::Error(pda::kInvalid, "gradient_create: null output");
    *out = new pda_gradient_s{new pda::GradientEvaluator(M(mesh), max_num_dof_per_cell)};
  });
}

pda_status pda_gradient_free(pda_gradient g) {
  if (g) { delete g->g; delete g; }
  return PDA_OK;
}

int32_t pda_gradient_num_faces(pda_gradient g) { return (g && g->g) ? g->g->numFaces() : -1; }
int64_t pda_gradient_launch_count(pda_gradient g) { return (g && g->g) ? g->g->launchCount() : -1; }

pda_status pda_gradient_faces(pda_gradient g, int32_t* cell_gid, int32_t* position, int32_t* parent_graph_row,
                              int32_t* normal_direction, double* centers) {
  return guarded([&] {
    pda::GradientEvaluator& e = GE(g);
    const size_t n = (size_t)e.numFaces();
    if (n == 0) return;
    if (cell_gid) std::memcpy(cell_gid, e.cellGid().data(), n * sizeof(int32_t));
    if (position) std::memcpy(position, e.position().data(), n * sizeof(int32_t));
    if (parent_graph_row) std::memcpy(parent_graph_row, e.parentRow().data(), n * sizeof(int32_t));
    if (normal_direction) std::memcpy(normal_direction, e.normalDirection().data(), n * sizeof(int32_t));
    if (centers) std::memcpy(centers, e.centers().data(), 3 * n * sizeof(double));
  });
}

pda_status pda_gradient_query_face(pda_gradient g, int32_t cell_gid, int position, int32_t* face_index) {
  return guarded([&] {
    if (!face_index) throw pda::Error(pda::kInvalid, "gradient_query_face: null output");
    const int32_t i = GE(g).findFace(cell_gid, position);
    if (i < 0) throw pda::Error(pda::kInvalid, "gradient_query_face: the mesh has no such boundary face");
    *face_index = i;
  });
}

pda_status pda_gradient_compute_host(pda_gradient g, const double* field, int num_dof_per_cell, double* normal_grad) {
  return guarded([&] { GE(g).computeHost(field, num_dof_per_cell, normal_grad); });
}

pda_status pda_gradient_compute_dev(pda_gradient g, const double* d_field, int num_dof_per_cell, double* d_normal_grad,
                                    void* stream) {
  return guarded([&] { GE(g).computeDev(d_field, num_dof_per_cell, d_normal_grad, stream); });
}

}  // extern "C"
