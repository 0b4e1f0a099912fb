/* pda_oracle.h -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's hot path
 * (pressio-demoapps velocity / Jacobian evaluation).  Never linked into the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef PDA_ORACLE_H_
#define PDA_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct or_problem_s or_problem;

const char* or_last_error(void);
int or_num_threads(void);

/* mesh read from the reference's text format (impl/mesh_read_{info,coords,connectivity}.hpp).
 * allow3dStencil7 = 1 enables the documented-but-unimplemented 3D stencil-7 layout. */
or_problem* or_create(const char* meshDir, int family, int probEnum, int recon, int icFlag, int nParams,
                      const char* const* names, const double* values);
/* same, from arrays (graph row-major [nSample][(stencil-1)*dim+1]); used for meshes too large for text files */
or_problem* or_create_from_arrays(int dim, int stencil, int32_t nSample, int32_t nStencil, const double dxyz[3],
                                  const double* x, const double* y, const double* z, const int32_t* graph,
                                  int family, int probEnum, int recon, int icFlag, int nParams,
                                  const char* const* names, const double* values);
/* lattice mode: a full mesh in natural ordering evaluated WITHOUT a stored graph (neighbours by index arithmetic, same
 * rows / columns / arithmetic as the stored-graph path): what makes the BASELINE sizes (512^3, 4096^2) checkable
 * against the oracle.  cx/cy/cz: per-axis cell-centre coordinates or NULL; no Jacobian pattern at these sizes
 * unless asked (or_pattern / or_velocity_and_jacobian build it on demand like the other constructors). */
or_problem* or_create_lattice(int dim, int stencil, const int32_t n[3], const double dxyz[3], const int32_t periodic[3],
                              const double* cx, const double* cy, const double* cz, int family, int probEnum,
                              int recon, int icFlag, int nParams, const char* const* names, const double* values);
void or_destroy(or_problem* p);
/* replaces the per-sample-row source table of a ProblemA family (nSample doubles) */
void or_set_source(or_problem* p, const double* values);
/* what: 0 dim, 1 stencilSize, 2 sampleMeshSize, 3 stencilMeshSize, 4 graph cols, 5 numInner, 6 numNearBd,
 *       7 isFullyPeriodic, 8 ndpc, 9 nDofStencil, 10 nDofSample, 11 nnz */
long long or_query(or_problem* p, int what);
void or_mesh_arrays(or_problem* p, int32_t* graph, double* x, double* y, double* z, int32_t* rowsInner,
                    int32_t* rowsNearBd, double* dxyz6);
void or_ic(or_problem* p, double* U);
int or_velocity(or_problem* p, const double* U, double t, double* V);
int or_velocity_and_jacobian(or_problem* p, const double* U, double t, double* V, double* vals);
void or_pattern(or_problem* p, int32_t* rowptr, int32_t* colidx);
int or_ghosts(or_problem* p, int side, double* out);
double or_time_velocity(or_problem* p, const double* U, double t, int warmup, int reps);
/* velocity of the inner rows [it0, it1) only, `reps` times (bounded sample of a big workload); returns seconds */
double or_time_velocity_inner_range(or_problem* p, const double* U, double t, double* V, int32_t it0, int32_t it1, int reps);

/* leaf functions exposed for the known-answer tests (tests_cpp/weno5/main.cc, weno3/main.cc) */
void or_weno5(double* uNeg, double* uPos, double qm2, double qm1, double q, double qp1, double qp2, double qp3);
void or_weno3(double* uNeg, double* uPos, double qm1, double q, double qp1, double qp2);
void or_weno5_grad(double* uNeg, double* uPos, double* gNeg6, double* gPos6, double qm2, double qm1, double q,
                   double qp1, double qp2, double qp3);
void or_weno3_grad(double* uNeg, double* uPos, double* gNeg4, double* gPos4, double qm1, double q, double qp1,
                   double qp2);
void or_euler_flux(int ndpc, double* F, const double* qL, const double* qR, const double* n, double gamma);
void or_euler_flux_jac(int ndpc, double* JL, double* JR, const double* qL, const double* qR, const double* n,
                       double gamma);
void or_swe_flux(double* F, const double* qL, const double* qR, const double* n, double g);
void or_swe_flux_jac(double* JL, double* JR, const double* qL, const double* qR, const double* n, double g);

/* GradientEvaluator restated (gradient.hpp:61-121, impl/gradient_2d.hpp:62-104, 164-286), on mesh arrays (2D):
 * or_gradient_faces lists the boundary faces -- rows with a first-layer neighbour missing among rowsNearBd
 * (mesh_ccu.hpp:298-326, 441-447), then Left, Front, Right, Back -- and returns their number (outputs may be NULL);
 * or_gradient_eval fills grad[face][dof] for field[stencilCell][dof]. */
int64_t or_gradient_faces(int stencil, const int32_t* graph, const int32_t* rowsNearBd, int32_t nNearBd, const double* x,
                          const double* y, const double* z, double dx, double dy, int32_t* cellGid, int32_t* position,
                          int32_t* parentRow, int32_t* normalDir, double* centers);
void or_gradient_eval(int stencil, const int32_t* graph, int64_t nFaces, const int32_t* position,
                      const int32_t* parentRow, double dx, double dy, const double* field, int ndpc, double* grad);

#ifdef __cplusplus
}
#endif
#endif
