/* pda_oracle.c -- TEST INFRASTRUCTURE ONLY (see pda_oracle.h).
 *
 * Plain-C restatement of the algorithm on pressio-demoapps' velocity/Jacobian hot path, written to follow the
 * reference's arithmetic literally (same expressions, same order; compiled with -ffp-contract=off) so that it can be
 * pinned against the reference itself (oracle/_ref/libpda_ref.so) and against the reference's golden files.  Every
 * function cites the reference file:line it follows (paths relative to /root/reference/include/pressiodemoapps).
 *
 * The one thing here the reference does NOT have is 3D WENO5: the third stencil layer of a 3D mesh
 * uses the documented column layout (functor_reconstruct_from_state.hpp:624-630, natural_order_mesh_3d.py:241-244)
 * and the same leaf functions.  That extension is labelled wherever it is used.
 *
 * Pinned by tests/test_oracle_cpu.py against: the known answers of tests_cpp/weno5/main.cc and weno3/main.cc, the
 * gold files of tests_cpp/eigen_1d_euler_sod_explicit, and oracle/_ref on every family (velocity, Jacobian, pattern).
 */
#include "pda_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

enum { F_EULER1D = 1, F_EULER2D = 2, F_EULER3D = 3, F_SWE2D = 4, F_DIFFREAC2D = 5, F_ADVDIFF2D = 6,
       F_ADVDIFFREAC2D = 7, F_ADVECTION1D = 8, F_DIFFREAC1D = 9 };
enum { E2_PERIODIC = 0, E2_KH = 1, E2_SEDOV_FULL = 2, E2_SEDOV_SYM = 3, E2_RIEMANN = 4, E2_NORMAL_SHOCK = 5,
       E2_DMR = 6, E2_CROSS_SHOCK = 7, E2_NEUMANN = 8 };

static char g_err[512];
const char* or_last_error(void) { return g_err; }
int or_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ================================================================================================== mesh */
typedef struct {
  int dim, stencil, ncols;
  int32_t nSample, nStencil;
  double d[3], dInv[3];
  double *x, *y, *z;
  int32_t* graph;
  int32_t *rowsInner, *rowsNearBd;
  int32_t nInner, nNearBd;
  int periodic;
  /* lattice mode (or_create_lattice): a full mesh in natural ordering whose connectivity is index arithmetic -- no
   * graph, no per-cell coordinates in memory (a 512^3 stencil-7 graph would be 10 GB).  Same rows, same columns, same
   * arithmetic as the stored-graph path; tests/test_oracle_cpu.py pins the two against each other bit for bit. */
  int lattice;
  int32_t n[3];
  int per[3];
  double *cx, *cy, *cz; /* per-axis cell-centre coordinates (as the mesh files carry them), may be NULL */
} or_mesh;

struct or_problem_s {
  or_mesh m;
  int family, prob, recon, icFlag, ndpc, S;
  double gamma;
  double icp[10];  /* euler2d: euler_2d_parametrization_helpers.hpp:59-72 ; swe: swe_2d_parametrization_helpers.hpp */
  double php[4];   /* swe: gravity, coriolis; burgers: diffusion; ADR: ux, uy, D, sigma; adv1d: velocity; ProblemA: D, k */
  double gs[4];
  double* src;     /* per-sample-row source table (ProblemA families) */
  double* ghost[6];     /* left, front, right, back, bottom, top : [nNearBd][gstride] */
  int gstride;
  int32_t *rowptr, *colidx;
  long long nnz;
};

/* neighbour `c` of lattice cell r (natural ordering gid = (k*ny + j)*nx + i, natural_order_mesh_{2,3}d.py):
 * column c-1 = (1D: 2, 2D: 4, 3D: 6) * layer + side; periodic axes wrap, others give -1 outside the domain */
static inline int32_t lattice_nb(const or_mesh* m, int32_t r, int c) {
  if (c == 0) return r;
  const int nside = (m->dim == 1) ? 2 : 2 * m->dim;
  const int layer = (c - 1) / nside;
  int side = (c - 1) % nside;
  if (m->dim == 1) side = side ? 2 : 0;
  const int32_t nx = m->n[0], ny = m->n[1];
  /* the (i,j,k) of the row asked last is remembered per thread: a cell's neighbours are asked for in a burst */
  static __thread int32_t lastRow = -1, lastNx = -1, lastNy = -1, lastIjk[3];
  if (r != lastRow || nx != lastNx || ny != lastNy) {
    lastRow = r; lastNx = nx; lastNy = ny;
    lastIjk[0] = r % nx; lastIjk[1] = (r / nx) % ny; lastIjk[2] = r / (nx * ny);
  }
  int32_t ijk[3] = {lastIjk[0], lastIjk[1], lastIjk[2]};
  /* side: 0 left (-x), 1 front (+y), 2 right (+x), 3 back (-y), 4 bottom (-z), 5 top (+z) */
  const int axis = (side == 0 || side == 2) ? 0 : ((side == 1 || side == 3) ? 1 : 2);
  const int dir = (side == 2 || side == 1 || side == 5) ? 1 : -1;
  int32_t v = ijk[axis] + dir * (layer + 1);
  if (v < 0 || v >= m->n[axis]) {
    if (!m->per[axis]) return -1;
    v = (v % m->n[axis] + m->n[axis]) % m->n[axis];
  }
  ijk[axis] = v;
  return (ijk[2] * ny + ijk[1]) * nx + ijk[0];
}
#define G(m, r, c) ((m)->graph ? (m)->graph[(size_t)(r) * (m)->ncols + (c)] : lattice_nb((m), (int32_t)(r), (c)))
/* cell-centre coordinates: stored per cell, or per axis in lattice mode */
static inline double mx(const or_mesh* m, int32_t i) { return m->x ? m->x[i] : (m->cx ? m->cx[i % m->n[0]] : 0.); }
static inline double my(const or_mesh* m, int32_t i) { return m->y ? m->y[i] : (m->cy ? m->cy[(i / m->n[0]) % m->n[1]] : 0.); }
static inline double mz(const or_mesh* m, int32_t i) { return m->z ? m->z[i] : (m->cz ? m->cz[i / (m->n[0] * m->n[1])] : 0.); }
/* row lists: identity when every row is an inner row of a lattice (fully periodic) */
#define ROW_INNER(m, it) ((m)->rowsInner ? (m)->rowsInner[it] : (int32_t)(it))

/* graph column conventions (functor_reconstruct_from_state.hpp:262-267,442-448,624-630):
 * 1D [l0 r0 | l1 r1 | l2 r2]; 2D [l0 f0 r0 b0 | ...]; 3D [l0 f0 r0 ba0 bot0 top0 | ...].
 * side: 0 left, 1 front, 2 right, 3 back, 4 bottom, 5 top */
static int gcol(int dim, int side, int layer) {
  if (dim == 1) return 1 + 2 * layer + (side == 2 ? 1 : 0);
  return 1 + (dim == 2 ? 4 : 6) * layer + side;
}
/* axis 1,2,3 -> (left,right) sides: x: left/right, y: back/front, z: bottom/top
 * (mixin_directional_flux_balance_jacobian.hpp:54-118) */
static int side_minus(int axis) { return axis == 1 ? 0 : (axis == 2 ? 3 : 4); }
static int side_plus(int axis) { return axis == 1 ? 2 : (axis == 2 ? 1 : 5); }

/* mesh_ccu.hpp:162-296: any -1 among the stencil layers of a side.  (3D tests layer 1 only for stencilSize==5 in the
 * reference; stencil 7 does not exist there -- the extension tests all layers.) */
static int has_bd(const or_mesh* m, int32_t row, int side) {
  int h = (m->stencil - 1) / 2;
  for (int L = 0; L < h; ++L)
    if (G(m, row, gcol(m->dim, side, L)) == -1) return 1;
  return 0;
}

static void classify(or_mesh* m) { /* mesh_ccu.hpp:385-439 + checkIfFullyPeriodic :331-342 */
  m->rowsInner = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m->nSample > 0 ? m->nSample : 1));
  m->rowsNearBd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m->nSample > 0 ? m->nSample : 1));
  m->nInner = m->nNearBd = 0;
  m->periodic = 1;
  for (int32_t r = 0; r < m->nSample; ++r) {
    int bd = 0;
    if (m->dim == 1) bd = has_bd(m, r, 0) || has_bd(m, r, 2);
    else for (int s = 0; s < 2 * m->dim; ++s) bd = bd || has_bd(m, r, s);
    if (bd) m->rowsNearBd[m->nNearBd++] = r; else m->rowsInner[m->nInner++] = r;
    for (int c = 0; c < m->ncols; ++c) if (G(m, r, c) < 0) m->periodic = 0;
  }
}

static int read_mesh(or_mesh* m, const char* dir) {
  char path[1024], line[4096];
  memset(m, 0, sizeof *m);
  snprintf(path, sizeof path, "%s/info.dat", dir);
  FILE* f = fopen(path, "r"); /* mesh_read_info.hpp:54-119 */
  if (!f) { snprintf(g_err, sizeof g_err, "file not found %s", path); return 1; }
  while (fgets(line, sizeof line, f)) {
    char key[64]; double val;
    if (sscanf(line, "%63s %lf", key, &val) != 2) continue;
    if (!strcmp(key, "dim")) m->dim = (int)val;
    else if (!strcmp(key, "dx")) { m->d[0] = val; m->dInv[0] = 1. / val; }
    else if (!strcmp(key, "dy")) { m->d[1] = val; m->dInv[1] = 1. / val; }
    else if (!strcmp(key, "dz")) { m->d[2] = val; m->dInv[2] = 1. / val; }
    else if (!strcmp(key, "sampleMeshSize")) m->nSample = (int32_t)val;
    else if (!strcmp(key, "stencilMeshSize")) m->nStencil = (int32_t)val;
    else if (!strcmp(key, "stencilSize")) m->stencil = (int)val;
  }
  fclose(f);
  m->ncols = (m->stencil - 1) * m->dim + 1;
  m->x = (double*)calloc((size_t)m->nStencil, sizeof(double));
  m->y = (double*)calloc((size_t)m->nStencil, sizeof(double));
  m->z = (double*)calloc((size_t)m->nStencil, sizeof(double));
  snprintf(path, sizeof path, "%s/coordinates.dat", dir); /* mesh_read_coords.hpp:54-87 */
  f = fopen(path, "r");
  if (!f) { snprintf(g_err, sizeof g_err, "file not found %s", path); return 1; }
  while (fgets(line, sizeof line, f)) {
    char* p = line; char* e;
    long gid = strtol(p, &e, 10);
    if (e == p) continue;
    p = e; m->x[gid] = strtod(p, &e);
    p = e; m->y[gid] = strtod(p, &e);
    if (m->dim == 3) { p = e; m->z[gid] = strtod(p, &e); }
  }
  fclose(f);
  m->graph = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->nSample * m->ncols);
  snprintf(path, sizeof path, "%s/connectivity.dat", dir); /* mesh_read_connectivity.hpp:54-83 */
  f = fopen(path, "r");
  if (!f) { snprintf(g_err, sizeof g_err, "file not found %s", path); return 1; }
  int32_t count = 0;
  while (fgets(line, sizeof line, f) && count < m->nSample) {
    char* p = line; char* e;
    int ok = 1;
    for (int c = 0; c < m->ncols; ++c) {
      long v = strtol(p, &e, 10);
      if (e == p) { ok = 0; break; }
      m->graph[(size_t)count * m->ncols + c] = (int32_t)v;
      p = e;
    }
    if (ok) ++count;
  }
  fclose(f);
  classify(m);
  return 0;
}

/* ================================================================================================== leaf math */
/* impl/weno5.hpp:56-178 */
void or_weno5(double* uNeg, double* uPos, double qim2, double qim1, double qi, double qip1, double qip2, double qip3) {
  const double epsilon = 1e-6, one = 1, two = 2, three = 3;
  const double four = two * two, five = three + two, six = three * two, seven = four + three, ten = five * two;
  const double eleven = five + six, twelve = six * two, thirteen = six + seven;
  const double oneOvfour = one / four, oneOvsix = one / six, oneOvten = one / ten, threeOvten = three / ten;
  const double sixOvten = six / ten, thirteenOvtwelve = thirteen / twelve;
  {
    const double p0 = (two * qim2 - seven * qim1 + eleven * qi) * oneOvsix;
    const double p1 = (-qim1 + five * qi + two * qip1) * oneOvsix;
    const double p2 = (two * qi + five * qip1 - qip2) * oneOvsix;
    const double B0 = thirteenOvtwelve * (qim2 - two * qim1 + qi) * (qim2 - two * qim1 + qi) +
                      oneOvfour * (qim2 - four * qim1 + three * qi) * (qim2 - four * qim1 + three * qi);
    const double B1 = thirteenOvtwelve * (qim1 - two * qi + qip1) * (qim1 - two * qi + qip1) +
                      oneOvfour * (qim1 - qip1) * (qim1 - qip1);
    const double B2 = thirteenOvtwelve * (qi - two * qip1 + qip2) * (qi - two * qip1 + qip2) +
                      oneOvfour * (three * qi - four * qip1 + qip2) * (three * qi - four * qip1 + qip2);
    const double alpha0 = oneOvten / (epsilon * epsilon + 2. * epsilon * B0 + B0 * B0);
    const double alpha1 = sixOvten / (epsilon * epsilon + 2. * epsilon * B1 + B1 * B1);
    const double alpha2 = threeOvten / (epsilon * epsilon + 2. * epsilon * B2 + B2 * B2);
    const double alphaSInv = one / (alpha0 + alpha1 + alpha2);
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv, w2 = alpha2 * alphaSInv;
    *uNeg = w0 * p0 + w1 * p1 + w2 * p2;
  }
  {
    const double p0 = (-qim1 + five * qi + two * qip1) * oneOvsix;
    const double p1 = (two * qi + five * qip1 - qip2) * oneOvsix;
    const double p2 = (eleven * qip1 - seven * qip2 + two * qip3) * oneOvsix;
    const double B0 = thirteenOvtwelve * (qim1 - two * qi + qip1) * (qim1 - two * qi + qip1) +
                      oneOvfour * (qim1 - four * qi + three * qip1) * (qim1 - four * qi + three * qip1);
    const double B1 = thirteenOvtwelve * (qi - two * qip1 + qip2) * (qi - two * qip1 + qip2) +
                      oneOvfour * (qi - qip2) * (qi - qip2);
    const double B2 = thirteenOvtwelve * (qip1 - two * qip2 + qip3) * (qip1 - two * qip2 + qip3) +
                      oneOvfour * (three * qip1 - four * qip2 + qip3) * (three * qip1 - four * qip2 + qip3);
    const double alpha0 = threeOvten / (epsilon * epsilon + 2. * epsilon * B0 + B0 * B0);
    const double alpha1 = sixOvten / (epsilon * epsilon + 2. * epsilon * B1 + B1 * B1);
    const double alpha2 = oneOvten / (epsilon * epsilon + 2. * epsilon * B2 + B2 * B2);
    const double alphaSInv = one / (alpha0 + alpha1 + alpha2);
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv, w2 = alpha2 * alphaSInv;
    *uPos = w0 * p0 + w1 * p1 + w2 * p2;
  }
}

/* impl/weno3.hpp:56-114 */
void or_weno3(double* uNeg, double* uPos, double qim1, double qi, double qip1, double qip2) {
  const double epsilon = 1.e-6, one = 1., two = 2., three = 3.;
  const double oneOvtwo = one / two, oneOvthree = one / three, twoOvthree = two / three;
  {
    const double p0 = (-qim1 + three * qi) * oneOvtwo;
    const double p1 = (qi + qip1) * oneOvtwo;
    const double B0 = (qim1 - qi) * (qim1 - qi);
    const double B1 = (qi - qip1) * (qi - qip1);
    const double alpha0 = oneOvthree / (epsilon * epsilon + two * epsilon * B0 + B0 * B0);
    const double alpha1 = twoOvthree / (epsilon * epsilon + two * epsilon * B1 + B1 * B1);
    const double alphaSInv = one / (alpha0 + alpha1);
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv;
    *uNeg = w0 * p0 + w1 * p1;
  }
  {
    const double p0 = (qi + qip1) * oneOvtwo;
    const double p1 = (three * qip1 - qip2) * oneOvtwo;
    const double B0 = (qi - qip1) * (qi - qip1);
    const double B1 = (qip1 - qip2) * (qip1 - qip2);
    const double alpha0 = twoOvthree / (epsilon * epsilon + two * epsilon * B0 + B0 * B0);
    const double alpha1 = oneOvthree / (epsilon * epsilon + two * epsilon * B1 + B1 * B1);
    const double alphaSInv = one / (alpha0 + alpha1);
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv;
    *uPos = w0 * p0 + w1 * p1;
  }
}

/* impl/weno3.hpp:116-246 (gradients w.r.t. qim1,qi,qip1,qip2; duNeg[3] = 0, duPos[0] = 0) */
void or_weno3_grad(double* uNeg, double* uPos, double* duNeg_dq, double* duPos_dq, double qim1, double qi,
                   double qip1, double qip2) {
  const double epsilon = 1e-6, one = 1, two = 2, three = 3;
  const double oneOvtwo = one / two, oneOvthree = one / three, twoOvthree = two / three;
  {
    double dp0_dq[3], dp1_dq[3], dalpha0_dq[3], dalpha1_dq[3], dalphaSInv_dq[3];
    const double p0 = (-qim1 + three * qi) * oneOvtwo;
    const double p1 = (qi + qip1) * oneOvtwo;
    dp0_dq[0] = -1. / 2.; dp0_dq[1] = 3. / 2.; dp0_dq[2] = 0.;
    dp1_dq[0] = 0.; dp1_dq[1] = 1. / 2.; dp1_dq[2] = 1. / 2.;
    const double B0 = (qim1 - qi) * (qim1 - qi);
    const double B1 = (qi - qip1) * (qi - qip1);
    const double alpha0 = oneOvthree / (epsilon * epsilon + two * epsilon * B0 + B0 * B0);
    const double alpha1 = twoOvthree / (epsilon * epsilon + two * epsilon * B1 + B1 * B1);
    const double alphaSInv = one / (alpha0 + alpha1);
    dalpha0_dq[0] = -4. * (qim1 - qi) / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 3));
    dalpha0_dq[1] = 4. * (qim1 - qi) / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 3.));
    dalpha0_dq[2] = 0.;
    dalpha1_dq[0] = 0.;
    dalpha1_dq[1] = -8. * (qi - qip1) / (3. * pow(pow(qi - qip1, 2.) + epsilon, 3.));
    dalpha1_dq[2] = 8. * (qi - qip1) / (3. * pow(pow(qi - qip1, 2.) + epsilon, 3.));
    dalphaSInv_dq[0] = (4. * (qim1 - qi)) / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 3.) *
                       pow(2. / (3. * pow(pow(qi - qip1, 2.) + epsilon, 2.)) + 1. / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 2.)), 2.));
    dalphaSInv_dq[1] = -((4. * (qim1 - qi)) / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 3.)) -
                         (8. * (qi - qip1)) / (3. * pow(pow(qi - qip1, 2.) + epsilon, 3.))) /
                       pow(2. / (3. * pow(pow(qi - qip1, 2.) + epsilon, 2.)) + 1. / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 2.)), 2.);
    dalphaSInv_dq[2] = -(8. * (qi - qip1)) /
                       (3. * pow(2. / (3. * pow(pow(qi - qip1, 2.) + epsilon, 2.)) + 1. / (3. * pow(pow(qim1 - qi, 2.) + epsilon, 2.)), 2.) *
                        pow(pow(qi - qip1, 2.) + epsilon, 3.));
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv;
    for (int i = 0; i < 3; i++) {
      duNeg_dq[i] = (dalpha0_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha0) * p0 + dp0_dq[i] * w0;
      duNeg_dq[i] += (dalpha1_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha1) * p1 + dp1_dq[i] * w1;
    }
    duNeg_dq[3] = 0.;
    *uNeg = w0 * p0 + w1 * p1;
  }
  {
    double dp0_dq[3], dp1_dq[3], dalpha0_dq[3], dalpha1_dq[3], dalphaSInv_dq[3];
    const double p0 = (qi + qip1) * oneOvtwo;
    const double p1 = (three * qip1 - qip2) * oneOvtwo;
    const double B0 = (qi - qip1) * (qi - qip1);
    const double B1 = (qip1 - qip2) * (qip1 - qip2);
    const double alpha0 = twoOvthree / (epsilon * epsilon + two * epsilon * B0 + B0 * B0);
    const double alpha1 = oneOvthree / (epsilon * epsilon + two * epsilon * B1 + B1 * B1);
    const double alphaSInv = one / (alpha0 + alpha1);
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv;
    dp0_dq[0] = 1. / 2.; dp0_dq[1] = 1 / 2.; dp0_dq[2] = 0.;
    dp1_dq[0] = 0.; dp1_dq[1] = 3. / 2.; dp1_dq[2] = -1. / 2.;
    dalpha0_dq[0] = -8. * (qi - qip1) / (3. * pow(pow(qi - qip1, 2.) + epsilon, 3.));
    dalpha0_dq[1] = 8. * (qi - qip1) / (3. * pow(pow(qi - qip1, 2.) + epsilon, 3.));
    dalpha0_dq[2] = 0.;
    dalpha1_dq[0] = 0.;
    dalpha1_dq[1] = -4. * (qip1 - qip2) / (3. * pow(pow(qip1 - qip2, 2.) + epsilon, 3.));
    dalpha1_dq[2] = 4. * (qip1 - qip2) / (3. * pow(pow(qip1 - qip2, 2) + epsilon, 3.));
    dalphaSInv_dq[0] = 8. * (qi - qip1) / (3. * pow(pow(qi - qip1, 2.) + epsilon, 3.) *
                       pow(1. / (3. * pow(pow(qip1 - qip2, 2.) + epsilon, 2.)) + 2. / (3. * pow(pow(qi - qip1, 2.) + epsilon, 2.)), 2.));
    dalphaSInv_dq[1] = -(8. * (qi - qip1) / (3 * pow(pow(qi - qip1, 2.) + epsilon, 3.)) -
                         4. * (qip1 - qip2) / (3. * pow(pow(qip1 - qip2, 2.) + epsilon, 3.))) /
                       pow(1. / (3. * pow(pow(qip1 - qip2, 2.) + epsilon, 2.)) + 2. / (3. * pow(pow(qi - qip1, 2.) + epsilon, 2.)), 2.);
    dalphaSInv_dq[2] = -4. * (qip1 - qip2) /
                       (3. * pow(1. / (3. * pow(pow(qip1 - qip2, 2.) + epsilon, 2.)) + 2. / (3. * pow(pow(qi - qip1, 2.) + epsilon, 2.)), 2.) *
                        pow(pow(qip1 - qip2, 2.) + epsilon, 3.));
    for (int i = 0; i < 3; i++) {
      duPos_dq[i + 1] = (dalpha0_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha0) * p0 + dp0_dq[i] * w0;
      duPos_dq[i + 1] += (dalpha1_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha1) * p1 + dp1_dq[i] * w1;
    }
    duPos_dq[0] = 0.;
    *uPos = w0 * p0 + w1 * p1;
  }
}

/* impl/weno5.hpp:180-434 (gradients w.r.t. qim2..qip3; duNeg[5] = 0, duPos[0] = 0) */
void or_weno5_grad(double* uNeg, double* uPos, double* duNeg_dq, double* duPos_dq, double qim2, double qim1,
                   double qi, double qip1, double qip2, double qip3) {
  const double epsilon = 1e-6, one = 1, two = 2, three = 3;
  const double four = two * two, five = three + two, six = three * two, seven = four + three, ten = five * two;
  const double eleven = five + six, twelve = six * two, thirteen = six + seven;
  const double oneOvfour = one / four, oneOvsix = one / six, oneOvten = one / ten, threeOvten = three / ten;
  const double sixOvten = six / ten, thirteenOvtwelve = thirteen / twelve;
  {
    const double p0 = (two * qim2 - seven * qim1 + eleven * qi) * oneOvsix;
    const double p1 = (-qim1 + five * qi + two * qip1) * oneOvsix;
    const double p2 = (two * qi + five * qip1 - qip2) * oneOvsix;
    const double dp0_dq[5] = {1. / 3., -7. / 6., 11. / 6., 0., 0.};
    const double dp1_dq[5] = {0., -1. / 6., 5. / 6., 1. / 3., 0.};
    const double dp2_dq[5] = {0., 0., 1. / 3., 5. / 6., -1. / 6.};
    const double B0 = thirteenOvtwelve * pow(qim2 - two * qim1 + qi, two) + oneOvfour * pow(qim2 - four * qim1 + three * qi, two);
    const double B1 = thirteenOvtwelve * pow(qim1 - two * qi + qip1, two) + oneOvfour * pow(qim1 - qip1, two);
    const double B2 = thirteenOvtwelve * pow(qi - two * qip1 + qip2, two) + oneOvfour * pow(three * qi - four * qip1 + qip2, two);
    double dB0_dq[5], dB1_dq[5], dB2_dq[5];
    dB0_dq[0] = (13. * (qim2 - 2. * qim1 + qi)) / 6. + (qim2 - 4. * qim1 + 3. * qi) / 2.;
    dB0_dq[1] = -(13. * (qim2 - 2. * qim1 + qi)) / 3. - 2. * (qim2 - 4. * qim1 + 3. * qi);
    dB0_dq[2] = (13. * (qim2 - 2. * qim1 + qi)) / 6. + (3. * (qim2 - 4. * qim1 + 3. * qi)) / 2.;
    dB0_dq[3] = 0.; dB0_dq[4] = 0.;
    dB1_dq[0] = 0.;
    dB1_dq[1] = (13. * (qip1 + qim1 - 2. * qi)) / 6. + (qim1 - qip1) / 2.;
    dB1_dq[2] = -(13. * (qip1 + qim1 - 2. * qi)) / 3.;
    dB1_dq[3] = (13. * (qip1 + qim1 - 2. * qi)) / 6. - (qim1 - qip1) / 2.;
    dB1_dq[4] = 0.;
    dB2_dq[0] = 0.; dB2_dq[1] = 0.;
    dB2_dq[2] = (13. * (qip2 - 2. * qip1 + qi)) / 6. + (3. * (qip2 - 4. * qip1 + 3. * qi)) / 2.;
    dB2_dq[3] = -(13. * (qip2 - 2. * qip1 + qi)) / 3. - 2. * (qip2 - 4. * qip1 + 3. * qi);
    dB2_dq[4] = (13. * (qip2 - 2. * qip1 + qi)) / 6. + (qip2 - 4. * qip1 + 3. * qi) / 2.;
    const double alpha0 = oneOvten / pow(epsilon + B0, two);
    const double alpha1 = sixOvten / pow(epsilon + B1, two);
    const double alpha2 = threeOvten / pow(epsilon + B2, two);
    const double alphaSInv = one / (alpha0 + alpha1 + alpha2);
    double dalpha0_dq[5], dalpha1_dq[5], dalpha2_dq[5], dalphaSInv_dq[5];
    for (int i = 0; i < 5; i++) {
      dalpha0_dq[i] = -1. / (5. * pow(B0 + epsilon, 3.)) * dB0_dq[i];
      dalpha1_dq[i] = -6. / (5. * pow(B1 + epsilon, 3.)) * dB1_dq[i];
      dalpha2_dq[i] = -3. / (5. * pow(B2 + epsilon, 3.)) * dB2_dq[i];
      dalphaSInv_dq[i] = -1. / pow(alpha2 + alpha1 + alpha0, 2.) * (dalpha0_dq[i] + dalpha1_dq[i] + dalpha2_dq[i]);
    }
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv, w2 = alpha2 * alphaSInv;
    for (int i = 0; i < 5; i++) {
      duNeg_dq[i] = (dalpha0_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha0) * p0 + dp0_dq[i] * w0;
      duNeg_dq[i] += (dalpha1_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha1) * p1 + dp1_dq[i] * w1;
      duNeg_dq[i] += (dalpha2_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha2) * p2 + dp2_dq[i] * w2;
    }
    duNeg_dq[5] = 0.;
    *uNeg = w0 * p0 + w1 * p1 + w2 * p2;
  }
  {
    const double p0 = (-qim1 + five * qi + two * qip1) * oneOvsix;
    const double p1 = (two * qi + five * qip1 - qip2) * oneOvsix;
    const double p2 = (eleven * qip1 - seven * qip2 + two * qip3) * oneOvsix;
    const double dp0_dq[5] = {-1. / 6., 5. / 6., 1. / 3., 0., 0.};
    const double dp1_dq[5] = {0., 1. / 3., 5. / 6., -1. / 6., 0.};
    const double dp2_dq[5] = {0., 0., 11. / 6., -7. / 6., 1. / 3.};
    const double B0 = thirteenOvtwelve * pow(qim1 - two * qi + qip1, two) + oneOvfour * pow(qim1 - four * qi + three * qip1, two);
    const double B1 = thirteenOvtwelve * pow(qi - two * qip1 + qip2, two) + oneOvfour * pow(qi - qip2, two);
    const double B2 = thirteenOvtwelve * pow(qip1 - two * qip2 + qip3, two) + oneOvfour * pow(three * qip1 - four * qip2 + qip3, two);
    double dB0_dq[5], dB1_dq[5], dB2_dq[5];
    dB0_dq[0] = (3. * qip1 + qim1 - 4. * qi) / 2. + (13. * (qip1 + qim1 - 2. * qi)) / 6.;
    dB0_dq[1] = -2. * (3. * qip1 + qim1 - 4. * qi) - (13. * (qip1 + qim1 - 2 * qi)) / 3.;
    dB0_dq[2] = (3. * (3. * qip1 + qim1 - 4. * qi)) / 2. + (13. * (qip1 + qim1 - 2. * qi)) / 6.;
    dB0_dq[3] = 0.; dB0_dq[4] = 0.;
    dB1_dq[0] = 0.;
    dB1_dq[1] = (13. * (qip2 - 2 * qip1 + qi)) / 6. + (qi - qip2) / 2.;
    dB1_dq[2] = -(13. * (qip2 - 2. * qip1 + qi)) / 3.;
    dB1_dq[3] = (13. * (qip2 - 2. * qip1 + qi)) / 6. - (qi - qip2) / 2.;
    dB1_dq[4] = 0.;
    dB2_dq[0] = 0.; dB2_dq[1] = 0.;
    dB2_dq[2] = (13. * (qip3 - 2. * qip2 + qip1)) / 6. + (3. * (qip3 - 4. * qip2 + 3. * qip1)) / 2.;
    dB2_dq[3] = -(13. * (qip3 - 2. * qip2 + qip1)) / 3. - 2. * (qip3 - 4. * qip2 + 3. * qip1);
    dB2_dq[4] = (13. * (qip3 - 2. * qip2 + qip1)) / 6. + (qip3 - 4. * qip2 + 3. * qip1) / 2.;
    const double alpha0 = threeOvten / pow(epsilon + B0, two);
    const double alpha1 = sixOvten / pow(epsilon + B1, two);
    const double alpha2 = oneOvten / pow(epsilon + B2, two);
    const double alphaSInv = one / (alpha0 + alpha1 + alpha2);
    double dalpha0_dq[5], dalpha1_dq[5], dalpha2_dq[5], dalphaSInv_dq[5];
    for (int i = 0; i < 5; i++) {
      dalpha0_dq[i] = -3. / (5. * pow(B0 + epsilon, 3.)) * dB0_dq[i];
      dalpha1_dq[i] = -6. / (5. * pow(B1 + epsilon, 3.)) * dB1_dq[i];
      dalpha2_dq[i] = -1. / (5. * pow(B2 + epsilon, 3.)) * dB2_dq[i];
      dalphaSInv_dq[i] = -1. / pow(alpha2 + alpha1 + alpha0, 2.) * (dalpha0_dq[i] + dalpha1_dq[i] + dalpha2_dq[i]);
    }
    const double w0 = alpha0 * alphaSInv, w1 = alpha1 * alphaSInv, w2 = alpha2 * alphaSInv;
    for (int i = 0; i < 5; i++) {
      duPos_dq[i + 1] = (dalpha0_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha0) * p0 + dp0_dq[i] * w0;
      duPos_dq[i + 1] += (dalpha1_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha1) * p1 + dp1_dq[i] * w1;
      duPos_dq[i + 1] += (dalpha2_dq[i] * alphaSInv + dalphaSInv_dq[i] * alpha2) * p2 + dp2_dq[i] * w2;
    }
    duPos_dq[0] = 0.;
    *uPos = w0 * p0 + w1 * p1 + w2 * p2;
  }
}

/* impl/euler_rusanov_flux_values_function.hpp:54-208 (ndpc = 3,4,5; n has ndpc-2 entries) */
void or_euler_flux(int ndpc, double* F, const double* qL, const double* qR, const double* n, double gamma) {
  const double half = 0.5, es = 1.e-30;
  const int nv = ndpc - 2, ie = ndpc - 1;
  double FL[5], FR[5], vL[3], vR[3];
  const double rL = qL[0], rR = qR[0];
  double unL = 0, unR = 0, kL = 0, kR = 0;
  for (int m = 0; m < nv; ++m) { vL[m] = qL[1 + m] / (rL + es); vR[m] = qR[1 + m] / (rR + es); }
  if (nv == 1) { unL = vL[0]; unR = vR[0]; kL = vL[0] * vL[0]; kR = vR[0] * vR[0]; }
  else if (nv == 2) {
    unL = vL[0] * n[0] + vL[1] * n[1]; unR = vR[0] * n[0] + vR[1] * n[1];
    kL = vL[0] * vL[0] + vL[1] * vL[1]; kR = vR[0] * vR[0] + vR[1] * vR[1];
  } else {
    unL = vL[0] * n[0] + vL[1] * n[1] + vL[2] * n[2]; unR = vR[0] * n[0] + vR[1] * n[1] + vR[2] * n[2];
    kL = vL[0] * vL[0] + vL[1] * vL[1] + vL[2] * vL[2]; kR = vR[0] * vR[0] + vR[1] * vR[1] + vR[2] * vR[2];
  }
  const double pL = (gamma - 1) * (qL[ie] - half * rL * (kL));
  const double HL = (qL[ie] + pL) / rL;
  const double pR = (gamma - 1) * (qR[ie] - half * rR * (kR));
  const double HR = (qR[ie] + pR) / rR;
  FL[0] = rL * unL; FR[0] = rR * unR;
  for (int m = 0; m < nv; ++m) {
    if (nv == 1) { FL[1] = rL * vL[0] * vL[0] + pL; FR[1] = rR * vR[0] * vR[0] + pR; }
    else { FL[1 + m] = rL * unL * vL[m] + pL * n[m]; FR[1 + m] = rR * unR * vR[m] + pR * n[m]; }
  }
  FL[ie] = rL * unL * HL; FR[ie] = rR * unR * HR;
  const double RT = sqrt(rR / (rL));
  double v[3], k = 0;
  for (int m = 0; m < nv; ++m) v[m] = (vL[m] + RT * vR[m]) / (1. + RT);
  if (nv == 1) k = v[0] * v[0]; else if (nv == 2) k = v[0] * v[0] + v[1] * v[1]; else k = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double H = (HL + RT * HR) / (1. + RT);
  const double a = sqrt((gamma - 1.) * (H - half * (k)));
  const double smax = (nv == 1) ? fabs(v[0]) + a : sqrt(k) + a;
  for (int i = 0; i < ndpc; ++i) F[i] = half * (FL[i] + FR[i] + smax * (qL[i] - qR[i]));
}

/* impl/euler_rusanov_flux_jacobian_function.hpp:54-406 ; JL, JR row-major [ndpc][ndpc] */
void or_euler_flux_jac(int N, double* JL, double* JR, const double* qL, const double* qR, const double* nIn, double gamma) {
  const double one = 1, two = 2, three = 3, half = 0.5, es = 1.e-30;
  const double gm1 = gamma - one;
  const int nv = N - 2, ie = N - 1;
  double n[3] = {0, 0, 0};
  if (nv == 1) n[0] = 1.0; else for (int m = 0; m < nv; ++m) n[m] = nIn[m];
  double vL[3] = {0, 0, 0}, vR[3] = {0, 0, 0}, v[3] = {0, 0, 0};
  const double rL = qL[0], rR = qR[0];
  for (int m = 0; m < nv; ++m) { vL[m] = qL[1 + m] / (rL + es); vR[m] = qR[1 + m] / (rR + es); }
  double unL, unR, kL, kR;
  if (nv == 1) { unL = vL[0]; unR = vR[0]; kL = vL[0] * vL[0]; kR = vR[0] * vR[0]; }
  else if (nv == 2) {
    unL = vL[0] * n[0] + vL[1] * n[1]; unR = vR[0] * n[0] + vR[1] * n[1];
    kL = vL[0] * vL[0] + vL[1] * vL[1]; kR = vR[0] * vR[0] + vR[1] * vR[1];
  } else {
    unL = vL[0] * n[0] + vL[1] * n[1] + vL[2] * n[2]; unR = vR[0] * n[0] + vR[1] * n[1] + vR[2] * n[2];
    kL = vL[0] * vL[0] + vL[1] * vL[1] + vL[2] * vL[2]; kR = vR[0] * vR[0] + vR[1] * vR[1] + vR[2] * vR[2];
  }
  const double pL = gm1 * (qL[ie] - half * rL * (kL));
  const double HL = (qL[ie] + pL) / rL;
  const double aL = sqrt(gm1 * (HL - half * (kL)));
  const double pR = gm1 * (qR[ie] - half * rR * (kR));
  const double HR = (qR[ie] + pR) / rR;
  const double aR = sqrt(gm1 * (HR - half * (kR)));
  const double r = sqrt(rR * rL);
  const double RT = sqrt(rR / (rL));
  for (int m = 0; m < nv; ++m) v[m] = (vL[m] + RT * vR[m]) / (one + RT);
  const double H = (HL + RT * HR) / (one + RT);
  double k;
  if (nv == 1) k = v[0] * v[0]; else if (nv == 2) k = v[0] * v[0] + v[1] * v[1]; else k = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const double a = sqrt(gm1 * (H - half * (k)));
  const double smax = (nv == 1) ? fabs(v[0]) + a : sqrt(k) + a;
  double gradL[5], gradR[5];
  const double VMagSqrRoe = k + es;
  const double VMagSqrL = kL, VMagSqrR = kR;
  double VVRoeL, VVRoeR;
  if (nv == 1) { VVRoeL = vL[0] * v[0]; VVRoeR = vR[0] * v[0]; }
  else if (nv == 2) { VVRoeL = vL[0] * v[0] + vL[1] * v[1]; VVRoeR = vR[0] * v[0] + vR[1] * v[1]; }
  else { VVRoeL = vL[0] * v[0] + vL[1] * v[1] + vL[2] * v[2]; VVRoeR = vR[0] * v[0] + vR[1] * v[1] + vR[2] * v[2]; }
  double rel[3];
  for (int m = 0; m < nv; ++m) rel[m] = v[m] / sqrt(VMagSqrRoe);
  double sL = 0, sR = 0;
  if (nv == 1) { sL = -half * (vL[0] + v[0]) * rel[0]; sR = -half * (vR[0] + v[0]) * rel[0]; }
  else if (nv == 2) {
    sL = -half * (vL[0] + v[0]) * rel[0] - half * (vL[1] + v[1]) * rel[1];
    sR = -half * (vR[0] + v[0]) * rel[0] - half * (vR[1] + v[1]) * rel[1];
  } else {
    sL = -half * (vL[0] + v[0]) * rel[0] - half * (vL[1] + v[1]) * rel[1] - half * (vL[2] + v[2]) * rel[2];
    sR = -half * (vR[0] + v[0]) * rel[0] - half * (vR[1] + v[1]) * rel[1] - half * (vR[2] + v[2]) * rel[2];
  }
  gradL[0] = one / (rL + r) * (sL + half * gm1 / a * (half * (VMagSqrRoe + VVRoeL) + half * (HL - H) - aL * aL / gm1 + half * (gamma - two) * VMagSqrL));
  gradR[0] = one / (rR + r) * (sR + half * gm1 / a * (half * (VMagSqrRoe + VVRoeR) + half * (HR - H) - aR * aR / gm1 + half * (gamma - two) * VMagSqrR));
  for (int m = 0; m < nv; ++m) {
    gradL[1 + m] = one / (rL + r) * (rel[m] - half * (gm1 * (v[m] + gm1 * vL[m])) / (a));
    gradR[1 + m] = one / (rR + r) * (rel[m] - half * (gm1 * (v[m] + gm1 * vR[m])) / (a));
  }
  gradL[ie] = half / (rL + r) * gamma * gm1 / (a);
  gradR[ie] = half / (rR + r) * gamma * gm1 / (a);

  for (int side = 0; side < 2; ++side) {
    double* J = side == 0 ? JL : JR;
    const double* vel = side == 0 ? vL : vR;
    const double un = side == 0 ? unL : unR, k2 = side == 0 ? kL : kR, Hs = side == 0 ? HL : HR;
    if (nv == 1) {
      const double u = vel[0];
      J[0] = 0; J[1] = half; J[2] = 0;
      J[3] = half * (half * gm1 * u * u - u * u);
      J[4] = half * ((three - gamma) * u);
      J[5] = half * gm1;
      J[6] = half * ((half * gm1 * u * u - Hs) * u);
      J[7] = half * (Hs - gm1 * u * u);
      J[8] = half * gamma * u;
    } else {
      J[0] = 0;
      for (int j = 0; j < nv; ++j) J[1 + j] = half * n[j];
      J[ie] = 0;
      for (int i = 0; i < nv; ++i) {
        J[(1 + i) * N] = half * (half * gm1 * k2 * n[i] - vel[i] * un);
        for (int j = 0; j < nv; ++j) {
          if (i == j) J[(1 + i) * N + 1 + j] = half * (vel[i] * n[j] - gm1 * vel[j] * n[i] + un);
          else J[(1 + i) * N + 1 + j] = half * (vel[i] * n[j] - gm1 * vel[j] * n[i]);
        }
        J[(1 + i) * N + ie] = half * gm1 * n[i];
      }
      J[ie * N] = half * ((half * gm1 * k2 - Hs) * un);
      for (int j = 0; j < nv; ++j) J[ie * N + 1 + j] = half * (Hs * n[j] - gm1 * vel[j] * un);
      J[ie * N + ie] = half * gamma * un;
    }
  }
  for (int i = 0; i < N; i++) {
    JL[i * N + i] += half * smax;
    for (int j = 0; j < N; j++) JL[i * N + j] += half * gradL[j] * (qL[i] - qR[i]);
  }
  for (int i = 0; i < N; i++) {
    JR[i * N + i] -= half * smax;
    for (int j = 0; j < N; j++) JR[i * N + j] += half * gradR[j] * (qL[i] - qR[i]);
  }
}

/* impl/swe_rusanov_flux_values_function.hpp:54-97 */
void or_swe_flux(double* F, const double* qL, const double* qR, const double* n, double gravity) {
  const double half = 0.5, es = 1.e-30;
  double FL[3], FR[3];
  const double hL = qL[0];
  const double uL = qL[1] / (hL + es), vL = qL[2] / (hL + es);
  const double unL = uL * n[0] + vL * n[1];
  const double pL = 0.5 * gravity * hL * hL;
  FL[0] = hL * unL; FL[1] = hL * unL * uL + pL * n[0]; FL[2] = hL * unL * vL + pL * n[1];
  const double hR = qR[0];
  const double uR = qR[1] / (hR + es), vR = qR[2] / (hR + es);
  const double unR = uR * n[0] + vR * n[1];
  const double pR = 0.5 * gravity * hR * hR;
  FR[0] = hR * unR; FR[1] = hR * unR * uR + pR * n[0]; FR[2] = hR * unR * vR + pR * n[1];
  const double hm = 0.5 * (hL + hR);
  const double um = (unL * pow(hL, 0.5) + unR * pow(hR, 0.5)) / (pow(hL, 0.5) + pow(hR, 0.5) + es);
  const double smax = sqrt(um * um) + sqrt(pow(sqrt(gravity * hm), 2.));
  for (int i = 0; i < 3; ++i) F[i] = half * (FL[i] + FR[i] + smax * (qL[i] - qR[i]));
}

/* impl/swe_rusanov_flux_jacobian_function.hpp:54-136 */
void or_swe_flux_jac(double* JL, double* JR, const double* qL, const double* qR, const double* n, double g) {
  const double es = 1.e-30;
  const double hL = qL[0], uL = qL[1] / (hL + es), vL = qL[2] / (hL + es);
  const double unL = uL * n[0] + vL * n[1];
  const double hR = qR[0], uR = qR[1] / (hR + es), vR = qR[2] / (hR + es);
  const double unR = uR * n[0] + vR * n[1];
  const double hm = 0.5 * (hL + hR);
  const double um = (unL * pow(hL, 0.5) + unR * pow(hR, 0.5)) / (pow(hL, 0.5) + pow(hR, 0.5) + es);
  const double smax = fabs(um) + fabs(pow(g * hm, 0.5));
  const double termL = (n[0] * qL[1] + n[1] * qL[2]) / pow(qL[0], 2.);
  const double termR = (n[0] * qR[1] + n[1] * qR[2]) / pow(qR[0], 2.);
  const double hL_sqrt = pow(hL, 0.5), hR_sqrt = pow(hR, 0.5);
  const double hsqrt_un = hL_sqrt * unL + hR_sqrt * unR + es;
  double dsmaxL[3], dsmaxR[3];
  dsmaxL[0] = -fabs(hsqrt_un) / (2. * hL_sqrt * pow(hL_sqrt + hR_sqrt, 2.)) +
              (0.5 * unL / hL_sqrt - hL_sqrt * termL) * hsqrt_un / ((hL_sqrt + hR_sqrt) * fabs(hsqrt_un)) +
              g / (pow(2., 3. / 2.) * pow(g * (hL + hR), 0.5));
  dsmaxL[1] = n[0] * hsqrt_un / (hL_sqrt * (hL_sqrt + hR_sqrt) * fabs(hsqrt_un));
  dsmaxL[2] = n[1] * hsqrt_un / (hL_sqrt * (hL_sqrt + hR_sqrt) * fabs(hsqrt_un));
  dsmaxR[0] = -fabs(hsqrt_un) / (2. * hR_sqrt * pow(hL_sqrt + hR_sqrt, 2.)) +
              (0.5 * unR / hR_sqrt - hR_sqrt * termR) * hsqrt_un / ((hL_sqrt + hR_sqrt) * fabs(hsqrt_un)) +
              g / (pow(2., 3. / 2.) * pow(g * (hL + hR), 0.5));
  dsmaxR[1] = n[0] * hsqrt_un / (hR_sqrt * (hL_sqrt + hR_sqrt) * fabs(hsqrt_un));
  dsmaxR[2] = n[1] * hsqrt_un / (hR_sqrt * (hL_sqrt + hR_sqrt) * fabs(hsqrt_un));
  JL[0] = -0.5 * dsmaxL[0] * (qR[0] - qL[0]) + 0.5 * (n[0] * uL + n[1] * vL - qL[0] * termL) + 0.5 * smax;
  JL[1] = 0.5 * n[0] - 0.5 * dsmaxL[1] * (qR[0] - qL[0]);
  JL[2] = 0.5 * n[1] - 0.5 * dsmaxL[2] * (qR[0] - qL[0]);
  JL[3] = 0.5 * (g * n[0] * qL[0] - qL[1] * termL) - 0.5 * dsmaxL[0] * (qR[1] - qL[1]);
  JL[4] = n[0] * uL + 0.5 * n[1] * vL + 0.5 * smax - 0.5 * dsmaxL[1] * (qR[1] - qL[1]);
  JL[5] = 0.5 * n[1] * uL - 0.5 * dsmaxL[2] * (qR[1] - qL[1]);
  JL[6] = 0.5 * (g * n[1] * qL[0] - qL[2] * termL) - 0.5 * dsmaxL[0] * (qR[2] - qL[2]);
  JL[7] = 0.5 * n[0] * vL - 0.5 * dsmaxL[1] * (qR[2] - qL[2]);
  JL[8] = n[1] * vL + 0.5 * n[0] * uL + 0.5 * smax - 0.5 * dsmaxL[2] * (qR[2] - qL[2]);
  JR[0] = -0.5 * dsmaxR[0] * (qR[0] - qL[0]) + 0.5 * (n[0] * uR + n[1] * vR - qR[0] * termR) - 0.5 * smax;
  JR[1] = 0.5 * n[0] - 0.5 * dsmaxR[1] * (qR[0] - qL[0]);
  JR[2] = 0.5 * n[1] - 0.5 * dsmaxR[2] * (qR[0] - qL[0]);
  JR[3] = 0.5 * (g * n[0] * qR[0] - qR[1] * termR) - 0.5 * dsmaxR[0] * (qR[1] - qL[1]);
  JR[4] = n[0] * uR + 0.5 * n[1] * vR - 0.5 * smax - 0.5 * dsmaxR[1] * (qR[1] - qL[1]);
  JR[5] = 0.5 * n[1] * uR - 0.5 * dsmaxR[2] * (qR[1] - qL[1]);
  JR[6] = 0.5 * (g * n[1] * qR[0] - qR[2] * termR) - 0.5 * dsmaxR[0] * (qR[2] - qL[2]);
  JR[7] = 0.5 * n[0] * vR - 0.5 * dsmaxR[1] * (qR[2] - qL[2]);
  JR[8] = n[1] * vR + 0.5 * n[0] * uR - 0.5 * smax - 0.5 * dsmaxR[2] * (qR[2] - qL[2]);
}

/* ================================================================================================== set-up */
static double energy2(double gm1Inv, int nv, const double* prim) { /* euler_compute_energy.hpp:52-125 */
  double k = 0;
  for (int m = 0; m < nv; ++m) k += prim[1 + m] * prim[1 + m];
  return prim[nv + 1] * gm1Inv + 0.5 * prim[0] * (k);
}

/* impl/euler_rankine_hugoniot.hpp:55-148 */
static void post_shock_at_rest(double post[4], const double pre[4], double angle, double mach, double gamma) {
  const double rho0 = pre[0], p0 = pre[3];
  const double m2 = mach * mach;
  const double rho1 = rho0 * (gamma + 1.) * m2 / (2. + (gamma - 1.) * m2);
  const double p1 = p0 * (1. + 2. * gamma / (gamma + 1.) * (m2 - 1.));
  const double a0 = sqrt(gamma * p0 / rho0), a1 = sqrt(gamma * p1 / rho1);
  const double num = (1. + 0.5 * (gamma - 1.) * m2), den = gamma * m2 - 0.5 * (gamma - 1.);
  const double mrel = sqrt(num / den);
  const double v1 = mrel * a1 - mach * a0 + 0.0;
  post[0] = rho1; post[1] = -v1 * cos(angle); post[2] = -v1 * sin(angle); post[3] = p1;
}

static int euler2d_ic_index(int prob, int icFlag, const char* s) {
  if (prob == E2_NORMAL_SHOCK && icFlag == 1 && !strcmp(s, "mach")) return 0;
  if (prob == E2_CROSS_SHOCK && icFlag == 1) {
    if (!strcmp(s, "crossShockDensity")) return 1;
    if (!strcmp(s, "crossShockInletXVel")) return 2;
    if (!strcmp(s, "crossShockBottomYVel")) return 3;
  }
  if (prob == E2_RIEMANN && icFlag == 1 && !strcmp(s, "riemannTopRightPressure")) return 4;
  if (prob == E2_RIEMANN && icFlag == 2) {
    if (!strcmp(s, "riemannTopRightPressure")) return 5;
    if (!strcmp(s, "riemannTopRightXVel")) return 6;
    if (!strcmp(s, "riemannTopRightYVel")) return 7;
    if (!strcmp(s, "riemannTopRightDensity")) return 8;
    if (!strcmp(s, "riemannBotLeftPressure")) return 9;
  }
  return -1;
}
static int swe_ic_index(int icFlag, const char* s) {
  static const char* n1[3] = {"pulseMagnitude", "pulseX", "pulseY"};
  static const char* n2[6] = {"pulseMagnitude1", "pulseX1", "pulseY1", "pulseMagnitude2", "pulseX2", "pulseY2"};
  if (icFlag == 1) for (int i = 0; i < 3; ++i) if (!strcmp(s, n1[i])) return i;
  if (icFlag == 2) for (int i = 0; i < 6; ++i) if (!strcmp(s, n2[i])) return 3 + i;
  return -1;
}

static int cmp_i32(const void* a, const void* b) {
  const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* initializeJacobian: euler_2d_prob_class.hpp:223-237,315-387 (same rule in euler_1d/3d, swe_2d);
 * diffusion_reaction_2d_prob_class.hpp:180-224.  setFromTriplets -> sorted columns, duplicates merged. */
static void build_pattern(or_problem* p) {
  if (p->rowptr) return; /* built on demand in lattice mode (the BASELINE-size lattices only evaluate velocities) */
  const or_mesh* m = &p->m;
  const int N = p->ndpc;
  const int nnbInner = (p->S - 1) * m->dim, nnbFirst = 2 * m->dim;
  char* isNb = (char*)calloc((size_t)m->nSample + 1, 1);
  for (int32_t i = 0; i < m->nNearBd; ++i) isNb[m->rowsNearBd[i]] = 1;
  p->rowptr = (int32_t*)malloc(sizeof(int32_t) * ((size_t)m->nSample * N + 1));
  long long nnz = 0;
  int32_t ids[20];
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) p->colidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    long long acc = 0;
    for (int32_t r = 0; r < m->nSample; ++r) {
      const int first = (p->family == F_DIFFREAC2D) || (p->family == F_DIFFREAC1D) || (isNb[r] && p->family != F_ADVECTION1D);
      const int ncand = first ? nnbFirst : nnbInner;
      int cnt = 0;
      ids[cnt++] = G(m, r, 0);
      for (int c = 1; c <= ncand; ++c) if (G(m, r, c) != -1) ids[cnt++] = G(m, r, c);
      qsort(ids, (size_t)cnt, sizeof(int32_t), cmp_i32);
      int u = 0;
      for (int i = 0; i < cnt; ++i) if (i == 0 || ids[i] != ids[i - 1]) ids[u++] = ids[i];
      for (int k = 0; k < N; ++k) {
        if (pass == 1) {
          p->rowptr[(size_t)r * N + k] = (int32_t)acc;
          for (int b = 0; b < u; ++b) for (int j = 0; j < N; ++j) p->colidx[acc + b * N + j] = ids[b] * N + j;
        }
        acc += (long long)u * N;
      }
    }
    nnz = acc;
    if (pass == 1) p->rowptr[(size_t)m->nSample * N] = (int32_t)acc;
  }
  p->nnz = nnz;
  free(isNb);
}

static or_problem* finish_create(or_problem* p, int family, int probEnum, int recon, int icFlag, int nParams,
                                 const char* const* names, const double* values) {
  or_mesh* m = &p->m;
  p->family = family; p->prob = probEnum; p->recon = recon; p->icFlag = icFlag;
  p->S = 3 + 2 * recon;
  p->gamma = 1.4;
  static const double e2def[10] = {9, 0.1, 10., 1., 0.4, 1.5, 0.0, 0.0, 1.5, 0.029};
  static const double swedef[9] = {1. / 8, 1, 1, 1. / 10, -2, -2, 1. / 8, 2, 2};
  p->gs[0] = 0.0002; p->gs[1] = 0.00005; p->gs[2] = 0.042; p->gs[3] = 0.062;
  switch (family) {
    case F_EULER1D: p->ndpc = 3; break;
    case F_EULER2D: p->ndpc = 4; memcpy(p->icp, e2def, sizeof e2def); break;
    case F_EULER3D: p->ndpc = 5; break;
    case F_SWE2D: p->ndpc = 3; memcpy(p->icp, swedef, sizeof swedef); p->php[0] = 9.8; p->php[1] = -3; break;
    case F_DIFFREAC2D: p->ndpc = (probEnum == 1) ? 2 : 1; p->S = 3; p->php[0] = 0.01; p->php[1] = 0.01; break;
    case F_DIFFREAC1D: p->ndpc = 1; p->S = 3; p->php[0] = 0.01; p->php[1] = 0.01; break;
    case F_ADVDIFF2D: /* advection_diffusion_2d_parametrization_helpers.hpp:69-82 */
      p->ndpc = 2; p->php[0] = 0.00001; p->icp[0] = 0.5; p->icp[1] = 0.15; p->icp[2] = 0.0; p->icp[3] = -0.2; break;
    case F_ADVDIFFREAC2D: /* advection_diffusion_reaction2d.hpp:124-128 */
      p->ndpc = 1; p->php[0] = 0.5 * cos(M_PI / 3); p->php[1] = 0.5 * sin(M_PI / 3); p->php[2] = 0.001; p->php[3] = 1.0; break;
    case F_ADVECTION1D: p->ndpc = 1; p->php[0] = 1.0; break; /* advection1d.hpp:88-92 */
    default: snprintf(g_err, sizeof g_err, "oracle: unknown family %d", family); or_destroy(p); return NULL;
  }
  for (int i = 0; i < nParams; ++i) {
    int idx = -1;
    if (family == F_EULER2D) {
      if (!strcmp(names[i], "gamma")) { p->gamma = values[i]; continue; }
      idx = euler2d_ic_index(probEnum, icFlag, names[i]);
      if (idx >= 0) p->icp[idx] = values[i];
    } else if (family == F_SWE2D) {
      if (!strcmp(names[i], "gravity")) { p->php[0] = values[i]; continue; }
      if (!strcmp(names[i], "coriolis")) { p->php[1] = values[i]; continue; }
      idx = swe_ic_index(icFlag, names[i]);
      if (idx >= 0) p->icp[idx] = values[i];
    } else if (family == F_DIFFREAC2D && probEnum == 1) {
      static const char* gn[4] = {"Du", "Dv", "F", "k"};
      for (int k = 0; k < 4; ++k) if (!strcmp(names[i], gn[k])) { p->gs[k] = values[i]; idx = k; }
    } else if (family == F_DIFFREAC2D || family == F_DIFFREAC1D) {
      static const char* pn[2] = {"diffusion", "reaction"};
      for (int k = 0; k < 2; ++k) if (!strcmp(names[i], pn[k])) { p->php[k] = values[i]; idx = k; }
    } else if (family == F_ADVDIFF2D) {
      static const char* bn[4] = {"pulseMagnitude", "pulseSpread", "pulseX", "pulseY"};
      if (!strcmp(names[i], "diffusion")) { p->php[0] = values[i]; continue; }
      for (int k = 0; k < 4; ++k) if (!strcmp(names[i], bn[k])) { p->icp[k] = values[i]; idx = k; }
    } else if (family == F_ADVDIFFREAC2D) {
      static const char* an[4] = {"ux", "uy", "diffusion", "sigma"};
      for (int k = 0; k < 4; ++k) if (!strcmp(names[i], an[k])) { p->php[k] = values[i]; idx = k; }
    } else if (family == F_ADVECTION1D) {
      if (!strcmp(names[i], "velocity")) { p->php[0] = values[i]; idx = 0; }
    }
    if (idx < 0) { snprintf(g_err, sizeof g_err, "oracle: invalid parameter %s", names[i]); or_destroy(p); return NULL; }
  }
  if (m->stencil < p->S) { snprintf(g_err, sizeof g_err, "oracle: mesh stencil too small"); or_destroy(p); return NULL; }
  /* allocateGhosts: euler_2d_prob_class.hpp:1226-1271 */
  p->gstride = p->ndpc * ((p->S - 1) / 2);
  for (int s = 0; s < 6; ++s) {
    p->ghost[s] = (double*)malloc(sizeof(double) * (size_t)(m->nNearBd > 0 ? m->nNearBd : 1) * p->gstride);
    for (size_t i = 0; i < (size_t)(m->nNearBd > 0 ? m->nNearBd : 1) * p->gstride; ++i) p->ghost[s][i] = 2.2250738585072014e-308;
  }
  /* default source functors tabulated per sample row: diffusion_reaction1d.hpp:86-97, diffusion_reaction2d.hpp:103-114,
   * advection_diffusion_reaction2d.hpp:87-97 (f = 1) */
  if (family == F_DIFFREAC1D || family == F_ADVDIFFREAC2D || (family == F_DIFFREAC2D && probEnum == 0)) {
    p->src = (double*)malloc(sizeof(double) * (size_t)(m->nSample > 0 ? m->nSample : 1));
    for (int32_t r = 0; r < m->nSample; ++r) {
      const int32_t c = G(m, r, 0);
      const double x = mx(m, c), y = my(m, c);
      if (family == F_DIFFREAC1D) p->src[r] = sin(M_PI * x) * x * x * 4. * cos(4. * M_PI * x);
      else if (family == F_DIFFREAC2D) p->src[r] = sin(M_PI * x * (y - 0.2)) * 4. * sin(4. * M_PI * y * x);
      else p->src[r] = 1.0;
    }
  }
  if (!m->lattice) build_pattern(p);
  return p;
}

void or_set_source(or_problem* p, const double* values) {
  if (p->src) memcpy(p->src, values, sizeof(double) * (size_t)p->m.nSample);
}

or_problem* or_create(const char* meshDir, int family, int probEnum, int recon, int icFlag, int nParams,
                      const char* const* names, const double* values) {
  or_problem* p = (or_problem*)calloc(1, sizeof *p);
  if (read_mesh(&p->m, meshDir)) { or_destroy(p); return NULL; }
  return finish_create(p, family, probEnum, recon, icFlag, nParams, names, values);
}

or_problem* or_create_from_arrays(int dim, int stencil, int32_t nSample, int32_t nStencil, const double dxyz[3],
                                  const double* x, const double* y, const double* z, const int32_t* graph,
                                  int family, int probEnum, int recon, int icFlag, int nParams,
                                  const char* const* names, const double* values) {
  or_problem* p = (or_problem*)calloc(1, sizeof *p);
  or_mesh* m = &p->m;
  m->dim = dim; m->stencil = stencil; m->nSample = nSample; m->nStencil = nStencil;
  m->ncols = (stencil - 1) * dim + 1;
  for (int a = 0; a < 3; ++a) { m->d[a] = a < dim ? dxyz[a] : 0.0; m->dInv[a] = a < dim ? 1. / dxyz[a] : 0.0; }
  m->x = (double*)calloc((size_t)nStencil, sizeof(double));
  m->y = (double*)calloc((size_t)nStencil, sizeof(double));
  m->z = (double*)calloc((size_t)nStencil, sizeof(double));
  memcpy(m->x, x, sizeof(double) * (size_t)nStencil);
  if (y) memcpy(m->y, y, sizeof(double) * (size_t)nStencil);
  if (z) memcpy(m->z, z, sizeof(double) * (size_t)nStencil);
  m->graph = (int32_t*)malloc(sizeof(int32_t) * (size_t)nSample * m->ncols);
  memcpy(m->graph, graph, sizeof(int32_t) * (size_t)nSample * m->ncols);
  classify(m);
  return finish_create(p, family, probEnum, recon, icFlag, nParams, names, values);
}

/* lattice mode: full mesh in natural ordering, n[3] cells (unused axes 1), connectivity by index arithmetic.
 * dxyz as the mesh files carry them (14-decimal roundings); cx/cy/cz per-axis centre coordinates or NULL (then the
 * initial condition, the DMR / cross-shock ghost rules and the Euler2d Jacobian factors, which read coordinates, are
 * not available).  Row classification = mesh_ccu.hpp:385-439 evaluated on indices: a cell is near the boundary when
 * one of its (mesh stencil-1)/2 layers leaves a non-periodic axis. */
or_problem* or_create_lattice(int dim, int stencil, const int32_t n[3], const double dxyz[3], const int32_t periodic[3],
                              const double* cx, const double* cy, const double* cz, int family, int probEnum,
                              int recon, int icFlag, int nParams, const char* const* names, const double* values) {
  or_problem* p = (or_problem*)calloc(1, sizeof(or_problem));
  or_mesh* m = &p->m;
  m->dim = dim; m->stencil = stencil; m->ncols = (stencil - 1) * dim + 1;
  m->lattice = 1;
  long long cells = 1;
  for (int a = 0; a < 3; ++a) {
    m->n[a] = (a < dim) ? n[a] : 1;
    m->per[a] = (a < dim) ? (periodic[a] != 0) : 1;
    m->d[a] = dxyz[a]; m->dInv[a] = (a < dim) ? 1. / dxyz[a] : 0.;
    cells *= m->n[a];
  }
  if (cells > 2147483647LL / 5) { snprintf(g_err, sizeof g_err, "lattice too large for int32 dof indices"); free(p); return NULL; }
  m->nSample = m->nStencil = (int32_t)cells;
  if (cx) { m->cx = (double*)malloc(sizeof(double) * (size_t)m->n[0]); memcpy(m->cx, cx, sizeof(double) * (size_t)m->n[0]); }
  if (cy) { m->cy = (double*)malloc(sizeof(double) * (size_t)m->n[1]); memcpy(m->cy, cy, sizeof(double) * (size_t)m->n[1]); }
  if (cz) { m->cz = (double*)malloc(sizeof(double) * (size_t)m->n[2]); memcpy(m->cz, cz, sizeof(double) * (size_t)m->n[2]); }
  const int h = (stencil - 1) / 2;
  m->periodic = 1;
  for (int a = 0; a < dim; ++a) if (!m->per[a]) m->periodic = 0;
  if (m->periodic) {
    m->nInner = m->nSample; m->nNearBd = 0; m->rowsInner = NULL;
    m->rowsNearBd = (int32_t*)malloc(sizeof(int32_t));
  } else {
    /* two passes: count, then fill (ascending row order, like classify()) */
    int32_t nNb = 0;
    for (int pass = 0; pass < 2; ++pass) {
      int32_t iIn = 0, iNb = 0;
      for (int32_t k = 0; k < m->n[2]; ++k)
        for (int32_t j = 0; j < m->n[1]; ++j)
          for (int32_t i = 0; i < m->n[0]; ++i) {
            const int32_t idx[3] = {i, j, k};
            int bd = 0;
            for (int a = 0; a < dim; ++a) if (!m->per[a] && (idx[a] < h || idx[a] >= m->n[a] - h)) bd = 1;
            const int32_t r = (k * m->n[1] + j) * m->n[0] + i;
            if (pass == 1) { if (bd) m->rowsNearBd[iNb] = r; else m->rowsInner[iIn] = r; }
            if (bd) ++iNb; else ++iIn;
          }
      if (pass == 0) {
        nNb = iNb;
        m->nNearBd = nNb; m->nInner = m->nSample - nNb;
        m->rowsNearBd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nNb > 0 ? nNb : 1));
        m->rowsInner = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m->nInner > 0 ? m->nInner : 1));
      }
    }
  }
  return finish_create(p, family, probEnum, recon, icFlag, nParams, names, values);
}

void or_destroy(or_problem* p) {
  if (!p) return;
  free(p->m.x); free(p->m.y); free(p->m.z); free(p->m.graph); free(p->m.rowsInner); free(p->m.rowsNearBd);
  free(p->m.cx); free(p->m.cy); free(p->m.cz);
  for (int s = 0; s < 6; ++s) free(p->ghost[s]);
  free(p->rowptr); free(p->colidx); free(p->src);
  free(p);
}

long long or_query(or_problem* p, int what) {
  const or_mesh* m = &p->m;
  switch (what) {
    case 0: return m->dim; case 1: return m->stencil; case 2: return m->nSample; case 3: return m->nStencil;
    case 4: return m->ncols; case 5: return m->nInner; case 6: return m->nNearBd; case 7: return m->periodic;
    case 8: return p->ndpc; case 9: return (long long)m->nStencil * p->ndpc; case 10: return (long long)m->nSample * p->ndpc;
    case 11: build_pattern(p); return p->nnz;
  }
  return -1;
}

void or_mesh_arrays(or_problem* p, int32_t* graph, double* x, double* y, double* z, int32_t* rowsInner,
                    int32_t* rowsNearBd, double* d6) {
  const or_mesh* m = &p->m;
  if (graph) memcpy(graph, m->graph, sizeof(int32_t) * (size_t)m->nSample * m->ncols);
  if (x) memcpy(x, m->x, sizeof(double) * (size_t)m->nStencil);
  if (y) memcpy(y, m->y, sizeof(double) * (size_t)m->nStencil);
  if (z) memcpy(z, m->z, sizeof(double) * (size_t)m->nStencil);
  if (rowsInner && m->rowsInner) memcpy(rowsInner, m->rowsInner, sizeof(int32_t) * (size_t)m->nInner);
  else if (rowsInner) for (int32_t i = 0; i < m->nInner; ++i) rowsInner[i] = i;
  if (rowsNearBd) memcpy(rowsNearBd, m->rowsNearBd, sizeof(int32_t) * (size_t)m->nNearBd);
  if (d6) for (int a = 0; a < 3; ++a) { d6[a] = m->d[a]; d6[3 + a] = m->dInv[a]; }
}

void or_pattern(or_problem* p, int32_t* rowptr, int32_t* colidx) {
  build_pattern(p);
  memcpy(rowptr, p->rowptr, sizeof(int32_t) * ((size_t)p->m.nSample * p->ndpc + 1));
  memcpy(colidx, p->colidx, sizeof(int32_t) * (size_t)p->nnz);
}

/* ================================================================================================== initial conditions */
void or_ic(or_problem* p, double* U) {
  const or_mesh* m = &p->m;
  const double gamma = p->gamma, gm1 = gamma - 1., gm1Inv = 1. / (gamma - 1.);
  const int32_t n = m->nStencil;
  double prim[5] = {0, 0, 0, 0, 0};
  if (p->family == F_EULER1D) { /* euler_1d_initial_condition.hpp:55-183 */
    for (int32_t i = 0; i < n; ++i) {
      const double x = mx(m, i);
      if (p->prob == 0) { prim[0] = 1. + 0.2 * sin(M_PI * x); prim[1] = 1.; prim[2] = 1.; }
      else if (p->prob == 1) {
        if (x <= 0.) { prim[0] = 1.; prim[1] = 0.; prim[2] = 1.; }
        if (x > 0.) { prim[0] = 0.125; prim[1] = 0.; prim[2] = 0.1; }
      } else if (p->prob == 2) {
        if (x <= 0.) { prim[0] = 0.445; prim[1] = 0.698; prim[2] = 3.528; }
        else if (x > 0.) { prim[0] = 0.5; prim[1] = 0.; prim[2] = 0.571; }
      } else {
        if (x <= -4.) { prim[0] = 27. / 7.; prim[1] = 2.629369; prim[2] = 31. / 3.; }
        else { prim[0] = 1. + (1. / 5.) * sin(5. * x); prim[1] = 0.; prim[2] = 1.; }
      }
      U[3 * i] = prim[0]; U[3 * i + 1] = prim[0] * prim[1]; U[3 * i + 2] = energy2(gm1Inv, 1, prim);
    }
    return;
  }
  if (p->family == F_EULER2D) { /* euler_2d_initial_condition.hpp:57-530 */
    double pre[4] = {gamma, 0., 0., 1.}, post[4] = {0, 0, 0, 0};
    const int dmr = p->prob == E2_DMR;
    if (dmr) post_shock_at_rest(post, pre, -(M_PI / 6.), 10., gamma);
    if (p->prob == E2_NORMAL_SHOCK) post_shock_at_rest(post, pre, 0., p->icp[0], gamma);
    /* Riemann IC 2 derived states (:319-352) */
    double r2s[4][4];
    if (p->prob == E2_RIEMANN && p->icFlag == 2) {
      const double p1 = p->icp[5], u1 = p->icp[6], v1 = p->icp[7], rho1 = p->icp[8], p3 = p->icp[9];
      const double eps = (gamma - 1.) / (gamma + 1.);
      const double fac13 = p1 / p3;
      const double fac43 = (1. / (2. * (1. + 2. * eps))) * (eps * (fac13 + 1.) + sqrt(pow(eps * (fac13 + 1.), 2) + 4. * (1. + 2. * eps) * fac13));
      const double p2 = fac43 * p3, p4 = p2, v2 = v1, u4 = u1;
      const double rho2 = rho1 * (p2 / p1 + eps) / (1 + eps * p2 / rho1), rho4 = rho2;
      const double psi21 = (p2 - p1) * (rho2 - rho1) / (rho2 * rho1);
      const double u2 = sqrt(psi21) + u1;
      const double psi41 = (p4 - p1) * (rho4 - rho1) / (rho4 * rho1);
      const double v4 = sqrt(psi41) + v1;
      const double u3 = u2, v3 = v4;
      const double rho3 = rho2 * (p3 - p2) / ((p3 - p2) - psi41 * rho2);
      const double t[4][4] = {{rho1, u1, v1, p1}, {rho2, u2, v2, p2}, {rho3, u3, v3, p3}, {rho4, u4, v4, p4}};
      memcpy(r2s, t, sizeof t);
    }
    for (int32_t i = 0; i < n; ++i) {
      const double x = mx(m, i), y = my(m, i);
      double* s = U + 4 * (size_t)i;
      int direct = 0;
      switch (p->prob) {
        case E2_PERIODIC: prim[0] = 1. + (1. / 5.) * sin(M_PI * (x + y)); prim[1] = 1.; prim[2] = 1.; prim[3] = 1.; break;
        case E2_KH: {
          const double pert = 0.025 * cos(2. * 3.14159265 / 10. * 4. * x);
          if (y > -2 + pert && y < 2 + pert) { s[0] = 2.; s[1] = s[0] * 0.5; } else { s[0] = 1.; s[1] = -s[0] * 0.5; }
          s[2] = 0.;
          s[3] = 2.5 / (gm1) + 0.5 / s[0] * (s[1] * s[1] + s[2] * s[2]);
          direct = 1;
          break;
        }
        case E2_SEDOV_FULL: case E2_SEDOV_SYM: {
          const int sym = p->prob == E2_SEDOV_SYM;
          const double sRad = (sym ? 3. : 2.) * (m->d[0] < m->d[1] ? m->d[0] : m->d[1]);
          const double myR = sqrt(x * x + y * y);
          prim[0] = 1.; prim[1] = 0.; prim[2] = 0.;
          if (myR <= sRad) prim[3] = sym ? gm1 * 0.851072 / (M_PI * sRad * sRad) : (gm1) / (M_PI * sRad * sRad);
          else prim[3] = sym ? 2.5e-5 : 5.e-5;
          break;
        }
        case E2_RIEMANN:
          if (p->icFlag == 1) {
            const double x0 = 0.5, y0 = 0.5;
            if (x >= x0 && y >= y0) { prim[0] = 0.5313; prim[1] = 0; prim[2] = 0; prim[3] = p->icp[4]; }
            else if (x < x0 && y >= y0) { prim[0] = 1; prim[1] = 0.7276; prim[2] = 0; prim[3] = 1; }
            else if (x < x0 && y < y0) { prim[0] = 0.8; prim[1] = 0; prim[2] = 0; prim[3] = 1; }
            else if (x > x0 && y < y0) { prim[0] = 1; prim[1] = 0; prim[2] = 0.7276; prim[3] = 1; }
          } else {
            const double x0 = 0.8, y0 = 0.8;
            int q = -1;
            if (x >= x0 && y >= y0) q = 0; else if (x < x0 && y >= y0) q = 1; else if (x < x0 && y < y0) q = 2; else if (x > x0 && y < y0) q = 3;
            if (q >= 0) memcpy(prim, r2s[q], sizeof(double) * 4);
          }
          break;
        case E2_NORMAL_SHOCK: memcpy(prim, (x < 1. / 6.) ? post : pre, sizeof(double) * 4); break;
        case E2_DMR: { const double xShock = 1. / 6. + tan(M_PI / 6.) * y; memcpy(prim, (x < xShock) ? post : pre, sizeof(double) * 4); break; }
        case E2_CROSS_SHOCK: prim[0] = p->icp[1]; prim[1] = p->icp[2]; prim[2] = 0.; prim[3] = 1.; break;
        default: s[0] = s[1] = s[2] = s[3] = 0.; direct = 1; break;
      }
      if (!direct) { s[0] = prim[0]; s[1] = prim[0] * prim[1]; s[2] = prim[0] * prim[2]; s[3] = energy2(gm1Inv, 2, prim); }
    }
    return;
  }
  if (p->family == F_EULER3D) { /* euler_3d_initial_condition.hpp:57-143 */
    const double dmin = fmin(m->d[0], fmin(m->d[1], m->d[2]));
    const double sRad = 3. * dmin;
    for (int32_t i = 0; i < n; ++i) {
      double* s = U + 5 * (size_t)i;
      if (p->prob == 0) { prim[0] = 1.0 + 0.2 * sin(M_PI * (mx(m, i) + my(m, i) + mz(m, i))); prim[1] = prim[2] = prim[3] = 1.0; prim[4] = 1.; }
      else {
        const double myR = sqrt(mx(m, i) * mx(m, i) + my(m, i) * my(m, i) + mz(m, i) * mz(m, i));
        prim[0] = 1.0; prim[1] = prim[2] = prim[3] = 0.0;
        prim[4] = (myR <= sRad) ? (3. * gm1 * 0.851072) / (4. * M_PI * sRad * sRad * sRad) : 2.5e-5;
      }
      s[0] = prim[0]; s[1] = prim[0] * prim[1]; s[2] = prim[0] * prim[2]; s[3] = prim[0] * prim[3]; s[4] = energy2(gm1Inv, 3, prim);
    }
    return;
  }
  if (p->family == F_SWE2D) { /* swe_2d_initial_condition.hpp:55-107 */
    for (int32_t i = 0; i < n; ++i) {
      double* s = U + 3 * (size_t)i;
      if (p->icFlag == 1) {
        const double dx1 = mx(m, i) - p->icp[1], dy1 = my(m, i) - p->icp[2];
        const double r = sqrt(dx1 * dx1 + dy1 * dy1);
        s[0] = 1. + p->icp[0] * exp(-(r * r));
      } else {
        const double dx1 = mx(m, i) - p->icp[4], dy1 = my(m, i) - p->icp[5];
        const double r1 = sqrt(dx1 * dx1 + dy1 * dy1);
        const double dx2 = mx(m, i) - p->icp[7], dy2 = my(m, i) - p->icp[8];
        const double r2 = sqrt(dx2 * dx2 + dy2 * dy2);
        s[0] = 1. + p->icp[3] * exp(-(r1 * r1)) + p->icp[6] * exp(-(r2 * r2));
      }
      s[1] = 0.; s[2] = 0.;
    }
    return;
  }
  if (p->family == F_DIFFREAC1D || p->family == F_ADVDIFFREAC2D || (p->family == F_DIFFREAC2D && p->prob == 0)) {
    /* zero state: diffusion_reaction_1d_prob_class.hpp:110-118, diffusion_reaction_2d_prob_class.hpp:144-149,
     * advection_diffusion_reaction_2d_initial_condition.hpp:57-61 */
    for (int32_t i = 0; i < n; ++i) U[i] = 0.;
    return;
  }
  if (p->family == F_ADVDIFF2D) { /* advection_diffusion_2d_initial_condition.hpp:54-78 */
    for (int32_t i = 0; i < n; ++i) {
      const double dx = mx(m, i) - p->icp[2], dy = my(m, i) - p->icp[3];
      const double dxSq = dx * dx, dySq = dy * dy;
      U[2 * (size_t)i] = p->icp[0] * exp(-(dxSq + dySq) / p->icp[1]);
      U[2 * (size_t)i + 1] = p->icp[0] * exp(-(dxSq + dySq) / p->icp[1]);
    }
    return;
  }
  if (p->family == F_ADVECTION1D) { /* advection_1d_prob_class.hpp:111-156 */
    for (int32_t i = 0; i < n; ++i) {
      const double x = mx(m, i);
      if (p->icFlag == 1) U[i] = sin(M_PI * x);
      else if (p->icFlag == 2) {
        const double dx1Sq = (x - 1.2) * (x - 1.2), dx2Sq = (x - 2.5) * (x - 2.5);
        U[i] = 0.8 * exp(-200.0 * dx1Sq / 16.0) + exp(-100.0 * dx2Sq / 36.0);
      } else if (p->icFlag == 3) {
        const double delta = 0.5 * 0.5;
        const double dx1Sq = (x - 2.) * (x - 2.), dx2Sq = (x - 3.0) * (x - 3.0);
        U[i] = exp(-dx1Sq / delta) + 0.5 * exp(-dx2Sq / delta);
      } else {
        U[i] = tanh(8. * (x - 1.)) - tanh(8. * (x - 3.));
      }
    }
    return;
  }
  if (p->family == F_DIFFREAC2D) { /* diffusion_reaction_2d_prob_class.hpp:141-178 */
    for (int32_t i = 0; i < n; ++i) {
      const int in = fabs(mx(m, i)) < 0.1 && fabs(my(m, i)) < 0.1;
      U[2 * (size_t)i] = in ? 0.5 : 1.; U[2 * (size_t)i + 1] = in ? 0.25 : 0.;
    }
  }
}

/* ================================================================================================== ghost fillers */
static const int kOpp[6] = {2, 3, 0, 1, 5, 4};

/* copy (with per-dof sign) U[src] into ghost[side](gRow, layer) */
static void ghost_copy(or_problem* p, int side, int32_t gRow, int layer, const double* U, int32_t src, int negDof) {
  const int N = p->ndpc;
  double* g = p->ghost[side] + (size_t)gRow * p->gstride + layer * N;
  for (int d = 0; d < N; ++d) g[d] = (d == negDof) ? -U[(size_t)src * N + d] : U[(size_t)src * N + d];
}
static void ghost_const(or_problem* p, int side, int32_t gRow, int layer, const double* v) {
  const int N = p->ndpc;
  memcpy(p->ghost[side] + (size_t)gRow * p->gstride + layer * N, v, sizeof(double) * N);
}

/* source cell of ghost `layer` on `side` for the two mirror styles found in the reference:
 * naive  (euler_1d_ghost_filler.hpp:104-148, euler_2d_ghost_filler_{sedov2d_sym,normal_shock,double_mach_reflection,
 *         cross_shock}.hpp, euler_3d_ghost_filler_sedov.hpp): layer 0 <- self, layer k <- opposite neighbour k-1,
 *         whatever the distance of the cell from the wall;
 * proper (euler_2d_ghost_filler_neumann.hpp:163-275, swe_2d_ghost_filler_inviscid_wall.hpp): branches on which
 *         layers are missing. */
static int32_t mirror_src(const or_mesh* m, int32_t row, int side, int layer, int proper) {
  const int opp = kOpp[side];
  const int32_t self = G(m, row, 0);
  if (layer == 0) return self;
  if (!proper) return G(m, row, gcol(m->dim, opp, layer - 1));
  const int32_t s0 = G(m, row, gcol(m->dim, side, 0));
  if (layer == 1) return (s0 == -1) ? G(m, row, gcol(m->dim, opp, 0)) : s0;
  const int32_t s1 = G(m, row, gcol(m->dim, side, 1));
  int32_t ind = self;
  if (s1 != -1 && s0 != -1) ind = s1;
  if (s1 == -1 && s0 != -1) ind = self;
  if (s1 == -1 && s0 == -1) ind = G(m, row, gcol(m->dim, opp, 1));
  return ind;
}

static void fill_ghosts(or_problem* p, const double* U, double t) {
  const or_mesh* m = &p->m;
  const int h = (p->S - 1) / 2;
  const int dim = m->dim;
  /* which filler */
  int style = 0; /* 0 none, 1 naive, 2 proper */
  int neg[6] = {-1, -1, -1, -1, -1, -1};
  if (p->family == F_EULER1D) style = (p->prob == 0) ? 0 : 1; /* euler_1d_prob_class.hpp:315-335 */
  else if (p->family == F_EULER2D) {
    switch (p->prob) { /* euler_2d_prob_class.hpp:465-565 */
      case E2_SEDOV_FULL: case E2_RIEMANN: case E2_NEUMANN: style = 2; break;
      case E2_SEDOV_SYM: style = 1; neg[0] = 1; neg[3] = 2; break;
      case E2_NORMAL_SHOCK: style = 1; neg[1] = 2; neg[3] = 2; break;
      case E2_DMR: case E2_CROSS_SHOCK: style = 1; break;
      default: style = 0;
    }
  } else if (p->family == F_EULER3D) { /* euler_3d_prob_class.hpp:296-329 */
    if (p->prob == 1) { style = 1; neg[0] = 1; neg[3] = 2; neg[4] = 3; }
  } else if (p->family == F_SWE2D) { /* swe_2d_prob_class.hpp:394-418 */
    style = 2; neg[0] = 1; neg[2] = 1; neg[1] = 2; neg[3] = 2;
  } else if (p->family == F_ADVDIFF2D) { /* advection_diffusion_2d_prob_class.hpp:296-315: outflow only */
    if (p->prob == 1) style = 2;
  } else if (p->family == F_ADVDIFFREAC2D) {
    style = 1;
  }
  if (!style) return;

  /* DMR constants: euler_2d_ghost_filler_double_mach_reflection.hpp:265-291,318-331 */
  double preS[4] = {0, 0, 0, 0}, postS[4] = {0, 0, 0, 0};
  const double wedge = 1. / 6., angle = M_PI / 6.;
  const double shockSpeed = 10. / cos(angle), shockSlope = tan(angle);
  if (p->family == F_EULER2D && p->prob == E2_DMR) {
    const double gm1Inv = 1. / (p->gamma - 1.);
    double pre[4] = {p->gamma, 0., 0., 1.}, post[4];
    post_shock_at_rest(post, pre, -angle, 10., p->gamma);
    preS[0] = pre[0]; preS[1] = pre[0] * pre[1]; preS[2] = pre[0] * pre[2]; preS[3] = energy2(gm1Inv, 2, pre);
    postS[0] = post[0]; postS[1] = post[0] * post[1]; postS[2] = post[0] * post[2]; postS[3] = energy2(gm1Inv, 2, post);
  }
  /* cross shock constants: euler_2d_ghost_filler_cross_shock.hpp:84-96 */
  double dirich[4] = {0, 0, 0, 0};
  const double csRho = p->icp[1], csU = p->icp[2], csV = p->icp[3];
  if (p->family == F_EULER2D && p->prob == E2_CROSS_SHOCK) {
    double prim[4] = {csRho, csU, 0, 1.};
    dirich[0] = prim[0]; dirich[1] = prim[0] * prim[1]; dirich[2] = prim[0] * prim[2];
    dirich[3] = energy2(1. / (p->gamma - 1.), 2, prim);
  }

  for (int32_t it = 0; it < m->nNearBd; ++it) {
    const int32_t row = m->rowsNearBd[it];
    const int32_t self = G(m, row, 0);
    const double myX = mx(m, self), myY = my(m, self);
    for (int side = 0; side < (dim == 1 ? 3 : 2 * dim); ++side) {
      if (dim == 1 && side == 1) continue;
      for (int L = 0; L < h; ++L) {
        if (G(m, row, gcol(dim, side, L)) != -1) continue;
        const int32_t src = mirror_src(m, row, side, L, style == 2);
        if (p->family == F_EULER2D && p->prob == E2_DMR) {
          if (side == 1) { /* top: Dirichlet, time dependent (:112-125,176-189,240-253) */
            const double yIn = myY + (double)(L + 1) * m->d[1];
            const double dist = (myX - wedge - shockSpeed * t - shockSlope * yIn);
            ghost_const(p, side, it, L, dist < 0. ? postS : preS);
          } else if (side == 3) ghost_copy(p, side, it, L, U, src, myX < wedge ? -1 : 2);
          else ghost_copy(p, side, it, L, U, src, -1);
        } else if (p->family == F_EULER2D && p->prob == E2_CROSS_SHOCK) {
          if (side == 0) ghost_const(p, side, it, L, dirich);
          else if (side == 3) { /* :128-143,186-200 */
            double v[4];
            v[0] = (L == 0) ? U[(size_t)self * 4] : csRho;
            v[1] = csRho * csU;
            v[2] = (myX < 0.5) ? 0. : csRho * csV;
            v[3] = U[(size_t)src * 4 + 3];
            ghost_const(p, side, it, L, v);
          } else ghost_copy(p, side, it, L, U, src, -1);
        } else if (p->family == F_ADVDIFF2D) { /* advection_diffusion_2d_ghost_filler_outflow.hpp:102-235 */
          const double zero2[2] = {0., 0.};
          if (side == 0 || side == 3) ghost_const(p, side, it, L, zero2);
          else ghost_copy(p, side, it, L, U, src, -1);
        } else if (p->family == F_ADVDIFFREAC2D) { /* advection_diffusion_reaction_2d_ghost_filler_problemA.hpp:92-147 */
          ghost_copy(p, side, it, L, U, src, 0);
        } else {
          ghost_copy(p, side, it, L, U, src, neg[side]);
        }
      }
    }
  }
}

/* ================================================================================================== evaluation */
/* stencil value: StencilFiller (functor_fill_stencil.hpp:66-1164): neighbour, or ghost(layer) if the neighbour is -1 */
static double sval(const or_problem* p, const double* U, int32_t cell, int side, int32_t gRow, int layer, int dof) {
  if (cell == -1) return p->ghost[side][(size_t)gRow * p->gstride + layer * p->ndpc + dof];
  return U[(size_t)cell * p->ndpc + dof];
}

/* advection_diffusion_2d_flux_functions.hpp:54-75 */
static void burgers_flux(double* F, const double* qL, const double* qR, const double* n) {
  const double fourInv = 1. / 4.;
  const double alpha_0 = fmax(fabs(qL[0]), fabs(qR[0]));
  const double alpha_1 = fmax(fabs(qL[1]), fabs(qR[1]));
  F[0] = alpha_0 * (qL[0] - qR[0]);
  F[0] += n[0] * (qL[0] * qL[0] + qR[0] * qR[0]);
  F[0] += n[1] * (qL[0] * qL[1] + qR[0] * qR[1]);
  F[0] *= fourInv;
  F[1] = alpha_1 * (qL[1] - qR[1]);
  F[1] += n[0] * (qL[0] * qL[1] + qR[0] * qR[1]);
  F[1] += n[1] * (qL[1] * qL[1] + qR[1] * qR[1]);
  F[1] *= fourInv;
}
/* advection_diffusion_2d_flux_functions.hpp:77-114; J row-major 2x2 */
static void burgers_flux_jac(double* JL, double* JR, const double* qL, const double* qR, const double* n) {
  const double two = 2., fourInv = 1. / 4.;
  if (fabs(qL[0]) > fabs(qR[0])) {
    JL[0] = (two * qL[0] - qR[0]) * copysign(1, qL[0]) + n[0] * two * qL[0] + n[1] * qL[1];
    JR[0] = n[0] * two * qR[0] + n[1] * qR[1] - fabs(qL[0]);
  } else {
    JL[0] = n[0] * two * qL[0] + n[1] * qL[1] + fabs(qR[0]);
    JR[0] = (qL[0] - two * qR[0]) * copysign(1, qR[0]) + n[0] * two * qR[0] + n[1] * qR[1];
  }
  JL[0] *= fourInv; JR[0] *= fourInv;
  if (fabs(qL[1]) > fabs(qR[1])) {
    JL[3] = (two * qL[1] - qR[1]) * copysign(1, qL[1]) + n[0] * qL[0] + n[1] * two * qL[1];
    JR[3] = n[0] * qR[0] + n[1] * two * qR[1] - fabs(qL[1]);
  } else {
    JL[3] = n[0] * qL[0] + n[1] * two * qL[1] + fabs(qR[1]);
    JR[3] = (qL[1] - two * qR[1]) * copysign(1, qR[1]) + n[0] * qR[0] + n[1] * two * qR[1];
  }
  JL[3] *= fourInv; JR[3] *= fourInv;
  JL[1] = n[1] * qL[0] * fourInv; JL[2] = n[0] * qL[1] * fourInv;
  JR[1] = n[1] * qR[0] * fourInv; JR[2] = n[0] * qR[1] * fourInv;
}

/* advection velocity of a scalar family along `axis` (advection_1d_mixins.hpp:79-94,
 * advection_diffusion_reaction_2d_flux_mixin.hpp: F = uNeg * a, dF/duNeg = a, dF/duPos = 0) */
static double adv_vel(const or_problem* p, int axis) {
  return (p->family == F_ADVECTION1D) ? p->php[0] : p->php[axis - 1];
}

static void flux(const or_problem* p, int axis, double* F, const double* qL, const double* qR) {
  double n[3] = {0, 0, 0};
  n[axis - 1] = 1.0;
  if (p->family == F_ADVDIFF2D) { burgers_flux(F, qL, qR, n); return; }
  if (p->family == F_ADVECTION1D || p->family == F_ADVDIFFREAC2D) { F[0] = qL[0] * adv_vel(p, axis); return; }
  if (p->family == F_SWE2D) or_swe_flux(F, qL, qR, n, p->php[0]);
  else or_euler_flux(p->ndpc, F, qL, qR, n, p->gamma);
}
static void flux_jac(const or_problem* p, int axis, double* JL, double* JR, const double* qL, const double* qR) {
  double n[3] = {0, 0, 0};
  n[axis - 1] = 1.0;
  if (p->family == F_ADVDIFF2D) { burgers_flux_jac(JL, JR, qL, qR, n); return; }
  if (p->family == F_ADVECTION1D || p->family == F_ADVDIFFREAC2D) { JL[0] = adv_vel(p, axis); JR[0] = 0.; return; }
  if (p->family == F_SWE2D) or_swe_flux_jac(JL, JR, qL, qR, n, p->php[0]);
  else or_euler_flux_jac(p->ndpc, JL, JR, qL, qR, n, p->gamma);
}

/* stencil cells of a row along an axis: [l_{h-1} .. l0, self, r0 .. r_{h-1}] (-1 = outside the domain) */
static void gather_cells(const or_problem* p, int32_t row, int axis, int S, int32_t* cells) {
  const or_mesh* m = &p->m;
  const int h = (S - 1) / 2;
  const int sm = side_minus(axis), sp = side_plus(axis);
  cells[h] = G(m, row, 0);
  for (int L = 0; L < h; ++L) {
    cells[h - 1 - L] = G(m, row, gcol(m->dim, sm, L));
    cells[h + 1 + L] = G(m, row, gcol(m->dim, sp, L));
  }
}
/* gather the S stencil values of one dof along an axis from those cells (ghost rows where a cell is missing) */
static void gather(const or_problem* p, const double* U, const int32_t* cells, int32_t gRow, int axis, int S, int dof,
                   double* q) {
  const int h = (S - 1) / 2;
  const int sm = side_minus(axis), sp = side_plus(axis);
  q[h] = U[(size_t)cells[h] * p->ndpc + dof];
  for (int L = 0; L < h; ++L) {
    q[h - 1 - L] = sval(p, U, cells[h - 1 - L], sm, gRow, L, dof);
    q[h + 1 + L] = sval(p, U, cells[h + 1 + L], sp, gRow, L, dof);
  }
}

/* edge states of a cell along one axis (functor_reconstruct_from_state.hpp / functor_reconstruct_from_stencil.hpp):
 * left face:  (uMinusHalfNeg, uMinusHalfPos); right face: (uPlusHalfNeg, uPlusHalfPos); optional gradients */
static void reconstruct(int S, const double* q, double* lN, double* lP, double* rN, double* rP, double* gLN, double* gLP,
                        double* gRN, double* gRP) {
  if (S == 3) {
    *lN = q[0]; *lP = q[1]; *rN = q[1]; *rP = q[2];
  } else if (S == 5) {
    if (gLN) { or_weno3_grad(lN, lP, gLN, gLP, q[0], q[1], q[2], q[3]); or_weno3_grad(rN, rP, gRN, gRP, q[1], q[2], q[3], q[4]); }
    else { or_weno3(lN, lP, q[0], q[1], q[2], q[3]); or_weno3(rN, rP, q[1], q[2], q[3], q[4]); }
  } else {
    if (gLN) { or_weno5_grad(lN, lP, gLN, gLP, q[0], q[1], q[2], q[3], q[4], q[5]); or_weno5_grad(rN, rP, gRN, gRP, q[1], q[2], q[3], q[4], q[5], q[6]); }
    else { or_weno5(lN, lP, q[0], q[1], q[2], q[3], q[4], q[5]); or_weno5(rN, rP, q[1], q[2], q[3], q[4], q[5], q[6]); }
  }
}

/* Eigen coeffRef on the fixed pattern: binary search in the row */
static void jadd(const or_problem* p, double* vals, int32_t row, int32_t col, double v) {
  int32_t lo = p->rowptr[row], hi = p->rowptr[row + 1] - 1;
  while (lo <= hi) {
    const int32_t mid = (lo + hi) / 2;
    if (p->colidx[mid] == col) { vals[mid] += v; return; }
    if (p->colidx[mid] < col) lo = mid + 1; else hi = mid - 1;
  }
  fprintf(stderr, "oracle: Jacobian entry (%d,%d) not in the pattern\n", row, col);
  abort();
}

/* first-order Jacobian factors: euler_2d_prob_class.hpp:1115-1224, swe_2d_prob_class.hpp:839-855,
 * euler_3d_prob_class.hpp:1017-1044 + :626-638, euler_1d_prob_class.hpp:608-616 */
static void jac_factors(const or_problem* p, int32_t row, int axis, double* f) {
  const or_mesh* m = &p->m;
  const int N = p->ndpc;
  for (int d = 0; d < N; ++d) f[d] = 1.;
  const int sm = side_minus(axis);
  if (p->family == F_SWE2D) { f[axis] = -1.; return; }
  if (p->family == F_ADVDIFFREAC2D) { f[0] = -1.; return; } /* advection_diffusion_reaction_2d_prob_class.hpp:634-635 */
  if (p->family == F_ADVDIFF2D) { /* advection_diffusion_2d_prob_class.hpp:1119-1155 */
    if (has_bd(m, row, sm)) for (int d = 0; d < N; ++d) f[d] = 0.;
    return;
  }
  if (p->family == F_EULER3D) { if (p->prob == 1 && has_bd(m, row, sm)) f[axis] = -1.; return; }
  if (p->family != F_EULER2D) return;
  const double myX = mx(m, G(m, row, 0));
  switch (p->prob) {
    case E2_SEDOV_SYM: if (has_bd(m, row, sm)) f[axis] = -1.; break;
    case E2_NORMAL_SHOCK: if (axis == 2) f[2] = -1.; break;
    case E2_DMR:
      if (axis == 1) break;
      if (has_bd(m, row, 3) && myX < 1. / 6.) break;
      if (has_bd(m, row, 3) && myX >= 1. / 6.) { f[2] = -1.; break; }
      for (int d = 0; d < N; ++d) f[d] = 0.;
      break;
    case E2_CROSS_SHOCK:
      if (axis == 1 && has_bd(m, row, 0)) { for (int d = 0; d < N; ++d) f[d] = 0.; break; }
      if (axis == 1 && has_bd(m, row, 2)) break;
      if (axis == 2 && has_bd(m, row, 3)) { f[1] = 0.; f[2] = 0.; break; }
      break;
    default: break;
  }
}

/* one cell: velocity (+ Jacobian) -- velocityAndOptionalJacobian for one graph row
 * (euler_2d_prob_class.hpp:633-720 inner, :723-989 near boundary; mixin_directional_flux_balance.hpp:68-84;
 *  mixin_directional_flux_balance_jacobian.hpp:142-371) */
static void eval_cell(const or_problem* p, const double* U, int32_t row, int32_t gRow, int nearBd, double* V, double* Jv) {
  const or_mesh* m = &p->m;
  const int N = p->ndpc, S = p->S, dim = m->dim;
  const int32_t vIdx = row * N;
  const int32_t selfCol = G(m, row, 0) * N;
  for (int axis = 1; axis <= dim; ++axis) {
    const double hInv = m->dInv[axis - 1];
    double lN[5], lP[5], rN[5], rP[5];
    double gLN[5][6], gLP[5][6], gRN[5][6], gRP[5][6];
    double q[7];
    int32_t cells[7];
    const int wantGrad = (Jv != NULL) && !nearBd && S > 3;
    double s3[5][3]; /* first-layer stencil (l0, self, r0) per dof: the diffusion term's operands */
    gather_cells(p, row, axis, S, cells);
    for (int d = 0; d < N; ++d) {
      gather(p, U, cells, gRow, axis, S, d, q);
      s3[d][0] = q[(S - 1) / 2 - 1]; s3[d][1] = q[(S - 1) / 2]; s3[d][2] = q[(S - 1) / 2 + 1];
      reconstruct(S, q, &lN[d], &lP[d], &rN[d], &rP[d], wantGrad ? gLN[d] : NULL, wantGrad ? gLP[d] : NULL,
                  wantGrad ? gRN[d] : NULL, wantGrad ? gRP[d] : NULL);
    }
    double FL[5], FR[5];
    flux(p, axis, FL, lN, lP);
    flux(p, axis, FR, rN, rP);
    if (V) for (int d = 0; d < N; ++d) V[vIdx + d] += hInv * (FL[d] - FR[d]);
    /* near-boundary rows of the advection-diffusion families add the axis' diffusion right after its flux balance
     * (advection_diffusion_2d_prob_class.hpp:729-753,959-1000; advection_diffusion_reaction_2d_prob_class.hpp:667-683) */
    if (V && nearBd && (p->family == F_ADVDIFF2D || p->family == F_ADVDIFFREAC2D)) {
      const double D = (p->family == F_ADVDIFF2D) ? p->php[0] : p->php[2];
      const double diffInvSq = D * (hInv * hInv);
      for (int d = 0; d < N; ++d) V[vIdx + d] += diffInvSq * (s3[d][2] - 2. * s3[d][1] + s3[d][0]);
    }
    if (!Jv) continue;

    double JLN[25], JLP[25], JRN[25], JRP[25];
    if (!nearBd) {
      flux_jac(p, axis, JLN, JLP, lN, lP);
      flux_jac(p, axis, JRN, JRP, rN, rP);
      const int h = (S - 1) / 2;
      if (S == 3) {
        const int32_t cim1 = cells[0] * N, cip1 = cells[2] * N;
        for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) {
          jadd(p, Jv, vIdx + k, cim1 + j, JLN[k * N + j] * hInv);
          jadd(p, Jv, vIdx + k, selfCol + j, (JLP[k * N + j] - JRN[k * N + j]) * hInv);
          jadd(p, Jv, vIdx + k, cip1 + j, -JRP[k * N + j] * hInv);
        }
      } else {
        (void)h;
        for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) {
          for (int mm = 0; mm < S - 1; ++mm) { /* sensitivity of the flux at i-1/2: stencil positions 0..S-2 */
            jadd(p, Jv, vIdx + k, cells[mm] * N + j, JLN[k * N + j] * gLN[j][mm] * hInv);
            jadd(p, Jv, vIdx + k, cells[mm] * N + j, JLP[k * N + j] * gLP[j][mm] * hInv);
          }
          for (int mm = 0; mm < S - 1; ++mm) { /* flux at i+1/2: stencil positions 1..S-1 */
            jadd(p, Jv, vIdx + k, cells[mm + 1] * N + j, -(JRN[k * N + j] * gRN[j][mm] * hInv));
            jadd(p, Jv, vIdx + k, cells[mm + 1] * N + j, -(JRP[k * N + j] * gRP[j][mm] * hInv));
          }
        }
      }
    } else {
      /* first-order Jacobian from the first-order stencil (ghost layer 0) */
      double qL[5], qC[5], qR[5], fac[5];
      const int sm = side_minus(axis), sp = side_plus(axis);
      const int32_t l0 = G(m, row, gcol(dim, sm, 0)), r0 = G(m, row, gcol(dim, sp, 0));
      for (int d = 0; d < N; ++d) {
        qL[d] = sval(p, U, l0, sm, gRow, 0, d);
        qC[d] = U[(size_t)G(m, row, 0) * N + d];
        qR[d] = sval(p, U, r0, sp, gRow, 0, d);
      }
      flux_jac(p, axis, JLN, JLP, qL, qC);
      flux_jac(p, axis, JRN, JRP, qC, qR);
      jac_factors(p, row, axis, fac);
      for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) jadd(p, Jv, vIdx + k, selfCol + j, (JLP[k * N + j] - JRN[k * N + j]) * hInv);
      if (l0 != -1) for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) jadd(p, Jv, vIdx + k, l0 * N + j, JLN[k * N + j] * hInv);
      if (r0 != -1) for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) jadd(p, Jv, vIdx + k, r0 * N + j, -JRP[k * N + j] * hInv);
      if (l0 == -1) for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) jadd(p, Jv, vIdx + k, selfCol + j, (fac[j] * JLN[k * N + j]) * hInv);
      if (r0 == -1) for (int k = 0; k < N; ++k) for (int j = 0; j < N; ++j) jadd(p, Jv, vIdx + k, selfCol + j, (fac[j] * -JRP[k * N + j]) * hInv);
    }
  }
  if (p->family == F_ADVDIFF2D || p->family == F_ADVDIFFREAC2D) {
    /* diffusion (+ source, reaction for ADR): advection_diffusion_2d_prob_class.hpp:755-792,1157-1201;
     * advection_diffusion_reaction_2d_prob_class.hpp:485-512,685-721,1057-1085 */
    const int adr = p->family == F_ADVDIFFREAC2D;
    const double D = adr ? p->php[2] : p->php[0];
    const double two = 2.;
    const double dxInvSq = m->dInv[0] * m->dInv[0], dyInvSq = m->dInv[1] * m->dInv[1];
    const double diffDxInvSq = D * dxInvSq, diffDyInvSq = D * dyInvSq;
    const int32_t iL = G(m, row, 1), iF = G(m, row, 2), iR = G(m, row, 3), iB = G(m, row, 4);
    if (!nearBd) {
      for (int d = 0; d < N; ++d) {
        if (V) {
          V[vIdx + d] += diffDxInvSq * (U[(size_t)iR * N + d] - two * U[selfCol + d] + U[(size_t)iL * N + d]);
          V[vIdx + d] += diffDyInvSq * (U[(size_t)iF * N + d] - two * U[selfCol + d] + U[(size_t)iB * N + d]);
        }
        if (Jv) {
          jadd(p, Jv, vIdx + d, selfCol + d, -two * diffDxInvSq - two * diffDyInvSq);
          jadd(p, Jv, vIdx + d, iL * N + d, diffDxInvSq);
          jadd(p, Jv, vIdx + d, iF * N + d, diffDyInvSq);
          jadd(p, Jv, vIdx + d, iR * N + d, diffDxInvSq);
          jadd(p, Jv, vIdx + d, iB * N + d, diffDyInvSq);
        }
      }
      if (adr) {
        if (V) { V[vIdx] += p->src[row]; V[vIdx] -= p->php[3] * U[selfCol]; }
        if (Jv) jadd(p, Jv, vIdx, selfCol, -p->php[3]);
      }
    } else {
      if (adr && V) { V[vIdx] += p->src[row]; V[vIdx] -= p->php[3] * U[selfCol]; }
      if (Jv) {
        const int32_t nb[4] = {iL, iF, iR, iB};
        const double dd[4] = {diffDxInvSq, diffDyInvSq, diffDxInvSq, diffDyInvSq};
        if (adr) {
          double selfValue = -two * diffDxInvSq - two * diffDyInvSq - p->php[3];
          for (int k = 0; k < 4; ++k) { if (nb[k] != -1) jadd(p, Jv, vIdx, nb[k], dd[k]); else selfValue += -dd[k]; }
          jadd(p, Jv, vIdx, selfCol, selfValue);
        } else {
          for (int d = 0; d < N; ++d) jadd(p, Jv, vIdx + d, selfCol + d, -two * diffDxInvSq - two * diffDyInvSq);
          for (int k = 0; k < 4; ++k)
            for (int d = 0; d < N; ++d) {
              if (nb[k] != -1) jadd(p, Jv, vIdx + d, nb[k] * N + d, dd[k]); else jadd(p, Jv, vIdx + d, selfCol + d, -dd[k]);
            }
        }
      }
    }
  }
  if (p->family == F_SWE2D) { /* swe_2d_prob_class.hpp:984-1012 */
    const double f = p->php[1];
    const double* u = U + selfCol;
    if (V) { V[vIdx + 1] -= f * u[2] / u[0]; V[vIdx + 2] += f * u[1] / u[0]; }
    if (Jv) {
      jadd(p, Jv, vIdx + 1, selfCol, f * u[2] / (u[0] * u[0]));
      jadd(p, Jv, vIdx + 1, selfCol + 2, -f / u[0]);
      jadd(p, Jv, vIdx + 2, selfCol + 1, f / u[0]);
      jadd(p, Jv, vIdx + 2, selfCol, -f * u[1] / (u[0] * u[0]));
    }
  }
}

/* diffusion_reaction_2d_prob_class.hpp:459-529 */
static void gray_scott(const or_problem* p, const double* U, double* V, double* Jv) {
  const or_mesh* m = &p->m;
  const double one = 1, two = 2;
  const double dxInvSq = m->dInv[0] * m->dInv[0], dyInvSq = m->dInv[1] * m->dInv[1];
  const double Du = p->gs[0], Dv = p->gs[1], F = p->gs[2], kk = p->gs[3];
  const double uDx = Du * dxInvSq, uDy = Du * dyInvSq, vDx = Dv * dxInvSq, vDy = Dv * dyInvSq;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int32_t smPt = 0; smPt < m->nSample; ++smPt) {
    const int32_t vi = smPt * 2;
    const int32_t si = G(m, smPt, 0) * 2, sl = G(m, smPt, 1) * 2, sf = G(m, smPt, 2) * 2, sr = G(m, smPt, 3) * 2, sb = G(m, smPt, 4) * 2;
    const double u = U[si], v = U[si + 1];
    const double uvSquared = u * v * v;
    if (V) {
      V[vi] = F * (one - u) - uvSquared + uDx * (U[sr] - two * U[si] + U[sl]) + uDy * (U[sb] - two * U[si] + U[sf]);
      V[vi + 1] = -(F + kk) * v + uvSquared + vDx * (U[sr + 1] - two * U[si + 1] + U[sl + 1]) + vDy * (U[sb + 1] - two * U[si + 1] + U[sf + 1]);
    }
    if (Jv) {
      jadd(p, Jv, vi, si, -two * uDx - two * uDy - v * v - F);
      jadd(p, Jv, vi, si + 1, -(two * u * v));
      jadd(p, Jv, vi, sl, uDx); jadd(p, Jv, vi, sf, uDy); jadd(p, Jv, vi, sr, uDx); jadd(p, Jv, vi, sb, uDy);
      jadd(p, Jv, vi + 1, si, v * v);
      jadd(p, Jv, vi + 1, si + 1, -two * vDx - two * vDy + two * u * v - (F + kk));
      jadd(p, Jv, vi + 1, sl + 1, vDx); jadd(p, Jv, vi + 1, sf + 1, vDy); jadd(p, Jv, vi + 1, sr + 1, vDx); jadd(p, Jv, vi + 1, sb + 1, vDy);
    }
  }
}

/* DiffusionReaction{1d,2d}::ProblemA: diffusion_reaction_1d_prob_class.hpp:211-302,
 * diffusion_reaction_2d_prob_class.hpp:306-456.  Ghost of a missing neighbour = -s(self)
 * (diffusion_reaction_1d_ghost_filler.hpp:85-94, diffusion_reaction_2d_ghost_filler.hpp:88-103). */
static void diffreac_problem_a(const or_problem* p, const double* U, double* V, double* Jv) {
  const or_mesh* m = &p->m;
  const double two = 2., three = 3.;
  const double dxInvSq = m->dInv[0] * m->dInv[0], dyInvSq = m->dInv[1] * m->dInv[1];
  const double D = p->php[0], kR = p->php[1];
  const double twoReacCoeff = kR * two;
  const double diffDxInvSq = D * dxInvSq, diffDyInvSq = D * dyInvSq;
  char* isNb = (char*)calloc((size_t)m->nSample + 1, 1);
  for (int32_t i = 0; i < m->nNearBd; ++i) isNb[m->rowsNearBd[i]] = 1;
  for (int32_t smPt = 0; smPt < m->nSample; ++smPt) {
    const int32_t uIndex = G(m, smPt, 0);
    const double u = U[uIndex];
    if (m->dim == 1) {
      const int32_t iL = G(m, smPt, 1), iR = G(m, smPt, 2);
      const double sL = (iL != -1) ? U[iL] : -u, sR = (iR != -1) ? U[iR] : -u;
      if (V) {
        V[smPt] = p->src[smPt];
        V[smPt] += kR * u * u;
        const double fd = sR - two * u + sL;
        V[smPt] += dxInvSq * D * fd;
      }
      if (Jv) {
        if (isNb[smPt]) {
          jadd(p, Jv, smPt, uIndex, -three * diffDxInvSq + twoReacCoeff * u);
          if (iL != -1) jadd(p, Jv, smPt, iL, diffDxInvSq);
          if (iR != -1) jadd(p, Jv, smPt, iR, diffDxInvSq);
        } else {
          jadd(p, Jv, smPt, uIndex, -two * diffDxInvSq + twoReacCoeff * u);
          jadd(p, Jv, smPt, iL, diffDxInvSq);
          jadd(p, Jv, smPt, iR, diffDxInvSq);
        }
      }
    } else {
      const int32_t iL = G(m, smPt, 1), iF = G(m, smPt, 2), iR = G(m, smPt, 3), iB = G(m, smPt, 4);
      const double sL = (iL != -1) ? U[iL] : -u, sR = (iR != -1) ? U[iR] : -u;
      const double sB = (iB != -1) ? U[iB] : -u, sF = (iF != -1) ? U[iF] : -u;
      if (V) {
        V[smPt] = p->src[smPt];
        V[smPt] += kR * u * u;
        V[smPt] += dxInvSq * D * (sR - two * u + sL);
        V[smPt] += dyInvSq * D * (sF - two * u + sB);
      }
      if (Jv) {
        double selfValue = -two * diffDxInvSq - two * diffDyInvSq + twoReacCoeff * u;
        if (iL != -1) jadd(p, Jv, smPt, iL, diffDxInvSq); else selfValue += -diffDxInvSq;
        if (iF != -1) jadd(p, Jv, smPt, iF, diffDyInvSq); else selfValue += -diffDyInvSq;
        if (iR != -1) jadd(p, Jv, smPt, iR, diffDxInvSq); else selfValue += -diffDxInvSq;
        if (iB != -1) jadd(p, Jv, smPt, iB, diffDyInvSq); else selfValue += -diffDyInvSq;
        jadd(p, Jv, smPt, uIndex, selfValue);
      }
    }
  }
  free(isNb);
}

/* velocityAndOptionalJacobian: euler_2d_prob_class.hpp:241-310 (zero V, zero J, ghosts, near-bd rows, inner rows) */
static int evaluate(or_problem* p, const double* U, double t, double* V, double* Jv) {
  const or_mesh* m = &p->m;
  const int N = p->ndpc;
  if (V) memset(V, 0, sizeof(double) * (size_t)m->nSample * N);
  if (Jv) { build_pattern(p); memset(Jv, 0, sizeof(double) * (size_t)p->nnz); }
  if (p->family == F_DIFFREAC2D && p->prob == 1) { gray_scott(p, U, V, Jv); return 0; }
  if (p->family == F_DIFFREAC2D || p->family == F_DIFFREAC1D) { diffreac_problem_a(p, U, V, Jv); return 0; }
  fill_ghosts(p, U, t);
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int32_t it = 0; it < m->nNearBd; ++it) eval_cell(p, U, m->rowsNearBd[it], it, 1, V, Jv);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int32_t it = 0; it < m->nInner; ++it) eval_cell(p, U, ROW_INNER(m, it), -1, 0, V, Jv);
  }
  return 0;
}

int or_velocity(or_problem* p, const double* U, double t, double* V) { return evaluate(p, U, t, V, NULL); }
int or_velocity_and_jacobian(or_problem* p, const double* U, double t, double* V, double* vals) {
  return evaluate(p, U, t, V, vals);
}

int or_ghosts(or_problem* p, int side, double* out) {
  const int n = p->m.nNearBd * p->gstride;
  if (out) memcpy(out, p->ghost[side], sizeof(double) * (size_t)n);
  return n;
}

/* bounded sample of a big workload: velocity of the inner rows [it0, it1) only (positions in graphRowsOfCellsAwayFromBd),
 * `reps` times, V rows zeroed first like the full evaluation does; returns the seconds spent.  For the CPU-baseline /
 * reference arm of bench.py: the 512^3 problem itself is evaluated, a slab of planes at a time. */
double or_time_velocity_inner_range(or_problem* p, const double* U, double t, double* V, int32_t it0, int32_t it1, int reps) {
  const or_mesh* m = &p->m;
  const int N = p->ndpc;
  if (it0 < 0) it0 = 0;
  if (it1 > m->nInner) it1 = m->nInner;
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  for (int rep = 0; rep < reps; ++rep) {
    fill_ghosts(p, U, t);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int32_t it = it0; it < it1; ++it) {
      const int32_t row = ROW_INNER(m, it);
      for (int d = 0; d < N; ++d) V[(size_t)row * N + d] = 0.;
      eval_cell(p, U, row, -1, 0, V, NULL);
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &b);
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

double or_time_velocity(or_problem* p, const double* U, double t, int warmup, int reps) {
  double* V = (double*)malloc(sizeof(double) * (size_t)p->m.nSample * p->ndpc);
  for (int i = 0; i < warmup; ++i) evaluate(p, U, t, V, NULL);
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  for (int i = 0; i < reps; ++i) evaluate(p, U, t, V, NULL);
  clock_gettime(CLOCK_MONOTONIC, &b);
  free(V);
  return ((double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec)) / reps;
}

/* ------------------------------------------------------------------------------------------------------------
 * GradientEvaluator (gradient.hpp:61-121): normal gradients at the faces on the domain boundary, 2D.
 * ------------------------------------------------------------------------------------------------------------ */
int64_t or_gradient_faces(int stencil, const int32_t* graph, const int32_t* rowsNearBd, int32_t nNearBd, const double* x,
                          const double* y, const double* z, double dx, double dy, int32_t* cellGid, int32_t* position,
                          int32_t* parentRow, int32_t* normalDir, double* centers) {
  /* initializeForStoringNormalGradsAtBoundaryFaces, gradient_2d.hpp:233-286 */
  const int nc = (stencil - 1) * 2 + 1;
  const double dxHalf = dx * 0.5, dyHalf = dy * 0.5;
  int64_t k = 0;
  for (int32_t it = 0; it < nNearBd; ++it) {
    const int32_t row = rowsNearBd[it];
    const int32_t* cc = graph + (size_t)row * nc;
    for (int pos = 0; pos < 4; ++pos) {       /* Left, Front, Right, Back = columns 1..4 */
      if (cc[1 + pos] != -1) continue;
      const int32_t gid = cc[0];
      if (cellGid) cellGid[k] = gid;
      if (position) position[k] = pos;
      if (parentRow) parentRow[k] = row;
      if (normalDir) normalDir[k] = (pos == 0 || pos == 2) ? 1 : 2;
      if (centers) {
        centers[3 * k + 0] = (pos == 0) ? x[gid] - dxHalf : ((pos == 2) ? x[gid] + dxHalf : x[gid]);
        centers[3 * k + 1] = (pos == 1) ? y[gid] + dyHalf : ((pos == 3) ? y[gid] - dyHalf : y[gid]);
        centers[3 * k + 2] = z[gid];
      }
      ++k;
    }
  }
  return k;
}

void or_gradient_eval(int stencil, const int32_t* graph, int64_t nFaces, const int32_t* position,
                      const int32_t* parentRow, double dx, double dy, const double* f, int ndpc, double* grad) {
  /* normalGradBoundaryFacesOneSidedFdAutoStencil (gradient_2d.hpp:164-229) calling
   * face_normal_gradient_for_cell_centered_function_2d (gradient_2d.hpp:62-104) */
  const int nc = (stencil - 1) * 2 + 1;
  for (int64_t k = 0; k < nFaces; ++k) {
    const int32_t* cc = graph + (size_t)parentRow[k] * nc;
    const int pos = position[k];
    const int alongX = (pos == 0 || pos == 2);
    const int rightSided = (pos == 0 || pos == 3);   /* Left and Back faces use the forward formula */
    const double h = alongX ? dx : dy;
    const int32_t i05 = cc[0];
    const int32_t ip15 = alongX ? cc[3] : cc[2];
    const int32_t im15 = alongX ? cc[1] : cc[4];
    const int32_t ip30 = (nc >= 9) ? (alongX ? cc[7] : cc[6]) : -1;
    const int32_t im30 = (nc >= 9) ? (alongX ? cc[5] : cc[8]) : -1;
    for (int j = 0; j < ndpc; ++j) {
      const double f05 = f[(size_t)i05 * ndpc + j];
      const double fp15 = (ip15 != -1) ? f[(size_t)ip15 * ndpc + j] : 0;
      const double fm15 = (im15 != -1) ? f[(size_t)im15 * ndpc + j] : 0;
      const double fp30 = (ip30 != -1) ? f[(size_t)ip30 * ndpc + j] : 0;
      const double fm30 = (im30 != -1) ? f[(size_t)im30 * ndpc + j] : 0;
      double g;
      if (stencil == 3) g = rightSided ? (-f05 + fp15) / h : (f05 - fm15) / h;
      else g = rightSided ? (-2 * f05 + 3 * fp15 - 1 * fp30) / h : (2 * f05 - 3 * fm15 + 1 * fm30) / h;
      grad[k * ndpc + j] = g;
    }
  }
}
