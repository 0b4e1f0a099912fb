/* Plain C99 program against include/pda_b200.h: the boundary is a real C ABI (no C++/torch types).
 * Builds a 2D lattice, creates Euler2d::Riemann WENO5, queries sizes and the CSR pattern (host side, no GPU needed),
 * then -- when a device is present -- evaluates velocity and Jacobian through the host-pointer entry points.
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lpressio-demoapps_b200/lib -lpda_b200 -Wl,-rpath,$PWD/pressio-demoapps_b200/lib -lm -o /tmp/c_abi_demo
 * exit code 0 = ok (prints "no device" and stops after the host part on a CPU-only box). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pda_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    pda_status s_ = (call);                                                  \
    if (s_ != PDA_OK) {                                                      \
      fprintf(stderr, "%s failed (%d): %s\n", #call, s_, pda_last_error()); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  const int32_t n[3] = {40, 30, 1};
  const double bounds[6] = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
  const int32_t periodic[3] = {0, 0, 0};
  pda_mesh mesh = NULL;
  pda_problem prob = NULL;
  CHECK(pda_mesh_make_lattice(2, n, bounds, periodic, 7, &mesh));
  CHECK(pda_problem_create(mesh, PDA_FAMILY_EULER2D, PDA_EULER2D_RIEMANN, PDA_WENO5, 1, 0, NULL, NULL, 0, &prob));
  const int32_t ndof = pda_problem_total_dof_stencil_mesh(prob);
  int64_t nnz = 0;
  CHECK(pda_problem_jacobian_nnz(prob, &nnz));
  printf("cells %d dofs %d nnz %lld\n", (int)pda_mesh_sample_mesh_size(mesh), (int)ndof, (long long)nnz);
  if (ndof != 40 * 30 * 4) return 2;
  double* U = (double*)malloc(sizeof(double) * (size_t)ndof);
  double* V = (double*)malloc(sizeof(double) * (size_t)ndof);
  double* J = (double*)malloc(sizeof(double) * (size_t)nnz);
  int32_t* rowptr = (int32_t*)malloc(sizeof(int32_t) * ((size_t)ndof + 1));
  int32_t* colidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)nnz);
  CHECK(pda_problem_initial_condition(prob, U));
  CHECK(pda_problem_jacobian_pattern(prob, rowptr, colidx));
  if (rowptr[ndof] != nnz) return 3;
  if (pda_device_count() < 1) {
    /* no CPU fallback: the evaluation must refuse, not compute */
    const pda_status s = pda_problem_velocity_host(prob, U, 0.0, V);
    printf("no device: velocity returned %d (%s)\n", s, pda_last_error());
    if (s != PDA_ERR_NO_DEVICE) return 4;
  } else {
    CHECK(pda_problem_velocity_and_jacobian_host(prob, U, 0.0, V, J));
    double vmax = 0.0, jmax = 0.0;
    int i;
    for (i = 0; i < ndof; ++i) if (fabs(V[i]) > vmax) vmax = fabs(V[i]);
    for (i = 0; i < nnz; ++i) if (fabs(J[i]) > jmax) jmax = fabs(J[i]);
    printf("max|V| %.6e max|J| %.6e launches %lld\n", vmax, jmax, (long long)pda_problem_launch_count(prob));
    if (!(vmax > 0.0) || !(jmax > 0.0) || vmax != vmax || jmax != jmax) return 5;
    /* 10 SSPRK3 steps with the state resident in HBM */
    CHECK(pda_problem_advance_host(prob, PDA_STEPPER_SSPRK3, U, 0.0, 1e-3, 10));
  }
  {
    /* normal gradients of the 4-dof state at the faces on the domain boundary (GradientEvaluator, gradient.hpp) */
    pda_gradient grad = NULL;
    int32_t nfaces, k = -1;
    CHECK(pda_gradient_create(mesh, 4, &grad));
    nfaces = pda_gradient_num_faces(grad);
    if (nfaces != 2 * (40 + 30)) return 6;
    CHECK(pda_gradient_query_face(grad, 0, 0 /* FacePosition::Left of cell 0 */, &k));
    if (k != 0) return 7;
    if (pda_device_count() >= 1) {
      double* g = (double*)malloc(sizeof(double) * 4 * (size_t)nfaces);
      CHECK(pda_gradient_compute_host(grad, U, 4, g));
      printf("boundary faces %d, d(rho)/dn at the first face %.6e\n", (int)nfaces, g[0]);
      free(g);
    } else if (pda_gradient_compute_host(grad, U, 4, V) != PDA_ERR_NO_DEVICE) {
      return 8;
    }
    CHECK(pda_gradient_free(grad));
  }
  free(U); free(V); free(J); free(rowptr); free(colidx);
  CHECK(pda_problem_free(prob));
  CHECK(pda_mesh_free(mesh));
  printf("c_abi_demo ok\n");
  return 0;
}
