/* bc_functors.c -- the custom-BC functors of the reference's Swe2d::CustomBCs test
 * (/root/reference/tests_cpp/eigen_2d_swe_custom_bcs/main.cc:6-58) written against the C-ABI's host-callback contract
 * (include/pda_b200.h: pda_bc_ghost_fn / pda_bc_factor_fn): a Dirichlet side whose ghost cells hold a fixed state and
 * homogeneous-Neumann sides whose ghost cells copy the boundary cell.  Built into examples/_bin/libbc_functors.so;
 * used by tools/bench_configs.py to time the host-callback path at 4096^2 and by the tests that compare it with the
 * device rule tables. */
#include <stdint.h>

void bc_ghost_dirichlet(void* user, int32_t near_bd_row, const int32_t* graph_row, double cell_x, double cell_y,
                        const double* U, int ndpc, double cell_width, double* ghost_values) {
  (void)user; (void)near_bd_row; (void)graph_row; (void)cell_x; (void)cell_y; (void)U; (void)cell_width;
  /* two ghost layers (WENO3): every layer holds the prescribed state (h, hu, hv) = (1, 0, 0) */
  for (int layer = 0; layer < 2; ++layer) {
    ghost_values[layer * ndpc + 0] = 1.0;
    for (int d = 1; d < ndpc; ++d) ghost_values[layer * ndpc + d] = 0.0;
  }
}
void bc_factor_dirichlet(void* user, const int32_t* graph_row, double cell_x, double cell_y, int ndpc, double* factors) {
  (void)user; (void)graph_row; (void)cell_x; (void)cell_y;
  for (int d = 0; d < ndpc; ++d) factors[d] = 0.0;
}
void bc_ghost_neumann(void* user, int32_t near_bd_row, const int32_t* graph_row, double cell_x, double cell_y,
                      const double* U, int ndpc, double cell_width, double* ghost_values) {
  (void)user; (void)near_bd_row; (void)cell_x; (void)cell_y; (void)cell_width;
  const int64_t self = graph_row[0];
  for (int layer = 0; layer < 2; ++layer)
    for (int d = 0; d < ndpc; ++d) ghost_values[layer * ndpc + d] = U[self * ndpc + d];
}
void bc_factor_neumann(void* user, const int32_t* graph_row, double cell_x, double cell_y, int ndpc, double* factors) {
  (void)user; (void)graph_row; (void)cell_x; (void)cell_y;
  for (int d = 0; d < ndpc; ++d) factors[d] = 1.0;
}
