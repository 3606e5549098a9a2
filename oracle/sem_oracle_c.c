/* sem_oracle_c.c -- C/OpenMP restatement of the reference's operator apply.
 *
 * TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 *
 * Algorithm (what the reference does, examples/squirmer-axisymmetric.py:268-295
 * with the dense 4-index local stiffness of examples/poisson.py:181-193 and the
 * scatter-add of sem/discrete.py:499):
 *     for every element e:  y[L2G[e]] += Lse[e] (NN x NN, dense) . u[L2G[e]]
 * restated with all host threads: elements are distributed over OpenMP threads
 * and the scatter uses atomic adds (shared nodes are touched by <= 4 elements).
 * Pinned against oracle/sem_oracle.py (itself pinned to the live reference's
 * golden vectors) in tests/test_oracle_golden.py.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int sem_oracle_c_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Use n threads from now on whatever OMP_NUM_THREADS says (torchrun exports
 * OMP_NUM_THREADS=1 to its workers; the CPU baseline is meant to use every host core). */
void sem_oracle_c_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* L: [E][NN][NN] row-major, l2g: [E][NN] uint32, u, y: [n_nodes]; y is overwritten */
void sem_oracle_c_apply_dense(int64_t E, int NN, int64_t n_nodes, const double *L,
                              const uint32_t *l2g, const double *u, double *y) {
  memset(y, 0, sizeof(double) * (size_t)n_nodes);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) {
    const double *Le = L + (size_t)e * NN * NN;
    const uint32_t *idx = l2g + (size_t)e * NN;
    double ul[289], yl[289];
    for (int k = 0; k < NN; ++k) ul[k] = u[idx[k]];
    for (int r = 0; r < NN; ++r) {
      const double *row = Le + (size_t)r * NN;
      double acc = 0.0;
#pragma omp simd reduction(+ : acc)
      for (int k = 0; k < NN; ++k) acc += row[k] * ul[k];
      yl[r] = acc;
    }
    for (int k = 0; k < NN; ++k) {
#pragma omp atomic
      y[idx[k]] += yl[k];
    }
  }
}
