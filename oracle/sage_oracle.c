/* sage_oracle.c -- plain-C CPU restatement of the SageBlock hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
 * The product library (libsldm_sage.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference (aledima00/sldm-gnn) has no tests, fixtures or
 * golden vectors, and the arithmetic of its hot path is torch-geometric 2.7.0
 * (uv.lock:1406-1407), which is neither vendored nor installable here.  This
 * file restates the published algorithm of that version as sequential loops;
 * oracle/sage_oracle.py restates it with the same ATen CPU operators PyG calls.
 * The two restatements are checked against each other and against hand-computed
 * known answers in tests/test_oracle.py.
 *
 * Two precisions are built from one body (sage_oracle_impl.h):
 *   *_f32 : fp32 arithmetic, edges visited in edge order -- the summation order
 *           of the reference's CPU scatter_add_, so the aggregation of the CUDA
 *           path (stable CSR) can be compared bit for bit;
 *   *_f64 : the same in double, used to adjudicate fp32 tolerance questions.
 *
 * oracle_csr_build follows the index semantics of PyG's scatter
 * (src/models/blocks/sageblock.py:18 -> MessagePassing.propagate): a stable
 * counting sort of the edges by destination (and by source for the transpose).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* rowptr_* have N+1 entries, col_* have E entries; returns 2 on an index outside [0,N) */
int oracle_csr_build(const int64_t* edge_index, int64_t E, int64_t N,
                     int32_t* rowptr_dst, int32_t* col_src,
                     int32_t* rowptr_src, int32_t* col_dst) {
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + E;
  for (int pass = 0; pass < 2; ++pass) {
    const int64_t* key = pass == 0 ? dst : src;
    const int64_t* val = pass == 0 ? src : dst;
    int32_t* rowptr = pass == 0 ? rowptr_dst : rowptr_src;
    int32_t* col = pass == 0 ? col_src : col_dst;
    memset(rowptr, 0, (size_t)(N + 1) * sizeof(int32_t));
    for (int64_t e = 0; e < E; ++e) {
      if (key[e] < 0 || key[e] >= N || val[e] < 0 || val[e] >= N) return 2;
      rowptr[key[e] + 1] += 1;
    }
    for (int64_t i = 0; i < N; ++i) rowptr[i + 1] += rowptr[i];
    int32_t* cursor = (int32_t*)malloc((size_t)(N > 0 ? N : 1) * sizeof(int32_t));
    if (!cursor) return 1;
    memcpy(cursor, rowptr, (size_t)N * sizeof(int32_t));
    for (int64_t e = 0; e < E; ++e) col[cursor[key[e]]++] = (int32_t)val[e]; /* edge order: stable */
    free(cursor);
  }
  return 0;
}

#define REAL float
#define SUFFIX _f32
#include "sage_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX _f64
#include "sage_oracle_impl.h"
#undef REAL
#undef SUFFIX
