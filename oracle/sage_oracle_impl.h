/* sage_oracle_impl.h -- body of the plain-C restatement, included twice by
 * sage_oracle.c with REAL = float (suffix _f32) and REAL = double (suffix _f64).
 * TEST INFRASTRUCTURE ONLY (see sage_oracle.c).
 *
 * Loops are the literal, sequential form of the reference's arithmetic:
 *   aggregate : PyG 2.7.0 utils/_scatter.py::scatter(reduce='mean') fed by
 *               x.index_select(0, edge_index[0])  -- one pass over the edges in edge
 *               order, acc[dst] += x[src]; count[dst] += 1; acc / max(count,1)
 *   project   : SAGEConv.forward: lin_l(agg) (+bias) + lin_r(x)
 *   post      : torch LayerNorm (biased variance, eps) -> LeakyReLU(slope) / ReLU
 *               (src/models/blocks/sageblock.py:10-14)
 *   backward  : the autograd of the above
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* x [N,Fin] fp32 in; everything computed in REAL; outputs REAL. */
int FN(oracle_layer_forward)(const float* x, int64_t N, int Fin, int Fout,
                             const int64_t* edge_index, int64_t E,
                             const float* W_l, const float* b_l, const float* W_r,
                             const float* gamma, const float* beta, double eps, double slope,
                             REAL* out, REAL* agg, REAL* xhat, REAL* rstd) {
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + E;
  REAL* cnt = (REAL*)calloc((size_t)(N > 0 ? N : 1), sizeof(REAL));
  if (!cnt) return 1;
  memset(agg, 0, (size_t)N * Fin * sizeof(REAL));
  for (int64_t e = 0; e < E; ++e) {
    int64_t s = src[e], d = dst[e];
    if (s < 0 || s >= N || d < 0 || d >= N) { free(cnt); return 2; }
    cnt[d] += (REAL)1;
    for (int f = 0; f < Fin; ++f) agg[d * Fin + f] += (REAL)x[s * Fin + f];
  }
  for (int64_t i = 0; i < N; ++i) {
    REAL c = cnt[i] < (REAL)1 ? (REAL)1 : cnt[i];
    for (int f = 0; f < Fin; ++f) agg[i * Fin + f] = agg[i * Fin + f] / c;
  }
  free(cnt);
  REAL* z = (REAL*)malloc((size_t)(Fout > 0 ? Fout : 1) * sizeof(REAL));
  if (!z) return 1;
  for (int64_t i = 0; i < N; ++i) {
    for (int o = 0; o < Fout; ++o) {
      REAL a = (REAL)0, r = (REAL)0;
      for (int k = 0; k < Fin; ++k) a += agg[i * Fin + k] * (REAL)W_l[(int64_t)o * Fin + k];
      for (int k = 0; k < Fin; ++k) r += (REAL)x[i * Fin + k] * (REAL)W_r[(int64_t)o * Fin + k];
      z[o] = (a + (REAL)b_l[o]) + r;
    }
    REAL mean = (REAL)0;
    for (int o = 0; o < Fout; ++o) mean += z[o];
    mean /= (REAL)Fout;
    REAL var = (REAL)0;
    for (int o = 0; o < Fout; ++o) var += (z[o] - mean) * (z[o] - mean);
    var /= (REAL)Fout;
    REAL rs = (REAL)1 / (REAL)sqrt((double)(var + (REAL)eps));
    if (rstd) rstd[i] = rs;
    for (int o = 0; o < Fout; ++o) {
      REAL h = (z[o] - mean) * rs;
      REAL y = h * (REAL)gamma[o] + (REAL)beta[o];
      if (xhat) xhat[i * Fout + o] = h;
      out[i * Fout + o] = y > (REAL)0 ? y : (REAL)slope * y;
    }
  }
  free(z);
  return 0;
}

/* backward from saved (agg, xhat, rstd) in REAL; x/W fp32; dout REAL.
 * dx may be NULL. Gradient outputs are overwritten. */
int FN(oracle_layer_backward)(const REAL* dout, const float* x, const REAL* agg,
                              const REAL* xhat, const REAL* rstd,
                              int64_t N, int Fin, int Fout,
                              const int64_t* edge_index, int64_t E,
                              const float* W_l, const float* W_r,
                              const float* gamma, const float* beta, double slope,
                              REAL* dx, REAL* dW_l, REAL* db_l, REAL* dW_r,
                              REAL* dgamma, REAL* dbeta) {
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + E;
  memset(dW_l, 0, (size_t)Fout * Fin * sizeof(REAL));
  memset(dW_r, 0, (size_t)Fout * Fin * sizeof(REAL));
  memset(db_l, 0, (size_t)Fout * sizeof(REAL));
  memset(dgamma, 0, (size_t)Fout * sizeof(REAL));
  memset(dbeta, 0, (size_t)Fout * sizeof(REAL));
  REAL* dz = (REAL*)malloc((size_t)(Fout > 0 ? Fout : 1) * sizeof(REAL));
  REAL* dagg = (REAL*)calloc((size_t)(N > 0 ? N : 1) * Fin, sizeof(REAL));
  REAL* cnt = (REAL*)calloc((size_t)(N > 0 ? N : 1), sizeof(REAL));
  if (!dz || !dagg || !cnt) { free(dz); free(dagg); free(cnt); return 1; }
  for (int64_t e = 0; e < E; ++e) cnt[dst[e]] += (REAL)1;
  if (dx) memset(dx, 0, (size_t)N * Fin * sizeof(REAL));
  for (int64_t i = 0; i < N; ++i) {
    REAL s1 = (REAL)0, s2 = (REAL)0;
    for (int o = 0; o < Fout; ++o) {
      REAL h = xhat[i * Fout + o];
      REAL y = h * (REAL)gamma[o] + (REAL)beta[o];
      REAL dy = y > (REAL)0 ? dout[i * Fout + o] : dout[i * Fout + o] * (REAL)slope;
      dgamma[o] += dy * h;
      dbeta[o] += dy;
      REAL t = dy * (REAL)gamma[o];
      dz[o] = t;
      s1 += t;
      s2 += t * h;
    }
    REAL c1 = s1 / (REAL)Fout, c2 = s2 / (REAL)Fout;
    for (int o = 0; o < Fout; ++o) {
      REAL h = xhat[i * Fout + o];
      dz[o] = rstd[i] * (dz[o] - c1 - h * c2);
      db_l[o] += dz[o];
      for (int k = 0; k < Fin; ++k) {
        dW_l[(int64_t)o * Fin + k] += dz[o] * agg[i * Fin + k];
        dW_r[(int64_t)o * Fin + k] += dz[o] * (REAL)x[i * Fin + k];
      }
    }
    if (dx) {
      REAL c = cnt[i] < (REAL)1 ? (REAL)1 : cnt[i];
      for (int k = 0; k < Fin; ++k) {
        REAL a = (REAL)0, r = (REAL)0;
        for (int o = 0; o < Fout; ++o) {
          a += dz[o] * (REAL)W_l[(int64_t)o * Fin + k];
          r += dz[o] * (REAL)W_r[(int64_t)o * Fin + k];
        }
        dagg[i * Fin + k] = a / c;  /* backward of out / count */
        dx[i * Fin + k] = r;        /* root term */
      }
    }
  }
  if (dx) {
    /* backward of scatter_add_ (gather by dst) then of index_select (index_add_ by src), edge order */
    for (int64_t e = 0; e < E; ++e) {
      int64_t s = src[e], d = dst[e];
      for (int k = 0; k < Fin; ++k) dx[s * Fin + k] += dagg[d * Fin + k];
    }
  }
  free(dz); free(dagg); free(cnt);
  return 0;
}

#undef FN
#undef CAT
#undef CAT_
