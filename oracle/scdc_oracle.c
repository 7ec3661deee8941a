/*
 * scdc_oracle.c — CPU ORACLE (test infrastructure, NOT the product path).
 *
 * Plain-C restatement of scDiffCom's permutation iteration on the integer-coded boundary of include/scdc.h, used
 *   (1) by tests/ as the bit-exact checker of the CUDA path under uploaded labels,
 *   (2) by bench.py's cpu_baseline / `--impl reference` legs as the timed CPU stand-in for the R path
 *       (R is not installed in this image).
 * Only tests/, __graft_entry__.smoke() and those bench legs may load it. Pinned against oracle_np.py (which is pinned
 * against the reference's fixtures and known-answer tests) in tests/test_oracle_pins.py.
 *
 * Follows, per replicate (citations relative to /root/reference):
 *   R/utils_cci.R:441-462   aggregate_cells: sums[group[cell]][gene] += x over each gene column in ascending cell
 *                           order (the C loop behind DelayedArray::rowsum on a dgCMatrix; Bioconductor DelayedArray,
 *                           unpinned, not vendored), then / table(group) recomputed from the permuted labels
 *   R/utils_cci.R:464-712   build_cci_or_drate("cci"): gather subunit means per row, pmin(na.rm=TRUE), sqrt(L*R) or (L+R)/2
 *   R/utils_cci.R:197-261   compute_fast outputs: cond1, cond2 scores, L_i / R_j cond2 - cond1 differences
 *   R/utils_permutation.R:543-602  two passes (conditions permuted; then cell types permuted inside permuted conditions),
 *                           columns [DE, cond1, cond2, L_1.., R_1..]
 *   R/utils_permutation.R:139-211  counting: >=, abs() two-sided for DE and subunit columns
 *
 * oracle_philox_labels restates THIS repo's device label generator (scdiffcom_b200/csrc/scdc_kernels.cuh,
 * k_draw_labels) — not a reference algorithm: the reference draws with base::sample on L'Ecuyer streams, which cannot
 * be reproduced without R. It lets tests check the device generator bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int N, G, T, C, L, R, nL, nR;
  const int32_t *colptr, *rowidx;
  const double* values;
  const int32_t *sub_gene, *row_lri, *row_emit, *row_recv;
  const double* observed; /* column-major [R][ncols] */
} oracle_problem;

/* aggregate_cells: means[k*G + g] */
static void aggregate_cells(const oracle_problem* p, const uint8_t* group, int K, double* means, double* ncell) {
  memset(means, 0, sizeof(double) * (size_t)K * p->G);
  memset(ncell, 0, sizeof(double) * (size_t)K);
  for (int i = 0; i < p->N; ++i) ncell[group[i]] += 1.0; /* table(group) */
  for (int g = 0; g < p->G; ++g)
    for (int e = p->colptr[g]; e < p->colptr[g + 1]; ++e) means[(size_t)group[p->rowidx[e]] * p->G + g] += p->values[e];
  for (int k = 0; k < K; ++k)
    for (int g = 0; g < p->G; ++g) means[(size_t)k * p->G + g] = means[(size_t)k * p->G + g] / ncell[k];
}

static double pmin_na_rm(const double* v, int n) {
  double m = NAN;
  for (int i = 0; i < n; ++i) {
    if (isnan(v[i])) continue;
    if (isnan(m) || v[i] < m) m = v[i];
  }
  return m;
}

/* build_cci_or_drate("cci") for one row and one condition: expr[s] = subunit means (NaN = NA), returns the score */
static double row_score(const oracle_problem* p, const double* means, int cond, int row, int score_type, double* expr) {
  const int nS = p->nL + p->nR, l = p->row_lri[row];
  for (int s = 0; s < nS; ++s) {
    int g = p->sub_gene[(size_t)l * nS + s];
    int ct = (s < p->nL) ? p->row_emit[row] : p->row_recv[row];
    expr[s] = (g < 0) ? NAN : means[(size_t)(cond * p->T + ct) * p->G + g];
  }
  double mL = pmin_na_rm(expr, p->nL), mR = pmin_na_rm(expr + p->nL, p->nR);
  return score_type == 0 ? sqrt(mL * mR) : (mL + mR) / 2;
}

/*
 * perm_cond / perm_ct: replicate-major [iterations][N] label codes (perm_cond NULL in detection mode).
 * counts: column-major [R][ncols] uint64 (zeroed here). dist: NULL or [iterations][ncols_d][R].
 * de_scale: NULL or [iterations][R] (differential mode): pass-1 score of cond2 + pass-1 score of cond1, the magnitude a
 * rounding error of the DE column scales with (used by the tests of the fp32-fast mode).
 * nthreads <= 0: all OpenMP threads.
 */
int oracle_perm_counts(int N, int G, int T, int C, int L, int R, int nL, int nR, const int32_t* colptr,
                       const int32_t* rowidx, const double* values, const int32_t* ct_code, const int32_t* sub_gene,
                       const int32_t* row_lri, const int32_t* row_emit, const int32_t* row_recv, const double* observed,
                       int64_t iterations, int score_type, const uint8_t* perm_cond, const uint8_t* perm_ct,
                       uint64_t* counts, double* dist, int nthreads, double* de_scale) {
  oracle_problem P = {N, G, T, C, L, R, nL, nR, colptr, rowidx, values, sub_gene, row_lri, row_emit, row_recv, observed};
  const int K = C * T, nS = nL + nR, ncols = (C == 2) ? 3 + nS : 1, ncols_d = (C == 2) ? 3 : 1;
  memset(counts, 0, sizeof(uint64_t) * (size_t)R * ncols);
  int bad = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* m1 = (double*)malloc(sizeof(double) * (size_t)K * G);
    double* m2 = (double*)malloc(sizeof(double) * (size_t)K * G);
    double* ncell = (double*)malloc(sizeof(double) * (size_t)K);
    uint8_t* grp = (uint8_t*)malloc((size_t)N);
    uint64_t* loc = (uint64_t*)calloc((size_t)R * ncols, sizeof(uint64_t));
    if (!m1 || !m2 || !ncell || !grp || !loc) {
#pragma omp atomic write
      bad = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int64_t it = 0; it < iterations; ++it) {
        const uint8_t* pct = perm_ct + (size_t)it * N;
        double e1[16], e2[16];
        if (C == 1) {
          for (int i = 0; i < N; ++i) grp[i] = pct[i];
          aggregate_cells(&P, grp, K, m2, ncell);
          for (int r = 0; r < R; ++r) {
            double sc = row_score(&P, m2, 0, r, score_type, e1);
            if (sc >= observed[r]) ++loc[r];
            if (dist) dist[(size_t)it * R + r] = sc;
          }
        } else {
          const uint8_t* pcd = perm_cond + (size_t)it * N;
          for (int i = 0; i < N; ++i) grp[i] = (uint8_t)(pcd[i] * T + ct_code[i]); /* pass 1: permuted condition, original type */
          aggregate_cells(&P, grp, K, m1, ncell);
          for (int i = 0; i < N; ++i) grp[i] = (uint8_t)(pcd[i] * T + pct[i]);     /* pass 2: permuted type inside permuted condition */
          aggregate_cells(&P, grp, K, m2, ncell);
          for (int r = 0; r < R; ++r) {
            double s1c1 = row_score(&P, m1, 0, r, score_type, e1);
            double s1c2 = row_score(&P, m1, 1, r, score_type, e2);
            double de = s1c2 - s1c1;
            double s2c1 = row_score(&P, m2, 0, r, score_type, e1 + 8); /* e1[8..] scratch */
            double s2c2 = row_score(&P, m2, 1, r, score_type, e2 + 8);
            if (fabs(de) >= fabs(observed[r])) ++loc[r];
            if (s2c1 >= observed[(size_t)R + r]) ++loc[(size_t)R + r];
            if (s2c2 >= observed[(size_t)2 * R + r]) ++loc[(size_t)2 * R + r];
            for (int s = 0; s < nS; ++s) {
              double d = e2[s] - e1[s]; /* cond2 - cond1, NaN when the subunit is NA */
              if (fabs(d) >= fabs(observed[(size_t)(3 + s) * R + r])) ++loc[(size_t)(3 + s) * R + r];
            }
            if (dist) {
              double* d = dist + ((size_t)it * ncols_d) * R + r;
              d[0] = de; d[(size_t)R] = s2c1; d[(size_t)2 * R] = s2c2;
            }
            if (de_scale) de_scale[(size_t)it * R + r] = s1c2 + s1c1;
          }
        }
      }
#pragma omp critical
      for (size_t i = 0; i < (size_t)R * ncols; ++i) counts[i] += loc[i];
    }
    free(m1); free(m2); free(ncell); free(grp); free(loc);
  }
  return bad;
}

/* aggregate_cells on given labels: out column-major [K][G] (out[g*K + k]); mode 1 = detection rate */
int oracle_group_means(int N, int G, int K, const int32_t* colptr, const int32_t* rowidx, const double* values,
                       const uint8_t* group, int mode, double* out) {
  double* ncell = (double*)calloc((size_t)K, sizeof(double));
  if (!ncell) return 1;
  for (int i = 0; i < N; ++i) ncell[group[i]] += 1.0;
  memset(out, 0, sizeof(double) * (size_t)K * G);
  for (int g = 0; g < G; ++g)
    for (int e = colptr[g]; e < colptr[g + 1]; ++e)
      out[(size_t)g * K + group[rowidx[e]]] += mode ? (values[e] > 0 ? 1.0 : 0.0) : values[e];
  for (int g = 0; g < G; ++g)
    for (int k = 0; k < K; ++k) out[(size_t)g * K + k] = out[(size_t)g * K + k] / ncell[k];
  free(ncell);
  return 0;
}

/* ---- restatement of the device label generator (scdc_kernels.cuh: k_draw_labels) ----------------------------------
 * Philox4x32-10, key = seed. A visit makes D draws (D = 1: cell type; D = 2: condition then cell type). Draw d of visit v
 * sits at word position pos = v * D + d: attempt 0 = element pos % 4 of block (pos / 4, 0, replicate lo, replicate hi);
 * attempt a >= 1 = element 0 of block (pos, a, replicate). Cells are visited in cell order (detection) or stably sorted
 * by cell type (differential). Sequential urn draws: next label k with probability remaining[k] / remaining total. */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t y0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, y1 = (uint32_t)p1, y2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, y3 = (uint32_t)p0;
    c0 = y0; c1 = y1; c2 = y2; c3 = y3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static uint32_t stream_word(uint32_t pos, int attempt, uint64_t rep, uint64_t seed) {
  uint32_t o[4];
  if (attempt == 0) {
    philox4x32_10(pos >> 2, 0u, (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return o[pos & 3];
  }
  philox4x32_10(pos, (uint32_t)attempt, (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return o[0];
}
/* Lemire's unbiased bounded integer: attempts a = 0, 1, .. until accepted */
static uint32_t bounded_draw(uint32_t pos, uint32_t range, uint64_t rep, uint64_t seed) {
  const uint32_t t = (0u - range) % range;
  for (int a = 0;; ++a) {
    uint64_t m = (uint64_t)stream_word(pos, a, rep, seed) * range;
    if ((uint32_t)m >= t) return (uint32_t)(m >> 32);
  }
}
/* smallest k with n[0] + .. + n[k] > u */
static int urn_pick(uint32_t* n, int T, uint32_t u) {
  uint32_t acc = 0;
  for (int k = 0; k < T; ++k) {
    acc += n[k];
    if (u < acc) { --n[k]; return k; }
  }
  return -1;
}

/* Detection mode, random-key generator (mirrors keygen_plan / k_draw_labels_keys of scdc_kernels.cuh): cell i has the key
 * (A_i, B_i, i), A_i = element i % 4 of Philox block (i / 4, 0x80000000, replicate), B_i the same with c1 = 0x80000001;
 * the cells sorted by key receive the labels 0 x n_0, 1 x n_1, ... in that order. */
static int keygen_applies(int N, int T, int C) {
  if (C != 1 || T < 2) return 0;
  int NB = 256;
  while (NB < 32768 && (long long)NB * 8 < N) NB <<= 1;
  const double mean = (double)N / NB;
  if (mean > 32.0) return 0;
  const int cap = (mean <= 8.0) ? 64 : 128;
  const long long smem = 4LL * (NB + 1) + 4LL * NB / 32 + (long long)(T - 1) * cap * 8 + 4LL * 8 * T + 256;
  return smem <= 180 * 1024;
}
typedef struct { uint32_t a, b, i; } keyrec;
static int keyrec_cmp(const void* x, const void* y) {
  const keyrec *p = (const keyrec*)x, *q = (const keyrec*)y;
  if (p->a != q->a) return p->a < q->a ? -1 : 1;
  if (p->b != q->b) return p->b < q->b ? -1 : 1;
  return p->i < q->i ? -1 : (p->i > q->i);
}
static int keygen_labels(int N, int T, const uint32_t* counts, uint64_t seed, uint64_t rep, uint8_t* ct_out) {
  keyrec* rec = (keyrec*)malloc(sizeof(keyrec) * (size_t)N);
  if (!rec) return 1;
  for (int i = 0; i < N; ++i) {
    uint32_t wa[4], wb[4];
    philox4x32_10((uint32_t)i >> 2, 0x80000000u, (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), wa);
    philox4x32_10((uint32_t)i >> 2, 0x80000001u, (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), wb);
    rec[i].a = wa[i & 3]; rec[i].b = wb[i & 3]; rec[i].i = (uint32_t)i;
  }
  qsort(rec, (size_t)N, sizeof(keyrec), keyrec_cmp);
  int pos = 0;
  for (int k = 0; k < T; ++k)
    for (uint32_t c = 0; c < counts[k]; ++c) ct_out[rec[pos++].i] = (uint8_t)k;
  free(rec);
  return 0;
}

/* Differential mode, random-key generator (mirrors keygen_plan2 / k_draw_labels_keys2): key words A1 / B1 / A2 / B2 of cell i
 * = element i % 4 of Philox block (i / 4, 0x80000000 + {0,1,2,3}, replicate). Pass 1: inside cell type t the n(t, cond1) cells
 * with the smallest (A1, B1, i) get cond' = 0, the others 1. Pass 2: inside {cond' = c} the cells sorted by (A2, B2, i)
 * receive the cell types 0 x n(0, c), 1 x n(1, c), ... */
static int keygen2_applies(int N, int T, int max_nt, int max_nc) {
  if (T < 2 || max_nt > 60000) return 0;
  int NB1 = 32;
  while (NB1 < 2048 && NB1 * 8 < max_nt) NB1 <<= 1;
  while (NB1 > 32 && (long long)T * NB1 * 2 > 65536) NB1 >>= 1;
  const double mean1 = (double)max_nt / NB1;
  if (mean1 > 32.0) return 0;
  const int cap1 = (mean1 <= 8.0) ? 64 : 128;
  int NB2 = 512;
  while (NB2 < 16384 && NB2 * 8 < max_nc) NB2 <<= 1;
  const double mean2 = (double)max_nc / NB2;
  if (mean2 > 32.0) return 0;
  const int cap2 = (mean2 <= 8.0) ? 64 : 128;
  const long long region1 = 2LL * T * NB1 + 8LL * T * cap1 + 4LL * 8 * T;
  const long long region2 = 4LL * (NB2 + 1) + 4LL * NB2 / 32 + 8LL * (T - 1) * cap2 + 4LL * 8 * T + 256;
  const long long smem = 4LL * ((N + 31) / 32) + (region1 > region2 ? region1 : region2) + 256;
  return smem <= 200 * 1024;
}
static void key_words(uint32_t i, uint32_t stream, uint64_t seed, uint64_t rep, uint32_t* a, uint32_t* b) {
  uint32_t wa[4], wb[4];
  philox4x32_10(i >> 2, 0x80000000u + stream, (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), wa);
  philox4x32_10(i >> 2, 0x80000001u + stream, (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), wb);
  *a = wa[i & 3]; *b = wb[i & 3];
}
static int keygen2_labels(int N, int T, const int32_t* ct_code, const uint32_t* counts /*[2][T]*/, uint64_t seed, uint64_t rep,
                          uint8_t* cond_out, uint8_t* ct_out) {
  keyrec* rec = (keyrec*)malloc(sizeof(keyrec) * (size_t)(N > 0 ? N : 1));
  if (!rec) return 1;
  for (int t = 0; t < T; ++t) { /* pass 1 */
    int m = 0;
    for (int i = 0; i < N; ++i)
      if (ct_code[i] == t) { key_words((uint32_t)i, 0u, seed, rep, &rec[m].a, &rec[m].b); rec[m].i = (uint32_t)i; ++m; }
    qsort(rec, (size_t)m, sizeof(keyrec), keyrec_cmp);
    for (int k = 0; k < m; ++k) cond_out[rec[k].i] = (uint8_t)((uint32_t)k < counts[t] ? 0 : 1);
  }
  for (int c = 0; c < 2; ++c) { /* pass 2 */
    int m = 0;
    for (int i = 0; i < N; ++i)
      if (cond_out[i] == c) { key_words((uint32_t)i, 2u, seed, rep, &rec[m].a, &rec[m].b); rec[m].i = (uint32_t)i; ++m; }
    qsort(rec, (size_t)m, sizeof(keyrec), keyrec_cmp);
    int pos = 0;
    for (int t = 0; t < T; ++t)
      for (uint32_t k = 0; k < counts[(size_t)c * T + t] && pos < m; ++k) ct_out[rec[pos++].i] = (uint8_t)t;
  }
  free(rec);
  return 0;
}
static int urn_forced(void) { /* SCDC_URN_LABELS=1 selects the sequential urn generator on both sides */
  const char* e = getenv("SCDC_URN_LABELS");
  return e && *e && atoi(e) != 0;
}

int oracle_philox_labels(int N, int T, int C, const int32_t* ct_code, const int32_t* cond_code, uint64_t seed,
                         int64_t perm_offset, int64_t n, uint8_t* cond_out, uint8_t* ct_out) {
  uint32_t* base = (uint32_t*)calloc((size_t)2 * T, sizeof(uint32_t));
  uint32_t* remc = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)2 * T);
  uint32_t* urn = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)2 * T);
  int32_t* visit = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  if (!base || !remc || !urn || !visit) return 1;
  for (int i = 0; i < N; ++i) ++base[(size_t)((C == 2) ? cond_code[i] : 0) * T + ct_code[i]];
  if (C == 2 && !urn_forced()) {
    int max_nt = 0, max_nc = 0;
    for (int t = 0; t < T; ++t) if ((int)(base[t] + base[(size_t)T + t]) > max_nt) max_nt = (int)(base[t] + base[(size_t)T + t]);
    for (int c = 0; c < 2; ++c) {
      uint32_t n_c = 0;
      for (int t = 0; t < T; ++t) n_c += base[(size_t)c * T + t];
      if ((int)n_c > max_nc) max_nc = (int)n_c;
    }
    if (keygen2_applies(N, T, max_nt, max_nc)) {
      int rc = 0;
      for (int64_t q = 0; q < n && !rc; ++q)
        rc = keygen2_labels(N, T, ct_code, base, seed, (uint64_t)(perm_offset + q), cond_out + (size_t)q * N, ct_out + (size_t)q * N);
      free(base); free(remc); free(urn); free(visit);
      return rc;
    }
  }
  if (keygen_applies(N, T, C) && !urn_forced()) {
    int rc = 0;
    for (int64_t q = 0; q < n && !rc; ++q)
      rc = keygen_labels(N, T, base, seed, (uint64_t)(perm_offset + q), ct_out + (size_t)q * N);
    free(base); free(remc); free(urn); free(visit);
    return rc;
  }
  if (C == 2) { /* stable sort of the cells by cell type */
    int v = 0;
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < N; ++i)
        if (ct_code[i] == t) visit[v++] = i;
  } else {
    for (int i = 0; i < N; ++i) visit[i] = i;
  }
  for (int64_t q = 0; q < n; ++q) {
    const uint64_t rep = (uint64_t)(perm_offset + q);
    memcpy(remc, base, sizeof(uint32_t) * (size_t)2 * T);
    memcpy(urn, base, sizeof(uint32_t) * (size_t)2 * T);
    uint32_t tot[2] = {0, 0};
    for (int t = 0; t < T; ++t) { tot[0] += base[t]; tot[1] += base[(size_t)T + t]; }
    for (int v = 0; v < N; ++v) {
      const int i = visit[v];
      int c = 0;
      if (C == 2) {
        const int t = ct_code[i];
        const uint32_t r0 = remc[t], r1 = remc[(size_t)T + t];
        const uint32_t u = bounded_draw((uint32_t)v * 2u, r0 + r1, rep, seed);
        c = (u < r0) ? 0 : 1;
        --remc[(size_t)c * T + t];
        cond_out[(size_t)q * N + i] = (uint8_t)c;
      }
      const uint32_t u = bounded_draw((uint32_t)v * (uint32_t)C + (uint32_t)(C - 1), tot[c], rep, seed);
      --tot[c];
      ct_out[(size_t)q * N + i] = (uint8_t)urn_pick(urn + (size_t)c * T, T, u);
    }
  }
  free(base); free(remc); free(urn); free(visit);
  return 0;
}
