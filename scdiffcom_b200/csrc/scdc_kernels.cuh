// scdc_kernels.cuh — sm_100a kernels of the permutation-test hot path (see DESIGN.md for the data layout).
//
//   K1   k_draw_labels<NG>   on-device Philox4x32-10 label shuffles (sequential urn draws, one thread per replicate)
//                             (replaces base::sample at reference R/utils_permutation.R:509,536-541,552-569)
//   K1u  k_pack_uploaded /    validation mode: re-layout and check uploaded label matrices
//        k_check_uploaded
//   K2a  k_pass1_cond         differential pass 1 (conditions shuffled inside fixed cell types): register accumulators
//   K1k2 k_draw_labels_keys2 differential-mode random-key shuffles (both passes), one block per replicate
//   K2b  k_scatter<J,U,PAIR>  general segmented sum over CSC, lane = replicate, shared-memory accumulators
//                             (replaces DelayedArray::rowsum + "/ table(group)", reference R/utils_cci.R:441-462)
//   K3   k_score_det /        gather + complex-min + score + diff + `>=` compare + counter bump, factorised per LRI
//        k_score_diff         (replaces build_cci_or_drate("cci"), reference R/utils_cci.R:464-712, the column assembly
//                             of R/utils_permutation.R:579-602 and the counting of R/utils_permutation.R:139-211)
//        k_score_count        generic warp-per-row variant (return_distributions, cross-check)
//        k_product_thresholds thresholds on the product for the one-sided geometric columns
//   obs  k_group_means /      observed pass on the original labels (reference R/utils_cci.R:171-283)
//        k_observed_rows
//
// Exactness contract: every (gene, group) sum is accumulated in fp64 by ONE lane walking the gene column in ascending
// cell order — the order of the C loop behind rowsum — so means, scores and therefore exceedance counts are bit-identical
// to the CPU oracle for the same labels. Compile with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scdc {

// ------------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based, key = seed. A visit makes D draws (D = 1: cell type, detection
// mode; D = 2: condition then cell type, differential mode). Draw d of visit v sits at word position pos = v * D + d of
// the replicate's stream: attempt 0 reads element pos % 4 of block (c0 = pos / 4, c1 = 0, replicate lo, replicate hi);
// a rejected attempt a >= 1 (probability < range / 2^32) reads element 0 of block (c0 = pos, c1 = a, replicate).
// Every draw is a pure function of (seed, replicate, visit, attempt): no sequential generator state, any replicate
// range can be regenerated anywhere, and one Philox block serves 4 / D consecutive visits.
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t y0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, y1 = (uint32_t)p1, y2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, y3 = (uint32_t)p0;
    c0 = y0; c1 = y1; c2 = y2; c3 = y3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform integer in [0, range): Lemire's multiply-shift with rejection (exactly unbiased). w0 = the attempt-0 word.
__host__ __device__ inline uint32_t bounded_draw(uint32_t w0, uint32_t pos, uint32_t range, uint32_t r_lo, uint32_t r_hi,
                                                 uint32_t k0, uint32_t k1) {
  uint64_t m = (uint64_t)w0 * range;
  uint32_t l = (uint32_t)m;
  if (l < range) {
    const uint32_t t = (0u - range) % range;
    uint32_t a = 0;
    while (l < t) {  // rejected: next attempt's word
      ++a;
      uint32_t o[4];
      philox4x32_10(pos, a, r_lo, r_hi, k0, k1, o);
      m = (uint64_t)o[0] * range;
      l = (uint32_t)m;
    }
  }
  return (uint32_t)(m >> 32);
}

// byte offset of replicate q's label inside one cell's label row (row length Bpad bytes), J replicates per lane
__host__ __device__ inline int lab2_offset(int q, int J) {
  int chunk = q / (32 * J), r = q % (32 * J);
  return chunk * 32 * J + (r % 32) * J + r / 32;
}

// ------------------------------------------------------------------------------------------------------------------
// K1: one thread = one replicate; sequential urn draws over the cells: a uniformly random arrangement of a multiset
// revealed cell by cell (next label k with probability remaining[k] / remaining total) — the distribution of
// base::sample on the same strata.
//   detection   : ct' = global shuffle of cell types                    (reference R/utils_permutation.R:509)
//   differential: cond' = shuffle of conditions inside each cell type   (:536-541)
//                 ct'   = shuffle of cell types inside each cond' class (:552-569)
// Cells are visited in `visit` order (differential: stably sorted by cell type, so the condition urn of the current
// cell type is two registers; detection: cell order). The cell-type urn is two-level: NG group sums (8 labels per
// group) in registers, the per-label remaining counts in shared memory laid out [label][thread] (bank = lane).
// Philox blocks depend only on (seed, replicate, visit), so they are computed off the urn's dependency chain.
// outputs, per reduction batch of Bb replicates (batch b = q / Bb holds replicates q' = q % Bb):
//   bits1[b][cell][Bb/32] (bit q'%32 of word q'/32 = cond'), lab2[b][cell][Bb] (group code cond' * T + ct', K2b layout).
// Batch-major blocks keep the row stride of the labels at Bb bytes however many replicates one launch draws. Optional replicate-major debug copies for scdc_draw_labels.
// ------------------------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(32) k_draw_labels(int N, int T, int C, const int* __restrict__ visit,
                                                    const int* __restrict__ ct_seg, const uint32_t* __restrict__ ngroup_u,
                                                    uint64_t seed, int64_t perm0, int Bb, int J,
                                                    uint32_t* __restrict__ bits1, uint8_t* __restrict__ lab2,
                                                    uint8_t* __restrict__ dbg_cond, uint8_t* __restrict__ dbg_ct) {
  extern __shared__ uint32_t sm_u32[];  // [C][NG*8][32]
  const int tx = threadIdx.x;
  const int q = blockIdx.x * 32 + tx;  // replicate inside the super-batch (a multiple of 32)
  const uint64_t rep = (uint64_t)(perm0 + q);
  const uint32_t r_lo = (uint32_t)rep, r_hi = (uint32_t)(rep >> 32), k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t gs[2][NG], tot[2] = {0, 0};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      uint32_t sum = 0;
      for (int k = 0; k < 8; ++k) {
        const int lab = j * 8 + k;
        const uint32_t n = (c < C && lab < T) ? ngroup_u[c * T + lab] : 0u;
        if (c < C) sm_u32[((c * NG * 8) + lab) * 32 + tx] = n;
        sum += n;
      }
      gs[c][j] = sum;
      tot[c] += sum;
    }
  }
  const int Bw = Bb / 32, qb = q % Bb;
  uint8_t* const lab2_q = lab2 + (size_t)(q / Bb) * N * Bb + lab2_offset(qb, J);
  uint32_t* const bits1_q = bits1 + (size_t)(q / Bb) * N * Bw + (qb >> 5);
  const int D = C;  // draws per visit
  uint32_t blk[4], nxt[4];
  philox4x32_10(0u, 0u, r_lo, r_hi, k0, k1, blk);
  philox4x32_10(1u, 0u, r_lo, r_hi, k0, k1, nxt);  // the next block is always computed one block ahead, off the urn's chain
  int v = 0;
  const int n_seg = (C == 2) ? T : 1;
  for (int t = 0; t < n_seg; ++t) {
    const int v_end = (C == 2) ? ct_seg[t + 1] : N;
    uint32_t r0 = 0, r1 = 0;
    if (C == 2) { r0 = ngroup_u[t]; r1 = ngroup_u[T + t]; }
    for (; v < v_end; ++v) {
      const uint32_t pos0 = (uint32_t)v * (uint32_t)D;
      const int el = (int)(pos0 & 3u);
      if (v > 0 && el == 0) {
#pragma unroll
        for (int w = 0; w < 4; ++w) blk[w] = nxt[w];
        philox4x32_10((pos0 >> 2) + 1u, 0u, r_lo, r_hi, k0, k1, nxt);
      }
      const int cell = visit ? visit[v] : v;
      int c = 0;
      uint32_t w_ct;
      if (C == 2) {
        const uint32_t w_cd = (el == 0) ? blk[0] : blk[2];  // pos0 is even
        w_ct = (el == 0) ? blk[1] : blk[3];
        const uint32_t u = bounded_draw(w_cd, pos0, r0 + r1, r_lo, r_hi, k0, k1);
        c = (u < r0) ? 0 : 1;
        if (c) --r1; else --r0;
      } else {
        w_ct = (el == 0) ? blk[0] : (el == 1) ? blk[1] : (el == 2) ? blk[2] : blk[3];
      }
      uint32_t u = bounded_draw(w_ct, pos0 + (uint32_t)(D - 1), tot[c], r_lo, r_hi, k0, k1);
      tot[c] -= 1;
      // level 1: group
      int grp = NG - 1;
      uint32_t run = 0, base = 0;
      bool found = false;
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        const uint32_t g = c ? gs[1][j] : gs[0][j];
        const uint32_t nr = run + g;
        if (!found && u < nr) { grp = j; base = run; found = true; }
        run = nr;
      }
      u -= base;
      // level 2: label inside the group
      uint32_t* cp = sm_u32 + ((c * NG * 8) + grp * 8) * 32 + tx;
      uint32_t n8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) n8[k] = cp[k * 32];
      int kk = 7;
      run = 0;
      found = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t nr = run + n8[k];
        if (!found && u < nr) { kk = k; found = true; }
        run = nr;
      }
      uint32_t nsel = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) nsel = (k == kk) ? n8[k] : nsel;
      cp[kk * 32] = nsel - 1;
#pragma unroll
      for (int j = 0; j < NG; ++j) {
        if (j == grp) { if (c) --gs[1][j]; else --gs[0][j]; }
      }
      const int pos = grp * 8 + kk;
      if (C == 2) {
        const uint32_t word = __ballot_sync(0xffffffffu, c == 1);
        if (tx == 0) bits1_q[(size_t)cell * Bw] = word;
      }
      lab2_q[(size_t)cell * Bb] = (uint8_t)(c * T + pos);
      if (dbg_ct) dbg_ct[(size_t)q * N + cell] = (uint8_t)pos;
      if (dbg_cond) dbg_cond[(size_t)q * N + cell] = (uint8_t)c;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1k: detection-mode label shuffles by RANDOM KEYS, one thread block per replicate (fully parallel over the cells).
// Cell i gets the key (A_i, B_i, i): A_i = element i % 4 of Philox block (i / 4, 0x80000000, replicate), B_i likewise with
// c1 = 0x80000001 (only evaluated to break a tie on A). Sorting the cells by key is a uniformly random permutation; the
// first n_0 cells of that order receive label 0, the next n_1 label 1, ... (the distribution of sample(cell_type),
// reference R/utils_permutation.R:509). No sort is needed: label(i) = #{j : theta_j <= key_i} where theta_j is the cell
// of rank C_j = n_0 + .. + n_j, found by histogram selection:
//   pass 1  histogram of the top bits of A (NB bins, shared-memory atomics), exclusive scan;
//   targets each rank C_j falls into one bin; the members of the target bins are collected (pass 2, <= cap per target) and
//           the member with the right rank inside the bin is the threshold theta_j;
//   pass 3  every cell compares its key with the T-1 thresholds (binary search) and writes its label.
// Output: replicate-major labels out[q][cell] (re-laid out by k_pack_uploaded). fail[0] is set if a target bin holds more
// than cap cells (probability < 1e-30 with the NB / cap chosen by keygen_plan).
// ------------------------------------------------------------------------------------------------------------------
struct KeygenPlan {
  int ok;     // 1: the random-key generator applies, 0: use the sequential urn generator
  int NB;     // histogram bins (power of two)
  int cap;    // members kept per target bin
  int smem;   // dynamic shared memory in bytes
};

__host__ __device__ inline KeygenPlan keygen_plan(int N, int T, int C) {
  KeygenPlan pl = {0, 0, 0, 0};
  if (C != 1 || T < 2) return pl;
  int NB = 256;
  while (NB < 32768 && (long long)NB * 8 < N) NB <<= 1;
  const double mean = (double)N / NB;
  if (mean > 32.0) return pl;
  const int cap = (mean <= 8.0) ? 64 : 128;
  const long long smem = 4LL * (NB + 1) + 4LL * NB / 32 + (long long)(T - 1) * cap * 8 + 4LL * 8 * T + 256;
  if (smem > 180 * 1024) return pl;
  pl.ok = 1; pl.NB = NB; pl.cap = cap; pl.smem = (int)smem;
  return pl;
}

__device__ __forceinline__ uint32_t keygen_word(uint32_t i, uint32_t stream, uint32_t r_lo, uint32_t r_hi, uint32_t k0, uint32_t k1) {
  uint32_t w[4];
  philox4x32_10(i >> 2, 0x80000000u | stream, r_lo, r_hi, k0, k1, w);
  return w[i & 3];
}

// (a_x, b_x, x) < (a_y, b_y, y) with the B words evaluated only on a tie of the A words
__device__ __forceinline__ bool keygen_less(uint32_t ax, uint32_t x, uint32_t ay, uint32_t y, uint32_t r_lo, uint32_t r_hi,
                                            uint32_t k0, uint32_t k1) {
  if (ax != ay) return ax < ay;
  const uint32_t bx = keygen_word(x, 1u, r_lo, r_hi, k0, k1), by = keygen_word(y, 1u, r_lo, r_hi, k0, k1);
  if (bx != by) return bx < by;
  return x < y;
}

__global__ void __launch_bounds__(256) k_draw_labels_keys(int N, int T, const uint32_t* __restrict__ ngroup_u, uint64_t seed,
                                                          int64_t perm0, int NB, int cap, uint8_t* __restrict__ out,
                                                          int* __restrict__ fail) {
  extern __shared__ uint32_t sm_k[];
  const int tid = threadIdx.x, nthr = blockDim.x, m = T - 1;
  const int q = blockIdx.x;
  const uint64_t rep = (uint64_t)(perm0 + q);
  const uint32_t r_lo = (uint32_t)rep, r_hi = (uint32_t)(rep >> 32), k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  int lg = 0;
  while ((1 << lg) < NB) ++lg;
  const int shift = 32 - lg;
  uint32_t* hist = sm_k;                       // [NB + 1] counts, then exclusive prefix
  uint32_t* mark = hist + NB + 1;              // [NB / 32] bitmap of target bins
  uint32_t* tbin = mark + NB / 32;             // [m] bin of target j (ascending in j)
  uint32_t* tq = tbin + m;                     // [m] rank of the threshold inside its bin
  uint32_t* lcnt = tq + m;                     // [m] members collected
  uint32_t* thrA = lcnt + m;                   // [m] thresholds: A word, B word, cell
  uint32_t* thrB = thrA + m;
  uint32_t* thrI = thrB + m;
  uint32_t* wsum = thrI + m;                   // [32] scan scratch
  uint32_t* lkey = wsum + 32;                  // [m][cap]
  uint32_t* lidx = lkey + m * cap;             // [m][cap]
  for (int b = tid; b <= NB; b += nthr) hist[b] = 0;
  for (int b = tid; b < NB / 32; b += nthr) mark[b] = 0;
  for (int j = tid; j < m; j += nthr) lcnt[j] = 0;
  __syncthreads();
  // ---- pass 1: histogram of the top bits of A ----
  for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {
    uint32_t w[4];
    philox4x32_10((uint32_t)i4 >> 2, 0x80000000u, r_lo, r_hi, k0, k1, w);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i4 + k < N) atomicAdd(&hist[w[k] >> shift], 1u);
  }
  __syncthreads();
  // ---- exclusive scan of the NB bins (NB / nthr consecutive bins per thread) ----
  {
    const int per = NB / nthr;  // NB >= 256 = nthr
    uint32_t loc = 0;
    for (int k = 0; k < per; ++k) loc += hist[tid * per + k];
    uint32_t inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
      if ((tid & 31) >= d) inc += v;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
      uint32_t v = (tid < nthr / 32) ? wsum[tid] : 0, iv = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, iv, d);
        if (tid >= d) iv += u;
      }
      wsum[tid] = iv - v;  // exclusive warp offsets
    }
    __syncthreads();
    uint32_t run = wsum[tid >> 5] + inc - loc;
    for (int k = 0; k < per; ++k) {
      const uint32_t c = hist[tid * per + k];
      hist[tid * per + k] = run;
      run += c;
    }
    if (tid == nthr - 1) hist[NB] = run;
  }
  __syncthreads();
  // ---- targets: rank C_j -> (bin, rank inside the bin) ----
  for (int j = tid; j < m; j += nthr) {
    uint32_t r = 0;
    for (int k = 0; k <= j; ++k) r += ngroup_u[k];
    int lo = 0, hi = NB;  // first bin with prefix > r
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (hist[mid + 1] > r) hi = mid; else lo = mid + 1;
    }
    tbin[j] = (uint32_t)lo;
    tq[j] = r - hist[lo];
    atomicOr(&mark[lo >> 5], 1u << (lo & 31));
  }
  __syncthreads();
  // ---- pass 2: collect the members of the target bins ----
  for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {
    uint32_t w[4];
    philox4x32_10((uint32_t)i4 >> 2, 0x80000000u, r_lo, r_hi, k0, k1, w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i4 + k >= N) continue;
      const uint32_t bin = w[k] >> shift;
      if (!((mark[bin >> 5] >> (bin & 31)) & 1u)) continue;
      int lo = 0, hi = m;  // first target with tbin >= bin
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tbin[mid] >= bin) hi = mid; else lo = mid + 1;
      }
      for (int j = lo; j < m && tbin[j] == bin; ++j) {
        const uint32_t pos = atomicAdd(&lcnt[j], 1u);
        if (pos < (uint32_t)cap) { lkey[j * cap + pos] = w[k]; lidx[j * cap + pos] = (uint32_t)(i4 + k); }
        else atomicExch(fail, 1);
      }
    }
  }
  __syncthreads();
  // ---- thresholds: the member with tq[j] smaller members ----
  for (int pair = tid; pair < m * cap; pair += nthr) {
    const int j = pair / cap, x = pair % cap;
    const int mj = min((int)lcnt[j], cap);
    if (x >= mj) continue;
    const uint32_t ax = lkey[j * cap + x], ix = lidx[j * cap + x];
    uint32_t less = 0;
    for (int y = 0; y < mj; ++y)
      if (y != x && keygen_less(lkey[j * cap + y], lidx[j * cap + y], ax, ix, r_lo, r_hi, k0, k1)) ++less;
    if (less == tq[j]) {
      thrA[j] = ax;
      thrI[j] = ix;
      thrB[j] = keygen_word(ix, 1u, r_lo, r_hi, k0, k1);
    }
  }
  __syncthreads();
  // ---- pass 3: label = number of thresholds <= key ----
  uint8_t* o = out + (size_t)q * N;
  const bool aligned = (reinterpret_cast<size_t>(o) & 3) == 0;
  for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {
    uint32_t w[4];
    philox4x32_10((uint32_t)i4 >> 2, 0x80000000u, r_lo, r_hi, k0, k1, w);
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t i = (uint32_t)(i4 + k), a = w[k];
      const uint32_t bin = a >> shift;
      int lo = 0, hi = m;  // first threshold > key
      if (!((mark[bin >> 5] >> (bin & 31)) & 1u)) {
        // not a target bin: every threshold lies in another bin, so only the bins need comparing (first tbin > bin)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (tbin[mid] < bin) lo = mid + 1; else hi = mid;
        }
        packed |= (uint32_t)lo << (8 * k);
        continue;
      }
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        bool thr_le;  // theta_mid <= key_i
        if (thrA[mid] != a) thr_le = thrA[mid] < a;
        else if (thrI[mid] == i) thr_le = true;
        else {
          const uint32_t b = keygen_word(i, 1u, r_lo, r_hi, k0, k1);
          thr_le = (thrB[mid] != b) ? (thrB[mid] < b) : (thrI[mid] < i);
        }
        if (thr_le) lo = mid + 1; else hi = mid;
      }
      packed |= (uint32_t)lo << (8 * k);
    }
    if (aligned && i4 + 3 < N) {
      *reinterpret_cast<uint32_t*>(o + i4) = packed;
    } else {
      for (int k = 0; k < 4 && i4 + k < N; ++k) o[i4 + k] = (uint8_t)(packed >> (8 * k));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1k2: differential-mode label shuffles by random keys, one thread block per replicate.
//   pass 1  cond' = shuffle of the conditions inside every cell type (reference R/utils_permutation.R:536-541): inside
//           cell type t the n(t, cond1) cells with the smallest keys (A1, B1, cell) get cond' = 0, the others cond' = 1;
//   pass 2  ct' = shuffle of the cell types inside every PERMUTED condition (:552-569): inside {cells with cond' = c} the
//           cells sorted by the keys (A2, B2, cell) receive the labels 0 x n(0, c), 1 x n(1, c), ...
// Key words: A1 / B1 / A2 / B2 of cell i = element i % 4 of Philox block (i / 4, 0x80000000 + {0, 1, 2, 3}, replicate).
// Both passes select rank thresholds by histogram (pass 1: T strata x NB1 bins, 16-bit counters packed in pairs, one
// threshold per stratum; pass 2: one stratum at a time, NB2 bins, T - 1 thresholds) and rank exactly only the members of
// the bins that contain a threshold. cond' lives in a shared-memory bitmap between the passes.
// Output: replicate-major combined labels out[q][cell] = cond' * T + ct' (or the two debug arrays).
// block = 512 or 1024 threads (a power of two <= NB2: the bin scan gives every thread NB2 / blockDim bins); the result does
// not depend on it.
// ------------------------------------------------------------------------------------------------------------------
struct KeygenPlan2 {
  int ok;
  int NB1, cap1;  // pass 1: bins per cell type, members kept per threshold bin
  int NB2, cap2;  // pass 2
  int smem;
};

__host__ __device__ inline KeygenPlan2 keygen_plan2(int N, int T, int max_nt, int max_nc) {
  KeygenPlan2 pl = {0, 0, 0, 0, 0, 0};
  if (T < 2 || max_nt > 60000) return pl;  // 16-bit bin counters
  int NB1 = 32;
  while (NB1 < 2048 && NB1 * 8 < max_nt) NB1 <<= 1;
  while (NB1 > 32 && (long long)T * NB1 * 2 > 65536) NB1 >>= 1;
  const double mean1 = (double)max_nt / NB1;
  if (mean1 > 32.0) return pl;
  const int cap1 = (mean1 <= 8.0) ? 64 : 128;
  int NB2 = 512;
  while (NB2 < 16384 && NB2 * 8 < max_nc) NB2 <<= 1;
  const double mean2 = (double)max_nc / NB2;
  if (mean2 > 32.0) return pl;
  const int cap2 = (mean2 <= 8.0) ? 64 : 128;
  const long long region1 = 2LL * T * NB1 + 8LL * T * cap1 + 4LL * 8 * T;
  const long long region2 = 4LL * (NB2 + 1) + 4LL * NB2 / 32 + 8LL * (T - 1) * cap2 + 4LL * 8 * T + 256;
  const long long smem = 4LL * ((N + 31) / 32) + (region1 > region2 ? region1 : region2) + 256;
  if (smem > 200 * 1024) return pl;
  pl.ok = 1; pl.NB1 = NB1; pl.cap1 = cap1; pl.NB2 = NB2; pl.cap2 = cap2; pl.smem = (int)smem;
  return pl;
}

// key order with explicit stream numbers: (a, b, cell), b evaluated only on a tie of a
__device__ __forceinline__ bool keygen_less_s(uint32_t ax, uint32_t x, uint32_t ay, uint32_t y, uint32_t sb, uint32_t r_lo,
                                              uint32_t r_hi, uint32_t k0, uint32_t k1) {
  if (ax != ay) return ax < ay;
  const uint32_t bx = keygen_word(x, sb, r_lo, r_hi, k0, k1), by = keygen_word(y, sb, r_lo, r_hi, k0, k1);
  if (bx != by) return bx < by;
  return x < y;
}

__global__ void __launch_bounds__(1024) k_draw_labels_keys2(int N, int T, const uint8_t* __restrict__ ct,
                                                           const uint32_t* __restrict__ ngroup_u /*[2][T]*/, uint64_t seed,
                                                           int64_t perm0, int NB1, int cap1, int NB2, int cap2,
                                                           uint8_t* __restrict__ out, uint8_t* __restrict__ dbg_cond,
                                                           uint8_t* __restrict__ dbg_ct, int* __restrict__ fail) {
  extern __shared__ uint32_t sm_k[];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int q = blockIdx.x;
  const uint64_t rep = (uint64_t)(perm0 + q);
  const uint32_t r_lo = (uint32_t)rep, r_hi = (uint32_t)(rep >> 32), k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const int nwords = (N + 31) / 32;
  uint32_t* bitmap = sm_k;            // [nwords] cond' of every cell
  uint32_t* region = bitmap + nwords;
  for (int w = tid; w < nwords; w += nthr) bitmap[w] = 0;
  // =========================== pass 1: conditions inside every cell type ===========================
  {
    int lg = 0;
    while ((1 << lg) < NB1) ++lg;
    const int shift = 32 - lg;
    uint32_t* hist = region;                  // [T * NB1 / 2] pairs of 16-bit counters
    uint32_t* tb = hist + T * NB1 / 2;        // [T] bin of the threshold (0xffffffff: every cell keeps cond' = 0)
    uint32_t* tq = tb + T;                    // [T] rank of the threshold inside its bin
    uint32_t* lcnt = tq + T;                  // [T]
    uint32_t* thrA = lcnt + T;                // [T] threshold key
    uint32_t* thrB = thrA + T;
    uint32_t* thrI = thrB + T;
    uint32_t* lkey = thrI + T + 2 * T;        // [T][cap1]   (2 T words of slack keep the layout equal to the plan's size)
    uint32_t* lidx = lkey + T * cap1;         // [T][cap1]
    for (int b = tid; b < T * NB1 / 2; b += nthr) hist[b] = 0;
    for (int t = tid; t < T; t += nthr) lcnt[t] = 0;
    __syncthreads();
    for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {
      uint32_t w[4];
      philox4x32_10((uint32_t)i4 >> 2, 0x80000000u, r_lo, r_hi, k0, k1, w);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i4 + k < N) {
          const uint32_t idx = (uint32_t)ct[i4 + k] * NB1 + (w[k] >> shift);
          atomicAdd(&hist[idx >> 1], 1u << ((idx & 1) * 16));
        }
    }
    __syncthreads();
    // thresholds: one warp per cell type; theta_t = the cell of rank n(t, cond1) = first cell that gets cond' = 1
    for (int t = warp; t < T; t += nwarp) {
      const uint32_t r = ngroup_u[t], n_t = ngroup_u[t] + ngroup_u[T + t];
      const int per = NB1 / 32;
      uint32_t loc = 0;
      for (int k = 0; k < per; ++k) {
        const uint32_t idx = (uint32_t)t * NB1 + lane * per + k;
        loc += (hist[idx >> 1] >> ((idx & 1) * 16)) & 0xffffu;
      }
      uint32_t inc = loc;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
      if (lane == 0) {
        if (total != n_t) atomicExch(fail, 3);  // a 16-bit bin counter overflowed
        if (r >= n_t) tb[t] = 0xffffffffu;
      }
      const uint32_t exc = inc - loc;
      if (r < n_t && exc <= r && r < inc) {
        uint32_t run = exc;
        for (int k = 0; k < per; ++k) {
          const uint32_t idx = (uint32_t)t * NB1 + lane * per + k;
          const uint32_t c = (hist[idx >> 1] >> ((idx & 1) * 16)) & 0xffffu;
          if (r < run + c) { tb[t] = lane * per + k; tq[t] = r - run; break; }
          run += c;
        }
      }
    }
    __syncthreads();
    for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {  // members of the threshold bins
      uint32_t w[4];
      philox4x32_10((uint32_t)i4 >> 2, 0x80000000u, r_lo, r_hi, k0, k1, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i4 + k >= N) continue;
        const uint32_t t = ct[i4 + k];
        if ((w[k] >> shift) != tb[t]) continue;
        const uint32_t pos = atomicAdd(&lcnt[t], 1u);
        if (pos < (uint32_t)cap1) { lkey[t * cap1 + pos] = w[k]; lidx[t * cap1 + pos] = (uint32_t)(i4 + k); }
        else atomicExch(fail, 1);
      }
    }
    __syncthreads();
    for (int pair = tid; pair < T * cap1; pair += nthr) {
      const int t = pair / cap1, x = pair % cap1;
      const int mj = min((int)lcnt[t], cap1);
      if (x >= mj || tb[t] == 0xffffffffu) continue;
      const uint32_t ax = lkey[t * cap1 + x], ix = lidx[t * cap1 + x];
      uint32_t less = 0;
      for (int y = 0; y < mj; ++y)
        if (y != x && keygen_less_s(lkey[t * cap1 + y], lidx[t * cap1 + y], ax, ix, 1u, r_lo, r_hi, k0, k1)) ++less;
      if (less == tq[t]) {
        thrA[t] = ax;
        thrI[t] = ix;
        thrB[t] = keygen_word(ix, 1u, r_lo, r_hi, k0, k1);
      }
    }
    __syncthreads();
    for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {  // cond' = (key >= theta of the cell's type)
      uint32_t w[4];
      philox4x32_10((uint32_t)i4 >> 2, 0x80000000u, r_lo, r_hi, k0, k1, w);
      uint32_t nib = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i4 + k >= N) continue;
        const uint32_t i = (uint32_t)(i4 + k), t = ct[i], a = w[k], bin = a >> shift, tbin = tb[t];
        bool one;
        if (tbin == 0xffffffffu) one = false;
        else if (bin != tbin) one = bin > tbin;
        else if (a != thrA[t]) one = a > thrA[t];
        else if (i == thrI[t]) one = true;
        else {
          const uint32_t b = keygen_word(i, 1u, r_lo, r_hi, k0, k1);
          one = (b != thrB[t]) ? (b > thrB[t]) : (i > thrI[t]);
        }
        nib |= (one ? 1u : 0u) << k;
      }
      if (nib) atomicOr(&bitmap[i4 >> 5], nib << (i4 & 31));
    }
    __syncthreads();
  }
  // =========================== pass 2: cell types inside every permuted condition ===========================
  {
    const int m = T - 1;
    int lg = 0;
    while ((1 << lg) < NB2) ++lg;
    const int shift = 32 - lg;
    uint32_t* hist = region;                     // [NB2 + 1] counts, then exclusive prefix
    uint32_t* mark = hist + NB2 + 1;             // [NB2 / 32] bitmap of threshold bins
    uint32_t* tbin = mark + NB2 / 32;            // [m]
    uint32_t* tq = tbin + m;
    uint32_t* lcnt = tq + m;
    uint32_t* thrA = lcnt + m;
    uint32_t* thrB = thrA + m;
    uint32_t* thrI = thrB + m;
    uint32_t* wsum = thrI + m;                   // [32]
    uint32_t* lkey = wsum + 32;                  // [m][cap2]
    uint32_t* lidx = lkey + m * cap2;
    for (int c = 0; c < 2; ++c) {
      const uint32_t* cnt = ngroup_u + c * T;    // n(t, c), t = 0..T-1
      for (int b = tid; b <= NB2; b += nthr) hist[b] = 0;
      for (int b = tid; b < NB2 / 32; b += nthr) mark[b] = 0;
      for (int j = tid; j < m; j += nthr) lcnt[j] = 0;
      __syncthreads();
      for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {
        const uint32_t member = ((c ? bitmap[i4 >> 5] : ~bitmap[i4 >> 5]) >> (i4 & 31)) & 0xfu;
        if (!member) continue;
        uint32_t w[4];
        philox4x32_10((uint32_t)i4 >> 2, 0x80000002u, r_lo, r_hi, k0, k1, w);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i4 + k < N && ((member >> k) & 1u)) atomicAdd(&hist[w[k] >> shift], 1u);
      }
      __syncthreads();
      {  // exclusive scan of the NB2 bins (NB2 / nthr consecutive bins per thread)
        const int per = NB2 / nthr;
        uint32_t loc = 0;
        for (int k = 0; k < per; ++k) loc += hist[tid * per + k];
        uint32_t inc = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += v;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (tid < 32) {
          uint32_t v = (tid < nwarp) ? wsum[tid] : 0, iv = v;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, iv, d);
            if (tid >= d) iv += u;
          }
          wsum[tid] = iv - v;
        }
        __syncthreads();
        uint32_t run = wsum[warp] + inc - loc;
        for (int k = 0; k < per; ++k) {
          const uint32_t cc = hist[tid * per + k];
          hist[tid * per + k] = run;
          run += cc;
        }
        if (tid == nthr - 1) hist[NB2] = run;
      }
      __syncthreads();
      for (int j = tid; j < m; j += nthr) {  // rank n(0,c) + .. + n(j,c) -> (bin, rank inside the bin)
        uint32_t r = 0;
        for (int k = 0; k <= j; ++k) r += cnt[k];
        int lo = 0, hi = NB2;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (hist[mid + 1] > r) hi = mid; else lo = mid + 1;
        }
        tbin[j] = (uint32_t)lo;  // lo == NB2: rank beyond the stratum (trailing empty labels), a threshold no key reaches
        tq[j] = (lo < NB2) ? r - hist[lo] : 0u;
        if (lo < NB2) atomicOr(&mark[lo >> 5], 1u << (lo & 31));
      }
      __syncthreads();
      for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {  // members of the threshold bins
        const uint32_t member = ((c ? bitmap[i4 >> 5] : ~bitmap[i4 >> 5]) >> (i4 & 31)) & 0xfu;
        if (!member) continue;
        uint32_t w[4];
        philox4x32_10((uint32_t)i4 >> 2, 0x80000002u, r_lo, r_hi, k0, k1, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (i4 + k >= N || !((member >> k) & 1u)) continue;
          const uint32_t bin = w[k] >> shift;
          if (!((mark[bin >> 5] >> (bin & 31)) & 1u)) continue;
          int lo = 0, hi = m;  // first threshold with tbin >= bin
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (tbin[mid] >= bin) hi = mid; else lo = mid + 1;
          }
          for (int j = lo; j < m && tbin[j] == bin; ++j) {
            const uint32_t pos = atomicAdd(&lcnt[j], 1u);
            if (pos < (uint32_t)cap2) { lkey[j * cap2 + pos] = w[k]; lidx[j * cap2 + pos] = (uint32_t)(i4 + k); }
            else atomicExch(fail, 1);
          }
        }
      }
      __syncthreads();
      for (int pair = tid; pair < m * cap2; pair += nthr) {
        const int j = pair / cap2, x = pair % cap2;
        const int mj = min((int)lcnt[j], cap2);
        if (x >= mj) continue;
        const uint32_t ax = lkey[j * cap2 + x], ix = lidx[j * cap2 + x];
        uint32_t less = 0;
        for (int y = 0; y < mj; ++y)
          if (y != x && keygen_less_s(lkey[j * cap2 + y], lidx[j * cap2 + y], ax, ix, 3u, r_lo, r_hi, k0, k1)) ++less;
        if (less == tq[j]) {
          thrA[j] = ax;
          thrI[j] = ix;
          thrB[j] = keygen_word(ix, 3u, r_lo, r_hi, k0, k1);
        }
      }
      __syncthreads();
      for (int i4 = tid * 4; i4 < N; i4 += nthr * 4) {  // ct' = number of thresholds <= key
        const uint32_t member = ((c ? bitmap[i4 >> 5] : ~bitmap[i4 >> 5]) >> (i4 & 31)) & 0xfu;
        if (!member) continue;
        uint32_t w[4];
        philox4x32_10((uint32_t)i4 >> 2, 0x80000002u, r_lo, r_hi, k0, k1, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (i4 + k >= N || !((member >> k) & 1u)) continue;
          const uint32_t i = (uint32_t)(i4 + k), a = w[k], bin = a >> shift;
          int lo = 0, hi = m;
          if (!((mark[bin >> 5] >> (bin & 31)) & 1u)) {
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (tbin[mid] < bin) lo = mid + 1; else hi = mid;
            }
          } else {
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              bool thr_le;  // theta_mid <= key_i
              if (tbin[mid] != bin) thr_le = tbin[mid] < bin;
              else if (thrA[mid] != a) thr_le = thrA[mid] < a;
              else if (thrI[mid] == i) thr_le = true;
              else {
                const uint32_t b = keygen_word(i, 3u, r_lo, r_hi, k0, k1);
                thr_le = (thrB[mid] != b) ? (thrB[mid] < b) : (thrI[mid] < i);
              }
              if (thr_le) lo = mid + 1; else hi = mid;
            }
          }
          if (out) out[(size_t)q * N + i] = (uint8_t)(c * T + lo);
          if (dbg_ct) dbg_ct[(size_t)q * N + i] = (uint8_t)lo;
          if (dbg_cond) dbg_cond[(size_t)q * N + i] = (uint8_t)c;
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1u: validation mode. up_ct / up_cond are replicate-major [nb][N]; writes the same device layouts as K1.
// Replicates q >= nb (padding) get the original labels (results for them are never counted).
// grid = (ceil(N/32), Bpad/32), block = (32, 32): threadIdx.x = replicate in word, threadIdx.y = cell in tile.
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_pack_uploaded(int N, int T, int C, int nb, int Bpad, int J, const uint8_t* __restrict__ ct,
                                const uint8_t* __restrict__ cond, const uint8_t* __restrict__ up_cond,
                                const uint8_t* __restrict__ up_ct, uint32_t* __restrict__ bits1,
                                uint8_t* __restrict__ lab2) {
  __shared__ uint8_t tile_ct[32][33], tile_cd[32][33];
  const int cell0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  {  // load: threadIdx.x walks cells (coalesced), threadIdx.y walks replicates
    int cell = cell0 + threadIdx.x, q = q0 + threadIdx.y;
    uint8_t vct = 0, vcd = 0;
    if (cell < N) {
      if (q < nb) {
        vct = up_ct[(size_t)q * N + cell];
        vcd = (C == 2) ? up_cond[(size_t)q * N + cell] : 0;
      } else {
        vct = ct[cell];
        vcd = (C == 2) ? cond[cell] : 0;
      }
    }
    tile_ct[threadIdx.y][threadIdx.x] = vct;
    tile_cd[threadIdx.y][threadIdx.x] = vcd;
  }
  __syncthreads();
  int cell = cell0 + threadIdx.y, q = q0 + threadIdx.x;
  uint8_t vct = tile_ct[threadIdx.x][threadIdx.y], vcd = tile_cd[threadIdx.x][threadIdx.y];
  if (vct >= T) vct = 0;  // out-of-range codes are reported by k_check_uploaded; keep the kernels in bounds meanwhile
  if (vcd >= C) vcd = 0;
  uint32_t word = __ballot_sync(0xffffffffu, vcd == 1);
  if (cell < N) {
    if (C == 2 && threadIdx.x == 0) bits1[(size_t)cell * (Bpad / 32) + (q >> 5)] = word;
    lab2[(size_t)cell * Bpad + lab2_offset(q, J)] = (uint8_t)(vcd * T + vct);
  }
}

// Re-layout of the random-key generator's replicate-major labels in[rep][cell] (detection mode) into one reduction batch
// block out[cell][Bb] (J labels of a lane adjacent). Tile = 128 replicates x 128 cells through shared memory: 128-byte
// coalesced reads of a replicate's cells, 128-byte coalesced writes of a cell's replicates. grid = (ceil(N/128), Bb/128).
// T_bits > 0 (differential mode, labels = cond' * T + ct'): also writes the pass-1 condition bits bits1[cell][Bb / 32].
__global__ void __launch_bounds__(256) k_relayout_labels(int N, int Bb, int J, const uint8_t* __restrict__ in,
                                                         uint8_t* __restrict__ out, int T_bits, uint32_t* __restrict__ bits1) {
  __shared__ __align__(16) uint8_t tile[128][132];
  const int cell0 = blockIdx.x * 128, rep0 = blockIdx.y * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fast = (N & 3) == 0;  // rows of `in` are then 4-byte aligned
  for (int rr = warp; rr < 128; rr += 8) {
    const uint8_t* row = in + (size_t)(rep0 + rr) * N + cell0;
    const int c = lane * 4;
    if (fast && cell0 + c + 3 < N) {
      *reinterpret_cast<uint32_t*>(&tile[rr][c]) = *reinterpret_cast<const uint32_t*>(row + c);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) tile[rr][c + k] = (cell0 + c + k < N) ? row[c + k] : 0;
    }
  }
  __syncthreads();
  const int cs = 32 * J;  // replicates per chunk
  for (int item = threadIdx.x; item < 128 * 32; item += 256) {
    const int c = item >> 5, wd = item & 31;
    if (cell0 + c >= N) continue;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = wd * 4 + k, chunk = b / cs, bb = b % cs;
      const int rep = chunk * cs + (bb % J) * 32 + bb / J;
      packed |= (uint32_t)tile[rep][c] << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(out + (size_t)(cell0 + c) * Bb + rep0 + wd * 4) = packed;
  }
  if (T_bits > 0) {
    for (int item = threadIdx.x; item < 128 * 4; item += 256) {
      const int c = item >> 2, w = item & 3;
      if (cell0 + c >= N) continue;
      uint32_t word = 0;
#pragma unroll 8
      for (int b = 0; b < 32; ++b) word |= (tile[w * 32 + b][c] >= T_bits ? 1u : 0u) << b;
      bits1[(size_t)(cell0 + c) * (Bb / 32) + (rep0 >> 5) + w] = word;
    }
  }
}

// validation of uploaded labels: every replicate must keep n(cell type, condition) of both passes
// (R/utils_permutation.R:536-569 guarantees it). One block per replicate. flag[0] != 0 on violation or bad code.
__global__ void k_check_uploaded(int N, int T, int C, const uint8_t* __restrict__ ct, const uint8_t* __restrict__ up_cond,
                                 const uint8_t* __restrict__ up_ct, const uint32_t* __restrict__ expect /*[C*T]*/,
                                 int* __restrict__ flag) {
  extern __shared__ int hist[];  // [2][C*T]
  const int K = C * T, q = blockIdx.x;
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    int c = (C == 2) ? up_cond[(size_t)q * N + i] : 0;
    int t2 = up_ct[(size_t)q * N + i];
    if (c >= C || t2 >= T) { atomicMax(flag, 2); continue; }
    atomicAdd(&hist[c * T + ct[i]], 1);       // pass 1 groups: (cond', original cell type)
    atomicAdd(&hist[K + c * T + t2], 1);      // pass 2 groups: (cond', ct')
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x)
    if ((uint32_t)hist[i] != expect[i % K]) atomicMax(flag, 1);
}

// ------------------------------------------------------------------------------------------------------------------
// staging helper: a warp loads 32 consecutive (cell, value) entries coalesced and re-reads them one by one as a
// broadcast from shared memory (fp64 values: one 16-byte LDS.128 = 2 wavefronts; fp32 values: one 8-byte LDS.64 = 1
// wavefront, instead of 3 shuffles). VT = double (SCDC_FP64_EXACT) or float (SCDC_FP32_FAST).
// ------------------------------------------------------------------------------------------------------------------
template <typename VT> struct Ent;
template <> struct Ent<double> {
  typedef uint4 pack;
  static __device__ __forceinline__ pack make(int cell, double v) {
    long long b = __double_as_longlong(v);
    return make_uint4((uint32_t)cell, 0u, (uint32_t)b, (uint32_t)(b >> 32));
  }
  static __device__ __forceinline__ double value(const pack& e) {
    return __longlong_as_double((long long)(((unsigned long long)e.w << 32) | e.z));
  }
};
template <> struct Ent<float> {
  typedef uint2 pack;
  static __device__ __forceinline__ pack make(int cell, float v) { return make_uint2((uint32_t)cell, __float_as_uint(v)); }
  static __device__ __forceinline__ float value(const pack& e) { return __uint_as_float(e.y); }
};

// label bytes of one lane (k_scatter: J group codes; k_pass1_cond: the byte with its 8 condition bits)
template <int J> struct LabWord;
template <> struct LabWord<1> { typedef uint32_t type; };  // one label byte, zero-extended by the load itself
template <> struct LabWord<2> { typedef uint16_t type; };
template <> struct LabWord<4> { typedef uint32_t type; };

// J label bytes of one lane. J = 1 loads the byte straight into a 32-bit register (ld.u8 zero-extends), so that no
// masking instruction is needed before the byte is used as an accumulator row number.
template <int J>
__device__ __forceinline__ typename LabWord<J>::type load_labels(const uint8_t* p);
template <>
__device__ __forceinline__ uint32_t load_labels<1>(const uint8_t* p) {
  uint32_t r;
  asm("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
template <>
__device__ __forceinline__ uint16_t load_labels<2>(const uint8_t* p) {
  uint16_t r;
  asm("ld.global.nc.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
template <>
__device__ __forceinline__ uint32_t load_labels<4>(const uint8_t* p) {
  uint32_t r;
  asm("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Pass-1 copy of the matrix, built ON THE DEVICE (replaces a host counting sort + a second H2D of the matrix): inside
// every gene column the entries are stably sorted by the ORIGINAL cell type of their cell, so that ascending cell order
// inside every (cell type, cond') group is kept and the pass-1 sums stay bit-identical to the reference's sequential sum.
// One warp per gene column: histogram of the column's cell types (shared memory), exclusive scan -> segptr[g][0..T],
// then a second sweep in column order, 32 entries at a time: lanes with the same cell type are ranked by lane id
// (__match_any_sync), so equal keys keep their order. The copy is written as 16-byte (cell, value) records, the unit the
// pass-1 kernel fetches with one load. grid = ceil(G / W), block = 32 * W, smem = W * (T + 1) * 4.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sort_by_celltype(int G, int T, const int* __restrict__ colptr,
                                                          const int* __restrict__ rowidx, const double* __restrict__ values,
                                                          const uint8_t* __restrict__ ct, const int* __restrict__ gene_order,
                                                          int* __restrict__ segptr, uint4* __restrict__ rec_s) {
  extern __shared__ int sm_sort[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const int gi = blockIdx.x * W + warp;
  if (gi >= G) return;
  const int g = gene_order ? gene_order[gi] : gi;   // heavy columns first when an order is given
  int* cur = sm_sort + warp * (T + 1);
  for (int t = lane; t <= T; t += 32) cur[t] = 0;
  __syncwarp();
  const int beg = colptr[g], end = colptr[g + 1];
  for (int e = beg + lane; e < end; e += 32) atomicAdd(&cur[ct[rowidx[e]]], 1);
  __syncwarp();
  {  // exclusive scan of cur[0..T): lane owns ceil(T / 32) consecutive labels
    const int per = (T + 31) / 32, t0 = lane * per;
    int loc = 0;
    for (int k = 0; k < per; ++k) loc += (t0 + k < T) ? cur[t0 + k] : 0;
    int inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    int run = beg + inc - loc;
    __syncwarp();
    for (int k = 0; k < per; ++k) {
      if (t0 + k < T) {
        const int c = cur[t0 + k];
        cur[t0 + k] = run;
        segptr[(size_t)g * (T + 1) + t0 + k] = run;
        run += c;
      }
    }
    if (lane == 31) segptr[(size_t)g * (T + 1) + T] = end;
  }
  __syncwarp();
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const bool valid = e < end;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const int cell = rowidx[e];
      const double v = values[e];
      const int t = ct[cell];
      const unsigned same = __match_any_sync(act, t);
      const int rank = __popc(same & ((1u << lane) - 1u));
      const int dst = cur[t] + rank;
      __syncwarp(act);
      if (rank == 0) cur[t] += __popc(same);
      rec_s[dst] = Ent<double>::make(cell, v);
    }
    __syncwarp();
  }
}

// fp32 copies of the matrix values (SCDC_FP32_FAST): round-to-nearest of every fp64 value
__global__ void k_to_float(size_t n, const double* __restrict__ in, float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
// ... and of the pass-1 records: (cell, fp64) 16 bytes -> (cell, fp32) 8 bytes
__global__ void k_rec_to_float(size_t n, const uint4* __restrict__ in, uint2* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint4 e = in[i];
    out[i] = Ent<float>::make((int)e.x, (float)Ent<double>::value(e));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K2a: differential pass 1. Cell types are fixed, only the condition bit varies per replicate, so the entries of a
// gene column, pre-sorted by (cell type, cell) [stable: ascending cell order inside every (cell type, cond') group is
// kept], need two register accumulators per replicate and no shared-memory scatter. A warp covers 256 replicates
// (8 per lane, k_pass1_cond below).
// M1 layout: [chunk32][gene][k][32 lanes] with k = cond * T + ct, value = sum / n_k.
// ------------------------------------------------------------------------------------------------------------------
// acc1 += v if the condition bit is set, else acc0 += v, for the 8 replicates of a lane, branch-free and bit-identical to
// the branchy form: vm = bit ? v : +0.0 (two FSEL on a predicate that R2P extracts for all 8 bits at once), acc1 += vm,
// acc0 += (v - vm). v - vm is exactly 0 or v, and adding +0.0 leaves an accumulator unchanged, so each accumulator sees
// exactly the additions of its own cells, in order. 5 issue slots per replicate (3 on the FP64 pipe); the multiplier form
// of round 1 (two DFMA with a 1.0 / 0.0 selector) needed 6.7: ptxas spends a shift pair + two LOP3 + a register move on
// every selector (ncu r2e: 72.6 instructions per entry x 256 replicates, ALU pipe 55 %, FP64 pipe 32 %).
template <typename VT>
__device__ __forceinline__ void sel_add8(VT (&acc1)[8], VT (&acc0)[8], uint32_t bits, VT v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const VT vm = ((bits >> j) & 1u) ? v : (VT)0;
    acc1[j] += vm;
    acc0[j] += (v - vm);
  }
}

// Pass-1 condition bits, re-packed in place for k_pass1_cond. The label generators write, per cell, one 32-bit word per
// 32 replicates (bit = lane). The pass-1 kernel gives every lane 8 replicates (q = 256 wc + 32 j + lane, j = 0..7) and
// wants THEIR 8 bits in one byte, so that a warp fetches the bits of a cell with one 32-byte request (one sector, one cycle
// of register write-back) instead of broadcasting 16 or 32 bytes to all 32 lanes. One thread per (cell, 256-replicate group):
// 8 words in, the same 32 bytes out: byte `lane`, bit j = bit `lane` of word j. n_groups = N * B / 256 over the whole buffer.
__global__ void k_bits_transpose8(size_t n_groups, uint32_t* __restrict__ bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_groups) return;
  uint4* p = reinterpret_cast<uint4*>(bits + i * 8);
  const uint4 lo = p[0], hi = p[1];
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint32_t out[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    uint32_t o = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int lane = 4 * m + b;
#pragma unroll
      for (int j = 0; j < 8; ++j) o |= ((w[j] >> lane) & 1u) << (8 * b + j);
    }
    out[m] = o;
  }
  p[0] = make_uint4(out[0], out[1], out[2], out[3]);
  p[1] = make_uint4(out[4], out[5], out[6], out[7]);
}

// No shared memory and no warp synchronisation. A warp covers 256 replicates (8 per lane, 16 register accumulators):
// every lane fetches the entry's 16-byte (fp32: 8-byte) record itself with a warp-uniform load and the byte that holds its
// own 8 condition bits of the entry's cell (k_bits_transpose8): 5 cycles of load-store-unit write-back per entry x 256
// replicates (fp32: 3), where round 1's layout paid 8 per entry x 128. That matters because the load-store pipe is what the
// pass-2 scatter saturates, and this kernel is meant to run NEXT TO it (k_scatter holds nearly all of an SM's shared
// memory at K > 100 but few registers and issue slots; this kernel needs only registers, FP64 and issue slots).
// The loads are software-pipelined three groups deep (records of group i + 2 and bits of group i + 1 in flight while
// group i is accumulated). Padding records are (cell 0, +0.0): bit-neutral.
// bits: [cell][Bb] bytes with Bb = B / 8 (one batch block). Work item = (active gene, 256-replicate warp chunk), heavy genes
// first. work_counter == nullptr: item = global warp index, grid.x = ceil(nGa * (B / 256) / W). work_counter != nullptr
// (zeroed before the launch): PERSISTENT form — a small grid (a block or two per SM) whose warps pull items from the
// counter until none is left. Launched before the pass-2 scatter, its few blocks sit on every SM for the whole batch and
// the scatter's blocks fill the rest of the SM (its shared memory, most of its registers): the two kernels run side by side
// by construction instead of at the block scheduler's discretion. block = 32 * W.
// STAGED (block <= 4 warps): the records of the gene column go through a 64-record ring per warp in shared memory — each
// lane loads ONE record of the next 32 (coalesced, a block ahead, held in a register until the ring half is free), and the
// groups read them back as broadcasts: 2 + 0.25 write-back cycles per record instead of the uniform load's 4 (fp32: 1.125
// instead of 2). The ring spans the column, not the segment: the segments of a gene are contiguous in `rec`. 4 KB (fp32:
// 2 KB) of static shared memory per block, so the launcher uses it only where that fits next to the scatter's blocks.
template <typename VT, int U, bool STAGED>
__global__ void __launch_bounds__(128) k_pass1_cond(int G, int nGa, int T, int nWC, const int* __restrict__ segptr,
                                                    const typename Ent<VT>::pack* __restrict__ rec,
                                                    const uint8_t* __restrict__ bits, int Bb,
                                                    const double* __restrict__ ngroup, const int* __restrict__ gene_order,
                                                    VT* __restrict__ M1, int* __restrict__ work_counter) {
  typedef typename Ent<VT>::pack pack_t;
  static_assert(5 * U <= 32, "a loop iteration reads at most 32 records ahead of its base");
  __shared__ pack_t ring_s[STAGED ? 4 * 64 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  pack_t* const ring = ring_s + (STAGED ? warp * 64 : 0);
  const long long n_tasks = (long long)nGa * nWC;  // nGa active genes (gene_order), G = layout stride of M1
  const int K = 2 * T;
  const pack_t zero = Ent<VT>::make(0, (VT)0);
  for (long long task = (long long)blockIdx.x * W + warp;;) {
  if (work_counter) {
    int t = 0;
    if (lane == 0) t = atomicAdd(work_counter, 1);
    task = __shfl_sync(0xffffffffu, t, 0);
  }
  if (task >= n_tasks) return;
  const int g = gene_order[(int)(task / nWC)], wc = (int)(task % nWC);
  const int* sp = segptr + (size_t)g * (T + 1);
  const uint8_t* bitbase = bits + wc * 32 + lane;  // + cell * Bb: the byte with this lane's 8 condition bits
  asm("" : "+l"(bitbase));
  const uint32_t ubb = (uint32_t)Bb;  // the address is one IMAD.WIDE (u32 cell x u32 stride + 64-bit lane base)
  // STAGED: record e of the column [col_beg, col_end) sits in ring[(e - col_beg) & 63] while filled - 64 <= e < filled;
  // `ahead` = this lane's record of the block [filled, filled + 32), already loaded
  int col_beg = 0, col_end = 0, filled = 0;
  pack_t ahead = zero;
  if (STAGED) {
    col_beg = sp[0]; col_end = sp[T]; filled = col_beg;
    if (col_beg + lane < col_end) ahead = __ldg(rec + col_beg + lane);
  }
  auto top_up = [&](int base) {  // makes [base, base + 32) resident; overwrites only records below base
    while (filled < base + 32 && filled < col_end) {
      __syncwarp();  // every lane has read what it needed of the half that is overwritten
      ring[(filled - col_beg + lane) & 63] = ahead;
      filled += 32;
      ahead = (filled + lane < col_end) ? __ldg(rec + filled + lane) : zero;
      __syncwarp();
    }
  };
  for (int t = 0; t < T; ++t) {
    const int beg = sp[t], end = sp[t + 1];
    VT a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = 0; a1[j] = 0; }
    if (end > beg) {
      pack_t R0[U], R1[U], R2[U];
      uint32_t B0[U], B1[U], B2[U];
      auto load_recs = [&](pack_t (&R)[U], int e0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (STAGED) R[u] = (e0 + u < end) ? ring[(e0 + u - col_beg) & 63] : zero;
          else R[u] = (e0 + u < end) ? __ldg(rec + e0 + u) : zero;
        }
      };
      auto load_bits = [&](uint32_t (&B)[U], const pack_t (&R)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) B[u] = load_labels<1>(bitbase + (uint64_t)R[u].x * ubb);
      };
      auto accumulate = [&](const pack_t (&R)[U], const uint32_t (&B)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          sel_add8<VT>(a1, a0, B[u], Ent<VT>::value(R[u]));
        }
      };
      if (STAGED) top_up(beg);
      load_recs(R0, beg);
      load_bits(B0, R0);
      load_recs(R1, beg + U);
      for (int base = beg; base < end; base += 3 * U) {
        if (STAGED) top_up(base);
        load_bits(B1, R1); load_recs(R2, base + 2 * U); accumulate(R0, B0);
        load_bits(B2, R2); load_recs(R0, base + 3 * U); accumulate(R1, B1);   // past the end: zero records, bit-neutral
        load_bits(B0, R0); load_recs(R1, base + 4 * U); accumulate(R2, B2);
      }
    }
    const VT n0 = (VT)ngroup[t], n1 = (VT)ngroup[T + t];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      size_t base = (((size_t)(wc * 8 + j) * G + g) * K) * 32 + lane;
      __stcs(&M1[base + (size_t)t * 32], a0[j] / n0);
      __stcs(&M1[base + (size_t)(T + t) * 32], a1[j] / n1);
    }
  }
  if (!work_counter) return;
  }  // next work item
}

// ------------------------------------------------------------------------------------------------------------------
// K2b: general segmented sum. One warp = one gene column x (32*J) replicates; lane owns J replicates and walks the
// column in ascending cell order doing acc[label][j][lane] += x in shared memory (VT = fp64, or fp32 in SCDC_FP32_FAST).
// The accumulator row stride is a multiple of 32 banks and the column index is the lane, so the read-modify-write is
// conflict-free and needs no atomics whatever the labels are. lab2 layout: [cell][Bpad] bytes, lane's J labels adjacent
// (one J-byte load). M2 layout: [chunk32][gene][k][32 lanes], value = sum / n_k.
// grid = (G, ceil(nCJ / W)), block = 32 * W, dynamic smem = W * (K * J * 32 * sizeof(VT) + 64 * sizeof(entry)).
// Shared-memory wavefronts per (entry x 32 J replicates): fp64 2 (LDS.128 entry broadcast) + 4 J (LDS.64 + STS.64);
// fp32 1 (LDS.64 broadcast) + 2 J (LDS.32 + STS.32): half the traffic through the pipe that bounds the kernel.
// ------------------------------------------------------------------------------------------------------------------
template <typename VT, int J, int U, int GRP>
__global__ void k_scatter(int G, int K, int nCJ, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                          const VT* __restrict__ values, const uint8_t* __restrict__ lab2, int lab_stride,
                          const double* __restrict__ ngroup, const int* __restrict__ gene_order, VT* __restrict__ M2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef typename LabWord<J>::type lab_t;
  typedef typename Ent<VT>::pack pack_t;
  constexpr uint32_t ROW = 32 * sizeof(VT);  // bytes of one accumulator row (32 lanes)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const int cJ = blockIdx.y * W + warp;
  if (cJ >= nCJ) return;
  const int g = gene_order[blockIdx.x];
  pack_t* stage = reinterpret_cast<pack_t*>(smem_raw) + warp * 64;  // two buffers of 32 entries
  VT* acc = reinterpret_cast<VT*>(smem_raw + (size_t)W * 64 * sizeof(pack_t)) + (size_t)warp * K * J * 32 + lane;
  for (int i = 0; i < K * J; ++i) acc[i * 32] = 0;
  const int beg = colptr[g], end = colptr[g + 1];
  const int nst = (end - beg + 31) >> 5;
  const uint8_t* labbase = lab2 + (size_t)cJ * 32 * J + lane * J;
  asm("" : "+l"(labbase));  // one opaque 64-bit value: keeps ptxas from re-adding the uniform base on every label load
  const uint32_t ustride = (uint32_t)lab_stride;  // label address = 64-bit lane base + u32 cell * u32 stride: one IMAD.WIDE
  char* const accb = reinterpret_cast<char*>(acc);  // J = 1: accumulator row of label l starts at byte l * ROW
  // Matrix entries and the flushed means are streamed (evict-first) so that they do not push the labels out of L2.
  // Software pipeline: U (value, label) slots live in registers. An entry is accumulated from its slot and the slot is
  // refilled AT ONCE with the entry U positions further down the column (value from the staged broadcast, label from
  // global memory / L2), so every label load has the accumulation of U - 1 other entries to complete in, with one set
  // of U registers (round 1 kept a current and a next group: half the distance for the same registers; ncu r2a/r2e showed
  // long-scoreboard stalls of 2.1 - 5.3 per issue with few resident warps). The entries two stages ahead are in flight
  // from global memory: the read-modify-write chain never waits on a load.
  int cell_n = 0;
  VT v_n = 0;
  if (beg + lane < end) { cell_n = __ldcs(rowidx + beg + lane); v_n = __ldcs(values + beg + lane); }
  stage[lane] = Ent<VT>::make(cell_n, v_n);
  cell_n = 0; v_n = 0;  // padding entries add +0.0 to cell 0's group: bit-neutral
  if (beg + 32 + lane < end) { cell_n = __ldcs(rowidx + beg + 32 + lane); v_n = __ldcs(values + beg + 32 + lane); }
  __syncwarp();
  VT val[U];
  lab_t lb[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const pack_t en = stage[u];
    val[u] = Ent<VT>::value(en);
    lb[u] = load_labels<J>(labbase + (uint64_t)en.x * ustride);
  }
  for (int s = 0; s < nst; ++s) {
    pack_t* cur_buf = stage + (s & 1) * 32;
    pack_t* nxt_buf = stage + ((s + 1) & 1) * 32;
    __syncwarp();  // every lane is done reading stage s-1 from nxt_buf
    nxt_buf[lane] = Ent<VT>::make(cell_n, v_n);
    {
      const int e = beg + (s + 2) * 32 + lane;
      cell_n = 0; v_n = 0;
      if (e < end) { cell_n = __ldcs(rowidx + e); v_n = __ldcs(values + e); }
    }
    __syncwarp();
#pragma unroll
    for (int gq = 0; gq < 32 / U; ++gq) {
      // the slots hold entries [gq * U, gq * U + U) of stage s; their successors sit U entries further: in this stage's
      // buffer, or (last group of the stage) at the head of the next stage's
      const pack_t* src = (gq + 1 < 32 / U) ? (cur_buf + (gq + 1) * U) : nxt_buf;
      if (GRP == 4 && J == 1) {
        // Four entries per step (few resident warps: more independent shared-memory loads in flight per lane). All four
        // accumulators are loaded first; an entry whose label repeats an earlier one of the group continues from that
        // entry's result, and the stores go out in entry order, so the value stored last for a label is the sequential sum.
#pragma unroll
        for (int u = 0; u < U; u += 4) {
          VT* const p1 = reinterpret_cast<VT*>(accb + lb[u] * ROW);
          VT* const p2 = reinterpret_cast<VT*>(accb + lb[u + 1] * ROW);
          VT* const p3 = reinterpret_cast<VT*>(accb + lb[u + 2] * ROW);
          VT* const p4 = reinterpret_cast<VT*>(accb + lb[u + 3] * ROW);
          const VT a1 = *p1, a2 = *p2, a3 = *p3, a4 = *p4;
          const VT r1 = a1 + val[u];
          const VT r2 = ((lb[u + 1] == lb[u]) ? r1 : a2) + val[u + 1];
          const VT r3 = ((lb[u + 2] == lb[u + 1]) ? r2 : ((lb[u + 2] == lb[u]) ? r1 : a3)) + val[u + 2];
          const VT r4 = ((lb[u + 3] == lb[u + 2]) ? r3 : ((lb[u + 3] == lb[u + 1]) ? r2 : ((lb[u + 3] == lb[u]) ? r1 : a4))) + val[u + 3];
          *p1 = r1; *p2 = r2; *p3 = r3; *p4 = r4;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const pack_t en = src[u + w];
            val[u + w] = Ent<VT>::value(en);
            lb[u + w] = load_labels<J>(labbase + (uint64_t)en.x * ustride);
          }
        }
      } else if (GRP == 2 && J == 1) {
        // Two entries per step: both accumulators are loaded before either add. If the two labels coincide the second
        // add continues from the first result, so the value stored last is ((acc + v1) + v2): the sequential sum, with a
        // dependency chain of one shared-memory round trip per PAIR (matters when few warps are resident).
#pragma unroll
        for (int u = 0; u < U; u += 2) {
          VT* const p1 = reinterpret_cast<VT*>(accb + lb[u] * ROW);
          VT* const p2 = reinterpret_cast<VT*>(accb + lb[u + 1] * ROW);
          const VT a1 = *p1, a2 = *p2;
          const VT r1 = a1 + val[u];
          const VT r2 = ((lb[u] == lb[u + 1]) ? r1 : a2) + val[u + 1];
          *p1 = r1;  // stores in entry order: for a repeated label the second store carries the sequential sum
          *p2 = r2;
          const pack_t e1 = src[u], e2 = src[u + 1];
          val[u] = Ent<VT>::value(e1);
          lb[u] = load_labels<J>(labbase + (uint64_t)e1.x * ustride);
          val[u + 1] = Ent<VT>::value(e2);
          lb[u + 1] = load_labels<J>(labbase + (uint64_t)e2.x * ustride);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (J == 1) {
            VT* const p1 = reinterpret_cast<VT*>(accb + lb[u] * ROW);
            *p1 = *p1 + val[u];
          } else {
            // row (label, j) starts at byte (label * J + j) * ROW: one byte extract (PRMT) + one multiply-add per row,
            // the j * ROW part folds into the LDS / STS immediate
            VT* row[J];
            VT cur[J];
#pragma unroll
            for (int j = 0; j < J; ++j)
              row[j] = reinterpret_cast<VT*>(accb + __byte_perm((uint32_t)lb[u], 0u, 0x4440u + j) * (J * ROW) + j * ROW);
#pragma unroll
            for (int j = 0; j < J; ++j) cur[j] = *row[j];
#pragma unroll
            for (int j = 0; j < J; ++j) *row[j] = cur[j] + val[u];
          }
          const pack_t en = src[u];
          val[u] = Ent<VT>::value(en);
          lb[u] = load_labels<J>(labbase + (uint64_t)en.x * ustride);
        }
      }
    }
  }
  __syncwarp();
  for (int k = 0; k < K; ++k) {
    const VT nk = (VT)ngroup[k];
#pragma unroll
    for (int j = 0; j < J; ++j)
      __stcs(&M2[(((size_t)(cJ * J + j) * G + g) * K + k) * 32 + lane], acc[(k * J + j) * 32] / nk);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K3 (v1, generic): one warp = one tested row, lane = replicate, loop over the batch's 32-replicate chunks.
// Reads the group means coalesced ([chunk][gene][k][lane]), forms the reference's per-iteration columns
//   differential: [cond2-cond1 score (pass 1), cond1 score (pass 2), cond2 score (pass 2), L_i diffs, R_j diffs (pass 1)]
//   detection   : [score]
// compares with the observed values (`>=`; two-sided abs() for DE and subunit columns) and adds ballots to the row's
// counters. A row is owned by exactly one warp per launch, so the counter update is a plain add.
// ------------------------------------------------------------------------------------------------------------------
struct ScoreArgs {
  int R, G, T, C, nL, nR, ncols, score_type;
  int nChunks;      // 32-replicate chunks in this batch
  int n_valid;      // replicates of this batch that count
  const int* sub_gene;   // [L][nL+nR]
  const int* row_lri;
  const int* row_emit;
  const int* row_recv;
  const double* obs;     // column-major [R][ncols]
  const void* M1;        // pass-1 means (differential) or nullptr; element type MT
  const void* M2;        // pass-2 means / detection means
  uint32_t* counts;      // column-major [R][ncols]
  uint32_t* band;        // BAND: comparisons inside the tolerance band, same shape
  double rel_tol;        // BAND
  double* dist;          // nullptr or [n_valid][ncols_d][R] (this batch's slice)
};

__device__ __forceinline__ double score_of(double a, double b, int score_type) {
  return score_type == 0 ? sqrt(a * b) : (a + b) / 2;
}

// ------------------------------------------------------------------------------------------------------------------
// Tolerance band of SCDC_FP32_FAST (include/scdc.h, scdc_options.rel_tol). The same fp64 expressions are evaluated by
// every score kernel, so band membership does not depend on which kernel counted a comparison.
//   one-sided geometric score, compared as a product x against observed o:  |x - o^2| <= (2 rel_tol) o^2
//   one-sided arithmetic score s:                                           |s - o|   <= rel_tol |o|
//   two-sided difference a - b of two non-negative magnitudes:              ||a - b| - |o|| <= rel_tol (a + b)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool band_product(double x, double o2, double rel_tol) { return fabs(x - o2) <= (2.0 * rel_tol) * o2; }
__device__ __forceinline__ bool band_value(double s, double o, double rel_tol) { return fabs(s - o) <= rel_tol * fabs(o); }
__device__ __forceinline__ bool band_diff(double a, double b, double abs_o, double rel_tol) {
  return fabs(fabs(a - b) - abs_o) <= rel_tol * (a + b);
}

template <typename MT, bool BAND>
__global__ void __launch_bounds__(256) k_score_count(ScoreArgs A) {
  const MT* const M1 = reinterpret_cast<const MT*>(A.M1);
  const MT* const M2 = reinterpret_cast<const MT*>(A.M2);
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int K = A.C * A.T, nS = A.nL + A.nR;
  const int ncols_d = (A.C == 2) ? 3 : 1;
  const double rt = A.rel_tol;
  for (int row = warp_global; row < A.R; row += n_warps) {
    const int l = A.row_lri[row], e = A.row_emit[row], r = A.row_recv[row];
    int gs[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) gs[s] = (s < nS) ? A.sub_gene[l * nS + s] : -1;
    double ob[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) ob[c] = (c < A.ncols) ? A.obs[(size_t)c * A.R + row] : 0.0;
    uint32_t cnt[16], bnd[BAND ? 16 : 1];
#pragma unroll
    for (int c = 0; c < 16; ++c) cnt[c] = 0;
#pragma unroll
    for (int c = 0; c < (BAND ? 16 : 1); ++c) bnd[c] = 0;
    for (int ch = 0; ch < A.nChunks; ++ch) {
      const bool valid = (ch * 32 + lane) < A.n_valid;
      const size_t cb = (size_t)ch * A.G;
      if (A.C == 1) {
        double mL = 0.0, mR = 0.0;
        bool hL = false, hR = false;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < nS && gs[s] >= 0) {
            const bool isL = s < A.nL;
            double v = (double)M2[((cb + gs[s]) * K + (isL ? e : r)) * 32 + lane];
            if (isL) { mL = hL ? fmin(mL, v) : v; hL = true; } else { mR = hR ? fmin(mR, v) : v; hR = true; }
          }
        }
        double sc = score_of(mL, mR, A.score_type);
        uint32_t b = __ballot_sync(0xffffffffu, valid && (sc >= ob[0]));
        cnt[0] += __popc(b);
        if (BAND) {
          const bool in = A.score_type == 0 ? band_product(mL * mR, ob[0] * ob[0], rt) : band_value(sc, ob[0], rt);
          bnd[0] += __popc(__ballot_sync(0xffffffffu, valid && in));
        }
        if (A.dist && valid) A.dist[((size_t)(ch * 32 + lane) * ncols_d) * A.R + row] = sc;
      } else {
        double s1[2], s2[2], x2[2];
        double dsub[8], ssub[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) { dsub[s] = 0.0; ssub[s] = 0.0; }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double mL = 0.0, mR = 0.0, qL = 0.0, qR = 0.0;
          bool hL = false, hR = false;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < nS && gs[s] >= 0) {
              const bool isL = s < A.nL;
              const size_t idx = ((cb + gs[s]) * K + c * A.T + (isL ? e : r)) * 32 + lane;
              double v1 = (double)M1[idx], v2 = (double)M2[idx];
              ssub[s] = (c == 0) ? v1 : (v1 + dsub[s]);    // cond2 + cond1: the scale of the band (dsub still holds cond1)
              dsub[s] = (c == 0) ? v1 : (v1 - dsub[s]);  // cond2 - cond1 (R/utils_cci.R:212-239)
              if (isL) {
                mL = hL ? fmin(mL, v1) : v1; qL = hL ? fmin(qL, v2) : v2; hL = true;
              } else {
                mR = hR ? fmin(mR, v1) : v1; qR = hR ? fmin(qR, v2) : v2; hR = true;
              }
            }
          }
          s1[c] = score_of(mL, mR, A.score_type);
          s2[c] = score_of(qL, qR, A.score_type);
          x2[c] = qL * qR;
        }
        const double de = s1[1] - s1[0];
        cnt[0] += __popc(__ballot_sync(0xffffffffu, valid && (fabs(de) >= fabs(ob[0]))));
        cnt[1] += __popc(__ballot_sync(0xffffffffu, valid && (s2[0] >= ob[1])));
        cnt[2] += __popc(__ballot_sync(0xffffffffu, valid && (s2[1] >= ob[2])));
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s < nS) cnt[3 + s] += __popc(__ballot_sync(0xffffffffu, valid && gs[s] >= 0 && (fabs(dsub[s]) >= fabs(ob[3 + s]))));
        }
        if (BAND) {
          bnd[0] += __popc(__ballot_sync(0xffffffffu, valid && band_diff(s1[1], s1[0], fabs(ob[0]), rt)));
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const bool in = A.score_type == 0 ? band_product(x2[c], ob[1 + c] * ob[1 + c], rt) : band_value(s2[c], ob[1 + c], rt);
            bnd[1 + c] += __popc(__ballot_sync(0xffffffffu, valid && in));
          }
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < nS) {
              // band_diff(cond2, cond1) with the difference and the sum already formed
              const bool in = fabs(fabs(dsub[s]) - fabs(ob[3 + s])) <= rt * ssub[s];
              bnd[3 + s] += __popc(__ballot_sync(0xffffffffu, valid && gs[s] >= 0 && in));
            }
          }
        }
        if (A.dist && valid) {
          double* d = A.dist + ((size_t)(ch * 32 + lane) * 3) * A.R + row;
          d[0] = de;
          d[(size_t)A.R] = s2[0];
          d[(size_t)2 * A.R] = s2[1];
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < A.ncols) {
          A.counts[(size_t)c * A.R + row] += cnt[c];
          if (BAND) A.band[(size_t)c * A.R + row] += bnd[BAND ? c : 0];
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K3 (production path): factorised per LRI. A block owns one LRI and a slice of its tested rows (thread = row, RPT rows
// per thread; observed values, thresholds and counters in registers for the whole launch) and walks the batch P replicates
// at a time:
//   tiles   : min over the LRI's present ligand subunits per cell type and the same for the receptor side
//             (pmin(na.rm = TRUE), reference R/utils_cci.R:613-653) in shared memory as [replicate][cell type] tiles, filled
//             by cp.async (LDGSTS) ONE TILE AHEAD of their use (double-buffered), so the global-memory latency hides behind
//             the row work of the previous tile; a fix-up pass folds the extra subunits of complexes in (rare);
//   rows    : every row combines its emitter / receiver entries: score, `>=` compare, register counter.
// No ballots, no atomics, no re-reading of the means per row. The one-sided geometric columns compare the PRODUCT with
// thr = min{y : sqrt_rn(y) >= observed} (k_product_thresholds); sqrt_rn is monotone, so
// (sqrt(a*b) >= observed) == (a*b >= thr) exactly and the fp64 sqrt disappears.
// k_score_det: detection mode (one column). k_score_diff (below): differential mode (3 + nL + nR columns).
// ------------------------------------------------------------------------------------------------------------------
template <typename MT>
__device__ __forceinline__ void cp_async_elem(MT* smem_dst, const MT* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(MT) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ScoreDetArgs {
  int G, T, Tp, nL, nS, score_type;
  int n_valid, P;
  int R;
  const int* task_lri;
  const int* task_beg;
  const int* task_end;
  const int* srow;      // [R] original row id of each LRI-sorted position
  const int* sub_gene;
  const int* row_emit;
  const int* row_recv;
  const double* obs;    // [R] observed CCI_SCORE, original row ids
  const double* thr;    // [R] product thresholds (geometric score)
  const void* M2;       // group means [chunk32][gene][T][32], element type MT
  uint32_t* counts;     // [R]
  uint32_t* band;       // BAND: [R]
  double rel_tol;       // BAND
};

// shared memory: RAW[2 buffers][2 sides][P][Tp] elements of MT (cp.async target, read by the rows; fp32 means stay fp32 in
// shared memory — half the bytes through the pipe that bounds this kernel — and are widened in registers).
template <int RPT, typename MT, bool BAND>
__global__ void __launch_bounds__(256, 2) k_score_det(ScoreDetArgs A) {
  extern __shared__ __align__(16) double sm_d[];
  const MT* const M2 = reinterpret_cast<const MT*>(A.M2);
  __shared__ int s_gs[8];
  __shared__ int s_extra;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int l = A.task_lri[blockIdx.x], rb = A.task_beg[blockIdx.x], re = A.task_end[blockIdx.x];
  const int T = A.T, Tp = A.Tp, P = A.P, nL = A.nL, nS = A.nS;
  const double rt = A.rel_tol;
  if (tid < 8) s_gs[tid] = (tid < nS) ? A.sub_gene[l * nS + tid] : -1;
  if (tid == 0) {
    int extra = 0;
    for (int s = 0; s < nS; ++s)
      if (s != 0 && s != nL && A.sub_gene[l * nS + s] >= 0) extra = 1;
    s_extra = extra;
  }
  int e_[RPT], r_[RPT], rid[RPT];
  double ob[RPT], th[RPT];
  uint32_t cnt[RPT], bnd[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int pos = rb + tid + q * nthr;
    rid[q] = (pos < re) ? A.srow[pos] : -1;
    e_[q] = r_[q] = 0;
    ob[q] = th[q] = 0.0;
    cnt[q] = bnd[q] = 0;
    if (rid[q] >= 0) {
      e_[q] = A.row_emit[rid[q]];
      r_[q] = A.row_recv[rid[q]];
      ob[q] = A.obs[rid[q]];
      th[q] = A.thr[rid[q]];
    }
  }
  const int tile = P * Tp;  // elements per (buffer, side)
  MT* const rawbase = reinterpret_cast<MT*>(sm_d);
  const int n_tiles = (A.n_valid + P - 1) / P;
  __syncthreads();
  const int g_first[2] = {s_gs[0], s_gs[nL]};  // LIGAND_1 / RECEPTOR_1: always present
  const bool has_extra = s_extra != 0;
  // thread -> (replicate p_t, cell type ct0 + k * ctstep) without integer divisions (P is a power of two)
  const int lgP = 31 - __clz(P);
  const int p_t = tid & (P - 1), ct0 = tid >> lgP, ctstep = nthr >> lgP;
  auto issue = [&](int t) {
    const int pb = t * P, ch = pb >> 5, lane0 = pb & 31;
    MT* dst0 = rawbase + (size_t)(t & 1) * 2 * tile;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const MT* src = M2 + (((size_t)ch * A.G + g_first[side]) * T) * 32 + lane0 + p_t;
      MT* dst = dst0 + side * tile + p_t * Tp;
      for (int ct = ct0; ct < T; ct += ctstep) cp_async_elem<MT>(dst + ct, src + (size_t)ct * 32);
    }
    cp_async_commit();
  };
  if (n_tiles > 0) issue(0);
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) { issue(t + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();  // tile t is in shared memory for every thread
    MT* raw = rawbase + (size_t)(t & 1) * 2 * tile;
    const int pb = t * P, ch = pb >> 5, lane0 = pb & 31;
    if (has_extra) {  // complexes: fold the extra subunits in (min of fp32 values is exact in fp32)
      for (int side = 0; side < 2; ++side) {
        const int s0 = side ? nL : 0, s1 = side ? nS : nL;
        for (int ct = ct0; ct < T; ct += ctstep) {
          MT m = raw[side * tile + p_t * Tp + ct];
          for (int s = s0 + 1; s < s1; ++s) {
            const int g = s_gs[s];
            if (g >= 0) m = (MT)fmin((double)m, (double)M2[(((size_t)ch * A.G + g) * T + ct) * 32 + lane0 + p_t]);
          }
          raw[side * tile + p_t * Tp + ct] = m;
        }
      }
      __syncthreads();
    }
    const int pmax = min(P, A.n_valid - pb);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (rid[q] < 0) continue;
      const MT* pa = raw + e_[q];
      const MT* pbp = raw + tile + r_[q];
      uint32_t c = 0, bc = 0;
      if (A.score_type == 0) {
        const double o2 = ob[q] * ob[q];
#pragma unroll 4
        for (int p = 0; p < pmax; ++p) {
          const double x = (double)pa[p * Tp] * (double)pbp[p * Tp];
          c += (x >= th[q]) ? 1u : 0u;
          if (BAND) bc += band_product(x, o2, rt) ? 1u : 0u;
        }
      } else {
#pragma unroll 4
        for (int p = 0; p < pmax; ++p) {
          const double sc = ((double)pa[p * Tp] + (double)pbp[p * Tp]) / 2;
          c += (sc >= ob[q]) ? 1u : 0u;
          if (BAND) bc += band_value(sc, ob[q], rt) ? 1u : 0u;
        }
      }
      cnt[q] += c;
      bnd[q] += bc;
    }
    __syncthreads();  // everyone is done with buffer t & 1 before it is overwritten
  }
#pragma unroll
  for (int q = 0; q < RPT; ++q)
    if (rid[q] >= 0) {
      A.counts[rid[q]] += cnt[q];
      if (BAND) A.band[rid[q]] += bnd[q];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K3, differential mode (production): same block / tile structure as k_score_det, for the 3 + nL + nR columns.
//   * fp64 tiles are four planes [pass][side][replicate][cell type][cond]: a row reads its emitter's ligand minima and its
//     receiver's receptor minima of one pass with ONE 16-byte load each, and the rows of a warp (consecutive receivers)
//     read consecutive 16-byte chunks: conflict-free. (With the four values of a cell type interleaved in 32 bytes, as in
//     round 1, consecutive cell types were 32 bytes apart and every such load was a 2-way bank conflict: ncu counted
//     17.4 G conflict wavefronts out of 35.0 G on config 5.) fp32 tiles keep [side][replicate][cell type][pass*2 + cond]:
//     16 bytes per cell type, one load per side, widened to fp64 in registers;
//   * FACT: the per-subunit columns |L_i / R_j cond2 - cond1| >= |observed| depend only on (LRI, emitter) resp.
//     (LRI, receiver). When the observed subunit differences are bitwise equal across the rows that share that pair
//     (always true for the reference's tables, checked on the host), they are counted once per (cell type, subunit) and
//     added to the rows at the end instead of once per row. The comparison happens in the fix-up pass, where the thread
//     that owns (replicate, cell type) has the subunit's two pass-1 means in registers: the P lanes that share a cell type
//     vote with one ballot and their first lane adds the count to a shared-memory counter that only it ever touches;
//   * the DE column |sqrt(x2) - sqrt(x1)| >= |observed| is decided from rsqrt.approx when the approximation's error band
//     cannot change the outcome (the band is 8x the documented 2^-20 relative error) and recomputed with the IEEE sqrt
//     otherwise, so the decision is always the exact one.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_approx(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// exact value of (|sqrt(x1) - sqrt(x0)| >= aob) for aob >= 0, for ANY fp64 inputs, decided without an IEEE sqrt whenever the
// approximation's error cannot change the outcome.
//   s = x * rsqrt.approx(x + 1e-300): the added constant keeps the approximation finite at x = 0 (s = 0 = sqrt(0) exactly)
//   and out of the flushed subnormal range; it perturbs s by less than 2^-60 relative for x >= 1e-260, and for smaller x
//   both s and sqrt(x) are below 1e-130, so the error is below 1e-130 ABSOLUTE. rsqrt.approx.f64 is accurate to 2^-22 on
//   normal inputs, so |s - sqrt(x)| <= 2^-20 sqrt(x) + 1e-130.
//   slack = (s1 + s0) * kslack + 1e-120, kslack >= 2^-17: 8x the relative and 1e10x the absolute bound.
//   Negative inputs, NaN and +inf make d NaN, so neither test passes and the IEEE path decides (as the reference would).
// kslack = 2^-17, plus 1.5 rel_tol with BAND, which widens the exactly-evaluated zone over the whole tolerance band: there
// bc also receives the exact band_diff(sqrt(x1), sqrt(x0)) decision, and outside of it a comparison is provably not in the
// band. (Round 1 guarded the approximation with fmax(x, 1e-300) and an exponent-range test per operand: ~55 instructions
// for this column alone; this form is 13 FP64 operations and two MUFU.)
constexpr double kDeSlack = 7.62939453125e-06;  // 2^-17
template <bool BAND>
__device__ __forceinline__ bool de_exceeds(double x1, double x0, double aob, double kslack, double rt, uint32_t& bc) {
  const double s1 = x1 * rsqrt_approx(x1 + 1e-300);
  const double s0 = x0 * rsqrt_approx(x0 + 1e-300);
  const double d = fabs(s1 - s0), slack = fma(s1 + s0, kslack, 1e-120);
  const bool yes = d - slack > aob, no = d + slack < aob;
  if (yes || no) return yes;
  const double e1 = sqrt(x1), e0 = sqrt(x0);
  if (BAND) bc += band_diff(e1, e0, aob, rt) ? 1u : 0u;
  return fabs(e1 - e0) >= aob;
}

// Self-test of de_exceeds against the plain IEEE expression on adversarial inputs (thresholds within a few ulps of the
// true difference, exact ties, zeros, tiny / huge products). out[0] = mismatches, out[1] = decisions taken by the fast path.
__global__ void k_selftest_de(unsigned long long n, unsigned long long seed, unsigned long long* out) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t w[4];
  philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x5eedu, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const int kind = w[3] & 15;
  // log-uniform magnitudes; kind selects the exponent range
  const double span = (kind < 9) ? 12.0 : (kind < 13 ? 300.0 : 40.0);
  double x0 = exp2((w[0] * (1.0 / 4294967296.0) - 0.5) * span), x1 = exp2((w[1] * (1.0 / 4294967296.0) - 0.5) * span);
  if (kind == 9) {   // tiny and subnormal products, far tails: 2^-1074 .. 2^-850 and 2^850 .. 2^1023
    x0 = exp2(-1074.0 + (w[0] * (1.0 / 4294967296.0)) * 224.0);
    x1 = (w[3] & 16) ? exp2(850.0 + (w[1] * (1.0 / 4294967296.0)) * 173.0) : exp2(-1074.0 + (w[1] * (1.0 / 4294967296.0)) * 224.0);
  }
  if (kind == 10) x0 = 0.0;
  if (kind == 11) x1 = x0;
  if (kind == 12) { x0 = 0.0; x1 = 0.0; }
  const double truth = fabs(sqrt(x1) - sqrt(x0));
  // threshold: the exact difference moved by -3..+3 ulps, or scaled by 1 +- 2^-k, or unrelated
  double aob = truth;
  const int mode = (w[2] >> 4) & 7, amt = (int)(w[2] & 7) - 3;
  if (mode < 3) aob = __longlong_as_double(__double_as_longlong(truth) + amt);
  else if (mode < 6) aob = truth * (1.0 + (amt >= 0 ? 1.0 : -1.0) * exp2(-(double)(10 + 6 * (w[2] >> 8 & 7))));
  else aob = exp2((w[2] * (1.0 / 4294967296.0) - 0.5) * span * 0.5);
  if (!(aob >= 0.0)) aob = 0.0;
  const bool want = truth >= aob;
  uint32_t bc = 0, bc2 = 0;
  const bool got = de_exceeds<false>(x1, x0, aob, kDeSlack, 0.0, bc);
  if (want != got) atomicAdd(&out[0], 1ull);
  // BAND variant (rel_tol = 1e-5): same decision, and the band flag equals the exact band test
  const double rt = 1e-5;
  const bool got2 = de_exceeds<true>(x1, x0, aob, kDeSlack + 1.5 * rt, rt, bc2);
  const bool want_band = band_diff(sqrt(x1), sqrt(x0), aob, rt);
  if (want != got2 || (bc2 != 0) != want_band) atomicAdd(&out[0], 1ull);
  // was it decidable without the IEEE sqrt?
  const double s1 = (x1 == 0.0) ? 0.0 : x1 * rsqrt_approx(x1), s0 = (x0 == 0.0) ? 0.0 : x0 * rsqrt_approx(x0);
  const double d = fabs(s1 - s0), slack = (s1 + s0) * 7.62939453125e-06 + 1e-120;
  if (d - slack > aob || d + slack < aob) atomicAdd(&out[1], 1ull);
  // the approximation itself must stay well inside the band used by the filter
  if (x1 > 1e-280 && x1 < 1e280 && fabs(s1 - sqrt(x1)) > sqrt(x1) * 9.5367431640625e-07) atomicAdd(&out[2], 1ull);  // 2^-20
}

struct ScoreDiffArgs {
  int G, T, Tp, K, nL, nR, score_type;
  int n_valid, P;
  int R;
  const int* task_lri;
  const int* task_beg;
  const int* task_end;
  const int* srow;
  const int* sub_gene;
  const int* row_emit;
  const int* row_recv;
  const double* obs;      // column-major [R][ncols], original row ids
  const double* thr;      // column-major [R][2]
  const double* obs_sub;  // FACT: [L][T][nS] common observed subunit differences (NaN = unused / absent)
  const void* M1;         // element type MT
  const void* M2;
  uint32_t* counts;
  uint32_t* band;         // BAND
  double rel_tol;         // BAND
};

// shared memory: RAW[2 buffers] of MT, each fp64: [pass][side][P][Tp][cond], fp32: [side][P][Tp][pass*2 + cond]; then (doubles)
// FACT: SO[T][nS] |observed subunit differences| + CNT[T][nS] (+ BND[T][nS]) uint32 counters; !FACT: DL[P][Tp][nL],
// DR[P][Tp][nR] subunit differences (+ WL, WR: cond2 + cond1, the band's scale).
template <int RPT, bool FACT, typename MT, bool BAND>
__global__ void __launch_bounds__(256, 2) k_score_diff(ScoreDiffArgs A) {
  extern __shared__ __align__(16) double sm_d[];
  constexpr bool F32 = sizeof(MT) == 4;
  const MT* const M1 = reinterpret_cast<const MT*>(A.M1);
  const MT* const M2 = reinterpret_cast<const MT*>(A.M2);
  __shared__ int s_gs[8];
  __shared__ int s_extra;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  const int l = A.task_lri[blockIdx.x], rb = A.task_beg[blockIdx.x], re = A.task_end[blockIdx.x];
  const int T = A.T, Tp = A.Tp, P = A.P, K = A.K, nL = A.nL, nR = A.nR, nS = nL + nR;
  const int ncols = 3 + nS;
  const double rt = A.rel_tol;
  const double kslack = kDeSlack + (BAND ? 1.5 * rt : 0.0);
  if (tid < 8) s_gs[tid] = (tid < nS) ? A.sub_gene[l * nS + tid] : -1;
  if (tid == 0) {
    int extra = 0;
    for (int s = 0; s < nS; ++s)
      if (s != 0 && s != nL && A.sub_gene[l * nS + s] >= 0) extra = 1;
    s_extra = extra;
  }
  constexpr int NOB = FACT ? 3 : 11;
  int e_[RPT], r_[RPT], rid[RPT];
  double ob[RPT][NOB], th[RPT][2];
  uint32_t cnt[RPT][NOB], bnd[RPT][BAND ? NOB : 1];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int pos = rb + tid + q * nthr;
    rid[q] = (pos < re) ? A.srow[pos] : -1;
    e_[q] = r_[q] = 0;
#pragma unroll
    for (int c = 0; c < NOB; ++c) { ob[q][c] = 0.0; cnt[q][c] = 0; }
#pragma unroll
    for (int c = 0; c < (BAND ? NOB : 1); ++c) bnd[q][c] = 0;
    th[q][0] = th[q][1] = 0.0;
    if (rid[q] >= 0) {
      e_[q] = A.row_emit[rid[q]];
      r_[q] = A.row_recv[rid[q]];
#pragma unroll
      for (int c = 0; c < NOB; ++c)
        if (c < ncols) ob[q][c] = A.obs[(size_t)c * A.R + rid[q]];
      th[q][0] = A.thr[rid[q]];
      th[q][1] = A.thr[(size_t)A.R + rid[q]];
    }
  }
  const int plane = P * Tp * 2;                 // fp64: elements of one [pass][side] plane
  const int sideStride = P * Tp * 4;            // fp32: elements per side
  const int rawSize = 2 * sideStride;           // elements per RAW buffer (both layouts)
  MT* const rawbase = reinterpret_cast<MT*>(sm_d);
  // element (side, p, ct, k = pass * 2 + cond) of a RAW buffer
  auto at = [&](int side, int p, int ct, int k) -> int {
    return F32 ? side * sideStride + (p * Tp + ct) * 4 + k : ((k >> 1) * 2 + side) * plane + (p * Tp + ct) * 2 + (k & 1);
  };
  double* const tail = sm_d + (F32 ? rawSize : 2 * rawSize);  // after the two RAW buffers
  // FACT tail: SO, CNT (BND); !FACT tail: DL, DR (WL, WR)
  double* const SO = tail;
  uint32_t* const CNT = reinterpret_cast<uint32_t*>(tail + T * nS);
  uint32_t* const BND = CNT + T * nS;
  double* const DL = tail;                      // [P][Tp][nL]
  double* const DR = DL + P * Tp * nL;          // [P][Tp][nR]
  double* const WL = DR + P * Tp * nR;
  double* const WR = WL + P * Tp * nL;
  if (FACT) {
    for (int i = tid; i < T * nS; i += nthr) {
      SO[i] = fabs(A.obs_sub[(size_t)l * T * nS + i]);  // NaN (absent / unused): never exceeded, never in the band
      CNT[i] = 0;
      if (BAND) BND[i] = 0;
    }
  }
  const int n_tiles = (A.n_valid + P - 1) / P;
  __syncthreads();
  const int g_first[2] = {s_gs[0], s_gs[nL]};
  const bool has_extra = s_extra != 0;
  // thread -> (replicate p, cell type ct0 + k * ctstep) without integer divisions (P is a power of two)
  const int lgP = 31 - __clz(P);
  const int p_t = tid & (P - 1), ct0 = tid >> lgP, ctstep = nthr >> lgP;
  const int n_ct_iter = (T + ctstep - 1) / ctstep;   // uniform trip count: the fix-up pass votes with full-warp ballots
  const uint32_t gmask = ((P >= 32) ? 0xffffffffu : ((1u << P) - 1u)) << (lane & ~(P - 1) & 31);
  const bool gleader = (lane & (P - 1) & 31) == 0;
  auto issue = [&](int t) {
    const int pb = t * P, ch = pb >> 5, lane0 = pb & 31;
    MT* dst0 = rawbase + (size_t)(t & 1) * rawSize;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const MT* src = ((k >> 1) ? M2 : M1) + (((size_t)ch * A.G + g_first[side]) * K + (k & 1) * T) * 32 + lane0 + p_t;
        MT* dst = dst0 + at(side, p_t, 0, k);
        const int cstep = F32 ? 4 : 2;
        for (int ct = ct0; ct < T; ct += ctstep) cp_async_elem<MT>(dst + ct * cstep, src + (size_t)ct * 32);
      }
    }
    cp_async_commit();
  };
  if (n_tiles > 0) issue(0);
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) { issue(t + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    MT* raw = rawbase + (size_t)(t & 1) * rawSize;
    const int pb = t * P, ch = pb >> 5, lane0 = pb & 31;
    const int pmax = min(P, A.n_valid - pb);
    const bool pvalid = p_t < pmax;
    // -------- fix-up: subunit differences (counted here when FACT), extra subunits of complexes folded in --------
    if (FACT || has_extra) {
      for (int side = 0; side < 2; ++side) {
        const int s0 = side ? nL : 0, nSide = side ? nR : nL;
        for (int it = 0; it < n_ct_iter; ++it) {
          const int ct_raw = ct0 + it * ctstep;
          const bool ok = ct_raw < T;
          const int ct = ok ? ct_raw : 0;
          const int p = p_t;
          MT m10 = raw[at(side, p, ct, 0)], m11 = raw[at(side, p, ct, 1)], m20 = raw[at(side, p, ct, 2)], m21 = raw[at(side, p, ct, 3)];
          if (FACT) {
            const double d10 = (double)m10, d11 = (double)m11;
            const double so = SO[ct * nS + s0];
            const uint32_t v = __ballot_sync(0xffffffffu, ok && pvalid && fabs(d11 - d10) >= so);  // cond2 - cond1 (R/utils_cci.R:212-239)
            uint32_t vb = 0;
            if (BAND) vb = __ballot_sync(0xffffffffu, ok && pvalid && band_diff(d11, d10, so, rt));
            if (gleader && ok) {
              CNT[ct * nS + s0] += __popc(v & gmask);
              if (BAND) BND[ct * nS + s0] += __popc(vb & gmask);
            }
          }
          for (int k = 1; k < nSide; ++k) {
            const int g = s_gs[s0 + k];  // block-uniform
            if (g >= 0) {
              const size_t i0 = (((size_t)ch * A.G + g) * K + ct) * 32 + lane0 + p, i1 = i0 + (size_t)T * 32;
              const MT a10 = M1[i0], a11 = M1[i1], a20 = M2[i0], a21 = M2[i1];
              m10 = a10 < m10 ? a10 : m10; m11 = a11 < m11 ? a11 : m11;   // means are never NaN: plain minimum
              m20 = a20 < m20 ? a20 : m20; m21 = a21 < m21 ? a21 : m21;
              if (FACT) {
                const double d10 = (double)a10, d11 = (double)a11;
                const double so = SO[ct * nS + s0 + k];
                const uint32_t v = __ballot_sync(0xffffffffu, ok && pvalid && fabs(d11 - d10) >= so);
                uint32_t vb = 0;
                if (BAND) vb = __ballot_sync(0xffffffffu, ok && pvalid && band_diff(d11, d10, so, rt));
                if (gleader && ok) {
                  CNT[ct * nS + s0 + k] += __popc(v & gmask);
                  if (BAND) BND[ct * nS + s0 + k] += __popc(vb & gmask);
                }
              }
            }
          }
          if (ok && has_extra) {
            raw[at(side, p, ct, 0)] = m10; raw[at(side, p, ct, 1)] = m11; raw[at(side, p, ct, 2)] = m20; raw[at(side, p, ct, 3)] = m21;
          }
        }
      }
    }
    if (!FACT) {  // subunit differences per (replicate, cell type) for the per-row comparison (tables not built the reference's way)
      for (int side = 0; side < 2; ++side) {
        const int s0 = side ? nL : 0, nSide = side ? nR : nL;
        for (int ct = ct0; ct < T; ct += ctstep) {
          const int p = p_t;
          double* dd = (side ? DR : DL) + (p * Tp + ct) * nSide;
          double* ww = (side ? WR : WL) + (p * Tp + ct) * nSide;
          for (int k = 0; k < nSide; ++k) {
            const int g = s_gs[s0 + k];
            double d = 0.0, w = 0.0;
            if (g >= 0) {
              const size_t i0 = (((size_t)ch * A.G + g) * K + ct) * 32 + lane0 + p, i1 = i0 + (size_t)T * 32;
              const double a10 = (double)M1[i0], a11 = (double)M1[i1];
              d = a11 - a10;
              w = a11 + a10;
            }
            dd[k] = d;
            if (BAND) ww[k] = w;
          }
        }
      }
    }
    if (has_extra || !FACT) __syncthreads();  // tile / D-tile writes of the fix-up are visible to the rows
    // ---------------- rows ----------------
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (rid[q] < 0) continue;
      const double aob0 = fabs(ob[q][0]);
      const double o2a = ob[q][1] * ob[q][1], o2b = ob[q][2] * ob[q][2];
      uint32_t c0 = 0, c1 = 0, c2 = 0, b0 = 0, b1 = 0, b2 = 0;
      // moving pointers (one add per replicate each) instead of p * Tp index arithmetic on every load
      const float4* fa_p = reinterpret_cast<const float4*>(raw) + e_[q];
      const float4* fb_p = reinterpret_cast<const float4*>(raw + (F32 ? sideStride : 0)) + r_[q];
      const double2* pl = reinterpret_cast<const double2*>(raw);
      const int pl2 = plane / 2;  // double2 per plane
      const double2* d0_p = pl + e_[q];             // pass 1, emitter side
      const double2* d1_p = pl + pl2 + r_[q];       // pass 1, receiver side
      const double2* d2_p = pl + 2 * pl2 + e_[q];   // pass 2, emitter side
      const double2* d3_p = pl + 3 * pl2 + r_[q];   // pass 2, receiver side
#pragma unroll 2
      for (int p = 0; p < pmax; ++p) {
        double2 a1, a2, b1v, b2v;   // emitter / receiver minima: pass 1 (c1, c2), pass 2 (c1, c2)
        if (F32) {
          const float4 fa = *fa_p, fb = *fb_p;
          fa_p += Tp; fb_p += Tp;
          a1 = make_double2((double)fa.x, (double)fa.y); a2 = make_double2((double)fa.z, (double)fa.w);
          b1v = make_double2((double)fb.x, (double)fb.y); b2v = make_double2((double)fb.z, (double)fb.w);
        } else {
          a1 = *d0_p; b1v = *d1_p; a2 = *d2_p; b2v = *d3_p;
          d0_p += Tp; d1_p += Tp; d2_p += Tp; d3_p += Tp;
        }
        if (A.score_type == 0) {
          c0 += de_exceeds<BAND>(a1.y * b1v.y, a1.x * b1v.x, aob0, kslack, rt, b0) ? 1u : 0u;
          const double xa = a2.x * b2v.x, xb = a2.y * b2v.y;
          c1 += (xa >= th[q][0]) ? 1u : 0u;
          c2 += (xb >= th[q][1]) ? 1u : 0u;
          if (BAND) {
            b1 += band_product(xa, o2a, rt) ? 1u : 0u;
            b2 += band_product(xb, o2b, rt) ? 1u : 0u;
          }
        } else {
          const double sc1 = (a1.y + b1v.y) / 2, sc0 = (a1.x + b1v.x) / 2;
          const double de = sc1 - sc0;
          const double sa = (a2.x + b2v.x) / 2, sb = (a2.y + b2v.y) / 2;
          c0 += (fabs(de) >= aob0) ? 1u : 0u;
          c1 += (sa >= ob[q][1]) ? 1u : 0u;
          c2 += (sb >= ob[q][2]) ? 1u : 0u;
          if (BAND) {
            b0 += band_diff(sc1, sc0, aob0, rt) ? 1u : 0u;
            b1 += band_value(sa, ob[q][1], rt) ? 1u : 0u;
            b2 += band_value(sb, ob[q][2], rt) ? 1u : 0u;
          }
        }
        if (!FACT) {
          const double* dl = DL + (p * Tp + e_[q]) * nL;
          const double* dr = DR + (p * Tp + r_[q]) * nR;
          const double* wl = WL + (p * Tp + e_[q]) * nL;
          const double* wr = WR + (p * Tp + r_[q]) * nR;
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            if (s < nS) {
              const double d = (s < nL) ? dl[s] : dr[s - nL];
              const double ao = fabs(ob[q][FACT ? 0 : 3 + s]);
              cnt[q][FACT ? 0 : 3 + s] += (fabs(d) >= ao) ? 1u : 0u;
              if (BAND) {
                const double w = (s < nL) ? wl[s] : wr[s - nL];
                bnd[q][(FACT || !BAND) ? 0 : 3 + s] += (fabs(fabs(d) - ao) <= rt * w) ? 1u : 0u;
              }
            }
          }
        }
      }
      cnt[q][0] += c0; cnt[q][1] += c1; cnt[q][2] += c2;
      if (BAND) { bnd[q][0] += b0; bnd[q][BAND ? 1 : 0] += b1; bnd[q][BAND ? 2 : 0] += b2; }
    }
    __syncthreads();  // everyone is done with RAW buffer t & 1 and the D tiles before they are overwritten
  }
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    if (rid[q] < 0) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A.counts[(size_t)c * A.R + rid[q]] += cnt[q][c];
      if (BAND) A.band[(size_t)c * A.R + rid[q]] += bnd[q][BAND ? c : 0];
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < nS) {
        const int fi = ((s < nL) ? e_[q] : r_[q]) * nS + s;
        A.counts[(size_t)(3 + s) * A.R + rid[q]] += FACT ? CNT[fi] : cnt[q][FACT ? 0 : 3 + s];
        if (BAND) A.band[(size_t)(3 + s) * A.R + rid[q]] += FACT ? BND[fi] : bnd[q][(FACT || !BAND) ? 0 : 3 + s];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K3, differential mode, tiles filled by the TMA unit (cp.async.bulk + mbarrier): the production kernel whenever the
// subunit columns factorise (FACT). The tile keeps the means' NATURAL layout: for every (side, pass, cond) plane and cell
// type, the P consecutive replicates of the tile are one contiguous piece (P x sizeof(MT) bytes, 16-byte aligned) of the
// 256-byte row [gene][k][32 lanes] in global memory, so ONE bulk copy issued by ONE thread moves it: 8 T copies per tile
// instead of 8 T P eight-byte LDGSTS (the cp.async kernel above spends ~9 % of its instructions on them at T = 60: ncu
// r2e source page), completion is signalled on an mbarrier (transaction bytes) instead of per-thread wait_group +
// __syncthreads. Shared-memory rows are padded to PP = P + 2 (fp64) / P + 4 (fp32) elements: the row stride in 16-byte chunks
// is odd, so the rows of a warp (consecutive receivers) read their 16-byte vectors — 2 (fp64) / 4 (fp32) replicates of one
// plane — from distinct bank groups. Everything else (fix-up votes of the FACT columns, exact DE decision, product
// thresholds, band counters) is the arithmetic of k_score_diff.
// shared memory: RAW[2 buffers][8 planes = side * 4 + pass * 2 + cond][T][PP] of MT, then SO / CNT (/ BND) as in k_score_diff.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
      ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

template <int RPT, typename MT, bool BAND>
__global__ void __launch_bounds__(256, 2) k_score_diff_bulk(ScoreDiffArgs A) {
  extern __shared__ __align__(16) double sm_d[];
  constexpr bool F32 = sizeof(MT) == 4;
  constexpr int V = F32 ? 4 : 2;               // replicates per 16-byte vector
  const MT* const M1 = reinterpret_cast<const MT*>(A.M1);
  const MT* const M2 = reinterpret_cast<const MT*>(A.M2);
  __shared__ int s_gs[8];
  __shared__ int s_extra;
  __shared__ __align__(8) uint64_t s_bar[2];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  const int l = A.task_lri[blockIdx.x], rb = A.task_beg[blockIdx.x], re = A.task_end[blockIdx.x];
  const int T = A.T, P = A.P, K = A.K, nL = A.nL, nR = A.nR, nS = nL + nR;
  const int PP = P + (F32 ? 4 : 2);
  const double rt = A.rel_tol;
  const double kslack = kDeSlack + (BAND ? 1.5 * rt : 0.0);
  if (tid < 8) s_gs[tid] = (tid < nS) ? A.sub_gene[l * nS + tid] : -1;
  if (tid == 0) {
    int extra = 0;
    for (int s = 0; s < nS; ++s)
      if (s != 0 && s != nL && A.sub_gene[l * nS + s] >= 0) extra = 1;
    s_extra = extra;
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int e_[RPT], r_[RPT], rid[RPT];
  double ob[RPT][3], th[RPT][2];
  uint32_t cnt[RPT][3], bnd[RPT][BAND ? 3 : 1];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int pos = rb + tid + q * nthr;
    rid[q] = (pos < re) ? A.srow[pos] : -1;
    e_[q] = r_[q] = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { ob[q][c] = 0.0; cnt[q][c] = 0; }
#pragma unroll
    for (int c = 0; c < (BAND ? 3 : 1); ++c) bnd[q][c] = 0;
    th[q][0] = th[q][1] = 0.0;
    if (rid[q] >= 0) {
      e_[q] = A.row_emit[rid[q]];
      r_[q] = A.row_recv[rid[q]];
#pragma unroll
      for (int c = 0; c < 3; ++c) ob[q][c] = A.obs[(size_t)c * A.R + rid[q]];
      th[q][0] = A.thr[rid[q]];
      th[q][1] = A.thr[(size_t)A.R + rid[q]];
    }
  }
  const int planeSz = T * PP;                   // elements of one plane
  const int rawSize = 8 * planeSz;              // elements of one RAW buffer
  MT* const rawbase = reinterpret_cast<MT*>(sm_d);
  double* const tail = sm_d + (2 * (size_t)rawSize * sizeof(MT) + 7) / 8;
  double* const SO = tail;
  uint32_t* const CNT = reinterpret_cast<uint32_t*>(tail + T * nS);
  uint32_t* const BND = CNT + T * nS;
  for (int i = tid; i < T * nS; i += nthr) {
    SO[i] = fabs(A.obs_sub[(size_t)l * T * nS + i]);  // NaN (absent / unused): never exceeded, never in the band
    CNT[i] = 0;
    if (BAND) BND[i] = 0;
  }
  const int n_tiles = (A.n_valid + P - 1) / P;
  __syncthreads();
  const int g_first[2] = {s_gs[0], s_gs[nL]};
  const bool has_extra = s_extra != 0;
  const int lgP = 31 - __clz(P);
  const int p_t = tid & (P - 1), ct0 = tid >> lgP, ctstep = nthr >> lgP;
  const int n_ct_iter = (T + ctstep - 1) / ctstep;   // uniform trip count: the fix-up pass votes with full-warp ballots
  const uint32_t gmask = ((P >= 32) ? 0xffffffffu : ((1u << P) - 1u)) << (lane & ~(P - 1) & 31);
  const bool gleader = (lane & (P - 1) & 31) == 0;
  const unsigned row_bytes = (unsigned)(P * sizeof(MT));
  // tile t -> buffer t & 1: thread 0 arms the buffer's mbarrier with the tile's bytes, the threads share the 8 T copies
  auto issue = [&](int t) {
    const int pb = t * P, ch = pb >> 5, lane0 = pb & 31;
    MT* dst0 = rawbase + (size_t)(t & 1) * rawSize;
    uint64_t* bar = &s_bar[t & 1];
    if (tid == 0) mbar_arrive_expect_tx(bar, (unsigned)(8 * T) * row_bytes);
    for (int i = tid; i < 8 * T; i += nthr) {
      const int q = i / T, ct = i - q * T, side = q >> 2, k = q & 3;
      const MT* src = ((k >> 1) ? M2 : M1) + (((size_t)ch * A.G + g_first[side]) * K + (k & 1) * T + ct) * 32 + lane0;
      bulk_g2s(dst0 + (size_t)q * planeSz + ct * PP, src, row_bytes, bar);
    }
  };
  if (n_tiles > 0) issue(0);
  for (int t = 0; t < n_tiles; ++t) {
    if (t + 1 < n_tiles) issue(t + 1);   // its buffer was released by the barrier that ended tile t - 1
    mbar_wait(&s_bar[t & 1], (unsigned)((t >> 1) & 1));
    MT* raw = rawbase + (size_t)(t & 1) * rawSize;
    const int pb = t * P, ch = pb >> 5, lane0 = pb & 31;
    const int pmax = min(P, A.n_valid - pb);
    const bool pvalid = p_t < pmax;
    // -------- fix-up: subunit columns voted per (replicate, cell type); extra subunits of complexes folded in --------
    for (int side = 0; side < 2; ++side) {
      const int s0 = side ? nL : 0, nSide = side ? nR : nL;
      MT* pl = raw + (size_t)side * 4 * planeSz;
      for (int it = 0; it < n_ct_iter; ++it) {
        const int ct_raw = ct0 + it * ctstep;
        const bool ok = ct_raw < T;
        const int ct = ok ? ct_raw : 0;
        const int p = p_t, o = ct * PP + p;
        MT m10 = pl[o], m11 = pl[planeSz + o], m20 = pl[2 * planeSz + o], m21 = pl[3 * planeSz + o];
        {
          const double d10 = (double)m10, d11 = (double)m11;
          const double so = SO[ct * nS + s0];
          const uint32_t v = __ballot_sync(0xffffffffu, ok && pvalid && fabs(d11 - d10) >= so);  // cond2 - cond1 (R/utils_cci.R:212-239)
          uint32_t vb = 0;
          if (BAND) vb = __ballot_sync(0xffffffffu, ok && pvalid && band_diff(d11, d10, so, rt));
          if (gleader && ok) {
            CNT[ct * nS + s0] += __popc(v & gmask);
            if (BAND) BND[ct * nS + s0] += __popc(vb & gmask);
          }
        }
        for (int k = 1; k < nSide; ++k) {
          const int g = s_gs[s0 + k];  // block-uniform
          if (g >= 0) {
            const size_t i0 = (((size_t)ch * A.G + g) * K + ct) * 32 + lane0 + p, i1 = i0 + (size_t)T * 32;
            const MT a10 = M1[i0], a11 = M1[i1], a20 = M2[i0], a21 = M2[i1];
            m10 = a10 < m10 ? a10 : m10; m11 = a11 < m11 ? a11 : m11;   // means are never NaN: plain minimum
            m20 = a20 < m20 ? a20 : m20; m21 = a21 < m21 ? a21 : m21;
            const double d10 = (double)a10, d11 = (double)a11;
            const double so = SO[ct * nS + s0 + k];
            const uint32_t v = __ballot_sync(0xffffffffu, ok && pvalid && fabs(d11 - d10) >= so);
            uint32_t vb = 0;
            if (BAND) vb = __ballot_sync(0xffffffffu, ok && pvalid && band_diff(d11, d10, so, rt));
            if (gleader && ok) {
              CNT[ct * nS + s0 + k] += __popc(v & gmask);
              if (BAND) BND[ct * nS + s0 + k] += __popc(vb & gmask);
            }
          }
        }
        if (ok && has_extra) { pl[o] = m10; pl[planeSz + o] = m11; pl[2 * planeSz + o] = m20; pl[3 * planeSz + o] = m21; }
      }
    }
    if (has_extra) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes to the tile before a later bulk copy into it
      __syncthreads();
    }
    // ---------------- rows: V replicates per 16-byte vector ----------------
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (rid[q] < 0) continue;
      const double aob0 = fabs(ob[q][0]);
      const double o2a = ob[q][1] * ob[q][1], o2b = ob[q][2] * ob[q][2];
      uint32_t c0 = 0, c1 = 0, c2 = 0, b0 = 0, b1 = 0, b2 = 0;
      const MT* le = raw + e_[q] * PP;                           // emitter side planes 0..3
      const MT* rr = raw + (size_t)4 * planeSz + r_[q] * PP;     // receiver side planes 4..7
      for (int p0 = 0; p0 < pmax; p0 += V) {
        MT Lm[4][V], Rm[4][V];   // [pass * 2 + cond][replicate], widened on use
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (F32) {
            const float4 a = *reinterpret_cast<const float4*>(le + (size_t)k * planeSz + p0);
            const float4 b = *reinterpret_cast<const float4*>(rr + (size_t)k * planeSz + p0);
            Lm[k][0] = a.x; Lm[k][1] = a.y; Lm[k][V - 2] = a.z; Lm[k][V - 1] = a.w;
            Rm[k][0] = b.x; Rm[k][1] = b.y; Rm[k][V - 2] = b.z; Rm[k][V - 1] = b.w;
          } else {
            const double2 a = *reinterpret_cast<const double2*>(le + (size_t)k * planeSz + p0);
            const double2 b = *reinterpret_cast<const double2*>(rr + (size_t)k * planeSz + p0);
            Lm[k][0] = a.x; Lm[k][V - 1] = a.y;
            Rm[k][0] = b.x; Rm[k][V - 1] = b.y;
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          if (p0 + v < pmax) {
            if (A.score_type == 0) {
              c0 += de_exceeds<BAND>((double)Lm[1][v] * (double)Rm[1][v], (double)Lm[0][v] * (double)Rm[0][v], aob0, kslack, rt, b0) ? 1u : 0u;
              const double xa = (double)Lm[2][v] * (double)Rm[2][v], xb = (double)Lm[3][v] * (double)Rm[3][v];
              c1 += (xa >= th[q][0]) ? 1u : 0u;
              c2 += (xb >= th[q][1]) ? 1u : 0u;
              if (BAND) {
                b1 += band_product(xa, o2a, rt) ? 1u : 0u;
                b2 += band_product(xb, o2b, rt) ? 1u : 0u;
              }
            } else {
              const double sc1 = ((double)Lm[1][v] + (double)Rm[1][v]) / 2, sc0 = ((double)Lm[0][v] + (double)Rm[0][v]) / 2;
              const double de = sc1 - sc0;
              const double sa = ((double)Lm[2][v] + (double)Rm[2][v]) / 2, sb = ((double)Lm[3][v] + (double)Rm[3][v]) / 2;
              c0 += (fabs(de) >= aob0) ? 1u : 0u;
              c1 += (sa >= ob[q][1]) ? 1u : 0u;
              c2 += (sb >= ob[q][2]) ? 1u : 0u;
              if (BAND) {
                b0 += band_diff(sc1, sc0, aob0, rt) ? 1u : 0u;
                b1 += band_value(sa, ob[q][1], rt) ? 1u : 0u;
                b2 += band_value(sb, ob[q][2], rt) ? 1u : 0u;
              }
            }
          }
        }
      }
      cnt[q][0] += c0; cnt[q][1] += c1; cnt[q][2] += c2;
      if (BAND) { bnd[q][0] += b0; bnd[q][BAND ? 1 : 0] += b1; bnd[q][BAND ? 2 : 0] += b2; }
    }
    __syncthreads();  // everyone is done with RAW buffer t & 1 before the next bulk copies overwrite it
  }
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    if (rid[q] < 0) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A.counts[(size_t)c * A.R + rid[q]] += cnt[q][c];
      if (BAND) A.band[(size_t)c * A.R + rid[q]] += bnd[q][BAND ? c : 0];
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < nS) {
        const int fi = ((s < nL) ? e_[q] : r_[q]) * nS + s;
        A.counts[(size_t)(3 + s) * A.R + rid[q]] += CNT[fi];
        if (BAND) A.band[(size_t)(3 + s) * A.R + rid[q]] += BND[fi];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Observed-pass helper (aggregate_cells on the original labels): one thread per gene, sequential over the column.
// out is column-major [K][G]: out[g*K + k]. mode 0: sum of x ; mode 1: sum of 1*(x > 0).
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_group_means(int G, int K, int T, const int* __restrict__ colptr, const int* __restrict__ rowidx,
                              const double* __restrict__ values, const uint8_t* __restrict__ ct,
                              const uint8_t* __restrict__ cond, const double* __restrict__ ngroup, int mode,
                              double* __restrict__ out) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double* o = out + (size_t)g * K;
  for (int k = 0; k < K; ++k) o[k] = 0.0;
  for (int e = colptr[g]; e < colptr[g + 1]; ++e) {
    int cell = rowidx[e];
    int k = (cond ? cond[cell] * T : 0) + ct[cell];
    double v = values[e];
    o[k] += (mode == 0) ? v : ((v > 0) ? 1.0 : 0.0);
  }
  for (int k = 0; k < K; ++k) o[k] = o[k] / ngroup[k];
}

// thr = min{y : sqrt_rn(y) >= o}: (sqrt(x) >= o) == (x >= thr) for every x because sqrt_rn is monotone non-decreasing.
__device__ __forceinline__ double next_up(double y) {  // y >= 0, finite
  return __longlong_as_double(__double_as_longlong(y) + 1);
}
__device__ __forceinline__ double next_down(double y) {  // y > 0
  return __longlong_as_double(__double_as_longlong(y) - 1);
}
__global__ void k_product_thresholds(size_t n, const double* __restrict__ obs, double* __restrict__ thr) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double o = obs[i];
  double y;
  if (isnan(o)) {
    y = o;                 // x >= NaN is false, like sqrt(x) >= NA
  } else if (o <= 0.0) {
    y = 0.0;               // sqrt(x) >= o for every x >= 0; x < 0 gives NaN >= o (false) and x >= 0 false alike
  } else if (isinf(o)) {
    y = o;
  } else {
    y = o * o;
    if (isinf(y)) {
      // sqrt(DBL_MAX) < o: only +inf passes
    } else {
      while (y > 0.0 && sqrt(next_down(y)) >= o) y = next_down(y);
      while (sqrt(y) < o) y = next_up(y);  // may step up to +inf
    }
  }
  thr[i] = y;
}

// ------------------------------------------------------------------------------------------------------------------
// Observed pass over template rows (reference run_simple_cci_analysis(compute_fast = FALSE), R/utils_cci.R:171-283):
// per row and condition gather the subunit means and detection rates of the emitter (ligands) / receiver (receptors),
// complex-min, CCI_SCORE and IS_CCI_EXPRESSED (is_detected_full, R/utils_cci.R:783-802). One thread per row; outputs are
// column-major over the n rows of this chunk. means / drate: [G][K] (K = C * T, k = cond * T + ct).
// ------------------------------------------------------------------------------------------------------------------
struct ObservedArgs {
  int n, T, C, nL, nS, score_type;
  double threshold_pct, threshold_min_cells;
  const int* row_lri;
  const int* row_emit;
  const int* row_recv;
  const int* sub_gene;
  const double* means;
  const double* drate;
  const double* ngroup;
  double* expr;           // [C * nS][n]
  double* dr;             // [C * nS][n]
  double* score;          // [C][n]
  unsigned char* is_expr; // [C][n]
};

__global__ void k_observed_rows(ObservedArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const int l = A.row_lri[i], e = A.row_emit[i], r = A.row_recv[i];
  const int K = A.C * A.T;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int c = 0; c < A.C; ++c) {
    double mL = 0.0, mR = 0.0, dL = 0.0, dR = 0.0;
    bool hL = false, hR = false;
    for (int s = 0; s < A.nS; ++s) {
      const int g = A.sub_gene[l * A.nS + s];
      const bool isL = s < A.nL;
      double v = nan, d = nan;
      if (g >= 0) {
        const size_t idx = (size_t)g * K + c * A.T + (isL ? e : r);
        v = A.means[idx];
        d = A.drate[idx];
        if (isL) { mL = hL ? fmin(mL, v) : v; dL = hL ? fmin(dL, d) : d; hL = true; }
        else { mR = hR ? fmin(mR, v) : v; dR = hR ? fmin(dR, d) : d; hR = true; }
      }
      A.expr[(size_t)(c * A.nS + s) * A.n + i] = v;
      A.dr[(size_t)(c * A.nS + s) * A.n + i] = d;
    }
    A.score[(size_t)c * A.n + i] = A.score_type == 0 ? sqrt(mL * mR) : (mL + mR) / 2;
    const double ne = A.ngroup[c * A.T + e], nr = A.ngroup[c * A.T + r];
    A.is_expr[(size_t)c * A.n + i] = (dL >= A.threshold_pct && dL * ne >= A.threshold_min_cells &&
                                      dR >= A.threshold_pct && dR * nr >= A.threshold_min_cells) ? 1 : 0;
  }
}

__global__ void k_widen_counts(size_t n, const uint32_t* __restrict__ in, long long* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (long long)in[i];
}

// Factorised subunit columns (k_score_diff FACT): the table obs_sub[L][T][nS] of the observed subunit differences shared
// by all rows with the same (LRI, emitter) resp. (LRI, receiver). Built and verified on the device: k_fact_fill lets every
// row store its values (the table starts as all-ones bit patterns = NaN; concurrent writers race, any winner will do),
// k_fact_check then compares every row with the table bit for bit: if the rows of one pair disagree, at least one of them
// differs from the winner and sets the flag (the host then takes the per-row path).
__global__ void k_fact_fill(int R, int T, int nL, int nS, const int* __restrict__ row_lri, const int* __restrict__ row_emit,
                            const int* __restrict__ row_recv, const double* __restrict__ obs, double* __restrict__ obs_sub) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int l = row_lri[r], e = row_emit[r], rc = row_recv[r];
  for (int s = 0; s < nS; ++s)
    obs_sub[((size_t)l * T + (s < nL ? e : rc)) * nS + s] = obs[(size_t)(3 + s) * R + r];
}
__global__ void k_fact_check(int R, int T, int nL, int nS, const int* __restrict__ row_lri, const int* __restrict__ row_emit,
                             const int* __restrict__ row_recv, const double* __restrict__ obs,
                             const double* __restrict__ obs_sub, int* __restrict__ mismatch) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int l = row_lri[r], e = row_emit[r], rc = row_recv[r];
  for (int s = 0; s < nS; ++s) {
    const double a = obs_sub[((size_t)l * T + (s < nL ? e : rc)) * nS + s], b = obs[(size_t)(3 + s) * R + r];
    if (__double_as_longlong(a) != __double_as_longlong(b) && !(isnan(a) && isnan(b))) *mismatch = 1;
  }
}

// counts[i] = 0xFFFFFFFF where the observed value is NaN (NA): the host maps it to SCDC_COUNT_NA (counts are <= 2e9)
__global__ void k_mark_na(size_t n, const double* __restrict__ obs, uint32_t* __restrict__ counts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && isnan(obs[i])) counts[i] = 0xFFFFFFFFu;
}

// multi-GPU merge without NCCL: acc += add (a peer's counters copied over NVLink)
__global__ void k_add_counts(size_t n, const uint32_t* __restrict__ add, uint32_t* __restrict__ acc) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += add[i];
}

}  // namespace scdc
