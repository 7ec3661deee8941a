// scdc_api.cu — host side of the C-ABI declared in include/scdc.h: validation, upload, batch loop, counter merge.
// The arithmetic lives in scdc_kernels.cuh. No CPU fallback: every compute entry point needs an sm_100 device.
//
// Environment knobs (experiments and tests only; the defaults are the measured best):
//   SCDC_TIMING=1            host phase times on stderr
//   SCDC_URN_LABELS=1        sequential urn label generator instead of the random-key kernels (also read by the oracle)
//   SCDC_BMAX, SCDC_SUPER_BATCH   cap of the reduction batch / of the label super-batch (replicates)
//   SCDC_J, SCDC_J32, SCDC_W, SCDC_U, SCDC_PAIR   k_scatter: replicates per lane (fp64 / fp32), warps per block, (value, label)
//                            slots per lane, entries per dependent step (0 / 1 / 2 -> 1 / 2 / 4)
//   SCDC_U1, SCDC_W1, SCDC_OVERLAP_PASS1, SCDC_P1_PERSISTENT, SCDC_P1_BLOCKS, SCDC_P1_STAGED   k_pass1_cond: records per pipeline stage,
//                            warps per block, run next to k_scatter (default 1) as a persistent grid (default 1) of n blocks per
//                            SM, records through a shared-memory ring (0 / 1; default by shape)
//   SCDC_SCORE_V1, SCDC_SCORE_P, SCDC_SCORE_SMEM_KB, SCDC_SCORE_THREADS, SCDC_NO_FACT, SCDC_OVERLAP_SCORE, SCDC_SCORE_BULK   K3 variants
//   SCDC_K1_THREADS          k_draw_labels_keys2: 512 / 1024 threads per replicate block (0 = by shared-memory footprint)
//   SCDC_NO_EARLY_REDUCE, SCDC_CHUNK_NNZ, SCDC_NO_PRELAUNCH   one-shot call: no reduction behind the upload, chunk size,
//                            no label pre-launch
//   SCDC_NO_NCCL=1, SCDC_BCAST_NCCL=1   multi-GPU call: peer copies + adds instead of ncclAllReduce; ncclBroadcast instead of the copy
//                            engines for the problem
//   SCDC_NO_POOL, SCDC_NO_STAGER   plain cudaMalloc / plain pageable copies
#include "../../include/scdc.h"
#include "scdc_kernels.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(e_ == cudaErrorMemoryAllocation ? SCDC_ENOMEM : SCDC_ECUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                           \
  } while (0)

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct PhaseTimer {  // SCDC_TIMING=1: host-side phase times on stderr
  bool on;
  double t;
  PhaseTimer() : on(getenv("SCDC_TIMING") != nullptr), t(now_ms()) {}
  void lap(const char* what) {
    if (!on) return;
    const double n = now_ms();
    fprintf(stderr, "[scdc timing] %-28s %8.2f ms\n", what, n - t);
    t = n;
  }
};

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

// Device buffers come from the device's default stream-ordered memory pool (cudaMallocAsync) with the release threshold
// lifted, so the GB-sized work buffers of one call are recycled by the next instead of going through cudaMalloc / cudaFree
// (measured 30-600 ms and erratic per call on a shared host). Allocation and free are ordered on the calling thread's
// "allocation stream" (the handle's stream, set by every API entry point); other streams are made to wait on it.
thread_local cudaStream_t g_alloc_stream = nullptr;
thread_local bool g_alloc_async = false;
// bytes this thread has asked the driver to copy (scdc_stats.h2d_bytes / d2h_bytes / p2p_bytes)
thread_local double g_h2d_bytes = 0, g_d2h_bytes = 0, g_p2p_bytes = 0;

template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DBuf() { release(); }
  void release() {
    if (p) {
      if (g_alloc_async) cudaFreeAsync(p, g_alloc_stream); else cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return g_alloc_async ? cudaMallocAsync(&p, count * sizeof(T), g_alloc_stream) : cudaMalloc(&p, count * sizeof(T));
  }
  cudaError_t ensure(size_t count) { return (count <= n && p) ? cudaSuccess : alloc(count); }
  cudaError_t upload(const T* src, size_t count, cudaStream_t s) {
    cudaError_t e = alloc(count);
    if (e != cudaSuccess || count == 0) return e;
    g_h2d_bytes += (double)(count * sizeof(T));
    return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

// Host-to-device copy of a large PAGEABLE buffer (the caller's matrix) through a pair of pinned staging buffers that are
// filled by several host threads: the driver's own staging of pageable memory is a single-threaded copy (~9 GB/s measured
// here); four threads into pinned memory plus asynchronous DMA roughly double that. The staging buffers are allocated once
// per process and reused (cudaHostAlloc is slow). Falls back to a plain cudaMemcpyAsync when pinned memory is unavailable.
class PinnedStager {
 public:
  static constexpr size_t kChunk = 16u << 20;
  static constexpr int kMaxThreads = 6;
  // min_bytes: smaller copies go through the driver's own (single-threaded) staging; n_thr: staging threads (<= 6)
  cudaError_t copy(void* dst, const void* src, size_t bytes, cudaStream_t s, size_t min_bytes = 2 * kChunk,
                   int n_thr = kMaxThreads) {
    if (bytes < min_bytes || getenv("SCDC_NO_STAGER")) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    Slot* slot = acquire();
    if (!slot) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    cudaError_t rc = cudaSuccess;
    n_thr = std::max(1, std::min(n_thr, kMaxThreads));
    const long n_chunks = (long)((bytes + kChunk - 1) / kChunk);
    // The helper threads live for the whole copy (not per chunk): chunk i may be filled once open_chunk >= i, i.e. once the
    // DMA that last read its staging buffer has finished; filled[b] counts the helpers that are done with buffer b.
    std::atomic<long> open_chunk{-1};
    std::atomic<int> filled[2];
    filled[0].store(0); filled[1].store(0);
    auto slice = [&](long i, int t) {
      const size_t off = (size_t)i * kChunk, n = std::min(kChunk, bytes - off);
      const size_t part = (n / n_thr + 63) & ~size_t(63);
      const size_t lo = std::min(n, part * t), hi = std::min(n, part * (t + 1));
      if (hi > lo) memcpy(slot->buf[i & 1] + lo, static_cast<const char*>(src) + off + lo, hi - lo);
    };
    std::vector<std::thread> helpers;
    try {
      for (int t = 1; t < n_thr; ++t)
        helpers.emplace_back([&, t]() {
          for (long i = 0; i < n_chunks; ++i) {
            while (open_chunk.load(std::memory_order_acquire) < i) std::this_thread::yield();
            if (open_chunk.load(std::memory_order_acquire) >= n_chunks) return;  // the copy was abandoned
            slice(i, t);
            filled[i & 1].fetch_add(1, std::memory_order_release);
          }
        });
    } catch (...) {  // fewer threads than planned: the slices are cut for n_thr, so the calling thread covers the missing ones
    }
    const int n_help = (int)helpers.size();
    for (long i = 0; i < n_chunks && rc == cudaSuccess; ++i) {
      const int b = (int)(i & 1);
      const size_t off = (size_t)i * kChunk, n = std::min(kChunk, bytes - off);
      if (i >= 2) rc = cudaEventSynchronize(slot->done[b]);  // the DMA that last read this staging buffer has finished
      if (rc != cudaSuccess) break;
      filled[b].store(0, std::memory_order_relaxed);
      open_chunk.store(i, std::memory_order_release);
      slice(i, 0);
      for (int t = n_help + 1; t < n_thr; ++t) slice(i, t);
      while (filled[b].load(std::memory_order_acquire) < n_help) std::this_thread::yield();
      rc = cudaMemcpyAsync(static_cast<char*>(dst) + off, slot->buf[b], n, cudaMemcpyHostToDevice, s);
      if (rc == cudaSuccess) rc = cudaEventRecord(slot->done[b], s);
    }
    open_chunk.store(n_chunks, std::memory_order_release);  // releases helpers that are still waiting (error path)
    for (auto& t : helpers) t.join();
    if (rc == cudaSuccess) { cudaEventSynchronize(slot->done[0]); cudaEventSynchronize(slot->done[1]); }
    release(slot);
    return rc;
  }

  // Device-to-host through the same pair of pinned buffers: chunk i is DMA'd into pinned memory while the caller's
  // `consume(chunk, offset, bytes)` works on chunk i - 1 (e.g. widens counters straight into the caller's array), so no
  // pageable temporary of the whole size is needed. Returns cudaErrorNotSupported when no pinned slot is available.
  template <typename F>
  cudaError_t fetch(const void* src_dev, size_t bytes, cudaStream_t s, F&& consume) {
    Slot* slot = acquire();
    if (!slot) return cudaErrorNotSupported;
    cudaError_t rc = cudaSuccess;
    const long n_chunks = (long)((bytes + kChunk - 1) / kChunk);
    auto issue = [&](long i) -> cudaError_t {
      const size_t off = (size_t)i * kChunk, n = std::min(kChunk, bytes - off);
      cudaError_t e = cudaMemcpyAsync(slot->buf[i & 1], static_cast<const char*>(src_dev) + off, n, cudaMemcpyDeviceToHost, s);
      return e == cudaSuccess ? cudaEventRecord(slot->done[i & 1], s) : e;
    };
    if (n_chunks > 0) rc = issue(0);
    for (long i = 0; i < n_chunks && rc == cudaSuccess; ++i) {
      if (i + 1 < n_chunks) rc = issue(i + 1);
      if (rc == cudaSuccess) rc = cudaEventSynchronize(slot->done[i & 1]);
      if (rc != cudaSuccess) break;
      const size_t off = (size_t)i * kChunk, n = std::min(kChunk, bytes - off);
      consume(slot->buf[i & 1], off, n);
    }
    if (rc != cudaSuccess) cudaStreamSynchronize(s);
    release(slot);
    return rc;
  }

 private:
  struct Slot {
    char* buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool busy = false;
    int device = -1;  // the events belong to this device: a slot is only reused by copies to the same device
  };
  std::mutex mu_;
  std::vector<Slot*> slots_;
  Slot* acquire() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu_);
    for (Slot* sl : slots_)
      if (!sl->busy && sl->device == dev) { sl->busy = true; return sl; }
    Slot* sl = new (std::nothrow) Slot;
    if (!sl) return nullptr;
    sl->device = dev;
    bool ok = true;
    for (int b = 0; b < 2 && ok; ++b) {
      ok = cudaHostAlloc(reinterpret_cast<void**>(&sl->buf[b]), kChunk, cudaHostAllocPortable) == cudaSuccess &&
           cudaEventCreateWithFlags(&sl->done[b], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
      cudaGetLastError();
      for (int b = 0; b < 2; ++b) { if (sl->buf[b]) cudaFreeHost(sl->buf[b]); if (sl->done[b]) cudaEventDestroy(sl->done[b]); }
      delete sl;
      return nullptr;
    }
    sl->busy = true;
    slots_.push_back(sl);
    return sl;
  }
  void release(Slot* sl) {
    std::lock_guard<std::mutex> lock(mu_);
    sl->busy = false;
  }
};
PinnedStager g_stager;

template <typename T>
cudaError_t upload_big(DBuf<T>& buf, const T* src, size_t count, cudaStream_t s) {
  cudaError_t e = buf.alloc(count);
  if (e != cudaSuccess || count == 0) return e;
  g_h2d_bytes += (double)(count * sizeof(T));
  return g_stager.copy(buf.p, src, count * sizeof(T), s);
}

// every API entry point that touches device memory of a handle opens one of these
struct AllocScope {
  cudaStream_t prev_s;
  bool prev_a;
  explicit AllocScope(cudaStream_t s) : prev_s(g_alloc_stream), prev_a(g_alloc_async) {
    g_alloc_stream = s;
    g_alloc_async = !getenv("SCDC_NO_POOL");
  }
  ~AllocScope() { g_alloc_stream = prev_s; g_alloc_async = prev_a; }
};

// Per-device facts and idle stream triplets, kept for the life of the process: a one-shot call then neither queries the
// driver (cudaMemGetInfo is occasionally slow) nor creates / destroys three streams.
struct DeviceCache {
  struct Streams { cudaStream_t s = nullptr, labels = nullptr, score = nullptr; };
  struct Entry {
    bool known = false;
    int num_sms = 0;
    size_t smem_optin = 0, mem_total = 0;
    std::vector<Streams> idle;
  };
  std::mutex mu;
  std::vector<Entry> dev;
  Entry& entry(int d) {
    if ((int)dev.size() <= d) dev.resize((size_t)d + 1);
    return dev[(size_t)d];
  }
};
DeviceCache g_devcache;

}  // namespace

struct scdc_handle {
  int device = 0;
  int N = 0, G = 0, T = 0, C = 0, L = 0, R = 0, nL = 0, nR = 0, K = 0, ncols = 0, nS = 0;
  size_t nnz = 0;
  cudaStream_t stream = nullptr, stream_labels = nullptr, stream_score = nullptr;
  int num_sms = 148;
  size_t smem_optin = 0;
  // problem
  DBuf<int> colptr, rowidx, segptr, gene_order, sub_gene, row_lri, row_emit, row_recv;
  DBuf<double> values, ngroup, obs;
  DBuf<uint4> rec_s;      // pass-1 copy of the matrix: (cell, value) records, inside each gene column sorted by cell type
  DBuf<float> values_f;   // SCDC_FP32_FAST: fp32 copies of the values / of the pass-1 records (made on the device at the
  DBuf<uint2> rec_sf;     //                 first fp32 run)
  DBuf<uint8_t> ct, cond;
  DBuf<uint32_t> expect;      // n(cell type, condition) as uint32 [K]
  DBuf<int> visit, ct_seg;     // K1 visit order (differential: cells stably sorted by cell type) and its segments
  // K3: rows grouped by LRI, split into block tasks; product thresholds for the one-sided geometric columns
  DBuf<int> task_lri, task_beg, task_end, srow;
  DBuf<double> thr, obs_sub;   // obs_sub: [L][T][nS] common observed subunit differences (factorised subunit columns)
  bool fact_ok = false;
  int n_tasks = 0, score_threads = 256;
  // per-run workspace (grown on demand)
  DBuf<uint32_t> bits1[2], counts, band;   // label buffers are double-buffered (K1 runs one batch ahead on stream_labels)
  DBuf<uint8_t> lab2[2], up_cond, up_ct;
  DBuf<double> M1[2], M2[2], dist;   // group means are double-buffered: K3 of batch b overlaps K2 of batch b+1
  DBuf<int> flag, early_genes, keyfail;   // keyfail: overflow flag of the random-key label generator
  DBuf<int> p1_counter;                   // work counter of the persistent pass-1 kernel
  DBuf<uint8_t> label_tmp;               // replicate-major labels of the random-key generator before re-layout
  double ms_upload = 0, h2d_bytes = 0;   // scdc_create: wall time and bytes copied host -> device
  size_t mem_total = 0;
  // host copies needed to (re)install tested rows: sub_gene table and per-gene nnz
  std::vector<int> sub_gene_h, gene_nnz_h;
  const std::atomic<bool>* abort = nullptr;  // multi-GPU: set by the calling thread when the interrupt callback fires
  bool one_shot = false;       // created by scdc_perm_counts for a single run (options known at creation)
  int max_nt = 0, max_nc = 0;  // largest cell type / largest condition (cells): inputs of keygen_plan2
  int n_active = 0;        // genes referenced by the installed rows (gene_order holds them, heaviest first)
  size_t nnz_active = 0;
  // labels of super-batch 0 drawn ahead of scdc_run (one-shot path: K1 overlaps the H2D of the matrix)
  struct Pre {
    bool valid = false;
    uint64_t seed = 0;
    int64_t perm_offset = 0, SB = 0;
    int J = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    // detection mode: the group sums of batch 0 were also launched ahead, gene chunk by gene chunk, behind the chunked
    // upload of the matrix (k2 events bracket them on the handle's stream)
    bool k2_done = false;
    int B = 0, k2_launches = 0;
    cudaEvent_t k2_t0 = nullptr, k2_t1 = nullptr;
    // differential mode: pass 1 of batch 0 was launched ahead too (next to that scatter); p1_ev marks its end
    bool p1_done = false;
    cudaEvent_t p1_ev = nullptr;
  } pre;
};

namespace {

constexpr int kScoreThreads = 256, kScoreRPT = 2;
// multi-GPU call: invoked by create_impl on the first device as soon as the matrix, the cell labels and the list of active
// genes are on the device and batch 0's scatter is running there (phase A of the cloning: the other devices get the matrix
// over NVLink and start their own labels / first scatter / pass-1 sort while the first device's host thread is still busy
// with the rows)
struct CreateHook {
  virtual int matrix_ready(scdc_handle* root) = 0;
  virtual ~CreateHook() {}
};
int prelaunch_labels(scdc_handle* h, const scdc_options* o);
int upload_matrix_with_early_reduction(scdc_handle* h, const scdc_problem* p, const scdc_options* o,
                                       std::atomic<int>& genes_valid, std::atomic<bool>& matrix_done, const int& matrix_rc);
int launch_early_scatter(scdc_handle* h, const scdc_options* o);
int launch_early_pass1(scdc_handle* h, const scdc_options* o, cudaStream_t sp);

int count_sm100_devices(std::vector<int>* ids) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int found = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) {
      ++found;
      if (ids) ids->push_back(d);
    }
  }
  return found;
}

// O(nnz) part of the validation: cell indices in range and strictly ascending inside every gene column.
// progress (optional) = number of leading gene columns already found valid.
int validate_matrix_range(const scdc_problem* p, int g_lo, int g_hi, std::atomic<int>* progress) {
  const int N = p->n_cells;
  for (int g = g_lo; g < g_hi; ++g) {
    if (progress && (g & 15) == 0) progress->store(g, std::memory_order_release);
    for (int e = p->colptr[g]; e < p->colptr[g + 1]; ++e) {
      int c = p->rowidx[e];
      if (c < 0 || c >= N) return fail(SCDC_EINVAL, "rowidx out of range at entry %d", e);
      if (e > p->colptr[g] && c <= p->rowidx[e - 1])
        return fail(SCDC_EINVAL, "rowidx not strictly ascending inside gene column %d", g);
    }
  }
  if (progress) progress->store(g_hi, std::memory_order_release);
  return SCDC_OK;
}

// n_thr > 1 (and no progress counter): gene ranges of about equal nnz are checked by n_thr host threads; the failure in the
// lowest gene range is the one reported, as the sequential scan would.
int validate_matrix(const scdc_problem* p, std::atomic<int>* progress = nullptr, int n_thr = 1) {
  const int G = p->n_genes;
  const long long nnz = p->colptr[G];
  if (progress || n_thr <= 1 || nnz < (1 << 22)) return validate_matrix_range(p, 0, G, progress);
  std::vector<int> bounds(1, 0);
  for (int w = 1; w <= n_thr; ++w) {
    const long long target = nnz * w / n_thr;
    int g = bounds.back();
    while (g < G && (w == n_thr || p->colptr[g + 1] <= target)) ++g;
    bounds.push_back(g);
  }
  std::vector<int> rcs((size_t)n_thr, SCDC_OK);
  std::vector<std::string> errs((size_t)n_thr);
  std::vector<std::thread> th;
  auto body = [&](int w) {
    rcs[w] = validate_matrix_range(p, bounds[w], bounds[w + 1], nullptr);
    if (rcs[w]) errs[w] = g_err;
  };
  try {
    for (int w = 1; w < n_thr; ++w) th.emplace_back(body, w);
  } catch (...) {  // no more threads: the calling thread checks the ranges that got none
    for (int w = (int)th.size() + 1; w < n_thr; ++w) body(w);
  }
  body(0);
  for (auto& t : th) t.join();
  for (int w = 0; w < n_thr; ++w)
    if (rcs[w]) { g_err = errs[w]; return rcs[w]; }
  return SCDC_OK;
}

// everything except the matrix entries (dimensions, labels, LRI table, rows): O(N + G + L + R)
int validate_meta(const scdc_problem* p) {
  if (!p) return fail(SCDC_EINVAL, "problem is NULL");
  if (p->n_cells < 1 || p->n_genes < 1 || p->n_celltypes < 1 || p->n_lri < 1 || p->n_rows < 0)
    return fail(SCDC_EINVAL, "non-positive dimension (N=%d G=%d T=%d L=%d R=%d)", p->n_cells, p->n_genes,
                p->n_celltypes, p->n_lri, p->n_rows);
  if (p->n_conditions != 1 && p->n_conditions != 2) return fail(SCDC_EINVAL, "n_conditions must be 1 or 2");
  if (p->n_celltypes * p->n_conditions > 254)
    return fail(SCDC_EINVAL, "n_celltypes * n_conditions = %d exceeds the 8-bit label range (254)",
                p->n_celltypes * p->n_conditions);
  if (p->max_nL < 1 || p->max_nR < 1 || p->max_nL + p->max_nR > 8)
    return fail(SCDC_EINVAL, "max_nL/max_nR must be >= 1 with max_nL + max_nR <= 8");
  if (!p->colptr || !p->rowidx || !p->values || !p->ct_code || !p->sub_gene || (p->n_rows > 0 && (!p->row_lri || !p->row_emit || !p->row_recv || !p->observed)))
    return fail(SCDC_EINVAL, "NULL array in problem");
  if (p->n_conditions == 2 && !p->cond_code) return fail(SCDC_EINVAL, "cond_code is NULL in differential mode");
  const int N = p->n_cells, G = p->n_genes, T = p->n_celltypes, C = p->n_conditions;
  if (p->colptr[0] != 0) return fail(SCDC_EINVAL, "colptr[0] != 0");
  for (int g = 0; g < G; ++g)
    if (p->colptr[g + 1] < p->colptr[g]) return fail(SCDC_EINVAL, "colptr not non-decreasing at gene %d", g);
  if ((long long)p->colptr[G] > 2147483647LL - 256) return fail(SCDC_EINVAL, "too many non-zeros for 32-bit entry indices");
  std::vector<int64_t> ng((size_t)C * T, 0);
  for (int i = 0; i < N; ++i) {
    int t = p->ct_code[i];
    int c = (C == 2) ? p->cond_code[i] : 0;
    if (t < 0 || t >= T) return fail(SCDC_EINVAL, "ct_code[%d] = %d out of range", i, t);
    if (c < 0 || c >= C) return fail(SCDC_EINVAL, "cond_code[%d] = %d out of range", i, c);
    ++ng[(size_t)c * T + t];
  }
  for (size_t k = 0; k < ng.size(); ++k)
    if (ng[k] < 1)
      return fail(SCDC_EINVAL, "group (condition %d, cell type %d) has no cells", (int)(k / T), (int)(k % T));
  const int nS = p->max_nL + p->max_nR;
  for (int l = 0; l < p->n_lri; ++l) {
    for (int s = 0; s < nS; ++s) {
      int g = p->sub_gene[(size_t)l * nS + s];
      if (g < -1 || g >= G) return fail(SCDC_EINVAL, "sub_gene[%d][%d] = %d out of range", l, s, g);
    }
  }
  for (int r = 0; r < p->n_rows; ++r) {
    if (p->row_lri[r] < 0 || p->row_lri[r] >= p->n_lri || p->row_emit[r] < 0 || p->row_emit[r] >= T ||
        p->row_recv[r] < 0 || p->row_recv[r] >= T)
      return fail(SCDC_EINVAL, "row %d has an out-of-range LRI / emitter / receiver code", r);
    const int l = p->row_lri[r];
    if (p->sub_gene[(size_t)l * nS] < 0 || p->sub_gene[(size_t)l * nS + p->max_nL] < 0)
      return fail(SCDC_EINVAL, "row %d: LRI %d lacks LIGAND_1 or RECEPTOR_1", r, l);
  }
  return SCDC_OK;
}

int validate_problem(const scdc_problem* p) {
  int rc = validate_meta(p);
  return rc ? rc : validate_matrix(p, nullptr, 4);
}

int validate_options(const scdc_problem* /*unused*/, int C, const scdc_options* o) {
  if (!o) return fail(SCDC_EINVAL, "options is NULL");
  if (o->iterations < 1) return fail(SCDC_EINVAL, "iterations must be >= 1");
  if (o->iterations > 2000000000LL) return fail(SCDC_EINVAL, "iterations must be <= 2e9 per call (32-bit counters)");
  if (o->perm_offset < 0) return fail(SCDC_EINVAL, "perm_offset must be >= 0");
  if (o->score_type != SCDC_SCORE_GEOMETRIC && o->score_type != SCDC_SCORE_ARITHMETIC)
    return fail(SCDC_EINVAL, "unknown score_type %d", o->score_type);
  if (o->precision != SCDC_FP64_EXACT && o->precision != SCDC_FP32_FAST)
    return fail(SCDC_EINVAL, "precision must be SCDC_FP64_EXACT or SCDC_FP32_FAST");
  if (std::isnan(o->rel_tol) || o->rel_tol > 0.5) return fail(SCDC_EINVAL, "rel_tol must be a number <= 0.5 (<= 0: the default 1e-5)");
  if (o->batch_size < 0) return fail(SCDC_EINVAL, "batch_size must be >= 0");
  if (C == 1 && o->perm_cond) return fail(SCDC_EINVAL, "perm_cond given in detection mode");
  if (C == 2 && ((o->perm_cond != nullptr) != (o->perm_ct != nullptr)))
    return fail(SCDC_EINVAL, "differential validation mode needs both perm_cond and perm_ct");
  return SCDC_OK;
}


}  // namespace

extern "C" {

const char* scdc_last_error(void) { return g_err.c_str(); }
const char* scdc_version(void) { return SCDC_VERSION_STRING; }
int scdc_device_count(void) { return count_sm100_devices(nullptr); }

void scdc_release(scdc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  const int h_device = h->device;
  cudaStream_t s0 = h->stream, s1 = h->stream_labels, s2 = h->stream_score;
  cudaEvent_t e0 = h->pre.t0, e1 = h->pre.t1, e2 = h->pre.k2_t0, e3 = h->pre.k2_t1, e4 = h->pre.p1_ev;
  if (s1) cudaStreamSynchronize(s1);  // e.g. labels drawn ahead of a run that never happened
  if (s2) cudaStreamSynchronize(s2);
  {
    AllocScope scope(s0);
    if (!s0) g_alloc_async = false;
    delete h;  // buffers go back to the pool in stream order
  }
  const int device = h_device;
  if (s0) cudaStreamSynchronize(s0);
  if (s0 && s1 && s2 && cudaGetLastError() == cudaSuccess) {  // idle and healthy: keep the triplet for the next handle
    std::lock_guard<std::mutex> lock(g_devcache.mu);
    DeviceCache::Entry& e = g_devcache.entry(device);
    if (e.idle.size() < 8) {
      DeviceCache::Streams st;
      st.s = s0; st.labels = s1; st.score = s2;
      e.idle.push_back(st);
      s0 = s1 = s2 = nullptr;
    }
  }
  if (s0) cudaStreamDestroy(s0);
  if (s1) cudaStreamDestroy(s1);
  if (s2) cudaStreamDestroy(s2);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e2) cudaEventDestroy(e2);
  if (e3) cudaEventDestroy(e3);
  if (e4) cudaEventDestroy(e4);
}

// Installs the tested rows (sub_cci_template_iter + observed vectors) on a handle: row tables, NA mask, K3 work list,
// product thresholds, factorisation table of the subunit columns, and the list of ACTIVE genes (the genes the rows
// reference — the reference's gene subset of data_tr, R/utils_permutation.R:79-84 — heaviest column first).
// Part 1 of install_rows: the list of active genes (all the group sums need from the rows).
static int install_active_genes(scdc_handle* h, int n_rows, const int32_t* row_lri) {
  cudaStream_t s = h->stream;
  const int G = h->G;
  std::vector<uint8_t> lri_used((size_t)h->L, 0), gene_used((size_t)G, 0);
  for (int r = 0; r < n_rows; ++r) lri_used[row_lri[r]] = 1;
  for (int l = 0; l < h->L; ++l)
    if (lri_used[l])
      for (int k = 0; k < h->nS; ++k) {
        const int g = h->sub_gene_h[(size_t)l * h->nS + k];
        if (g >= 0) gene_used[g] = 1;
      }
  std::vector<int> order;
  h->nnz_active = 0;
  for (int g = 0; g < G; ++g)
    if (gene_used[g] || n_rows == 0) { order.push_back(g); h->nnz_active += (size_t)h->gene_nnz_h[g]; }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->gene_nnz_h[a] > h->gene_nnz_h[b]; });
  h->n_active = (int)order.size();
  CU(h->gene_order.upload(order.data(), order.size(), s));
  CU(cudaStreamSynchronize(s));  // host staging vector dies at scope exit
  return SCDC_OK;
}

// `up` != NULL: the active genes are installed already and the group sums of batch 0 are running on the handle's streams;
// allocate, copy and launch on `up` instead (an idle stream), and make the handle's stream wait for it at the end.
static int install_rows(scdc_handle* h, int n_rows, const int32_t* row_lri, const int32_t* row_emit, const int32_t* row_recv,
                        const double* observed, cudaStream_t up = nullptr) {
  cudaStream_t s = up ? up : h->stream;
  AllocScope alloc_scope(s);
  const int T = h->T, C = h->C;
  h->R = n_rows;
  PhaseTimer pt;
  if (!up) {
    int rc = install_active_genes(h, n_rows, row_lri);
    if (rc) return rc;
  }
  CU(upload_big(h->row_lri, row_lri, (size_t)n_rows, s));
  CU(upload_big(h->row_emit, row_emit, (size_t)n_rows, s));
  CU(upload_big(h->row_recv, row_recv, (size_t)n_rows, s));
  CU(upload_big(h->obs, observed, (size_t)n_rows * h->ncols, s));   // NaN = NA: marked in the counters on the device (k_mark_na)
  // K3 work list: rows stably sorted by LRI (counting sort), each LRI cut into slices of score_threads * kScoreRPT rows
  std::vector<int> srow(n_rows), t_lri, t_beg, t_end;
  {
    std::vector<int> start((size_t)h->L + 1, 0);
    for (int r = 0; r < n_rows; ++r) ++start[(size_t)row_lri[r] + 1];
    for (int l = 0; l < h->L; ++l) start[(size_t)l + 1] += start[l];
    for (int r = 0; r < n_rows; ++r) srow[start[row_lri[r]]++] = r;
  }
  h->score_threads = env_int("SCDC_SCORE_THREADS", kScoreThreads);
  if (h->score_threads < 32 || h->score_threads > kScoreThreads || (h->score_threads & 31)) h->score_threads = kScoreThreads;
  const int rpb = h->score_threads * kScoreRPT;
  for (int b = 0; b < n_rows;) {
    int e = b;
    while (e < n_rows && row_lri[srow[e]] == row_lri[srow[b]]) ++e;
    for (int c = b; c < e; c += rpb) {
      t_lri.push_back(row_lri[srow[b]]);
      t_beg.push_back(c);
      t_end.push_back(std::min(e, c + rpb));
    }
    b = e;
  }
  h->n_tasks = (int)t_lri.size();
  pt.lap("rows: K3 work list");
  CU(h->thr.alloc((size_t)n_rows * C + 1));
  if (n_rows > 0) {  // product thresholds of the one-sided geometric columns, computed on the device (IEEE sqrt)
    scdc::k_product_thresholds<<<(unsigned)(((size_t)n_rows * C + 255) / 256), 256, 0, s>>>(
        (size_t)n_rows * C, h->obs.p + (C == 2 ? (size_t)n_rows : 0), h->thr.p);
    CU(cudaGetLastError());
  }
  h->fact_ok = false;
  bool fact_pending = false;
  if (C == 2 && n_rows > 0 && !env_int("SCDC_NO_FACT", 0)) {
    // The subunit columns of a row depend only on (LRI, emitter) / (LRI, receiver). If the observed subunit
    // differences agree bitwise across the rows sharing such a pair (true for tables built by the reference, where they
    // are joins of the same group means), K3 counts them once per (cell type, subunit). Table and check on the device.
    const size_t n_tab = (size_t)h->L * T * h->nS;
    CU(h->obs_sub.alloc(n_tab));
    CU(h->flag.ensure(1));
    CU(cudaMemsetAsync(h->obs_sub.p, 0xFF, n_tab * sizeof(double), s));  // all-ones = NaN: unused / absent
    CU(cudaMemsetAsync(h->flag.p, 0, sizeof(int), s));
    const unsigned nb = (unsigned)((n_rows + 255) / 256);
    scdc::k_fact_fill<<<nb, 256, 0, s>>>(n_rows, T, h->nL, h->nS, h->row_lri.p, h->row_emit.p, h->row_recv.p, h->obs.p, h->obs_sub.p);
    scdc::k_fact_check<<<nb, 256, 0, s>>>(n_rows, T, h->nL, h->nS, h->row_lri.p, h->row_emit.p, h->row_recv.p, h->obs.p,
                                          h->obs_sub.p, h->flag.p);
    CU(cudaGetLastError());
    fact_pending = true;
  }
  CU(h->srow.upload(srow.data(), srow.size(), s));
  CU(h->task_lri.upload(t_lri.data(), t_lri.size(), s));
  CU(h->task_beg.upload(t_beg.data(), t_beg.size(), s));
  CU(h->task_end.upload(t_end.data(), t_end.size(), s));
  pt.lap("rows: enqueue rest");
  CU(cudaStreamSynchronize(s));  // host staging vectors die at scope exit
  if (fact_pending) {
    int mismatch = 1;
    CU(cudaMemcpy(&mismatch, h->flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    h->fact_ok = mismatch == 0;
  }
  if (up) {
    cudaEvent_t done;
    CU(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(done, up);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, done, 0);
    cudaEventDestroy(done);
    CU(e);
  }
  return SCDC_OK;
}

// segptr / rec_s from colptr / rowidx / values / ct, all on the device (k_sort_by_celltype), enqueued on `s`
static int build_pass1_copy(scdc_handle* h, cudaStream_t s) {
  const int G = h->G, T = h->T;
  CU(h->segptr.alloc((size_t)G * (T + 1)));
  CU(h->rec_s.alloc(h->nnz ? h->nnz : 1));
  const int W = 8;
  scdc::k_sort_by_celltype<<<(G + W - 1) / W, W * 32, (size_t)W * (T + 1) * sizeof(int), s>>>(
      G, T, h->colptr.p, h->rowidx.p, h->values.p, h->ct.p, nullptr, h->segptr.p, h->rec_s.p);
  CU(cudaGetLastError());
  return SCDC_OK;
}

// streams (from the per-device cache when idle ones exist), pool settings and device facts of a new handle
static int open_device_context(scdc_handle* h, int device) {
  CU(cudaSetDevice(device));
  h->device = device;
  {
    std::lock_guard<std::mutex> lock(g_devcache.mu);
    DeviceCache::Entry& e = g_devcache.entry(device);
    if (!e.idle.empty()) {
      h->stream = e.idle.back().s; h->stream_labels = e.idle.back().labels; h->stream_score = e.idle.back().score;
      e.idle.pop_back();
    }
    if (e.known) { h->num_sms = e.num_sms; h->smem_optin = e.smem_optin; h->mem_total = e.mem_total; }
  }
  if (!h->stream) CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (!getenv("SCDC_NO_POOL")) {  // keep freed work buffers in the device's default pool for the next call
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  if (!h->stream_labels) {  // label generation runs ahead with few, long-lived warps: give its blocks scheduling priority
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&h->stream_labels, cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&h->stream_score, cudaStreamNonBlocking, lo));
  }
  if (h->mem_total == 0) {
    int v = 0;
    CU(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    h->num_sms = v;
    CU(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    h->smem_optin = (size_t)v;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    h->mem_total = total_b;
    std::lock_guard<std::mutex> lock(g_devcache.mu);
    DeviceCache::Entry& e = g_devcache.entry(device);
    e.known = true; e.num_sms = h->num_sms; e.smem_optin = h->smem_optin; e.mem_total = h->mem_total;
  }
  return SCDC_OK;
}

// Cell labels and everything derived from them alone (8-bit codes, group sizes, K1 visit order, largest strata): what
// the label generators need. Synchronises `s` (the staging vectors are local).
static int install_labels(scdc_handle* h, const scdc_problem* p, cudaStream_t s) {
  const int N = h->N, T = h->T, C = h->C, K = h->K;
  try {
    std::vector<uint8_t> ct8(N), cond8(N, 0);
    std::vector<uint32_t> ng((size_t)K, 0);
    for (int i = 0; i < N; ++i) {
      ct8[i] = (uint8_t)p->ct_code[i];
      if (C == 2) cond8[i] = (uint8_t)p->cond_code[i];
      ++ng[(size_t)cond8[i] * T + ct8[i]];
    }
    std::vector<double> ngd(K);
    for (int k = 0; k < K; ++k) ngd[k] = (double)ng[k];
    h->max_nt = h->max_nc = 0;
    for (int t = 0; t < T; ++t) h->max_nt = std::max(h->max_nt, (int)(ng[t] + (C == 2 ? ng[(size_t)T + t] : 0u)));
    for (int c = 0; c < C; ++c) {
      uint32_t n_c = 0;
      for (int t = 0; t < T; ++t) n_c += ng[(size_t)c * T + t];
      h->max_nc = std::max(h->max_nc, (int)n_c);
    }
    // K1 visit order: differential mode walks the cells grouped by cell type (stable), detection mode in cell order
    std::vector<int> visit, ct_seg;
    if (C == 2) {
      visit.resize(N);
      ct_seg.assign((size_t)T + 1, 0);
      for (int i = 0; i < N; ++i) ++ct_seg[(size_t)ct8[i] + 1];
      for (int t = 0; t < T; ++t) ct_seg[(size_t)t + 1] += ct_seg[t];
      std::vector<int> cur(ct_seg.begin(), ct_seg.end() - 1);
      for (int i = 0; i < N; ++i) visit[cur[ct8[i]]++] = i;
      CU(h->visit.upload(visit.data(), visit.size(), s));
      CU(h->ct_seg.upload(ct_seg.data(), ct_seg.size(), s));
    }
    CU(h->ct.upload(ct8.data(), N, s));
    CU(h->cond.upload(cond8.data(), N, s));
    CU(h->ngroup.upload(ngd.data(), K, s));
    CU(h->expect.upload(ng.data(), K, s));
    CU(cudaStreamSynchronize(s));
  } catch (const std::bad_alloc&) {
    return fail(SCDC_ENOMEM, "host allocation failed while staging the labels");
  }
  return SCDC_OK;
}

// o_hint != NULL (one-shot path): the options of the run that follows. The cheap part of the validation has been done by
// the caller; the O(nnz) matrix check runs on a helper thread and K1 is launched ahead, both overlapping the H2D copies.
static int create_impl(const scdc_problem* p, int32_t device, scdc_handle** out, const scdc_options* o_hint,
                       CreateHook* hook = nullptr) {
  if (!out) return fail(SCDC_EINVAL, "out is NULL");
  *out = nullptr;
  PhaseTimer pt;
  int rc = o_hint ? SCDC_OK : validate_problem(p);
  if (rc) return rc;
  pt.lap("create: validate");
  int matrix_rc = SCDC_OK;
  std::string matrix_err;
  std::thread matrix_check;
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } joiner{matrix_check};
  std::atomic<int> genes_valid{0};
  std::atomic<bool> matrix_done{false};
  if (o_hint) {
    // detection mode consumes the matrix chunk by chunk behind a progress counter (sequential scan); differential mode only
    // needs the verdict, so the scan is split over a few threads
    auto check = [&, p]() {
      matrix_rc = (p->n_conditions == 1) ? validate_matrix(p, &genes_valid) : validate_matrix(p, nullptr, 4);
      if (matrix_rc) matrix_err = g_err;
      matrix_done.store(true, std::memory_order_release);
    };
    try {
      matrix_check = std::thread(check);
    } catch (...) {
      check();  // no thread available: validate here, before the upload
    }
  }
  std::vector<int> ids;
  if (count_sm100_devices(&ids) == 0)
    return fail(SCDC_ENOGPU, "no visible CUDA device with compute capability 10.x (B200); there is no CPU fallback");
  if (device < 0) device = ids[0];
  if (std::find(ids.begin(), ids.end(), (int)device) == ids.end())
    return fail(SCDC_ENOGPU, "device %d is not a visible compute-capability-10.x device", device);
  const double t0 = now_ms(), h2d0 = g_h2d_bytes;
  CU(cudaSetDevice(device));
  std::unique_ptr<scdc_handle, void (*)(scdc_handle*)> h(new (std::nothrow) scdc_handle, scdc_release);
  if (!h) return fail(SCDC_ENOMEM, "host allocation failed");
  h->device = device;
  h->one_shot = o_hint != nullptr;
  h->N = p->n_cells; h->G = p->n_genes; h->T = p->n_celltypes; h->C = p->n_conditions; h->L = p->n_lri;
  h->R = p->n_rows; h->nL = p->max_nL; h->nR = p->max_nR; h->K = h->C * h->T; h->nS = h->nL + h->nR;
  h->ncols = (h->C == 2) ? 3 + h->nS : 1;
  h->nnz = (size_t)p->colptr[h->G];
  rc = open_device_context(h.get(), device);
  if (rc) return rc;
  AllocScope alloc_scope(h->stream);
  const int N = h->N, G = h->G, T = h->T, C = h->C, K = h->K;
  cudaStream_t s = h->stream;
  try {
    rc = install_labels(h.get(), p, s);
    if (rc) return rc;
    CU(h->colptr.upload(p->colptr, (size_t)G + 1, s));
    pt.lap("create: context + small prep");
    if (o_hint) {  // K1 needs only the labels: draw the first super-batch while the matrix is copied
      rc = prelaunch_labels(h.get(), o_hint);
      if (rc) return rc;
    }
    CU(h->sub_gene.upload(p->sub_gene, (size_t)h->L * h->nS, s));
    h->sub_gene_h.assign(p->sub_gene, p->sub_gene + (size_t)h->L * h->nS);
    h->gene_nnz_h.resize(G);
    for (int g = 0; g < G; ++g) h->gene_nnz_h[g] = p->colptr[g + 1] - p->colptr[g];
    bool rows_installed = false;
    if (o_hint && C == 1 && h->pre.valid && !o_hint->return_distributions && !env_int("SCDC_NO_EARLY_REDUCE", 0) &&
        p->n_rows > 0) {
      // detection mode, one-shot call: the rows go first (they define the active genes), then the matrix is copied gene
      // chunk by gene chunk and the group sums of batch 0 are launched behind every chunk, so the copy hides behind compute
      h->R = p->n_rows;
      rc = install_active_genes(h.get(), p->n_rows, p->row_lri);
      if (rc) return rc;
      rc = upload_matrix_with_early_reduction(h.get(), p, o_hint, genes_valid, matrix_done, matrix_rc);
      if (rc) {
        if (matrix_check.joinable()) matrix_check.join();
        cudaDeviceSynchronize();
        if (matrix_rc) { g_err = matrix_err; return matrix_rc; }
        return rc;
      }
      // the rest of the rows (K3's inputs) goes up on the copy stream while the group sums run
      rc = install_rows(h.get(), p->n_rows, p->row_lri, p->row_emit, p->row_recv, p->observed, h->stream_score);
      if (rc) return rc;
      rows_installed = true;
    } else {
      CU(upload_big(h->rowidx, p->rowidx, h->nnz, s));
      CU(upload_big(h->values, p->values, h->nnz, s));
    }
    if (matrix_check.joinable()) {
      matrix_check.join();
      if (matrix_rc) {
        cudaDeviceSynchronize();
        g_err = matrix_err;
        return matrix_rc;
      }
    }
    pt.lap("create: enqueue uploads");
    // Differential mode, one-shot call: the pass-2 group sums of batch 0 need only the matrix, the labels and the list of
    // active genes. They are launched now; the pass-1 copy of the matrix and the rows are built and uploaded on the copy
    // stream (and allocated in its order) while that kernel runs. run_device then skips batch 0's k_scatter.
    cudaStream_t s_up = s;
    struct TmpStream { cudaStream_t s = nullptr; ~TmpStream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } } rows_stream;
    cudaEvent_t ev_matrix;   // the matrix, the labels and the small tables are on the device
    CU(cudaEventCreateWithFlags(&ev_matrix, cudaEventDisableTiming));
    struct EvDel { cudaEvent_t e; ~EvDel() { cudaEventDestroy(e); } } ev_matrix_del{ev_matrix};
    CU(cudaEventRecord(ev_matrix, s));
    cudaEvent_t ev_sort;     // the pass-1 copy is built (when it was built on another stream)
    CU(cudaEventCreateWithFlags(&ev_sort, cudaEventDisableTiming));
    EvDel ev_sort_del{ev_sort};
    if (o_hint && C == 2 && h->pre.valid && !o_hint->return_distributions && !env_int("SCDC_NO_EARLY_REDUCE", 0) &&
        p->n_rows > 0) {
      h->R = p->n_rows;
      rc = install_active_genes(h.get(), p->n_rows, p->row_lri);
      if (rc) return rc;
      rc = launch_early_scatter(h.get(), o_hint);
      if (rc) return rc;
      // the rows' uploads and small kernels go to the HIGH-priority label stream (idle by now): on a low-priority stream
      // their blocks starved behind the scatter's 24 k blocks and create waited ~350 ms for them (SCDC_TIMING, config 5)
      if (h->pre.k2_done) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&rows_stream.s, cudaStreamNonBlocking, hi));
        s_up = rows_stream.s;   // (the label stream gets the sort and batch 0's pass 1 below)
      }
      if (hook && h->pre.k2_done) {
        rc = hook->matrix_ready(h.get());
        if (rc) return rc;
        CU(cudaSetDevice(device));
      }
    }
    if (C == 2) {
      // pass-1 copy of the matrix (inside each gene column, entries stably sorted by cell type), built on the device from
      // the CSC that is already there: no host sort, no second H2D of the matrix. With batch 0's scatter already running on
      // the compute stream, the sort goes to the high-priority label stream (idle by now: the labels were drawn during the
      // matrix upload), so that its few blocks are scheduled at once next to the scatter's; the rows' pageable copies
      // below use the third stream and do not queue behind it.
      cudaStream_t s_sort = (s_up != s) ? h->stream_labels : s;
      AllocScope sort_scope(s_sort);
      if (s_sort != s) CU(cudaStreamWaitEvent(s_sort, ev_matrix, 0));
      rc = build_pass1_copy(h.get(), s_sort);
      if (rc) return rc;
      if (s_sort != s) {
        rc = launch_early_pass1(h.get(), o_hint, s_sort);   // batch 0's pass 1 next to batch 0's scatter
        if (rc) return rc;
        CU(cudaEventRecord(ev_sort, s_sort));
        CU(cudaStreamWaitEvent(s, ev_sort, 0));  // everything later on the compute stream sees the pass-1 copy (and M1 of batch 0)
      }
    }
    AllocScope up_scope(s_up);
    pt.lap("create: pass-1 matrix copy");
    if (!rows_installed) {
      // s_up != s: everything queued on the copy stream so far is covered by the event install_rows records at its end
      rc = install_rows(h.get(), p->n_rows, p->row_lri, p->row_emit, p->row_recv, p->observed, s_up != s ? s_up : nullptr);
      if (rc) return rc;
    }
    CU(h->flag.alloc(1));
    pt.lap("create: enqueue rest");
    if (s_up != s) {
      // Batch 0's scatter is still running on the compute stream: do not wait for it. Everything that read host staging
      // memory on that stream was enqueued before ev_matrix; the rows went through the copy stream.
      CU(cudaStreamSynchronize(s_up));
      CU(cudaEventSynchronize(ev_matrix));
    } else {
      CU(cudaStreamSynchronize(s));  // host staging vectors die at scope exit
    }
    pt.lap("create: wait for H2D");
  } catch (const std::bad_alloc&) {
    return fail(SCDC_ENOMEM, "host allocation failed while staging the problem");
  }
  h->ms_upload = now_ms() - t0;
  h->h2d_bytes = g_h2d_bytes - h2d0;
  *out = h.release();
  return SCDC_OK;
}

int scdc_create(const scdc_problem* p, int32_t device, scdc_handle** out) { return create_impl(p, device, out, nullptr); }

}  // extern "C"

namespace {

template <int NG>
cudaError_t launch_draw_ng(const scdc_handle* h, cudaStream_t s, uint64_t seed, int64_t perm0, int B, int Bb, int J,
                           uint32_t* bits1, uint8_t* lab2, uint8_t* dbg_cond, uint8_t* dbg_ct) {
  auto kern = scdc::k_draw_labels<NG>;
  const size_t smem = (size_t)h->C * NG * 8 * 32 * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<B / 32, 32, smem, s>>>(h->N, h->T, h->C, h->C == 2 ? h->visit.p : nullptr, h->ct_seg.p, h->expect.p, seed, perm0, Bb, J,
                               bits1, lab2, dbg_cond, dbg_ct);
  return cudaGetLastError();
}

// The generators write the pass-1 condition bits as one word per 32 replicates; k_pass1_cond reads them byte-packed per
// lane (k_bits_transpose8, in place). B replicates in blocks of Bb (a multiple of 256 whenever pass 1 runs).
cudaError_t repack_condition_bits(const scdc_handle* h, cudaStream_t s, uint32_t* bits1, int64_t B, int Bb) {
  if (Bb % 256 != 0) return cudaSuccess;  // debug draws (scdc_draw_labels): the words are never consumed
  const size_t n_groups = (size_t)h->N * (size_t)(B / 256);
  if (n_groups) scdc::k_bits_transpose8<<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(n_groups, bits1);
  return cudaGetLastError();
}

// B = replicates drawn by this launch (a multiple of Bb), Bb = replicates per reduction batch (label block size).
// Detection mode uses the parallel random-key generator (one block per replicate) when keygen_plan allows it, else the
// sequential urn generator (one thread per replicate); oracle_philox_labels mirrors the same rule.
cudaError_t launch_draw(scdc_handle* h, cudaStream_t s, uint64_t seed, int64_t perm0, int B, int Bb, int J,
                        uint32_t* bits1, uint8_t* lab2, uint8_t* dbg_cond, uint8_t* dbg_ct) {
  const int T = h->T, N = h->N;
  const bool keys_allowed = !env_int("SCDC_URN_LABELS", 0);
  const scdc::KeygenPlan2 kp2 = scdc::keygen_plan2(N, T, h->max_nt, h->max_nc);
  const scdc::KeygenPlan kp = scdc::keygen_plan(N, T, h->C);
  const bool keys2 = keys_allowed && h->C == 2 && kp2.ok, keys1 = keys_allowed && !keys2 && kp.ok;
  if (keys1 || keys2) {
    // Random-key generators: one block per replicate writes replicate-major labels (label_tmp, or the caller's debug
    // arrays), k_relayout_labels transposes them into the reduction's [batch][cell][Bb] blocks (+ condition bit-planes).
    const bool dbg = dbg_ct || dbg_cond;
    cudaError_t e = cudaSuccess;
    if (!h->keyfail.p) {  // overflow flag of the generators, checked at the end of the run
      e = h->keyfail.alloc(1);
      if (e == cudaSuccess) e = cudaMemsetAsync(h->keyfail.p, 0, sizeof(int), g_alloc_async ? g_alloc_stream : s);
      if (e != cudaSuccess) return e;
    }
    uint8_t* out = nullptr;
    if (!dbg) {
      e = h->label_tmp.ensure((size_t)N * B);
      if (e != cudaSuccess) return e;
      out = h->label_tmp.p;
    }
    if (s != g_alloc_stream && g_alloc_async) {  // buffers are allocated in the order of the allocation stream
      cudaEvent_t dep;
      e = cudaEventCreateWithFlags(&dep, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
      cudaError_t e1 = cudaEventRecord(dep, g_alloc_stream), e2 = cudaStreamWaitEvent(s, dep, 0);
      cudaEventDestroy(dep);
      if (e1 != cudaSuccess) return e1;
      if (e2 != cudaSuccess) return e2;
    }
    if (keys2) {
      auto kern = scdc::k_draw_labels_keys2;
      if (kp2.smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kp2.smem);
      if (e != cudaSuccess) return e;
      // one block per SM when the histograms need more than half of its shared memory (T = 60, 250 k cells: 156 KB): then
      // 1024 threads instead of 512, so that the SM still has 32 warps to hide Philox's dependent multiply chain with
      int k1_threads = (kp2.smem > 112 * 1024 && kp2.NB2 >= 1024) ? 1024 : 512;
      const int k1_env = env_int("SCDC_K1_THREADS", 0);
      if (k1_env == 512 || (k1_env == 1024 && kp2.NB2 >= 1024)) k1_threads = k1_env;
      kern<<<B, k1_threads, kp2.smem, s>>>(N, T, h->ct.p, h->expect.p, seed, perm0, kp2.NB1, kp2.cap1, kp2.NB2, kp2.cap2, out,
                                           dbg_cond, dbg_ct, h->keyfail.p);
    } else {
      auto kern = scdc::k_draw_labels_keys;
      if (kp.smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kp.smem);
      if (e != cudaSuccess) return e;
      kern<<<B, 256, kp.smem, s>>>(N, T, h->expect.p, seed, perm0, kp.NB, kp.cap, dbg ? dbg_ct : out, h->keyfail.p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess || dbg) return e;
    for (int b0 = 0; b0 < B; b0 += Bb) {
      dim3 grid((N + 127) / 128, Bb / 128);
      scdc::k_relayout_labels<<<grid, 256, 0, s>>>(N, Bb, J, out + (size_t)b0 * N, lab2 + (size_t)b0 * N, keys2 ? T : 0,
                                                   keys2 ? bits1 + (size_t)b0 / 32 * N : nullptr);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess && keys2) e = repack_condition_bits(h, s, bits1, B, Bb);
    return e;
  }
  cudaError_t e;
  if (T <= 16) e = launch_draw_ng<2>(h, s, seed, perm0, B, Bb, J, bits1, lab2, dbg_cond, dbg_ct);
  else if (T <= 32) e = launch_draw_ng<4>(h, s, seed, perm0, B, Bb, J, bits1, lab2, dbg_cond, dbg_ct);
  else if (T <= 64) e = launch_draw_ng<8>(h, s, seed, perm0, B, Bb, J, bits1, lab2, dbg_cond, dbg_ct);
  else if (T <= 128) e = launch_draw_ng<16>(h, s, seed, perm0, B, Bb, J, bits1, lab2, dbg_cond, dbg_ct);
  else e = launch_draw_ng<32>(h, s, seed, perm0, B, Bb, J, bits1, lab2, dbg_cond, dbg_ct);
  if (e == cudaSuccess && h->C == 2 && !dbg_ct && !dbg_cond) e = repack_condition_bits(h, s, bits1, B, Bb);
  return e;
}

struct ScatterPlan {
  int J, W, blocks_per_sm;
  size_t smem;
  bool f32;
};

// f32 (SCDC_FP32_FAST): fp32 accumulators are half the size, so twice the replicates per lane fit the same footprint
ScatterPlan plan_scatter(const scdc_handle* h, int Bpad, bool f32) {
  const int K = h->K;
  const size_t vs = f32 ? 4 : 8, es = f32 ? 8 : 16;  // bytes per accumulator / per staged entry
  int J = f32 ? ((K <= 32) ? 4 : (K <= 64 ? 2 : 1)) : ((K <= 16) ? 4 : (K <= 32 ? 2 : 1));
  J = env_int(f32 ? "SCDC_J32" : "SCDC_J", J);
  if (J != 1 && J != 2 && J != 4) J = 1;
  while (J > 1 && (Bpad % (32 * J)) != 0) J >>= 1;
  const size_t per_warp = (size_t)K * J * 32 * vs + 64 * es;
  const int nCJ = Bpad / (32 * J);
  const size_t cap = h->smem_optin ? h->smem_optin : 232448;
  int bestW = 1, best_warps = 0, best_bps = 1;
  for (int W = 1; W <= 8 && W <= nCJ; ++W) {
    size_t need = per_warp * W;
    if (need > cap) break;
    int bps = (int)std::min<size_t>(32, (233472 - 0) / (need + 1024));
    if (bps < 1) bps = 1;
    int warps = std::min(64, bps * W);
    if (warps > best_warps || (warps == best_warps && W > bestW)) { best_warps = warps; bestW = W; best_bps = bps; }
  }
  int W = env_int("SCDC_W", bestW);
  if (W < 1 || W > 32 || per_warp * W > cap) W = bestW;
  ScatterPlan pl;
  pl.J = J; pl.W = W; pl.blocks_per_sm = best_bps; pl.smem = per_warp * W; pl.f32 = f32;
  return pl;
}

template <typename VT> const VT* matrix_values(const scdc_handle* h);
template <> const double* matrix_values<double>(const scdc_handle* h) { return h->values.p; }
template <> const float* matrix_values<float>(const scdc_handle* h) { return h->values_f.p; }

template <typename VT, int J, int U, int GRP>
cudaError_t launch_scatter_u(const scdc_handle* h, const ScatterPlan& pl, int Bpad, const uint8_t* lab2, int lab_stride,
                             void* M2, cudaStream_t s, const int* genes, int n_genes) {
  auto kern = scdc::k_scatter<VT, J, U, GRP>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return e;
  const int nCJ = Bpad / (32 * J);
  dim3 grid(n_genes, (nCJ + pl.W - 1) / pl.W);  // active genes only (gene_order, or one upload chunk of it)
  kern<<<grid, pl.W * 32, pl.smem, s>>>(h->G, h->K, nCJ, h->colptr.p, h->rowidx.p, matrix_values<VT>(h), lab2, lab_stride,
                                        h->ngroup.p, genes, static_cast<VT*>(M2));
  return cudaGetLastError();
}

template <typename VT, int J>
cudaError_t launch_scatter_j(const scdc_handle* h, const ScatterPlan& pl, int Bpad, const uint8_t* lab2, int lab_stride,
                             void* M2, cudaStream_t s, const int* genes, int n_genes) {
  // U = (value, label) slots per lane = how many entries ahead the label loads run; PAIR = two entries per dependent step.
  // Few resident warps (large K) need the deepest pipeline and the pair mode.
  const int warps_per_sm = pl.blocks_per_sm * pl.W;
  if constexpr (J == 1) {
    // registers per thread (ptxas): fp64 168 / 96 / 60 and fp32 132 / 80 / 48 for U = 32 / 16 / 8. Measured on config 5
    // (K = 120): fp64, 7 warps / SM: U = 32 343 ms vs U = 16 430 ms per 2560 replicates; fp32, 14 warps / SM: U = 16
    // 486 ms per 6400 vs U = 32 695 ms (U = 32 halves the resident warps there).
    const int U = env_int("SCDC_U", warps_per_sm <= 8 ? 32 : (warps_per_sm <= 24 ? 16 : 8));
    // entries per dependent step (SCDC_PAIR = 0 / 1 / 2 -> 1 / 2 / 4)
    const int pair = env_int("SCDC_PAIR", (warps_per_sm <= 16) ? 1 : 0);
#define SCDC_GO(UU, GG) return launch_scatter_u<VT, 1, UU, GG>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes)
    if (U >= 32) { if (pair >= 2) SCDC_GO(32, 4); if (pair) SCDC_GO(32, 2); SCDC_GO(32, 1); }
    if (U >= 16) { if (pair >= 2) SCDC_GO(16, 4); if (pair) SCDC_GO(16, 2); SCDC_GO(16, 1); }
    if (pair) SCDC_GO(8, 2);
    SCDC_GO(8, 1);
#undef SCDC_GO
  } else {
    const int U = env_int("SCDC_U", 16);
    if (U >= 16) return launch_scatter_u<VT, J, 16, 1>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
    return launch_scatter_u<VT, J, 8, 1>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
  }
}

// pass-2 group sums of one batch: plan -> kernel instance (value type, replicates per lane, prefetch depth, pair mode)
cudaError_t launch_scatter(const scdc_handle* h, const ScatterPlan& pl, int Bpad, const uint8_t* lab2, int lab_stride,
                           void* M2, cudaStream_t s, const int* genes = nullptr, int n_genes = -1) {
  if (!genes) { genes = h->gene_order.p; n_genes = h->n_active; }
  if (n_genes <= 0) return cudaSuccess;
  if (pl.f32) {
    if (pl.J == 4) return launch_scatter_j<float, 4>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
    if (pl.J == 2) return launch_scatter_j<float, 2>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
    return launch_scatter_j<float, 1>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
  }
  if (pl.J == 4) return launch_scatter_j<double, 4>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
  if (pl.J == 2) return launch_scatter_j<double, 2>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
  return launch_scatter_j<double, 1>(h, pl, Bpad, lab2, lab_stride, M2, s, genes, n_genes);
}

// SCDC_FP32_FAST: fp32 copies of the matrix values (pass 2) / of the pass-1 records, made on the device once per handle
// (enqueued on `s`, which must be ordered after the fp64 originals)
int ensure_f32_values(scdc_handle* h, cudaStream_t s, bool pass1) {
  const size_t n = h->nnz, need = std::max<size_t>(n, 1);
  if (pass1) {
    if (h->rec_sf.p && h->rec_sf.n >= need) return SCDC_OK;
    CU(h->rec_sf.alloc(need));
    if (n) scdc::k_rec_to_float<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, h->rec_s.p, h->rec_sf.p);
  } else {
    if (h->values_f.p && h->values_f.n >= need) return SCDC_OK;
    CU(h->values_f.alloc(need));
    if (n) scdc::k_to_float<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, h->values.p, h->values_f.p);
  }
  CU(cudaGetLastError());
  return SCDC_OK;
}

int pick_batch(const scdc_handle* h, const scdc_options* o) {
  const double per_perm = (double)h->C * h->G * h->K * (o->precision == SCDC_FP32_FAST ? 4.0 : 8.0) + (double)h->N * 1.125 +
                          ((o->perm_ct != nullptr) ? 2.0 * h->N : 0.0) +
                          (o->return_distributions ? (double)h->R * ((h->C == 2) ? 3 : 1) * 8.0 : 0.0);
  double budget = std::min((double)h->mem_total * 0.25, 24.0e9);
  long long B = (long long)(budget / per_perm);
  // Large batches waste fewer padded replicates (10 000 -> 1 x 10 112 instead of 3 x 3456). The one-shot detection call
  // keeps batches <= 4096: its first batch is reduced chunk by chunk behind the matrix upload, which is ~10 % less
  // efficient than the monolithic launch, so only about a third of the work should go that way.
  B = std::min<long long>(B, std::max(128, env_int("SCDC_BMAX", (h->one_shot && h->C == 1) ? 4096 : 16384)));
  if (o->batch_size > 0) B = o->batch_size;
  const long long q = (h->C == 2) ? 256 : 128;  // granularity: k_pass1_cond covers 256 replicates per warp, k_scatter 128
  B = std::max<long long>(q, (B / q) * q);
  if (o->batch_size <= 0) {
    // equal batches: the reduction always runs whole batches, so n batches of the smallest multiple of q that covers
    // iterations / n waste at most q - 1 replicates per batch (10 000 -> 1 x 10 112 instead of 3 x 4096); among the batch
    // counts that fit the memory budget (and up to twice the minimum) take the one that pads least
    // "Pads" counts what the scatter really runs: whole blocks of W warps x 32 J replicates (config 5, 12 500 replicates:
    // 5 x 2560 runs 5 x 84 chunks for 80 useful ones each, 7 x 1792 runs 7 x 56 full ones). SCDC_BATCH_RULE=0: plain padding.
    const bool block_aware = env_int("SCDC_BATCH_RULE", 1) != 0;
    const long long n_min = (o->iterations + B - 1) / B;
    long long best_B = 0, best_total = 0;
    for (long long n = n_min; n <= (block_aware ? 3 * n_min + 1 : 2 * n_min); ++n) {
      const long long Bn = (((o->iterations + n - 1) / n + q - 1) / q) * q;
      if (Bn > B) continue;
      long long total = n * Bn;
      if (block_aware) {
        const ScatterPlan pn = plan_scatter(h, (int)Bn, o->precision == SCDC_FP32_FAST);
        const long long slice = 32LL * pn.J * pn.W;
        total = n * (((Bn + slice - 1) / slice) * slice);
      }
      if (!best_B || total < best_total) { best_B = Bn; best_total = total; }
    }
    if (best_B) B = best_B;
  } else {
    B = std::min<long long>(B, ((o->iterations + q - 1) / q) * q);
  }
  return (int)B;
}

// replicates per label super-batch (device-drawn labels): as many as the label memory cap allows, normally all
int64_t pick_super_batch(const scdc_handle* h, const scdc_options* o, int B) {
  const int64_t iters_pad = ((o->iterations + B - 1) / B) * B;
  const double per_rep = (double)h->N * (h->C == 2 ? 1.125 : 1.0);
  const double cap = std::min((double)h->mem_total * 0.1, 8.0e9);
  int64_t SB = std::max<int64_t>(B, ((int64_t)(cap / per_rep) / B) * B);
  SB = std::min<int64_t>(SB, iters_pad);
  const int sb_env = env_int("SCDC_SUPER_BATCH", 0);
  if (sb_env > 0) SB = std::min<int64_t>(iters_pad, std::max<int64_t>(B, ((int64_t)sb_env / B) * B));
  return SB;
}

// One-shot path: draw the labels of super-batch 0 on the label stream right after the small uploads, so that K1 (which
// needs only the cell labels) overlaps the H2D of the expression matrix. run_device picks them up if the plan matches.
int prelaunch_labels(scdc_handle* h, const scdc_options* o) {
  if (o->perm_ct != nullptr || env_int("SCDC_NO_PRELAUNCH", 0)) return SCDC_OK;
  const int B = pick_batch(h, o);
  const ScatterPlan pl = plan_scatter(h, B, o->precision == SCDC_FP32_FAST);
  const int64_t SB = pick_super_batch(h, o, B);
  CU(h->lab2[0].ensure((size_t)h->N * SB));
  if (h->C == 2) CU(h->bits1[0].ensure((size_t)h->N * (SB / 32)));
  if (!h->pre.t0) { CU(cudaEventCreate(&h->pre.t0)); CU(cudaEventCreate(&h->pre.t1)); }
  cudaEvent_t dep;
  CU(cudaEventCreateWithFlags(&dep, cudaEventDisableTiming));
  cudaError_t e1 = cudaEventRecord(dep, h->stream), e2 = cudaStreamWaitEvent(h->stream_labels, dep, 0);
  cudaEventDestroy(dep);
  CU(e1); CU(e2);
  CU(cudaEventRecord(h->pre.t0, h->stream_labels));
  CU(launch_draw(h, h->stream_labels, o->seed, o->perm_offset, (int)SB, B, pl.J, h->bits1[0].p, h->lab2[0].p, nullptr, nullptr));
  CU(cudaEventRecord(h->pre.t1, h->stream_labels));
  h->pre.valid = true; h->pre.seed = o->seed; h->pre.perm_offset = o->perm_offset; h->pre.SB = SB; h->pre.J = pl.J;
  return SCDC_OK;
}

// One-shot call, detection mode: copy the matrix gene chunk by gene chunk and launch the group sums of batch 0 for the
// active genes of every chunk right behind its copy (same stream), so the pageable H2D copy hides behind compute. A chunk's
// kernel is launched only once the helper thread has validated that chunk's indices. run_device then skips the group-sum
// launch of batch 0 (h->pre.k2_done).
int upload_matrix_with_early_reduction(scdc_handle* h, const scdc_problem* p, const scdc_options* o,
                                       std::atomic<int>& genes_valid, std::atomic<bool>& matrix_done, const int& matrix_rc) {
  cudaStream_t s = h->stream;
  const int G = h->G, K = h->K;
  const int B = pick_batch(h, o);
  const bool f32 = o->precision == SCDC_FP32_FAST;
  const ScatterPlan pl = plan_scatter(h, B, f32);
  if (pl.J != h->pre.J) {  // cannot happen (same planning inputs); fall back to the plain copy
    CU(upload_big(h->rowidx, p->rowidx, h->nnz, s));
    CU(upload_big(h->values, p->values, h->nnz, s));
    return SCDC_OK;
  }
  CU(h->rowidx.alloc(h->nnz));
  CU(h->values.alloc(h->nnz));
  if (f32) CU(h->values_f.alloc(h->nnz ? h->nnz : 1));
  CU(h->M2[0].ensure((size_t)(B / 32) * G * K * 32));
  // chunks: contiguous gene ranges of about equal nnz
  const int n_chunks = (int)std::max<size_t>(1, std::min<size_t>(16, h->nnz / (size_t)std::max(1, env_int("SCDC_CHUNK_NNZ", 1 << 19))));
  std::vector<int> bounds(1, 0);
  for (int c = 1; c <= n_chunks; ++c) {
    const long long target = (long long)h->nnz * c / n_chunks;
    int g = bounds.back();
    while (g < G && (c == n_chunks || p->colptr[g + 1] <= target)) ++g;
    bounds.push_back(g);
  }
  // active genes of every chunk, in the handle's heavy-first order
  std::vector<int> order((size_t)h->n_active), by_chunk, off(1, 0);
  CU(cudaMemcpyAsync(order.data(), h->gene_order.p, order.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (int c = 0; c < n_chunks; ++c) {
    for (int g : order)
      if (g >= bounds[c] && g < bounds[c + 1]) by_chunk.push_back(g);
    off.push_back((int)by_chunk.size());
  }
  DBuf<int>& d_list = h->early_genes;
  CU(d_list.upload(by_chunk.data(), by_chunk.size(), s));
  if (!h->pre.k2_t0) { CU(cudaEventCreate(&h->pre.k2_t0)); CU(cudaEventCreate(&h->pre.k2_t1)); }
  CU(cudaStreamWaitEvent(s, h->pre.t1, 0));  // labels of super-batch 0
  // The copies go on their own stream: a pageable cudaMemcpyAsync waits for the work queued before it on ITS stream, so on
  // the compute stream every chunk's copy would wait for the previous chunk's kernel.
  cudaStream_t sc = h->stream_score;
  cudaEvent_t copied;
  CU(cudaEventCreateWithFlags(&copied, cudaEventDisableTiming));
  struct EvDel { cudaEvent_t e; ~EvDel() { cudaEventDestroy(e); } } evdel{copied};
  {  // the copy stream must see the allocations made on the compute stream
    CU(cudaEventRecord(copied, s));
    CU(cudaStreamWaitEvent(sc, copied, 0));
  }
  bool first = true;
  for (int c = 0; c < n_chunks; ++c) {
    const size_t e0 = (size_t)p->colptr[bounds[c]], e1 = (size_t)p->colptr[bounds[c + 1]];
    if (e1 > e0) {
      // (staging the chunks through pinned memory with four host threads, g_stager.copy(.., 1 MB, 4), measured 0.8 ms
      // slower per call than the driver's own staging of these 3-5 MB pieces on an idle host)
      CU(cudaMemcpyAsync(h->rowidx.p + e0, p->rowidx + e0, (e1 - e0) * sizeof(int), cudaMemcpyHostToDevice, sc));
      CU(cudaMemcpyAsync(h->values.p + e0, p->values + e0, (e1 - e0) * sizeof(double), cudaMemcpyHostToDevice, sc));
      g_h2d_bytes += (double)(e1 - e0) * 12.0;
    }
    // chunk kernels alternate between two streams so that the tail of one chunk overlaps the head of the next
    cudaStream_t sk = (c & 1) ? h->stream_labels : s;
    CU(cudaEventRecord(copied, sc));
    CU(cudaStreamWaitEvent(sk, copied, 0));
    while (genes_valid.load(std::memory_order_acquire) < bounds[c + 1] && !matrix_done.load(std::memory_order_acquire))
      std::this_thread::yield();
    if (matrix_done.load(std::memory_order_acquire) && matrix_rc != SCDC_OK)
      return fail(SCDC_EINVAL, "matrix validation failed");  // the caller reports the helper thread's message
    if (first) {
      CU(cudaEventRecord(h->pre.k2_t0, s));
      CU(cudaStreamWaitEvent(h->stream_labels, h->pre.k2_t0, 0));  // allocations, labels, gene list
      first = false;
    }
    const int n_g = off[c + 1] - off[c];
    if (f32 && e1 > e0)  // this chunk's fp32 values
      scdc::k_to_float<<<(unsigned)((e1 - e0 + 255) / 256), 256, 0, sk>>>(e1 - e0, h->values.p + e0, h->values_f.p + e0);
    CU(launch_scatter(h, pl, B, h->lab2[0].p, B, h->M2[0].p, sk, d_list.p + off[c], n_g));
  }
  if (first) CU(cudaEventRecord(h->pre.k2_t0, s));
  CU(cudaEventRecord(copied, h->stream_labels));  // join the second kernel stream
  CU(cudaStreamWaitEvent(s, copied, 0));
  CU(cudaEventRecord(h->pre.k2_t1, s));
  h->pre.k2_done = true;
  h->pre.B = B;
  h->pre.k2_launches = n_chunks;
  return SCDC_OK;
}

// One-shot call, differential mode: k_scatter of batch 0 over the whole (already uploaded and validated) matrix, launched
// from scdc_create so that it runs while the host builds and uploads the pass-1 copy and the rows.
int launch_early_scatter(scdc_handle* h, const scdc_options* o) {
  cudaStream_t s = h->stream;
  const int B = pick_batch(h, o);
  const bool f32 = o->precision == SCDC_FP32_FAST;
  const ScatterPlan pl = plan_scatter(h, B, f32);
  if (pl.J != h->pre.J || h->n_active == 0) return SCDC_OK;  // (cannot happen / nothing to reduce): the normal path runs
  if (f32) {
    int rc = ensure_f32_values(h, s, false);
    if (rc) return rc;
  }
  CU(h->M2[0].ensure((size_t)(B / 32) * h->G * h->K * 32));
  if (!h->pre.k2_t0) { CU(cudaEventCreate(&h->pre.k2_t0)); CU(cudaEventCreate(&h->pre.k2_t1)); }
  CU(cudaStreamWaitEvent(s, h->pre.t1, 0));  // labels of super-batch 0
  CU(cudaEventRecord(h->pre.k2_t0, s));
  CU(launch_scatter(h, pl, B, h->lab2[0].p, B, h->M2[0].p, s));
  CU(cudaEventRecord(h->pre.k2_t1, s));
  h->pre.k2_done = true;
  h->pre.B = B;
  h->pre.k2_launches = 1;
  return SCDC_OK;
}

// Pass 1 of one batch (k_pass1_cond). persistent = a small grid whose warps pull (gene, chunk) items from a counter; used
// when pass 1 runs next to the scatter.
cudaError_t launch_pass1_kernel(scdc_handle* h, bool f32, int B, const uint8_t* bits1_b, void* M1b, cudaStream_t sp, int W,
                                bool persistent) {
  const int nWC = B / 256;
  const long long tasks = (long long)h->n_active * nWC;
  if (tasks <= 0) return cudaSuccess;
  unsigned grid1 = (unsigned)((tasks + W - 1) / W);
  int* counter = nullptr;
  if (persistent) {
    // blocks per SM: one next to the fp64 scatter (168 registers per thread at U = 32), two next to the fp32 one
    // (measured, config 5 / config 3: fp64 1895 vs 2206 / 279 vs 286 ms of reduction per step, fp32 1219 vs 1158 / 162 vs 142)
    grid1 = (unsigned)std::min<long long>(grid1, (long long)h->num_sms * std::max(1, env_int("SCDC_P1_BLOCKS", f32 ? 2 : 1)));
    counter = h->p1_counter.p;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), sp);
    if (e != cudaSuccess) return e;
  }
  const int U1 = env_int("SCDC_U1", 4);  // records per pipeline stage
  // Records staged through a shared-memory ring (fewer load-store write-back cycles, which is what pass 1 takes from the
  // scatter next to it): 4 warps at most, and — persistent form — only where the ring (+ 1 KB the system reserves per block)
  // fits into what the scatter's resident blocks leave of the SM's 228 KB: launched first, this kernel would otherwise cost
  // the scatter a block per SM (fp64 K = 60: two 7-warp blocks leave 2 KB)
  // Measured (r2z, reduction per step, uniform loads -> ring): config 3 (fp64, K = 50, scatter at 99 % of the load-store pipe
  // with 16 resident warps) 280.2 -> 237.1 ms; config 5 (fp64, K = 120, 7 warps, latency-bound at 79 %) 1803.9 -> 1804.5;
  // config 5 fp32 (14 warps) 1137.9 -> 1143.3. Default: fp64 with more than 8 resident scatter warps per SM.
  const int staged_env = env_int("SCDC_P1_STAGED", -1);
  bool staged = staged_env != 0 && W <= 4 && persistent;
  if (staged) {
    const ScatterPlan pl = plan_scatter(h, B, f32);
    const size_t blocks1 = (size_t)std::max(1, env_int("SCDC_P1_BLOCKS", f32 ? 2 : 1));
    const size_t left = 233472 - std::min<size_t>(233472, (size_t)pl.blocks_per_sm * (pl.smem + 1024));
    staged = left >= blocks1 * ((f32 ? 2048 : 4096) + 1024);
    if (staged_env < 0) staged = staged && !f32 && pl.blocks_per_sm * pl.W > 8;
  }
  auto go = [&](auto kern, auto* rec, auto* M1t) {
    kern<<<grid1, W * 32, 0, sp>>>(h->G, h->n_active, h->T, nWC, h->segptr.p, rec, bits1_b, B / 8, h->ngroup.p, h->gene_order.p,
                                   M1t, counter);
  };
  if (f32) {
    if (U1 == 2) go(scdc::k_pass1_cond<float, 2, false>, h->rec_sf.p, static_cast<float*>(M1b));
    else if (staged) go(scdc::k_pass1_cond<float, 4, true>, h->rec_sf.p, static_cast<float*>(M1b));
    else go(scdc::k_pass1_cond<float, 4, false>, h->rec_sf.p, static_cast<float*>(M1b));
  } else {
    if (U1 == 2) go(scdc::k_pass1_cond<double, 2, false>, h->rec_s.p, static_cast<double*>(M1b));
    else if (staged) go(scdc::k_pass1_cond<double, 4, true>, h->rec_s.p, static_cast<double*>(M1b));
    else go(scdc::k_pass1_cond<double, 4, false>, h->rec_s.p, static_cast<double*>(M1b));
  }
  return cudaGetLastError();
}

// One-shot call, differential mode: pass 1 of batch 0 right behind the device-side sort, on the same high-priority stream
// `sp` (the pass-1 copy of the matrix is its input), so that it runs next to batch 0's scatter instead of after it.
// run_device then skips it and waits for h->pre.p1_ev.
int launch_early_pass1(scdc_handle* h, const scdc_options* o, cudaStream_t sp) {
  if (!h->pre.k2_done || h->C != 2 || env_int("SCDC_NO_EARLY_PASS1", 0) || !env_int("SCDC_OVERLAP_PASS1", 1) ||
      !env_int("SCDC_P1_PERSISTENT", 1))
    return SCDC_OK;
  const bool f32 = o->precision == SCDC_FP32_FAST;
  const int B = h->pre.B;
  AllocScope scope(sp);
  CU(h->M1[0].ensure((size_t)(B / 32) * h->G * h->K * 32));
  CU(h->p1_counter.ensure(1));
  if (f32) {
    int rc = ensure_f32_values(h, sp, true);
    if (rc) return rc;
  }
  if (!h->pre.p1_ev) CU(cudaEventCreateWithFlags(&h->pre.p1_ev, cudaEventDisableTiming));
  CU(launch_pass1_kernel(h, f32, B, reinterpret_cast<const uint8_t*>(h->bits1[0].p), h->M1[0].p, sp, env_int("SCDC_W1", 4), true));
  CU(cudaEventRecord(h->pre.p1_ev, sp));
  h->pre.p1_done = true;
  return SCDC_OK;
}

// Runs replicates [o->perm_offset, o->perm_offset + o->iterations) on the handle's device; leaves uint32 counters in
// h->counts (column-major [R][ncols]). Uploaded labels (if any) are indexed from 0 = first replicate of this call.
int run_device(scdc_handle* h, const scdc_options* o, double* dist_host, scdc_stats* st, cudaStream_t s) {
  CU(cudaSetDevice(h->device));
  AllocScope alloc_scope(h->stream);
  const int N = h->N, G = h->G, T = h->T, C = h->C, K = h->K, R = h->R;
  const int B = pick_batch(h, o);
  const int Bw = B / 32;
  const bool f32 = o->precision == SCDC_FP32_FAST;
  const double rel_tol = (o->rel_tol > 0.0) ? o->rel_tol : 1e-5;
  const ScatterPlan pl = plan_scatter(h, B, f32);
  const bool uploaded = o->perm_ct != nullptr;
  // Device-drawn labels: K1 is one thread per replicate and sequential over the cells, so its run time hardly depends on
  // how many replicates it draws. It therefore draws a whole SUPER-batch (SB replicates, as many as the label memory cap
  // allows — normally all of them) in one launch; the reduction then walks it in batches of B. With several
  // super-batches the label buffers are double-buffered and K1 runs one super-batch ahead on a second stream.
  const int64_t iters_pad = ((o->iterations + B - 1) / B) * B;
  const int64_t SB = uploaded ? B : pick_super_batch(h, o, B);
  const bool pre_ok = !uploaded && h->pre.valid && h->pre.seed == o->seed && h->pre.perm_offset == o->perm_offset &&
                      h->pre.SB == SB && h->pre.J == pl.J;
  const bool early_k2 = pre_ok && h->pre.k2_done && h->pre.B == B && s == h->stream;
  const bool early_p1 = early_k2 && h->pre.p1_done;
  h->pre.valid = false;  // consumed (or stale)
  h->pre.k2_done = false;
  h->pre.p1_done = false;
  const int SBw = (int)(SB / 32);
  const bool pipelined = !uploaded && SB < iters_pad;
  const int nbuf = pipelined ? 2 : 1;
  cudaStream_t sL = pipelined ? h->stream_labels : s;
  const int ncols_d = (C == 2) ? 3 : 1;
  const size_t msz = (size_t)Bw * G * K * 32;
  for (int i = 0; i < nbuf; ++i) {
    CU(h->lab2[i].ensure((size_t)N * SB));
    if (C == 2) CU(h->bits1[i].ensure((size_t)N * SBw));
  }
  // K3 (issue-bound) of batch b runs on its own stream next to K2 (shared-memory-pipe-bound) of batch b + 1
  // (measured: no gain on configs 2/3 — both kernels want the same SM resources — so it is opt-in: SCDC_OVERLAP_SCORE=1)
  const bool overlap_score = !uploaded && !dist_host && env_int("SCDC_OVERLAP_SCORE", 0) && o->iterations > B;
  const int nm = overlap_score ? 2 : 1;
  cudaStream_t sS = overlap_score ? h->stream_score : s;
  for (int i = 0; i < nm; ++i) {
    if (C == 2) CU(h->M1[i].ensure(msz));
    CU(h->M2[i].ensure(msz));
  }
  CU(h->counts.ensure((size_t)R * h->ncols + 1));
  if (C == 2) CU(h->p1_counter.ensure(1));
  if (f32) {
    CU(h->band.ensure((size_t)R * h->ncols + 1));
    int rc = ensure_f32_values(h, h->stream, false);
    if (!rc && C == 2) rc = ensure_f32_values(h, h->stream, true);
    if (rc) return rc;
  }
  if (uploaded) {
    CU(h->up_ct.ensure((size_t)B * N));
    if (C == 2) CU(h->up_cond.ensure((size_t)B * N));
  }
  if (dist_host) CU(h->dist.ensure((size_t)B * ncols_d * R));
  if (s != h->stream) {  // buffers are (re)allocated in the order of the handle's stream: the caller's stream waits for it
    cudaEvent_t dep;
    CU(cudaEventCreateWithFlags(&dep, cudaEventDisableTiming));
    cudaError_t e1 = cudaEventRecord(dep, h->stream), e2 = cudaStreamWaitEvent(s, dep, 0);
    cudaEventDestroy(dep);
    CU(e1); CU(e2);
  }
  CU(cudaMemsetAsync(h->counts.p, 0, ((size_t)R * h->ncols + 1) * sizeof(uint32_t), s));
  if (f32) CU(cudaMemsetAsync(h->band.p, 0, ((size_t)R * h->ncols + 1) * sizeof(uint32_t), s));
  if (uploaded) CU(cudaMemsetAsync(h->flag.p, 0, sizeof(int), s));
  std::vector<cudaEvent_t> evs, evl;  // timing events on the compute stream (3 per batch) / label stream (2 per batch)
  auto mark = [&](std::vector<cudaEvent_t>& v, cudaStream_t st_) -> cudaError_t {
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreate(&e);
    if (rc != cudaSuccess) return rc;
    v.push_back(e);
    return cudaEventRecord(e, st_);
  };
  struct EvGuard {
    std::vector<cudaEvent_t>& v;
    ~EvGuard() { for (auto e : v) cudaEventDestroy(e); }
  } guard{evs}, guard2{evl};
  cudaEvent_t ready[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  struct SyncGuard {
    cudaEvent_t* a; cudaEvent_t* b;
    ~SyncGuard() { for (int i = 0; i < 2; ++i) { if (a[i]) cudaEventDestroy(a[i]); if (b[i]) cudaEventDestroy(b[i]); } }
  } guard3{ready, freed};
  cudaEvent_t m_ready[2] = {nullptr, nullptr}, m_free[2] = {nullptr, nullptr};
  SyncGuard guard4{m_ready, m_free};
  bool m_free_set[2] = {false, false};
  if (overlap_score) {
    for (int i = 0; i < 2; ++i) {
      CU(cudaEventCreateWithFlags(&m_ready[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&m_free[i], cudaEventDisableTiming));
    }
    // the score stream must not start before earlier work on the compute stream (counter memset)
    CU(cudaEventRecord(m_ready[0], s));
    CU(cudaStreamWaitEvent(sS, m_ready[0], 0));
  }
  cudaEvent_t ov_ev[2] = {nullptr, nullptr}, ov_none[2] = {nullptr, nullptr};  // fork / join of the pass-1 overlap
  SyncGuard guard6{ov_ev, ov_none};
  if (C == 2) {
    CU(cudaEventCreateWithFlags(&ov_ev[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ov_ev[1], cudaEventDisableTiming));
  }
  cudaEvent_t& ov_fork = ov_ev[0];
  cudaEvent_t& ov_join = ov_ev[1];
  std::vector<cudaEvent_t> evk;  // timing events on the score stream when it is a separate stream
  EvGuard guard5{evk};
  bool freed_set[2] = {false, false};
  if (pipelined) {
    for (int i = 0; i < 2; ++i) {
      CU(cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming));
    }
  }
  int64_t launches = 0, reduce_launches = 0;
  double reduce_bytes = 0, reduce_wavefronts = 0;
  const double sv = f32 ? 4.0 : 8.0;
  CU(mark(evs, s));  // evs[0]: start of the whole loop on the compute stream
  auto issue_labels = [&](int64_t sb_start, int buf) -> int {
    if (pipelined && freed_set[buf]) CU(cudaStreamWaitEvent(sL, freed[buf], 0));
    CU(mark(evl, sL));
    CU(launch_draw(h, sL, o->seed, o->perm_offset + sb_start, (int)SB, B, pl.J, h->bits1[buf].p, h->lab2[buf].p, nullptr, nullptr));
    CU(mark(evl, sL));
    if (pipelined) CU(cudaEventRecord(ready[buf], sL));
    ++launches;
    return SCDC_OK;
  };
  if (pipelined) {
    // the label stream must not run ahead of work already queued on the compute stream (e.g. a previous run's readers)
    cudaEvent_t start;
    CU(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    cudaError_t e1 = cudaEventRecord(start, s), e2 = cudaStreamWaitEvent(sL, start, 0);
    cudaEventDestroy(start);
    CU(e1); CU(e2);
    if (!pre_ok) {
      int rc = issue_labels(0, 0);
      if (rc) return rc;
    }
  }
  // Interruptible runs (scdc_options.interrupt, or the abort flag of a multi-GPU call): at most two batches are in flight;
  // while the host waits for the older one it polls the stop source, and a stop ends the call after the enqueued work.
  const bool can_stop = o->interrupt != nullptr || h->abort != nullptr;
  cudaEvent_t batch_ev[2] = {nullptr, nullptr}, batch_none[2] = {nullptr, nullptr};
  SyncGuard guard7{batch_ev, batch_none};
  if (can_stop)
    for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&batch_ev[i], cudaEventDisableTiming));
  double last_poll = now_ms();
  auto stop_requested = [&]() -> bool {
    if (h->abort && h->abort->load(std::memory_order_relaxed)) return true;
    if (!o->interrupt) return false;
    const double t = now_ms();
    if (t - last_poll < 20.0) return false;  // the callback may be expensive (R_ToplevelExec): at most 50 polls / s
    last_poll = t;
    return o->interrupt(o->interrupt_user) != 0;
  };
  auto wait_polling = [&](cudaEvent_t ev) -> int {  // 0: event reached, 1: stop requested, < 0: CUDA error
    for (;;) {
      const cudaError_t q = ev ? cudaEventQuery(ev) : cudaStreamQuery(s);
      if (q == cudaSuccess) return 0;
      if (q != cudaErrorNotReady) { fail(SCDC_ECUDA, "cudaEventQuery failed: %s", cudaGetErrorString(q)); return -1; }
      if (stop_requested()) return 1;
      std::this_thread::sleep_for(std::chrono::microseconds(500));
    }
  };
  auto interrupted = [&](int64_t done) -> int {
    cudaStreamSynchronize(s);
    if (pipelined) cudaStreamSynchronize(sL);
    cudaStreamSynchronize(h->stream_score);
    return fail(SCDC_EINTERRUPT, "interrupted by the caller after %lld of %lld replicates were enqueued", (long long)done,
                (long long)o->iterations);
  };
  int64_t batch_no = 0;
  for (int64_t done = 0; done < o->iterations; done += B, ++batch_no) {
    if (can_stop && batch_no >= 2) {
      const int w = wait_polling(batch_ev[batch_no & 1]);  // batch_no - 2 has finished: one batch stays queued behind the running one
      if (w < 0) return SCDC_ECUDA;
      if (w > 0) return interrupted(done);
    }
    if (can_stop && batch_no > 0 && stop_requested()) return interrupted(done);  // (rate-limited) also when the GPU keeps up
    const int nb = (int)std::min<int64_t>(B, o->iterations - done);
    const int64_t sb_idx = done / SB, in_sb = done % SB;  // position inside the current super-batch
    const int buf = pipelined ? (int)(sb_idx & 1) : 0;
    if (uploaded) {
      CU(mark(evl, s));
      CU(cudaMemcpyAsync(h->up_ct.p, o->perm_ct + (size_t)done * N, (size_t)nb * N, cudaMemcpyHostToDevice, s));
      if (C == 2)
        CU(cudaMemcpyAsync(h->up_cond.p, o->perm_cond + (size_t)done * N, (size_t)nb * N, cudaMemcpyHostToDevice, s));
      g_h2d_bytes += (double)nb * N * C;
      scdc::k_check_uploaded<<<nb, 256, 2 * K * sizeof(int), s>>>(N, T, C, h->ct.p, h->up_cond.p, h->up_ct.p, h->expect.p,
                                                                  h->flag.p);
      dim3 grid((N + 31) / 32, B / 32), block(32, 32);
      scdc::k_pack_uploaded<<<grid, block, 0, s>>>(N, T, C, nb, B, pl.J, h->ct.p, h->cond.p, h->up_cond.p, h->up_ct.p,
                                                   h->bits1[0].p, h->lab2[0].p);
      CU(cudaGetLastError());
      if (C == 2) CU(repack_condition_bits(h, s, h->bits1[0].p, B, B));
      CU(mark(evl, s));
      launches += 2;
    } else if (in_sb == 0) {
      if (pre_ok && done == 0) {
        CU(cudaStreamWaitEvent(s, h->pre.t1, 0));  // super-batch 0 was drawn ahead by prelaunch_labels (buffer 0)
        if (pipelined && SB < o->iterations) {
          int rc = issue_labels(SB, 1);
          if (rc) return rc;
        }
      } else if (!pipelined) {
        int rc = issue_labels(done, 0);
        if (rc) return rc;
      } else {
        if (done + SB < o->iterations) {  // next super-batch's labels, enqueued first so that they overlap this one's compute
          int rc = issue_labels(done + SB, buf ^ 1);
          if (rc) return rc;
        }
        CU(cudaStreamWaitEvent(s, ready[buf], 0));
      }
    }
    const uint8_t* lab2_b = h->lab2[buf].p + (size_t)in_sb * N;                      // block [in_sb / B]: [cell][B] bytes
    const uint8_t* bits1_b =   // [cell][B / 8] bytes (k_bits_transpose8)
        (C == 2) ? reinterpret_cast<const uint8_t*>(h->bits1[buf].p + (size_t)(in_sb / 32) * N) : nullptr;
    const int mb = overlap_score ? (int)((done / B) & 1) : 0;  // group-means buffer of this batch
    void* M1b = h->M1[mb].p;   // element type: double, or float in SCDC_FP32_FAST (half of the buffer used)
    void* M2b = h->M2[mb].p;
    if (overlap_score && m_free_set[mb]) CU(cudaStreamWaitEvent(s, m_free[mb], 0));  // K3 of batch b-2 has read it
    CU(mark(evs, s));
    // Pass 1 (issue-slot bound, little shared memory) can run NEXT TO the pass-2 scatter (shared-memory-pipe bound): it is
    // then launched after the scatter on a second stream, in small blocks that fit into the registers / shared memory the
    // scatter leaves free on every SM, and joined before K3.
    const int ov1 = (C == 2) ? env_int("SCDC_OVERLAP_PASS1", 1) : 0;
    auto launch_pass1 = [&](cudaStream_t sp, int W, bool persistent) -> cudaError_t {
      return launch_pass1_kernel(h, f32, B, bits1_b, M1b, sp, W, persistent);
    };
    if (C == 2 && !ov1) {
      CU(launch_pass1(s, 4, false));
      ++launches; ++reduce_launches;
    }
    if (C == 2 && early_p1 && done == 0) {
      // pass 1 of batch 0 was launched from scdc_create next to the early scatter (launch_early_pass1)
      CU(cudaStreamWaitEvent(s, h->pre.p1_ev, 0));
      ++launches; ++reduce_launches;
    } else if (C == 2 && ov1) {
      // Pass 1 (FP64 / issue-slot bound, registers only) runs NEXT TO the pass-2 scatter (shared-memory-pipe bound): its
      // persistent grid of a block per SM is launched first, on a high-priority stream, so that its blocks are resident on
      // every SM before the scatter's blocks fill the rest; joined before K3.
      cudaStream_t sp1 = pipelined ? h->stream_score : h->stream_labels;  // (the label stream draws ahead when pipelined)
      CU(cudaEventRecord(ov_fork, s));  // labels, means buffers and everything before are ready
      CU(cudaStreamWaitEvent(sp1, ov_fork, 0));
      CU(launch_pass1(sp1, env_int("SCDC_W1", 4), env_int("SCDC_P1_PERSISTENT", 1) != 0));
      CU(cudaEventRecord(ov_join, sp1));
      ++launches; ++reduce_launches;
    }
    if (early_k2 && done == 0) {
      launches += h->pre.k2_launches;  // batch 0 was reduced behind the matrix upload (upload_matrix_with_early_reduction)
      reduce_launches += h->pre.k2_launches;
    } else {
      CU(launch_scatter(h, pl, B, lab2_b, B, M2b, s));
      ++launches; ++reduce_launches;
    }
    if (C == 2 && ov1 && !(early_p1 && done == 0)) CU(cudaStreamWaitEvent(s, ov_join, 0));
    if (pipelined && (in_sb + B >= SB || done + B >= o->iterations)) { CU(cudaEventRecord(freed[buf], s)); freed_set[buf] = true; }
    // algorithmic bytes of this batch's reduction launches: matrix once per pass-kernel, labels, means out
    reduce_bytes += (double)C * ((double)h->nnz_active * (sv + 4.0) + 4.0 * (h->n_active + 1)) +
                    (double)B * N * (C == 2 ? 1.125 : 1.0) + (double)C * B * K * h->n_active * sv;
    // load-store data-pipe wavefronts of the scatter by design, per (entry x 32 J replicates): entry broadcast (fp64 LDS.128 =
    // 2, fp32 LDS.64 = 1), the write-back of the warp's label load (32 J bytes = 1), J x (LDS + STS) (fp64 4 J, fp32 2 J) and
    // the staging store (4/32 resp. 2/32). Pass 1 (k_pass1_cond) uses no shared memory and is not counted.
    const double wf_b = f32 ? 1.0 : 2.0, wf_rmw = f32 ? 2.0 : 4.0, wf_st = f32 ? 0.0625 : 0.125;
    reduce_wavefronts += (double)h->nnz_active * (B / (32.0 * pl.J)) * (wf_b + 1.0 + wf_rmw * pl.J + wf_st);
    CU(mark(evs, s));
    if (overlap_score) {
      CU(cudaEventRecord(m_ready[mb], s));
      CU(cudaStreamWaitEvent(sS, m_ready[mb], 0));
      CU(mark(evk, sS));
    }
    if (R > 0 && !dist_host && !env_int("SCDC_SCORE_V1", 0)) {
      const int Tp = T | 1;
      // shared memory of the production score kernels, in doubles per replicate of a tile: RAW tiles double-buffered
      // (differential 2 x 2 sides x 4 values, detection 2 x 2 sides x 1), + the subunit-difference tiles when the subunit
      // columns are not factorised
      const bool fact = C == 2 && h->fact_ok;
      const int n_arr = (C == 2) ? (f32 ? 8 : 16) + (fact ? 0 : h->nS * (f32 ? 2 : 1)) : (f32 ? 2 : 4);
      const size_t tail = fact ? (size_t)T * h->nS * 8 + (size_t)T * h->nS * 4 * (f32 ? 2 : 1) + 8 : 0;
      int P = 32;
      const size_t smem_target = (size_t)env_int("SCDC_SCORE_SMEM_KB", 100) * 1024;
      while (P > 4 && (size_t)n_arr * P * Tp * 8 + tail > smem_target) P >>= 1;
      P = env_int("SCDC_SCORE_P", P);
      const size_t smem = (size_t)n_arr * P * Tp * 8 + tail;
      if (C == 2) {
        scdc::ScoreDiffArgs D;
        D.G = G; D.T = T; D.Tp = Tp; D.K = K; D.nL = h->nL; D.nR = h->nR; D.score_type = o->score_type;
        D.n_valid = nb; D.P = P; D.R = R;
        D.task_lri = h->task_lri.p; D.task_beg = h->task_beg.p; D.task_end = h->task_end.p; D.srow = h->srow.p;
        D.sub_gene = h->sub_gene.p; D.row_emit = h->row_emit.p; D.row_recv = h->row_recv.p; D.obs = h->obs.p;
        D.thr = h->thr.p; D.obs_sub = h->obs_sub.p; D.M1 = M1b; D.M2 = M2b; D.counts = h->counts.p;
        D.band = h->band.p; D.rel_tol = rel_tol;
        auto go = [&](auto kern, size_t smem_k) -> cudaError_t {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k);
          if (e != cudaSuccess) return e;
          kern<<<h->n_tasks, h->score_threads, smem_k, sS>>>(D);
          return cudaGetLastError();
        };
        if (fact && env_int("SCDC_SCORE_BULK", 0)) {
          // OPT-IN (SCDC_SCORE_BULK=1): tiles in the means' natural layout, filled by cp.async.bulk + mbarrier
          // (k_score_diff_bulk): rows padded to P + 2 (fp64) / P + 4 (fp32) elements, two buffers of 8 planes x T rows.
          // Correct (the whole GPU suite passes with it) but measured SLOWER than the cp.async kernel: config 5 score
          // 436 vs 307 ms per 12.8 k replicates, config 3 60.7 vs 50.2 ms per 10 k (profiles/NOTES_r2.md, r2j): a tile is
          // 8 T pieces of only P x 8 = 64 bytes, and 480 tiny bulk copies per tile cost more than they save.
          const size_t es = f32 ? 4 : 8;
          const int pad = f32 ? 4 : 2;
          int Pb = 32;
          while (Pb > 4 && 2 * 8 * (size_t)T * (Pb + pad) * es + tail > smem_target) Pb >>= 1;
          Pb = env_int("SCDC_SCORE_P", Pb);
          D.P = Pb;
          const size_t smem_b = ((2 * 8 * (size_t)T * (Pb + pad) * es + 7) / 8) * 8 + tail;
          if (f32) CU(go(scdc::k_score_diff_bulk<kScoreRPT, float, true>, smem_b));
          else CU(go(scdc::k_score_diff_bulk<kScoreRPT, double, false>, smem_b));
        } else if (f32) {
          CU(fact ? go(scdc::k_score_diff<kScoreRPT, true, float, true>, smem) : go(scdc::k_score_diff<kScoreRPT, false, float, true>, smem));
        } else {
          CU(fact ? go(scdc::k_score_diff<kScoreRPT, true, double, false>, smem) : go(scdc::k_score_diff<kScoreRPT, false, double, false>, smem));
        }
      } else {
        scdc::ScoreDetArgs A;
        A.G = G; A.T = T; A.Tp = Tp; A.nL = h->nL; A.nS = h->nS; A.score_type = o->score_type;
        A.n_valid = nb; A.P = P; A.R = R;
        A.task_lri = h->task_lri.p; A.task_beg = h->task_beg.p; A.task_end = h->task_end.p; A.srow = h->srow.p;
        A.sub_gene = h->sub_gene.p; A.row_emit = h->row_emit.p; A.row_recv = h->row_recv.p; A.obs = h->obs.p;
        A.thr = h->thr.p; A.M2 = M2b; A.counts = h->counts.p; A.band = h->band.p; A.rel_tol = rel_tol;
        auto go = [&](auto kern) -> cudaError_t {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
          kern<<<h->n_tasks, h->score_threads, smem, sS>>>(A);
          return cudaGetLastError();
        };
        CU(f32 ? go(scdc::k_score_det<kScoreRPT, float, true>) : go(scdc::k_score_det<kScoreRPT, double, false>));
      }
      ++launches;
    } else if (R > 0) {
      scdc::ScoreArgs A;
      A.R = R; A.G = G; A.T = T; A.C = C; A.nL = h->nL; A.nR = h->nR; A.ncols = h->ncols; A.score_type = o->score_type;
      A.nChunks = (nb + 31) / 32; A.n_valid = nb;
      A.sub_gene = h->sub_gene.p; A.row_lri = h->row_lri.p; A.row_emit = h->row_emit.p; A.row_recv = h->row_recv.p;
      A.obs = h->obs.p; A.M1 = M1b; A.M2 = M2b; A.counts = h->counts.p; A.band = h->band.p; A.rel_tol = rel_tol;
      A.dist = dist_host ? h->dist.p : nullptr;
      const int blocks = std::max(1, std::min((R + 7) / 8, h->num_sms * 8));
      if (f32) scdc::k_score_count<float, true><<<blocks, 256, 0, sS>>>(A);
      else scdc::k_score_count<double, false><<<blocks, 256, 0, sS>>>(A);
      CU(cudaGetLastError());
      ++launches;
      if (dist_host) {
        CU(cudaMemcpyAsync(dist_host + (size_t)done * ncols_d * R, h->dist.p, (size_t)nb * ncols_d * R * sizeof(double),
                           cudaMemcpyDeviceToHost, sS));
        g_d2h_bytes += (double)nb * ncols_d * R * sizeof(double);
      }
    }
    if (overlap_score) {
      CU(mark(evk, sS));
      CU(cudaEventRecord(m_free[mb], sS));
      m_free_set[mb] = true;
      if (done + B >= o->iterations) CU(cudaStreamWaitEvent(s, m_free[mb], 0));  // the compute stream ends after the last K3
    }
    CU(mark(evs, s));
    if (uploaded || dist_host) CU(cudaStreamSynchronize(s));  // host source / destination buffers are reused per batch
    if (can_stop) CU(cudaEventRecord(batch_ev[batch_no & 1], s));
  }
  if (can_stop) {
    const int w = wait_polling(nullptr);
    if (w < 0) return SCDC_ECUDA;
    if (w > 0) return interrupted(o->iterations);
  }
  CU(cudaStreamSynchronize(s));
  if (pipelined) CU(cudaStreamSynchronize(sL));
  if (h->keyfail.p) {
    int kf = 0;
    CU(cudaMemcpy(&kf, h->keyfail.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (kf) return fail(SCDC_ECUDA, "internal: the random-key label generator overflowed a target bin (probability < 1e-30)");
  }
  if (uploaded) {
    int flag = 0;
    CU(cudaMemcpy(&flag, h->flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag == 2) return fail(SCDC_EINVAL, "uploaded labels contain a code outside [0, n_conditions) x [0, n_celltypes)");
    if (flag == 1)
      return fail(SCDC_EINVAL,
                  "uploaded labels do not preserve the number of cells per (cell type, condition): not a "
                  "run_stat_iteration shuffle (reference R/utils_permutation.R:536-569)");
  }
  if (st) {
    for (size_t i = 1; i + 2 < evs.size(); i += 3) {  // evs[0] = loop start; then per batch: ready, reduced, scored
      float b = 0, c = 0;
      cudaEventElapsedTime(&b, evs[i], evs[i + 1]);
      cudaEventElapsedTime(&c, evs[i + 1], evs[i + 2]);
      st->ms_reduce += b;
      if (!overlap_score) st->ms_score += c;
    }
    for (size_t i = 0; i + 1 < evk.size(); i += 2) {  // score stream: overlaps the next batch's ms_reduce
      float c = 0;
      cudaEventElapsedTime(&c, evk[i], evk[i + 1]);
      st->ms_score += c;
    }
    if (pre_ok) {
      float a = 0;
      cudaEventElapsedTime(&a, h->pre.t0, h->pre.t1);
      st->ms_labels += a;  // drawn during scdc_create, overlapped with the matrix upload
    }
    if (early_k2) {
      float a = 0;
      cudaEventElapsedTime(&a, h->pre.k2_t0, h->pre.k2_t1);
      st->ms_reduce += a;  // includes the interleaved H2D chunks it hides
    }
    for (size_t i = 0; i + 1 < evl.size(); i += 2) {
      float a = 0;
      cudaEventElapsedTime(&a, evl[i], evl[i + 1]);
      st->ms_labels += a;  // on the label stream: overlaps ms_reduce / ms_score when pipelined
    }
    if (evs.size() >= 2) {
      float d = 0;
      cudaEventElapsedTime(&d, evs.front(), evs.back());
      st->ms_device += d;
    }
    st->reduce_bytes += reduce_bytes;
    st->reduce_wavefronts += reduce_wavefronts;
    st->reduce_launches += reduce_launches;
    st->kernel_launches += launches;
    st->batch_size = B;
  }
  return SCDC_OK;
}

// uint32 device counters -> the caller's uint64 array. The (row, column) entries whose observed value is NA are first
// marked on the device (k_mark_na writes 0xFFFFFFFF, no counter can reach it) and become SCDC_COUNT_NA here.
int counts_to_host(const scdc_handle* h, uint32_t* dev, uint64_t* out, cudaStream_t s) {
  const size_t n = (size_t)h->R * h->ncols;
  if (!n) return SCDC_OK;
  PhaseTimer pt;
  scdc::k_mark_na<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, h->obs.p, dev);
  CU(cudaGetLastError());
  g_d2h_bytes += (double)n * sizeof(uint32_t);
  const int n_thr = (n > (1u << 20)) ? 4 : 1;
  // widen [lo, hi) of a block of uint32 counters that starts at element `first` of the table
  auto widen = [&](const uint32_t* raw, size_t first, size_t count) {
    auto body = [&](int w) {
      const size_t lo = count * w / n_thr, hi = count * (w + 1) / n_thr;
      for (size_t i = lo; i < hi; ++i) out[first + i] = (raw[i] == 0xFFFFFFFFu) ? SCDC_COUNT_NA : (uint64_t)raw[i];
    };
    std::vector<std::thread> th;
    try {
      for (int w = 1; w < n_thr; ++w) th.emplace_back(body, w);
    } catch (...) {
      for (int w = (int)th.size() + 1; w < n_thr; ++w) body(w);
    }
    body(0);
    for (auto& t : th) t.join();
  };
  // pinned staging chunks: the DMA of chunk i + 1 runs while chunk i is widened into the caller's array
  cudaError_t e = (n * sizeof(uint32_t) >= (4u << 20) && !getenv("SCDC_NO_STAGER"))
                      ? g_stager.fetch(dev, n * sizeof(uint32_t), s,
                                       [&](const char* chunk, size_t off, size_t bytes) {
                                         widen(reinterpret_cast<const uint32_t*>(chunk), off / sizeof(uint32_t), bytes / sizeof(uint32_t));
                                       })
                      : cudaErrorNotSupported;
  if (e == cudaErrorNotSupported) {  // small tables, or no pinned memory: one pageable copy
    std::unique_ptr<uint32_t[]> raw32(new (std::nothrow) uint32_t[n]);
    if (!raw32) return fail(SCDC_ENOMEM, "host allocation failed");
    CU(cudaMemcpyAsync(raw32.get(), dev, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    widen(raw32.get(), 0, n);
  } else {
    CU(e);
  }
  pt.lap("counts: NA marks + D2H + widen");
  return SCDC_OK;
}

// ---- NCCL, loaded lazily (only the multi-GPU broadcast / merge use it; without it peer copies do the same job) ------
struct NcclApi {
  void* lib = nullptr;
  bool tried = false;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (tried) return lib != nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    CommInitAll = (int (*)(void**, int, const int*))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclBroadcast");
    GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
    GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!(CommInitAll && CommDestroy && AllReduce && Broadcast && GroupStart && GroupEnd)) lib = nullptr;
    return lib != nullptr;
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;            // one multi-GPU call at a time uses the cached communicators
std::vector<int> g_comm_devs;    // device set of the cached communicators
std::vector<void*> g_comms;
constexpr int kNcclUint8 = 1, kNcclUint32 = 3;  // ncclUint8, ncclUint32
constexpr int kNcclSum = 0;                     // ncclSum

// communicators for `devs`, created once per device set and kept for the life of the process (ncclCommInitAll ~1.5 s).
// false: NCCL is unavailable (not loadable, SCDC_NO_NCCL=1, or the init failed) and the peer-copy path is used.
bool nccl_comms_for(const std::vector<int>& devs) {
  if (env_int("SCDC_NO_NCCL", 0) || !g_nccl.load()) return false;
  if (g_comm_devs == devs && !g_comms.empty()) return true;
  for (void* c : g_comms) if (c) g_nccl.CommDestroy(c);
  g_comms.assign(devs.size(), nullptr);
  g_comm_devs.clear();
  if (g_nccl.CommInitAll(g_comms.data(), (int)devs.size(), devs.data()) != 0) {
    g_comms.clear();
    return false;
  }
  g_comm_devs = devs;
  return true;
}

// One device buffer of the root handle and its (already allocated) twins on the other devices
struct Mirror {
  const void* src;
  std::vector<void*> dst;  // [W], dst[0] == src
  size_t bytes;
};

// Copies every Mirror from device hs[0] to the devices of hs[1..] on the given per-device streams: ncclBroadcast over
// NVLink when NCCL is available, else cudaMemcpyPeerAsync from the root.
int broadcast_mirrors(const std::vector<scdc_handle*>& hs, const std::vector<Mirror>& ms, bool use_nccl,
                      const std::vector<cudaStream_t>& streams) {
  const int W = (int)hs.size();
  for (const Mirror& m : ms) {
    if (!m.bytes) continue;
    if (use_nccl) {
      int nrc = g_nccl.GroupStart();
      for (int r = 0; r < W && nrc == 0; ++r)
        nrc = g_nccl.Broadcast(m.src, m.dst[r], m.bytes, kNcclUint8, 0, g_comms[r], streams[r]);
      const int erc = g_nccl.GroupEnd();
      if (nrc != 0 || erc != 0)
        return fail(SCDC_ENCCL, "ncclBroadcast failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc ? nrc : erc) : "?");
    } else {
      for (int r = 1; r < W; ++r)
        CU(cudaMemcpyPeerAsync(m.dst[r], hs[r]->device, m.src, hs[0]->device, m.bytes, streams[r]));
    }
    g_p2p_bytes += (double)m.bytes * (W - 1);
  }
  return SCDC_OK;
}

// ---- the twins of the first device's handle on the other devices of a multi-GPU call -----------------------------------
// Nothing big is read from the caller's host arrays twice: the problem travels device to device over NVLink by the copy
// engines (cudaMemcpyPeerAsync, every twin pulls from the first device; SCDC_BCAST_NCCL=1: ncclBroadcast) on a dedicated
// high-priority stream per device, because the devices' compute streams are busy with their first batch by then. The copy
// engines are the default because they take no SMs: the ncclBroadcast kernels held SMs of every device during the first
// batch and cost the 2-GPU call 0-40 ms (r2u). Three phases, so that the twins lag the first device as little as possible:
//   phase 0 (side thread, while the calling thread uploads the matrix): contexts, the cell labels (N bytes, from the host)
//            and each twin's own label draw;
//   phase A (from create_impl's hook, as soon as the matrix is on the first device and its first scatter is running, while
//            its host thread still has the rows to do): CSC matrix, LRI table, active genes; every twin then starts the
//            scatter of its first batch and builds its pass-1 copy of the matrix locally (k_sort_by_celltype);
//   phase B (after create): tested rows, observed values, thresholds, K3 work list, factorisation table.
struct MultiGpuCall : CreateHook {
  std::vector<scdc_handle*> hs;
  std::vector<int> devs;
  std::vector<scdc_options> os;
  const scdc_problem* prob = nullptr;
  bool use_nccl = false, phase_a_done = false;
  // SCDC_BCAST_NCCL=1: ncclBroadcast; 0: the devices' copy engines (cudaMemcpyPeerAsync), which take no SMs away from the
  // first batch's kernels that are running on every device by then
  bool bcast_nccl = env_int("SCDC_BCAST_NCCL", 0) != 0;
  std::vector<cudaStream_t> bs;    // broadcast stream of every device (high priority; the compute streams are busy)

  int open_broadcast_streams() {
    bs.assign(devs.size(), nullptr);
    for (size_t r = 0; r < devs.size(); ++r) {
      CU(cudaSetDevice(devs[r]));
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&bs[r], cudaStreamNonBlocking, hi));
    }
    return SCDC_OK;
  }
  void close_broadcast_streams() {
    for (size_t r = 0; r < bs.size(); ++r)
      if (bs[r]) { cudaSetDevice(devs[r]); cudaStreamSynchronize(bs[r]); cudaStreamDestroy(bs[r]); bs[r] = nullptr; }
  }

  template <typename M>
  int mirror(std::vector<Mirror>& ms, M member) {   // member: pointer to a DBuf member of scdc_handle
    scdc_handle* h0 = hs[0];
    const int W = (int)hs.size();
    auto& src = h0->*member;
    ms.emplace_back();
    Mirror& m = ms.back();
    m.src = src.p;
    m.bytes = src.p ? src.n * sizeof(*src.p) : 0;
    m.dst.assign(W, nullptr);
    m.dst[0] = src.p;
    for (int r = 1; r < W && m.bytes; ++r) {
      CU(cudaSetDevice(hs[r]->device));
      AllocScope scope(bs[r]);   // allocated in the BROADCAST stream's order: the twin's compute stream may be busy
      auto& d = hs[r]->*member;
      CU(d.alloc(src.n));
      m.dst[r] = d.p;
    }
    return SCDC_OK;
  }
#define SCDC_MIRROR(member)                          \
  do {                                               \
    int rc_ = mirror(ms, &scdc_handle::member);      \
    if (rc_) return rc_;                             \
  } while (0)

  // device `to` may read / write the stream-ordered allocations of device `of` (pool memory needs cudaMemPoolSetAccess;
  // cudaDeviceEnablePeerAccess covers plain cudaMalloc memory, SCDC_NO_POOL=1)
  static void allow_peer(int of, int to) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, of) == cudaSuccess) {
      cudaMemAccessDesc d;
      memset(&d, 0, sizeof d);
      d.location.type = cudaMemLocationTypeDevice;
      d.location.id = to;
      d.flags = cudaMemAccessFlagsProtReadWrite;
      cudaMemPoolSetAccess(pool, &d, 1);
    }
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(to);
    cudaDeviceEnablePeerAccess(of, 0);
    cudaSetDevice(cur);
    cudaGetLastError();  // "already enabled" / unsupported: peer copies then fall back to staging, still correct
  }

  // Broadcast on the devices' broadcast streams. Nothing has to be waited for first: the root's buffers of a phase are
  // complete on the host's clock when the phase runs (create_impl synchronises its uploads; what still runs on the root's
  // compute stream is its first scatter, which must NOT be waited for), and the twins' buffers were allocated in the
  // broadcast streams' own order. The twins' compute streams then wait for the arrival.
  int broadcast_ordered(const std::vector<Mirror>& ms) {
    const int W = (int)hs.size();
    int rc = broadcast_mirrors(hs, ms, use_nccl && bcast_nccl, bs);
    if (rc) return rc;
    for (int r = 1; r < W; ++r) {
      CU(cudaSetDevice(hs[r]->device));
      cudaEvent_t ev;
      CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      cudaError_t e = cudaEventRecord(ev, bs[r]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(hs[r]->stream, ev, 0);
      cudaEventDestroy(ev);
      CU(e);
    }
    return SCDC_OK;
  }

  int phase_0(int root_device) {
    const int W = (int)hs.size();
    const scdc_problem* p = prob;
    std::vector<int> rcs((size_t)W, SCDC_OK);
    std::vector<std::string> errs((size_t)W);
    auto one = [&](int r) {   // one twin: context, labels from the host, its own label draw (a host thread per twin)
      auto body = [&]() -> int {
        std::unique_ptr<scdc_handle, void (*)(scdc_handle*)> h(new (std::nothrow) scdc_handle, scdc_release);
        if (!h) return fail(SCDC_ENOMEM, "host allocation failed");
        int rc = open_device_context(h.get(), devs[r]);
        if (rc) return rc;
        allow_peer(root_device, devs[r]);
        allow_peer(devs[r], root_device);
        if (getenv("SCDC_TIMING")) {
          int can = -1;
          cudaDeviceCanAccessPeer(&can, devs[r], root_device);
          fprintf(stderr, "[scdc timing] device %d can access device %d directly: %d; broadcast / merge by %s\n", devs[r],
                  root_device, can, use_nccl ? "NCCL" : "peer copies");
        }
        h->N = p->n_cells; h->G = p->n_genes; h->T = p->n_celltypes; h->C = p->n_conditions; h->L = p->n_lri;
        h->R = p->n_rows; h->nL = p->max_nL; h->nR = p->max_nR; h->K = h->C * h->T; h->nS = h->nL + h->nR;
        h->ncols = (h->C == 2) ? 3 + h->nS : 1;
        h->nnz = (size_t)p->colptr[h->G];
        h->one_shot = true;
        hs[r] = h.release();
        scdc_handle* hr = hs[r];
        AllocScope scope(hr->stream);
        rc = install_labels(hr, p, hr->stream);
        if (!rc && os[r].iterations > 0) rc = prelaunch_labels(hr, &os[r]);
        return rc;
      };
      try {
        rcs[r] = body();
      } catch (...) {
        rcs[r] = fail(SCDC_ENOMEM, "exception while preparing device %d", devs[r]);
      }
      if (rcs[r]) errs[r] = g_err;
    };
    std::vector<std::thread> th;
    try {
      for (int r = 2; r < W; ++r) th.emplace_back(one, r);
    } catch (...) {
      for (int r = (int)th.size() + 2; r < W; ++r) one(r);
    }
    if (W > 1) one(1);
    for (auto& t : th) t.join();
    for (int r = 1; r < W; ++r)
      if (rcs[r]) { g_err = errs[r]; return rcs[r]; }
    return SCDC_OK;
  }

  int phase_a(bool early_launches) {
    scdc_handle* h0 = hs[0];
    const int W = (int)hs.size();
    for (int r = 1; r < W; ++r) {
      hs[r]->sub_gene_h = h0->sub_gene_h; hs[r]->gene_nnz_h = h0->gene_nnz_h;
      hs[r]->n_active = h0->n_active; hs[r]->nnz_active = h0->nnz_active;
    }
    std::vector<Mirror> ms;
    SCDC_MIRROR(colptr); SCDC_MIRROR(rowidx); SCDC_MIRROR(values); SCDC_MIRROR(sub_gene); SCDC_MIRROR(gene_order);
    PhaseTimer pt;
    int rc = broadcast_ordered(ms);
    if (rc) return rc;
    if (pt.on && env_int("SCDC_TIMING", 1) >= 2) {  // SCDC_TIMING=2: wait for the copies so that their duration shows (changes the overlap it reports on)
      for (int r = 1; r < W; ++r) { cudaSetDevice(hs[r]->device); cudaStreamSynchronize(hs[r]->stream); }
      pt.lap("twins: matrix over NVLink");
    }
    for (int r = 1; r < W; ++r) {
      scdc_handle* h = hs[r];
      CU(cudaSetDevice(h->device));
      AllocScope scope(h->stream);
      cudaEvent_t ev;
      CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      struct EvDel { cudaEvent_t e; ~EvDel() { cudaEventDestroy(e); } } evdel{ev};
      CU(cudaEventRecord(ev, h->stream));  // the matrix has arrived
      if (early_launches && h->C == 2 && h->pre.valid && h->n_active > 0) {
        rc = launch_early_scatter(h, &os[r]);
        if (rc) return rc;
      }
      if (h->C == 2) {  // pass-1 copy of the matrix, built locally on the high-priority stream next to the first scatter
        cudaStream_t s_sort = h->stream_labels;
        AllocScope sort_scope(s_sort);
        CU(cudaStreamWaitEvent(s_sort, ev, 0));
        rc = build_pass1_copy(h, s_sort);
        if (!rc && early_launches) rc = launch_early_pass1(h, &os[r], s_sort);
        if (rc) return rc;
        CU(cudaEventRecord(ev, s_sort));
        CU(cudaStreamWaitEvent(h->stream, ev, 0));
      }
    }
    CU(cudaSetDevice(h0->device));
    phase_a_done = true;
    return SCDC_OK;
  }

  // phase 0 on a side thread (the calling thread is busy with the first device's upload)
  int root_device = 0, p0_rc = SCDC_OK;
  std::string p0_err;
  std::thread p0_thread;
  void start_phase_0() {
    auto run = [this]() {
      PhaseTimer pt;
      p0_rc = phase_0(root_device);
      if (p0_rc) p0_err = g_err;
      pt.lap("twins: contexts + labels (side thread)");
    };
    try {
      p0_thread = std::thread(run);
    } catch (...) {
      run();
    }
  }
  int join_phase_0() {
    if (p0_thread.joinable()) p0_thread.join();
    if (p0_rc) g_err = p0_err;
    return p0_rc;
  }
  ~MultiGpuCall() { if (p0_thread.joinable()) p0_thread.join(); }

  int matrix_ready(scdc_handle* root) override {
    int rc = join_phase_0();
    if (rc) return rc;
    hs[0] = root;
    return phase_a(true);
  }

  int phase_b() {
    scdc_handle* h0 = hs[0];
    const int W = (int)hs.size();
    for (int r = 1; r < W; ++r) {
      hs[r]->R = h0->R; hs[r]->fact_ok = h0->fact_ok; hs[r]->n_tasks = h0->n_tasks; hs[r]->score_threads = h0->score_threads;
      hs[r]->n_active = h0->n_active; hs[r]->nnz_active = h0->nnz_active;
    }
    std::vector<Mirror> ms;
    SCDC_MIRROR(row_lri); SCDC_MIRROR(row_emit); SCDC_MIRROR(row_recv); SCDC_MIRROR(obs); SCDC_MIRROR(task_lri);
    SCDC_MIRROR(task_beg); SCDC_MIRROR(task_end); SCDC_MIRROR(srow); SCDC_MIRROR(thr); SCDC_MIRROR(obs_sub);
    for (int r = 1; r < W; ++r) {
      CU(cudaSetDevice(hs[r]->device));
      AllocScope scope(hs[r]->stream);
      CU(hs[r]->flag.alloc(1));
    }
    return broadcast_ordered(ms);
  }
#undef SCDC_MIRROR
};

}  // namespace

extern "C" {

int scdc_run(scdc_handle* h, const scdc_options* o, scdc_result* res) {
  if (!h || !res) return fail(SCDC_EINVAL, "handle or result is NULL");
  int rc = validate_options(nullptr, h->C, o);
  if (rc) return rc;
  if (!res->counts && !res->counts_device) return fail(SCDC_EINVAL, "result.counts and result.counts_device are both NULL");
  if (o->return_distributions && !res->distributions)
    return fail(SCDC_EINVAL, "return_distributions set but result.distributions is NULL");
  const double t0 = now_ms(), h2d0 = g_h2d_bytes, d2h0 = g_d2h_bytes;
  memset(&res->stats, 0, sizeof res->stats);
  cudaStream_t s = o->stream ? (cudaStream_t)o->stream : h->stream;
  rc = run_device(h, o, o->return_distributions ? res->distributions : nullptr, &res->stats, s);
  if (rc) return rc;
  const double t1 = now_ms();
  const size_t n = (size_t)h->R * h->ncols;
  const bool f32 = o->precision == SCDC_FP32_FAST;
  if (n && (res->counts_device || res->band_device)) {
    if (res->counts_device)
      scdc::k_widen_counts<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, h->counts.p, (long long*)res->counts_device);
    if (res->band_device) {
      if (f32) scdc::k_widen_counts<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, h->band.p, (long long*)res->band_device);
      else CU(cudaMemsetAsync(res->band_device, 0, n * sizeof(long long), s));
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    res->stats.kernel_launches += 1;
  }
  // (the widened copies above were taken before the NA marks go into the uint32 counters)
  if (res->counts || res->band_counts) {
    rc = counts_to_host(h, h->counts.p, res->counts ? res->counts : res->band_counts, s);
    if (rc) return rc;
  }
  if (res->band_counts) {
    if (f32) {
      rc = counts_to_host(h, h->band.p, res->band_counts, s);
      if (rc) return rc;
    } else {  // nothing is approximated in fp64: zero, with the NA pattern of the counts
      const uint64_t* src = res->counts ? res->counts : res->band_counts;
      for (size_t i = 0; i < n; ++i) res->band_counts[i] = (src[i] == SCDC_COUNT_NA) ? SCDC_COUNT_NA : 0;
    }
  }
  res->stats.ms_merge = now_ms() - t1;
  res->stats.ms_total = now_ms() - t0;
  res->stats.n_gpus = 1;
  res->stats.h2d_bytes = g_h2d_bytes - h2d0;
  res->stats.d2h_bytes = g_d2h_bytes - d2h0;
  return SCDC_OK;
}

int scdc_perm_counts(const scdc_problem* p, const scdc_options* o, scdc_result* res) {
  if (!res) return fail(SCDC_EINVAL, "result is NULL");
  PhaseTimer pt;
  int rc = validate_meta(p);  // the O(nnz) matrix check overlaps the upload (create_impl)
  if (rc) return rc;
  rc = validate_options(p, p->n_conditions, o);
  if (rc) return rc;
  if (!res->counts) return fail(SCDC_EINVAL, "result.counts is NULL");
  if (o->return_distributions && !res->distributions)
    return fail(SCDC_EINVAL, "return_distributions set but result.distributions is NULL");
  std::vector<int> ids;
  const int avail = count_sm100_devices(&ids);
  if (avail == 0) {
    rc = validate_matrix(p);  // argument errors are reported before the missing device
    if (rc) return rc;
    return fail(SCDC_ENOGPU, "no visible CUDA device with compute capability 10.x (B200); there is no CPU fallback");
  }
  if (o->n_gpus < 0) return fail(SCDC_EINVAL, "n_gpus must be >= 0");
  if (o->device_base < 0 || o->device_base >= avail)
    return fail(SCDC_ENOGPU, "device_base = %d but %d compute-capability-10.x devices are visible", o->device_base, avail);
  ids.erase(ids.begin(), ids.begin() + o->device_base);
  const int usable = (int)ids.size();
  int W = (o->n_gpus == 0) ? usable : o->n_gpus;
  if (W > usable)
    return fail(SCDC_ENOGPU, "n_gpus = %d but only %d compute-capability-10.x devices are visible from device_base %d", W,
                usable, o->device_base);
  if (o->return_distributions) W = 1;  // <= 1000 replicates (R/utils_permutation.R:103-109): single GPU
  W = (int)std::min<int64_t>(W, std::max<int64_t>(1, o->iterations / 128));
  const double t0 = now_ms();
  memset(&res->stats, 0, sizeof res->stats);
  if (W == 1) {
    scdc_handle* h = nullptr;
    rc = create_impl(p, ids[0], &h, o);
    if (rc) return rc;
    pt.lap("perm_counts: create");
    scdc_result r1 = *res;
    r1.counts_device = nullptr;
    r1.band_device = nullptr;
    rc = scdc_run(h, o, &r1);
    pt.lap("perm_counts: run");
    res->stats = r1.stats;
    res->stats.ms_upload = h->ms_upload;
    res->stats.h2d_bytes += h->h2d_bytes;
    scdc_release(h);
    pt.lap("perm_counts: release");
    res->stats.ms_total = now_ms() - t0;
    return rc;
  }
  // ---- multi-GPU: replicate ranges sharded over W devices -----------------------------------------------------------
  // The problem is validated and uploaded ONCE (device ids[0], with the same overlaps as the single-GPU call), broadcast to
  // the other devices over NVLink (ncclBroadcast; peer copies when NCCL is unavailable), every device runs its replicate
  // range on its own host thread, and the uint32 counters are merged with one ncclAllReduce (or peer copies + adds).
  const double h2d0 = g_h2d_bytes, d2h0 = g_d2h_bytes, p2p0 = g_p2p_bytes;
  const bool f32 = o->precision == SCDC_FP32_FAST;
  MultiGpuCall mg;
  mg.hs.assign(W, nullptr);
  mg.os.assign(W, *o);
  mg.prob = p;
  std::vector<scdc_handle*>& hs = mg.hs;
  std::vector<scdc_options>& os = mg.os;
  std::vector<int> rcs(W, 0);
  std::vector<std::string> errs(W);
  std::vector<scdc_stats> sts(W);
  const int64_t per = (o->iterations + W - 1) / W;
  for (int r = 0; r < W; ++r) {
    memset(&sts[r], 0, sizeof(scdc_stats));
    os[r].interrupt = nullptr;  // the callback belongs to the calling thread; the workers watch abort_flag
    const int64_t b = std::min<int64_t>(o->iterations, per * r), e = std::min<int64_t>(o->iterations, per * (r + 1));
    os[r].iterations = e - b;
    os[r].perm_offset = o->perm_offset + b;
    if (o->perm_ct) os[r].perm_ct = o->perm_ct + (size_t)b * p->n_cells;
    if (o->perm_cond) os[r].perm_cond = o->perm_cond + (size_t)b * p->n_cells;
  }
  auto cleanup = [&]() {
    mg.close_broadcast_streams();
    for (auto h : hs) scdc_release(h);
  };
  std::lock_guard<std::mutex> comm_lock(g_nccl_mu);
  mg.devs.assign(ids.begin(), ids.begin() + W);
  const bool use_nccl = mg.use_nccl = nccl_comms_for(mg.devs);
  rc = mg.open_broadcast_streams();
  if (rc) { const std::string keep = g_err; cleanup(); g_err = keep; return rc; }
  // the twins' contexts, labels and label draws are prepared on a side thread while this one validates and uploads the
  // matrix to the first device; the hook (matrix_ready) joins it before the broadcast
  mg.root_device = ids[0];
  mg.start_phase_0();
  {
    scdc_handle* h0 = nullptr;
    rc = create_impl(p, ids[0], &h0, &os[0], &mg);  // W <= iterations / 128: every shard has replicates
    const std::string keep = g_err;
    const int rc0 = mg.join_phase_0();  // no-op when the hook has joined it already
    if (rc || rc0) {  // twins made by phase 0 (after a failed create mg.hs[0] is a handle create_impl has released itself)
      if (rc) { hs[0] = nullptr; g_err = keep; } else { hs[0] = h0; rc = rc0; }
      const std::string keep2 = g_err;
      cleanup();
      g_err = keep2;
      return rc;
    }
    hs[0] = h0;
  }
  pt.lap("perm_counts: create on the first device");
  if (!mg.phase_a_done) rc = mg.phase_a(false);
  if (!rc) rc = mg.phase_b();
  if (rc) { const std::string keep = g_err; cleanup(); g_err = keep; return rc; }
  pt.lap("perm_counts: broadcast");
  std::atomic<bool> abort_flag{false};
  std::atomic<int> workers_left{W};
  {
    std::vector<std::thread> th;
    auto body = [&](int r) {
      struct Done { std::atomic<int>& n; ~Done() { n.fetch_sub(1); } } done_guard{workers_left};
      try {
        const double h2d_before = g_h2d_bytes, w0 = now_ms();
        if (o->interrupt) hs[r]->abort = &abort_flag;
        rcs[r] = run_device(hs[r], &os[r], nullptr, &sts[r], hs[r]->stream);
        if (pt.on)
          fprintf(stderr, "[scdc timing] device %d: run %.1f ms on the host clock (device %.1f: labels %.1f, reduce %.1f, score %.1f)\n",
                  hs[r]->device, now_ms() - w0, sts[r].ms_device, sts[r].ms_labels, sts[r].ms_reduce, sts[r].ms_score);
        if (rcs[r]) errs[r] = g_err;
        sts[r].h2d_bytes = g_h2d_bytes - h2d_before;
      } catch (const std::bad_alloc&) {
        rcs[r] = SCDC_ENOMEM; errs[r] = "host allocation failed in a device worker";
      } catch (...) {
        rcs[r] = SCDC_ECUDA; errs[r] = "unexpected exception in a device worker";
      }
    };
    try {
      for (int r = 1; r < W; ++r) th.emplace_back(body, r);
    } catch (...) {  // thread creation failed: the remaining devices run one after the other on this thread
      for (int r = (int)th.size() + 1; r < W; ++r) body(r);
    }
    if (o->interrupt) {
      std::thread t0;
      bool own_thread = true;
      try { t0 = std::thread(body, 0); } catch (...) { own_thread = false; }
      if (!own_thread) body(0);
      while (workers_left.load() > 0) {
        if (!abort_flag.load() && o->interrupt(o->interrupt_user) != 0) abort_flag.store(true);
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
      }
      if (t0.joinable()) t0.join();
    } else {
      body(0);
    }
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < W; ++r)
    if (rcs[r]) { cleanup(); g_err = errs[r]; return rcs[r]; }
  pt.lap("perm_counts: run");
  const double t1 = now_ms();
  const size_t n = (size_t)p->n_rows * hs[0]->ncols;
  // total replicates <= 2e9 (validate_options), so the uint32 counters cannot overflow in the sum
  int bad = 0;
  const int n_sets = f32 ? 2 : 1;  // exceedance counters (+ band counters)
  if (n) {
    if (use_nccl) {
      for (int k = 0; k < n_sets && !bad; ++k) {
        int nrc = g_nccl.GroupStart();
        for (int r = 0; r < W && nrc == 0; ++r) {
          uint32_t* buf = k ? hs[r]->band.p : hs[r]->counts.p;
          nrc = g_nccl.AllReduce(buf, buf, n, kNcclUint32, kNcclSum, g_comms[r], hs[r]->stream);
        }
        if (g_nccl.GroupEnd() != 0 || nrc != 0) bad = 2;
      }
    } else {
      // no NCCL: the root pulls every peer's counters over NVLink (cudaMemcpyPeer) and adds them
      cudaSetDevice(hs[0]->device);
      AllocScope scope(hs[0]->stream);
      DBuf<uint32_t> tmp;
      if (tmp.alloc(n) != cudaSuccess) bad = 1;
      for (int k = 0; k < n_sets && !bad; ++k) {
        for (int r = 1; r < W && !bad; ++r) {
          uint32_t* src = k ? hs[r]->band.p : hs[r]->counts.p;
          uint32_t* dst = k ? hs[0]->band.p : hs[0]->counts.p;
          if (cudaMemcpyPeerAsync(tmp.p, hs[0]->device, src, hs[r]->device, n * sizeof(uint32_t), hs[0]->stream) != cudaSuccess) { bad = 3; break; }
          scdc::k_add_counts<<<(unsigned)((n + 255) / 256), 256, 0, hs[0]->stream>>>(n, tmp.p, dst);
          g_p2p_bytes += (double)n * sizeof(uint32_t);
        }
      }
      if (!bad && cudaStreamSynchronize(hs[0]->stream) != cudaSuccess) bad = 3;
    }
  }
  for (int r = 0; r < W; ++r) {
    cudaSetDevice(hs[r]->device);
    if (cudaStreamSynchronize(hs[r]->stream) != cudaSuccess) bad = bad ? bad : 3;
  }
  pt.lap("perm_counts: counters merged across devices");
  if (!bad) {
    cudaSetDevice(hs[0]->device);
    if (counts_to_host(hs[0], hs[0]->counts.p, res->counts, hs[0]->stream)) bad = 3;
    if (!bad && res->band_counts) {
      if (f32) { if (counts_to_host(hs[0], hs[0]->band.p, res->band_counts, hs[0]->stream)) bad = 3; }
      else {
        for (size_t i = 0; i < n; ++i) res->band_counts[i] = (res->counts[i] == SCDC_COUNT_NA) ? SCDC_COUNT_NA : 0;
      }
    }
  }
  if (bad) {
    cleanup();
    return fail(bad == 1 ? SCDC_ENOMEM : bad == 2 ? SCDC_ENCCL : SCDC_ECUDA, "counter merge across %d GPUs failed (stage %d)", W, bad);
  }
  for (int r = 0; r < W; ++r) {
    res->stats.ms_labels = std::max(res->stats.ms_labels, sts[r].ms_labels);
    res->stats.ms_reduce = std::max(res->stats.ms_reduce, sts[r].ms_reduce);
    res->stats.ms_score = std::max(res->stats.ms_score, sts[r].ms_score);
    res->stats.ms_device = std::max(res->stats.ms_device, sts[r].ms_device);
    res->stats.reduce_bytes += sts[r].reduce_bytes;
    res->stats.reduce_wavefronts += sts[r].reduce_wavefronts;
    res->stats.reduce_launches += sts[r].reduce_launches;
    res->stats.kernel_launches += sts[r].kernel_launches + 1;
    res->stats.batch_size = sts[r].batch_size;
    res->stats.h2d_bytes += sts[r].h2d_bytes;   // uploaded label slices, counted by each worker thread
  }
  res->stats.ms_upload = hs[0]->ms_upload;
  res->stats.h2d_bytes += hs[0]->h2d_bytes;
  (void)h2d0;
  res->stats.d2h_bytes = g_d2h_bytes - d2h0;
  res->stats.p2p_bytes = g_p2p_bytes - p2p0;
  pt.lap("perm_counts: counters to the host");
  cleanup();
  pt.lap("perm_counts: release");
  res->stats.ms_merge = now_ms() - t1;
  res->stats.ms_total = now_ms() - t0;
  res->stats.n_gpus = W;
  res->stats.reserved = use_nccl ? 1 : 0;
  return SCDC_OK;
}

int scdc_group_means(scdc_handle* h, double* means, double* detection_rate) {
  if (!h) return fail(SCDC_EINVAL, "handle is NULL");
  CU(cudaSetDevice(h->device));
  AllocScope alloc_scope(h->stream);
  DBuf<double> out;
  const size_t n = (size_t)h->K * h->G;
  CU(out.alloc(n));
  for (int mode = 0; mode < 2; ++mode) {
    double* dst = mode == 0 ? means : detection_rate;
    if (!dst) continue;
    scdc::k_group_means<<<(h->G + 127) / 128, 128, 0, h->stream>>>(h->G, h->K, h->T, h->colptr.p, h->rowidx.p, h->values.p,
                                                                 h->ct.p, h->C == 2 ? h->cond.p : nullptr, h->ngroup.p, mode,
                                                                 out.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst, out.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return SCDC_OK;
}

int scdc_set_rows(scdc_handle* h, int32_t n_rows, const int32_t* row_lri, const int32_t* row_emit, const int32_t* row_recv,
                  const double* observed) {
  if (!h) return fail(SCDC_EINVAL, "handle is NULL");
  if (n_rows < 0 || (n_rows > 0 && (!row_lri || !row_emit || !row_recv || !observed)))
    return fail(SCDC_EINVAL, "bad rows");
  for (int r = 0; r < n_rows; ++r) {
    if (row_lri[r] < 0 || row_lri[r] >= h->L || row_emit[r] < 0 || row_emit[r] >= h->T || row_recv[r] < 0 || row_recv[r] >= h->T)
      return fail(SCDC_EINVAL, "row %d has an out-of-range LRI / emitter / receiver code", r);
    const int l = row_lri[r];
    if (h->sub_gene_h[(size_t)l * h->nS] < 0 || h->sub_gene_h[(size_t)l * h->nS + h->nL] < 0)
      return fail(SCDC_EINVAL, "row %d: LRI %d lacks LIGAND_1 or RECEPTOR_1", r, l);
  }
  CU(cudaSetDevice(h->device));
  AllocScope alloc_scope(h->stream);
  h->pre.valid = false;
  try {
    return install_rows(h, n_rows, row_lri, row_emit, row_recv, observed);
  } catch (const std::bad_alloc&) {
    return fail(SCDC_ENOMEM, "host allocation failed while staging the rows");
  }
}

int scdc_observed_pass(scdc_handle* h, int64_t n, const int32_t* row_lri, const int32_t* row_emit, const int32_t* row_recv,
                       double threshold_pct, double threshold_min_cells, int32_t score_type, scdc_observed* out) {
  if (!h || !out) return fail(SCDC_EINVAL, "handle or out is NULL");
  if (n < 0 || (n > 0 && (!row_lri || !row_emit || !row_recv))) return fail(SCDC_EINVAL, "bad rows");
  if (score_type != SCDC_SCORE_GEOMETRIC && score_type != SCDC_SCORE_ARITHMETIC)
    return fail(SCDC_EINVAL, "unknown score_type %d", score_type);
  for (int64_t r = 0; r < n; ++r) {
    if (row_lri[r] < 0 || row_lri[r] >= h->L || row_emit[r] < 0 || row_emit[r] >= h->T || row_recv[r] < 0 || row_recv[r] >= h->T)
      return fail(SCDC_EINVAL, "row %lld has an out-of-range LRI / emitter / receiver code", (long long)r);
  }
  CU(cudaSetDevice(h->device));
  AllocScope alloc_scope(h->stream);
  cudaStream_t s = h->stream;
  const int K = h->K, G = h->G, C = h->C, nS = h->nS;
  DBuf<double> means, drate;
  CU(means.alloc((size_t)K * G));
  CU(drate.alloc((size_t)K * G));
  for (int mode = 0; mode < 2; ++mode) {
    scdc::k_group_means<<<(G + 127) / 128, 128, 0, s>>>(G, K, h->T, h->colptr.p, h->rowidx.p, h->values.p, h->ct.p,
                                                        C == 2 ? h->cond.p : nullptr, h->ngroup.p, mode,
                                                        mode == 0 ? means.p : drate.p);
    CU(cudaGetLastError());
  }
  const int64_t chunk = 1 << 20;  // rows per device chunk
  const int nc = (int)std::min<int64_t>(chunk, std::max<int64_t>(n, 1));
  DBuf<int> d_lri, d_emit, d_recv;
  DBuf<double> d_expr, d_dr, d_score;
  DBuf<unsigned char> d_is;
  CU(d_lri.alloc(nc)); CU(d_emit.alloc(nc)); CU(d_recv.alloc(nc));
  CU(d_expr.alloc((size_t)nc * C * nS)); CU(d_dr.alloc((size_t)nc * C * nS)); CU(d_score.alloc((size_t)nc * C));
  CU(d_is.alloc((size_t)nc * C));
  for (int64_t b = 0; b < n; b += chunk) {
    const int m = (int)std::min<int64_t>(chunk, n - b);
    CU(cudaMemcpyAsync(d_lri.p, row_lri + b, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_emit.p, row_emit + b, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(d_recv.p, row_recv + b, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, s));
    scdc::ObservedArgs A;
    A.n = m; A.T = h->T; A.C = C; A.nL = h->nL; A.nS = nS; A.score_type = score_type;
    A.threshold_pct = threshold_pct; A.threshold_min_cells = threshold_min_cells;
    A.row_lri = d_lri.p; A.row_emit = d_emit.p; A.row_recv = d_recv.p; A.sub_gene = h->sub_gene.p;
    A.means = means.p; A.drate = drate.p; A.ngroup = h->ngroup.p;
    A.expr = d_expr.p; A.dr = d_dr.p; A.score = d_score.p; A.is_expr = d_is.p;
    scdc::k_observed_rows<<<(m + 255) / 256, 256, 0, s>>>(A);
    CU(cudaGetLastError());
    // device chunks are column-major over m rows; the caller's arrays are column-major over n rows
    for (int col = 0; col < C * nS; ++col) {
      if (out->expression)
        CU(cudaMemcpyAsync(out->expression + (size_t)col * n + b, d_expr.p + (size_t)col * m, (size_t)m * sizeof(double),
                           cudaMemcpyDeviceToHost, s));
      if (out->detection_rate)
        CU(cudaMemcpyAsync(out->detection_rate + (size_t)col * n + b, d_dr.p + (size_t)col * m, (size_t)m * sizeof(double),
                           cudaMemcpyDeviceToHost, s));
    }
    for (int c = 0; c < C; ++c) {
      if (out->score)
        CU(cudaMemcpyAsync(out->score + (size_t)c * n + b, d_score.p + (size_t)c * m, (size_t)m * sizeof(double),
                           cudaMemcpyDeviceToHost, s));
      if (out->is_expressed)
        CU(cudaMemcpyAsync(out->is_expressed + (size_t)c * n + b, d_is.p + (size_t)c * m, (size_t)m, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
  }
  CU(cudaStreamSynchronize(s));
  return SCDC_OK;
}

}  // extern "C"

// ---- f3: integer-coded template and merge-back (host side of the shim) ----------------------------------------------
namespace {
// body(lo, hi) over [0, n) on a few host threads (falls back to the calling thread when none can be created)
template <typename F>
void parallel_rows(int64_t n, F&& body) {
  const int n_thr = (n > (1 << 20)) ? (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1;
  std::vector<std::thread> th;
  auto part = [&](int w) { body(n * w / n_thr, n * (w + 1) / n_thr); };
  try {
    for (int w = 1; w < n_thr; ++w) th.emplace_back(part, w);
  } catch (...) {
    for (int w = (int)th.size() + 1; w < n_thr; ++w) part(w);
  }
  part(0);
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

int scdc_cci_template(int32_t T, int32_t C, int32_t L, const int32_t* lri_order, int32_t N, const int32_t* ct_code,
                      const int32_t* cond_code, int32_t* emitter, int32_t* receiver, int32_t* lri_idx,
                      int32_t* ncells_emitter, int32_t* ncells_receiver) {
  if (T < 1 || L < 1 || (C != 1 && C != 2) || !lri_order) return fail(SCDC_EINVAL, "bad template dimensions");
  const int64_t n = (int64_t)T * T * L;
  if (n > 2147483647LL) return fail(SCDC_EINVAL, "template of %lld rows exceeds R's 32-bit row index", (long long)n);
  std::vector<uint8_t> seen;
  try {
    seen.assign((size_t)L, 0);
  } catch (const std::bad_alloc&) {
    return fail(SCDC_ENOMEM, "host allocation failed");
  }
  for (int l = 0; l < L; ++l) {
    if (lri_order[l] < 0 || lri_order[l] >= L || seen[lri_order[l]]) return fail(SCDC_EINVAL, "lri_order is not a permutation of 0..L-1");
    seen[lri_order[l]] = 1;
  }
  std::vector<int32_t> ncell((size_t)C * T, 0);
  if (ncells_emitter || ncells_receiver) {  // add_cell_number: table(cell_type[, condition])
    if (N < 1 || !ct_code || (C == 2 && !cond_code)) return fail(SCDC_EINVAL, "cell labels are needed for NCELLS");
    for (int i = 0; i < N; ++i) {
      const int t = ct_code[i], c = (C == 2) ? cond_code[i] : 0;
      if (t < 0 || t >= T || c < 0 || c >= C) return fail(SCDC_EINVAL, "label of cell %d out of range", i);
      ++ncell[(size_t)c * T + t];
    }
  }
  const int64_t TL = (int64_t)T * L;
  parallel_rows(n, [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      const int e = (int)(i / TL), r = (int)((i / L) % T), l = lri_order[i % L];
      if (emitter) emitter[i] = e;
      if (receiver) receiver[i] = r;
      if (lri_idx) lri_idx[i] = l;
      for (int c = 0; c < C; ++c) {
        if (ncells_emitter) ncells_emitter[(size_t)c * n + i] = ncell[(size_t)c * T + e];
        if (ncells_receiver) ncells_receiver[(size_t)c * n + i] = ncell[(size_t)c * T + r];
      }
    }
  });
  return SCDC_OK;
}

int scdc_merge_pvalues(int64_t n_full, int64_t n_tested, const int64_t* tested_rows, int32_t ncols, const uint64_t* counts,
                       int64_t iterations, int32_t L, int32_t nL, int32_t nR, const int32_t* sub_gene, const int32_t* lri_idx,
                       double* pvalues) {
  if (n_full < 0 || n_tested < 0 || n_tested > n_full || ncols < 1 || iterations < 1 || !pvalues ||
      (n_tested > 0 && (!tested_rows || !counts)))
    return fail(SCDC_EINVAL, "bad merge arguments");
  const bool differential = ncols > 1;
  if (differential && (ncols != 3 + nL + nR || !sub_gene || !lri_idx || L < 1))
    return fail(SCDC_EINVAL, "differential merge needs ncols = 3 + max_nL + max_nR, sub_gene and lri_idx");
  for (int64_t k = 0; k < n_tested; ++k)
    if (tested_rows[k] < 0 || tested_rows[k] >= n_full) return fail(SCDC_EINVAL, "tested_rows[%lld] out of range", (long long)k);
  const double nan = std::nan(""), it = (double)iterations;
  const int nS = nL + nR;
  parallel_rows(n_full, [&](int64_t lo, int64_t hi) {  // rows outside the tested subset: 1 (:434-441); absent subunits: NA (:442-469)
    for (int c = 0; c < ncols; ++c) {
      double* col = pvalues + (size_t)c * n_full;
      if (!differential || c < 3) {
        for (int64_t i = lo; i < hi; ++i) col[i] = 1.0;
      } else {
        const int sidx = c - 3;
        for (int64_t i = lo; i < hi; ++i) col[i] = (sub_gene[(size_t)lri_idx[i] * nS + sidx] < 0) ? nan : 1.0;
      }
    }
  });
  parallel_rows(n_tested, [&](int64_t lo, int64_t hi) {  // p = counts / iterations (:215-259), NA where the count is NA
    for (int c = 0; c < ncols; ++c) {
      double* col = pvalues + (size_t)c * n_full;
      const uint64_t* cnt = counts + (size_t)c * n_tested;
      for (int64_t k = lo; k < hi; ++k) col[tested_rows[k]] = (cnt[k] == SCDC_COUNT_NA) ? nan : (double)cnt[k] / it;
    }
  });
  return SCDC_OK;
}

int scdc_bh_adjust(int64_t n, const double* p, const int32_t* group, double* out) {
  if (n < 0 || (n > 0 && (!p || !out))) return fail(SCDC_EINVAL, "bad p-value arrays");
  try {
    std::vector<int64_t> idx;
    idx.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i)
      if (!std::isnan(p[i])) idx.push_back(i);
    // by group, then by DEcreasing p (stable: ties keep their order, as order(p, decreasing = TRUE) does)
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
      if (group && group[a] != group[b]) return group[a] < group[b];
      return p[a] > p[b];
    });
    std::vector<double> adj(idx.size());
    for (size_t b = 0; b < idx.size();) {
      size_t e = b;
      while (e < idx.size() && (!group || group[idx[e]] == group[idx[b]])) ++e;
      const double m = (double)(e - b);
      double run = INFINITY;  // cummin(n / (n:1) * p[o]) over decreasing p
      for (size_t k = b; k < e; ++k) {
        const double rank = m - (double)(k - b);  // n, n - 1, ..., 1
        const double v = m / rank * p[idx[k]];
        run = std::min(run, v);
        adj[k] = std::min(1.0, run);
      }
      b = e;
    }
    for (int64_t i = 0; i < n; ++i)
      if (std::isnan(p[i])) out[i] = p[i];
    for (size_t k = 0; k < idx.size(); ++k) out[idx[k]] = adj[k];
  } catch (const std::bad_alloc&) {
    return fail(SCDC_ENOMEM, "host allocation failed");
  }
  return SCDC_OK;
}

int scdc_trim(int32_t device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail(SCDC_ENOGPU, "no CUDA device"); }
  for (int d = 0; d < n; ++d) {
    if (device >= 0 && d != device) continue;
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, d));
    CU(cudaSetDevice(d));
    CU(cudaDeviceSynchronize());
    CU(cudaMemPoolTrimTo(pool, 0));
  }
  return SCDC_OK;
}

int scdc_selftest_de_filter(uint64_t n, uint64_t seed, uint64_t out[3]) {
  if (!out || n == 0) return fail(SCDC_EINVAL, "out is NULL or n == 0");
  std::vector<int> ids;
  if (count_sm100_devices(&ids) == 0) return fail(SCDC_ENOGPU, "no visible CUDA device with compute capability 10.x");
  CU(cudaSetDevice(ids[0]));
  DBuf<unsigned long long> d;
  CU(d.alloc(3));
  CU(cudaMemset(d.p, 0, 3 * sizeof(unsigned long long)));
  scdc::k_selftest_de<<<(unsigned)((n + 255) / 256), 256>>>(n, seed, d.p);
  CU(cudaGetLastError());
  unsigned long long hres[3];
  CU(cudaMemcpy(hres, d.p, sizeof hres, cudaMemcpyDeviceToHost));
  out[0] = hres[0]; out[1] = hres[1]; out[2] = hres[2];
  return SCDC_OK;
}

int scdc_draw_labels(scdc_handle* h, uint64_t seed, int64_t perm_offset, int64_t n, uint8_t* cond_out, uint8_t* ct_out) {
  if (!h || !ct_out) return fail(SCDC_EINVAL, "handle or ct_out is NULL");
  if (n < 1 || perm_offset < 0) return fail(SCDC_EINVAL, "n must be >= 1 and perm_offset >= 0");
  CU(cudaSetDevice(h->device));
  AllocScope alloc_scope(h->stream);
  cudaStream_t s = h->stream;
  const int N = h->N, C = h->C;
  const int B = 128;
  DBuf<uint8_t> lab2, dct, dcond;
  DBuf<uint32_t> bits1;
  CU(lab2.alloc((size_t)N * B));
  CU(bits1.alloc((size_t)N * (B / 32)));
  CU(dct.alloc((size_t)B * N));
  CU(dcond.alloc((size_t)B * N));
  for (int64_t done = 0; done < n; done += B) {
    const int nb = (int)std::min<int64_t>(B, n - done);
    CU(launch_draw(h, s, seed, perm_offset + done, B, B, 1, bits1.p, lab2.p, dcond.p, dct.p));
    CU(cudaMemcpyAsync(ct_out + (size_t)done * N, dct.p, (size_t)nb * N, cudaMemcpyDeviceToHost, s));
    if (cond_out && C == 2)
      CU(cudaMemcpyAsync(cond_out + (size_t)done * N, dcond.p, (size_t)nb * N, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return SCDC_OK;
}

}  // extern "C"
