// Error plumbing, device queries and the K0 host-side batch packers of libhhfm_sm100.so.
// Packers replace the numpy slicing that builds `feed_dict` batches (FM.py:251-256, OurModel7.py:373-385):
// id columns are narrowed to int32 and laid out as 16-byte aligned per-sample records in (pinned) host
// memory, ready for one cudaMemcpyAsync.  Plain std::thread fan-out; no device work here.
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace hhfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return HHFM_ERR_LAUNCH;
  }
  return HHFM_OK;
}

int sm_count() {
  static thread_local int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;   // B200
  }
  return cached;
}

template <typename Fn>
static void parallel_rows(int64_t rows, int nthreads, Fn fn) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  const int64_t min_rows = 1 << 14;
  int nt = (int)std::min<int64_t>(nthreads, (rows + min_rows - 1) / min_rows);
  if (nt <= 1) {
    fn(0, rows);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (rows + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const int64_t a = t * chunk, b = std::min(rows, a + chunk);
    if (a >= b) break;
    th.emplace_back([=] { fn(a, b); });
  }
  for (auto& x : th) x.join();
}

template <typename T>
static int pack_ids(const T* src, int64_t rows, int64_t cols, int64_t src_row_stride, int32_t* dst,
                    int64_t dst_row_stride, int64_t dst_col0, int64_t id_limit, int nthreads) {
  HHFM_REQUIRE(src && dst, "pack_ids: NULL argument");
  HHFM_REQUIRE(rows >= 0 && cols >= 0 && src_row_stride >= cols && dst_col0 >= 0 && dst_row_stride >= dst_col0 + cols,
               "pack_ids: bad shape rows=%lld cols=%lld", (long long)rows, (long long)cols);
  std::atomic<int64_t> bad{-1};
  parallel_rows(rows, nthreads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; r++) {
      const T* s = src + r * src_row_stride;
      int32_t* d = dst + r * dst_row_stride + dst_col0;
      for (int64_t c = 0; c < cols; c++) {
        const int64_t v = (int64_t)s[c];
        if (v < 0 || v >= id_limit) bad.store(r);
        d[c] = (int32_t)v;
      }
    }
  });
  if (bad.load() >= 0) {
    set_error("pack_ids: id out of range [0,%lld) in row %lld", (long long)id_limit, (long long)bad.load());
    return HHFM_ERR_BAD_ARG;
  }
  return HHFM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent worker pool for the pipelined packer (thread creation per call costs more than packing a chunk).
// Not fork-safe by itself: the child of a fork() gets a fresh pool (pthread_atfork).
// ---------------------------------------------------------------------------------------------------------------
class Pool {
 public:
  // pin_base >= 0: worker i is pinned to core (pin_base + i) % cores, so the pools of several ranks on one box work on
  // disjoint cores instead of migrating over all of them
  Pool(int n, int pin_base) : n_(n) {
    for (int i = 0; i < n_; i++) th_.emplace_back([this, i] { loop(i); });
    if (pin_base >= 0) {
      const int cores = (int)std::thread::hardware_concurrency();
      for (int i = 0; i < n_ && cores > 0; i++) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET((pin_base + i) % cores, &set);
        pthread_setaffinity_np(th_[i].native_handle(), sizeof(set), &set);      // best effort
      }
    }
  }
  int size() const { return n_; }
  // run fn(worker) on every worker and wait
  void run(const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> serial(run_mu_);      // one parallel region at a time
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn;
    pending_ = n_;
    gen_++;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* f;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        f = fn_;
      }
      (*f)(id);
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t gen_ = 0;
};

static Pool* g_pool = nullptr;
static std::mutex g_pool_mu;
static void pool_forget() { g_pool = nullptr; new (&g_pool_mu) std::mutex(); }   // child after fork(): threads are gone

static Pool* get_pool(int nthreads) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (g_pool == nullptr) {
    static bool hooked = false;
    if (!hooked) { pthread_atfork(nullptr, nullptr, pool_forget); hooked = true; }
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    const char* pin = getenv("HHFM_PACK_PIN_BASE");
    g_pool = new Pool(nthreads, pin ? atoi(pin) : -1);      // lives for the process (workers block on a condition variable when idle)
  }
  return g_pool;
}

// uint16 wire records -> int32 device records (0xFFFF = padding -> -1)
__global__ void __launch_bounds__(256) widen_u16_kernel(const uint16_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const uint2 v = *reinterpret_cast<const uint2*>(src + i);
    int4 o;
    o.x = (int)(v.x & 0xFFFFu); o.y = (int)(v.x >> 16); o.z = (int)(v.y & 0xFFFFu); o.w = (int)(v.y >> 16);
    o.x = o.x == 0xFFFF ? -1 : o.x; o.y = o.y == 0xFFFF ? -1 : o.y; o.z = o.z == 0xFFFF ? -1 : o.z; o.w = o.w == 0xFFFF ? -1 : o.w;
    *reinterpret_cast<int4*>(dst + i) = o;
  } else {
    for (int64_t k = i; k < n; k++) { const int v = src[k]; dst[k] = v == 0xFFFF ? -1 : v; }
  }
}

template <typename D>
static inline void pack_row_block_scalar(const hhfm_pack_part* parts, int n_parts, int64_t r0, int64_t r1, D* dst, int64_t stride,
                                         int64_t width, int64_t id_limit, std::atomic<int64_t>& bad) {
  const D pad = (D)-1;          // int32: -1, uint16: 0xFFFF
  const uint64_t lim = (uint64_t)id_limit;
  for (int64_t r = r0; r < r1; r++) {
    D* d = dst + r * stride;
    uint64_t viol = 0;          // branch-free range check: a negative id is a huge unsigned value
    for (int p = 0; p < n_parts; p++) {
      const hhfm_pack_part& pt = parts[p];
      if (pt.elem_bytes == 8) {
        const int64_t* s = reinterpret_cast<const int64_t*>(pt.data) + r * pt.row_stride;
        for (int64_t c = 0; c < pt.cols; c++) {
          const uint64_t v = (uint64_t)s[c];
          viol |= (uint64_t)(v >= lim);
          d[c] = (D)v;
        }
      } else {
        const int32_t* s = reinterpret_cast<const int32_t*>(pt.data) + r * pt.row_stride;
        for (int64_t c = 0; c < pt.cols; c++) {
          const uint64_t v = (uint64_t)(int64_t)s[c];
          viol |= (uint64_t)(v >= lim);
          d[c] = (D)v;
        }
      }
      d += pt.cols;
    }
    if (viol) bad.store(r);
    for (int64_t c = width; c < stride; c++) dst[r * stride + c] = pad;
  }
}

// int64 ids -> 16-bit wire records, 8 ids per instruction (vpmovqw) with masked tails.  Alone, 16 threads on the B200 box pack
// 168 MB of ids in 1.6 ms (scalar 2.1 ms, a read-only sweep 1.5 ms); inside partial_fit, where the H2D DMA shares the host
// memory, both versions end at the same 3.5 ms per 2^20-positive step (profiles/r1_epoch_timing.md): the step is bound by
// host memory traffic (int64 ids in, wire records out, DMA read), not by the packing arithmetic.
__attribute__((target("avx512f,avx512bw,avx512vl")))
static void pack_row_block_u16_avx512(const hhfm_pack_part* parts, int n_parts, int64_t r0, int64_t r1, uint16_t* dst,
                                      int64_t stride, int64_t width, int64_t id_limit, std::atomic<int64_t>& bad) {
  const __m512i lim = _mm512_set1_epi64(id_limit);
  for (int64_t r = r0; r < r1; r++) {
    uint16_t* d = dst + r * stride;
    __mmask8 viol = 0;
    for (int p = 0; p < n_parts; p++) {
      const hhfm_pack_part& pt = parts[p];
      const int64_t* s = reinterpret_cast<const int64_t*>(pt.data) + r * pt.row_stride;
      int64_t c = 0;
      for (; c + 8 <= pt.cols; c += 8) {
        const __m512i v = _mm512_loadu_si512(s + c);
        viol |= _mm512_cmpge_epu64_mask(v, lim);
        _mm_storeu_si128(reinterpret_cast<__m128i*>(d + c), _mm512_cvtepi64_epi16(v));
      }
      if (c < pt.cols) {
        const __mmask8 m = (__mmask8)((1u << (pt.cols - c)) - 1);
        const __m512i v = _mm512_maskz_loadu_epi64(m, s + c);
        viol |= _mm512_cmpge_epu64_mask(v, lim);           // masked-off lanes are 0: in range
        _mm512_mask_cvtepi64_storeu_epi16(d + c, m, v);
      }
      d += pt.cols;
    }
    if (viol) bad.store(r);
    for (int64_t c = width; c < stride; c++) dst[r * stride + c] = 0xFFFF;
  }
}

static bool cpu_has_avx512() {
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
  if (!ok) return false;
  const char* e = getenv("HHFM_PACK_SIMD");          // 0 = scalar packer (A/B measurements)
  return !(e && e[0] == '0');
}

template <typename D>
static inline void pack_row_block(const hhfm_pack_part* parts, int n_parts, int64_t r0, int64_t r1, D* dst, int64_t stride,
                                  int64_t width, int64_t id_limit, std::atomic<int64_t>& bad) {
  if (sizeof(D) == 2 && cpu_has_avx512()) {
    bool all64 = true;
    for (int p = 0; p < n_parts; p++) all64 = all64 && parts[p].elem_bytes == 8;
    if (all64) {
      pack_row_block_u16_avx512(parts, n_parts, r0, r1, reinterpret_cast<uint16_t*>(dst), stride, width, id_limit, bad);
      return;
    }
  }
  pack_row_block_scalar<D>(parts, n_parts, r0, r1, dst, stride, width, id_limit, bad);
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_abi_version(void) { return 1; }
extern "C" const char* hhfm_last_error(void) { return g_err; }
extern "C" int64_t hhfm_partials_len(void) { return kPartials; }

extern "C" int hhfm_pack_ids_i64(const int64_t* src, int64_t rows, int64_t cols, int64_t src_row_stride, int32_t* dst,
                                 int64_t dst_row_stride, int64_t dst_col0, int64_t id_limit, int nthreads) {
  return pack_ids<int64_t>(src, rows, cols, src_row_stride, dst, dst_row_stride, dst_col0, id_limit, nthreads);
}

extern "C" int hhfm_pack_ids_i32(const int32_t* src, int64_t rows, int64_t cols, int64_t src_row_stride, int32_t* dst,
                                 int64_t dst_row_stride, int64_t dst_col0, int64_t id_limit, int nthreads) {
  return pack_ids<int32_t>(src, rows, cols, src_row_stride, dst, dst_row_stride, dst_col0, id_limit, nthreads);
}

extern "C" int hhfm_pack_fill_i32(int32_t* dst, int64_t rows, int64_t cols, int64_t dst_row_stride, int64_t dst_col0,
                                  int32_t value, int nthreads) {
  HHFM_REQUIRE(dst && rows >= 0 && cols >= 0 && dst_row_stride >= dst_col0 + cols, "pack_fill: bad arguments");
  parallel_rows(rows, nthreads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; r++)
      for (int64_t c = 0; c < cols; c++) dst[r * dst_row_stride + dst_col0 + c] = value;
  });
  return HHFM_OK;
}

extern "C" int hhfm_pack_csr_i64(const int64_t* src, const float* src_val, int64_t rows, int64_t cols,
                                 int64_t src_row_stride, int32_t* row_ptr, int32_t* col, float* val, int64_t id_limit,
                                 int nthreads) {
  HHFM_REQUIRE(src && row_ptr && col, "pack_csr: NULL argument");
  HHFM_REQUIRE((src_val == nullptr) == (val == nullptr), "pack_csr: src_val and val must both be given or both NULL");
  HHFM_REQUIRE(rows * cols < ((int64_t)1 << 31), "pack_csr: nnz does not fit int32 row_ptr");
  int rc = pack_ids<int64_t>(src, rows, cols, src_row_stride, col, cols, 0, id_limit, nthreads);
  if (rc) return rc;
  parallel_rows(rows + 1, nthreads, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; r++) row_ptr[r] = (int32_t)(r * cols);
  });
  if (val)
    parallel_rows(rows, nthreads, [&](int64_t a, int64_t b) {
      for (int64_t r = a; r < b; r++) std::memcpy(val + r * cols, src_val + r * src_row_stride, sizeof(float) * cols);
    });
  return HHFM_OK;
}

extern "C" int64_t hhfm_pack_upload_staging_bytes(int64_t rows, int64_t stride, int64_t id_limit) {
  return rows * stride * (id_limit <= 65535 ? 2 : 4);
}

extern "C" int hhfm_pack_upload_records(const hhfm_pack_part* parts, int32_t n_parts, int64_t rows, int64_t stride,
                                        int64_t id_limit, void* host_staging, void* dev_staging, int32_t* dev_records,
                                        int32_t nthreads, hhfm_stream_t stream) {
  HHFM_REQUIRE(parts && n_parts >= 1 && host_staging && dev_records, "pack_upload_records: NULL argument");
  HHFM_REQUIRE(rows >= 0 && stride >= 1 && id_limit >= 1, "pack_upload_records: bad sizes");
  int64_t width = 0;
  for (int p = 0; p < n_parts; p++) {
    HHFM_REQUIRE(parts[p].data && parts[p].cols >= 0 && parts[p].row_stride >= parts[p].cols &&
                     (parts[p].elem_bytes == 8 || parts[p].elem_bytes == 4),
                 "pack_upload_records: bad part %d", p);
    width += parts[p].cols;
  }
  HHFM_REQUIRE(width <= stride, "pack_upload_records: stride %lld < total width %lld", (long long)stride, (long long)width);
  if (rows == 0) return HHFM_OK;
  const bool narrow = id_limit <= 65535;           // 0xFFFF is the padding code
  HHFM_REQUIRE(!narrow || dev_staging, "pack_upload_records: the 16-bit wire format needs dev_staging");
  cudaStream_t st = (cudaStream_t)stream;
  Pool* pool = get_pool(nthreads);
  const int nt = pool->size();
  // chunk so that the H2D copy of chunk i overlaps the packing of chunk i+1 (>= 64K rows per chunk, <= 16 chunks)
  int64_t n_chunks = rows / (64 * 1024);
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > 16) n_chunks = 16;
  const int64_t chunk = (rows + n_chunks - 1) / n_chunks;
  std::atomic<int64_t> bad{-1};
  const size_t esz = narrow ? 2 : 4;
  n_chunks = (rows + chunk - 1) / chunk;
  // ONE parallel region for the whole batch: every worker packs its slice of chunk 0, 1, ... without a barrier in between,
  // and whichever worker completes a chunk queues that chunk's H2D copy (waking the pool once per chunk cost as much as
  // packing the chunk).  The copies all go to `stream`; their order among themselves does not matter.
  int device = 0;
  cudaGetDevice(&device);
  std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[n_chunks]);
  for (int64_t i = 0; i < n_chunks; i++) done[i].store(0);
  std::atomic<int> copy_failed{0};
  const std::function<void(int)> job = [&](int w) {
    bool device_set = false;
    for (int64_t ci = 0; ci < n_chunks; ci++) {
      const int64_t c0 = ci * chunk, c1 = std::min(rows, c0 + chunk);
      const int64_t per = (c1 - c0 + nt - 1) / nt;
      const int64_t a = c0 + (int64_t)w * per, b = std::min(c1, a + per);
      if (a < b) {
        if (narrow) pack_row_block<uint16_t>(parts, n_parts, a, b, reinterpret_cast<uint16_t*>(host_staging), stride, width, id_limit, bad);
        else pack_row_block<int32_t>(parts, n_parts, a, b, reinterpret_cast<int32_t*>(host_staging), stride, width, id_limit, bad);
      }
      if (done[ci].fetch_add(1, std::memory_order_acq_rel) + 1 != nt) continue;
      if (bad.load() >= 0 || copy_failed.load()) continue;
      if (!device_set) { cudaSetDevice(device); device_set = true; }
      const size_t off = (size_t)c0 * stride * esz, bytes = (size_t)(c1 - c0) * stride * esz;
      void* dst = narrow ? (void*)((char*)dev_staging + off) : (void*)((char*)dev_records + off);
      if (cudaMemcpyAsync(dst, (const char*)host_staging + off, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) copy_failed.store(1);
    }
  };
  pool->run(job);
  if (copy_failed.load()) {
    set_error("pack_upload_records: cudaMemcpyAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    return HHFM_ERR_LAUNCH;
  }
  if (bad.load() >= 0) {
    set_error("pack_ids: id out of range [0,%lld) in row %lld", (long long)id_limit, (long long)bad.load());
    return HHFM_ERR_BAD_ARG;
  }
  if (narrow) {
    const int64_t n = rows * stride;
    widen_u16_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(dev_staging), dev_records, n);
    return check_launch("widen_u16_kernel");
  }
  return HHFM_OK;
}
