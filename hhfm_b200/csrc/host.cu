// Error plumbing, device queries and the K0 host-side batch packers of libhhfm_sm100.so.
// Packers replace the numpy slicing that builds `feed_dict` batches (FM.py:251-256, OurModel7.py:373-385):
// id columns are narrowed to int32 and laid out as 16-byte aligned per-sample records in (pinned) host
// memory, ready for one cudaMemcpyAsync.  Plain std::thread fan-out; no device work here.
#include <algorithm>
#include <atomic>
#include <cstring>
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
