// Host-side helpers of the usflow_b200 library (no CUDA): the fp32 -> bf16 narrowing of input rows on the host cores,
// so that the PCIe copy of a scoring call moves 2 bytes per coordinate instead of 4.  The bf16 precision tier rounds its
// input to bf16 as its first device step anyway (usf_convert_rows_kernel, round-to-nearest-even), so narrowing before
// the copy gives bit-identical scores.
#include <stdint.h>
#include <string.h>
#include <pthread.h>
#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/usflow_b200.h"

namespace {

// round-to-nearest-even, NaN -> 0x7FFF: the bit pattern __float2bfloat16_rn produces on the device
static inline uint16_t bf16_rne(uint32_t u) {
  const uint32_t r = (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
  return (u & 0x7FFFFFFFu) > 0x7F800000u ? (uint16_t)0x7FFF : (uint16_t)r;
}

#define USF_NARROW_BODY                                            \
  for (int64_t i = 0; i < n; ++i) {                                \
    uint32_t u;                                                    \
    memcpy(&u, src + i, 4);                                        \
    dst[i] = bf16_rne(u);                                          \
  }

// 16 values per iteration; the 32-byte results go out with non-temporal stores once dst is 32-byte aligned (the staging
// buffer is read next by the PCIe DMA engine, not by this core: no read-for-ownership, no cache pollution)
__attribute__((target("avx512f,avx512bw,avx512vl"))) void narrow_avx512(const float* src, uint16_t* dst, int64_t n) {
  int64_t i = 0;
  for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31) != 0; ++i) {
    uint32_t u;
    memcpy(&u, src + i, 4);
    dst[i] = bf16_rne(u);
  }
  const __m512i bias = _mm512_set1_epi32(0x7FFF), one = _mm512_set1_epi32(1);
  const __m512i absmask = _mm512_set1_epi32(0x7FFFFFFF), inf = _mm512_set1_epi32(0x7F800000);
  const __m512i qnan = _mm512_set1_epi32(0x7FFF);
  for (; i + 16 <= n; i += 16) {
    const __m512i u = _mm512_loadu_si512(src + i);
    __m512i r = _mm512_add_epi32(_mm512_add_epi32(u, bias), _mm512_and_si512(_mm512_srli_epi32(u, 16), one));
    r = _mm512_srli_epi32(r, 16);
    const __mmask16 isnan = _mm512_cmpgt_epu32_mask(_mm512_and_si512(u, absmask), inf);
    r = _mm512_mask_mov_epi32(r, isnan, qnan);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm512_cvtepi32_epi16(r));
  }
  _mm_sfence();
  for (; i < n; ++i) {
    uint32_t u;
    memcpy(&u, src + i, 4);
    dst[i] = bf16_rne(u);
  }
}
__attribute__((target("avx2"))) void narrow_avx2(const float* src, uint16_t* dst, int64_t n) { USF_NARROW_BODY }
void narrow_base(const float* src, uint16_t* dst, int64_t n) { USF_NARROW_BODY }

typedef void (*narrow_fn)(const float*, uint16_t*, int64_t);
narrow_fn pick_narrow() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl"))
    return narrow_avx512;
  if (__builtin_cpu_supports("avx2")) return narrow_avx2;
  return narrow_base;
}

struct Job {
  const float* src;
  int64_t lds;
  void* dst;          // bf16 rows (narrow) or fp32 rows (plain staged copy)
  int64_t ldd, rows, cols;
  int narrow;
};

// A small persistent pool: workers sleep on a condition variable between calls and pull row blocks from an atomic
// counter during one.  One job at a time (calls are serialised by `call_mu`).
class Pool {
 public:
  // Leaked on purpose (workers may outlive static destruction order).  A forked child has none of the parent's worker
  // threads and possibly a locked mutex: it gets a fresh pool.
  static Pool& get() {
    static std::once_flag once;
    std::call_once(once, [] {
      instance() = new Pool();
      pthread_atfork(nullptr, nullptr, [] { instance() = new Pool(); });
    });
    return *instance();
  }

  void run(const Job& job, int threads) {
    std::lock_guard<std::mutex> call(call_mu_);
    const int64_t kBlock = block_rows(job);
    const int64_t blocks = (job.rows + kBlock - 1) / kBlock;
    int helpers = threads - 1;
    if (helpers > blocks - 1) helpers = (int)(blocks - 1);
    if (helpers < 0) helpers = 0;
    ensure(helpers);
    job_ = job;
    next_.store(0, std::memory_order_relaxed);
    blocks_ = blocks;
    block_rows_ = kBlock;
    {
      std::lock_guard<std::mutex> l(mu_);
      active_ = helpers;
      wanted_ = helpers;
      ++generation_;
    }
    if (helpers > 0) cv_.notify_all();
    work();
    if (helpers > 0) {
      std::unique_lock<std::mutex> l(mu_);
      done_cv_.wait(l, [&] { return active_ == 0; });
    }
  }

 private:
  Pool() : fn_(pick_narrow()) {}
  static Pool*& instance() {
    static Pool* p = nullptr;
    return p;
  }

  static int64_t block_rows(const Job& j) {
    int64_t r = (int64_t)(1 << 16) / (j.cols > 0 ? j.cols : 1);   // ~64K elements per block
    return r < 1 ? 1 : r;
  }

  void ensure(int n) {
    while ((int)workers_.size() < n) {
      const int id = (int)workers_.size();
      workers_.emplace_back([this, id] { loop(id); });
      workers_.back().detach();
    }
  }

  void work() {
    for (;;) {
      const int64_t b = next_.fetch_add(1, std::memory_order_relaxed);
      if (b >= blocks_) return;
      const int64_t r0 = b * block_rows_;
      const int64_t r1 = r0 + block_rows_ < job_.rows ? r0 + block_rows_ : job_.rows;
      const bool flat = job_.lds == job_.cols && job_.ldd == job_.cols;
      if (job_.narrow) {
        uint16_t* d = static_cast<uint16_t*>(job_.dst);
        if (flat) {
          fn_(job_.src + r0 * job_.lds, d + r0 * job_.ldd, (r1 - r0) * job_.cols);
        } else {
          for (int64_t r = r0; r < r1; ++r) fn_(job_.src + r * job_.lds, d + r * job_.ldd, job_.cols);
        }
      } else {
        float* d = static_cast<float*>(job_.dst);
        if (flat) {
          memcpy(d + r0 * job_.ldd, job_.src + r0 * job_.lds, sizeof(float) * (size_t)((r1 - r0) * job_.cols));
        } else {
          for (int64_t r = r0; r < r1; ++r) memcpy(d + r * job_.ldd, job_.src + r * job_.lds, sizeof(float) * (size_t)job_.cols);
        }
      }
    }
  }

  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return generation_ != seen; });
        seen = generation_;
        if (id >= wanted_) continue;      // this call uses fewer helpers
      }
      work();
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--active_ == 0) done_cv_.notify_one();
      }
    }
  }

  narrow_fn fn_;
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  Job job_{};
  std::atomic<int64_t> next_{0};
  int64_t blocks_ = 0, block_rows_ = 1;
  int active_ = 0, wanted_ = 0;
  uint64_t generation_ = 0;
};

}  // namespace

static int host_rows(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols, int threads,
                     int narrow) {
  if (rows < 0 || cols < 0 || (rows > 0 && cols > 0 && (src == nullptr || dst == nullptr)) || lds < cols || ldd < cols)
    return USF_E_ARG;
  if (rows == 0 || cols == 0) return USF_OK;
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
  }
  if (threads > 64) threads = 64;
  Pool::get().run(Job{src, lds, dst, ldd, rows, cols, narrow}, threads);
  return USF_OK;
}

extern "C" int usf_host_f32_to_bf16(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols,
                                    int threads) {
  return host_rows(src, lds, dst, ldd, rows, cols, threads, 1);
}

extern "C" int usf_host_copy_f32(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int64_t cols,
                                 int threads) {
  return host_rows(src, lds, dst, ldd, rows, cols, threads, 0);
}
