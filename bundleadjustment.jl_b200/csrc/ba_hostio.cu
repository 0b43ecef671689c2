// ba_hostio.cu -- device-to-host delivery of the large per-observation outputs (cx, vals, rows, cols) of the
// host-pointer entry points, i.e. the arrays a Julia `ccall` passes (include/bagpu.h).
//
// The reference hands `cons!` / `jac_coord!` ordinary Julia arrays (src/lm.jl:39,54,341: pageable memory).  A plain
// cudaMemcpy into pageable memory is staged by the driver through one internal bounce buffer and a single copying
// thread -- a fraction of the PCIe rate -- and page-locking the caller's array behind its back is not an option (it
// may be freed while registered).  Instead: a ring of pinned chunks owned by the library; the GPU fills chunk i
// (full PCIe rate) while a small pool of host threads copies the finished chunks i-1, i-2, ... into the caller's
// array.  Buffers that already are page-locked (ba_alloc_pinned, cudaHostRegister by the caller) take one direct copy.
#include <immintrin.h>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "ba_internal.h"

namespace ba {
namespace {

constexpr size_t CHUNK = (size_t)8 << 20;  // 8 MiB per pinned chunk
constexpr int RING = 12;                   // chunks in flight (96 MiB of pinned memory per process and device)

// pinned chunk -> caller's array with non-temporal stores: the destination is written once and not read here, so
// skipping the read-for-ownership of its cache lines saves a quarter of the host-memory traffic of the staged path
// (DMA write + copy read + RFO + write-back -> DMA write + copy read + streaming write)
__attribute__((target("avx2"))) void stream_copy_avx2(char* dst, const char* src, size_t n) {
  size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head > n) head = n;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  const size_t body = n & ~(size_t)127;
  for (size_t i = 0; i < body; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  memcpy(dst + body, src + body, n - body);
}
void stream_copy(void* dst, const void* src, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2") && getenv("BAGPU_COPY_PLAIN") == nullptr;
  if (avx2) stream_copy_avx2(static_cast<char*>(dst), static_cast<const char*>(src), n);
  else memcpy(dst, src, n);
}

struct job {
  cudaEvent_t ev;
  const void* src;
  void* dst;
  size_t bytes;
  std::atomic<int>* slot_busy;
  int device;
};

class copy_pool {
 public:
  explicit copy_pool(int nthreads) {
    for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { run(); });
  }
  ~copy_pool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void push(const job& j) {
    {
      std::lock_guard<std::mutex> g(m_);
      q_.push_back(j);
    }
    cv_.notify_one();
  }

 private:
  void run() {
    int dev = -1;
    for (;;) {
      job j;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;
        j = q_.front();
        q_.pop_front();
      }
      if (dev != j.device) {
        cudaSetDevice(j.device);
        dev = j.device;
      }
      cudaEventSynchronize(j.ev);  // the GPU has filled the chunk
      stream_copy(j.dst, j.src, j.bytes);
      j.slot_busy->store(0, std::memory_order_release);
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<job> q_;
  std::vector<std::thread> workers_;
  bool stop_ = false;
};

struct ring_state {
  char* base = nullptr;
  cudaEvent_t ev[RING] = {};
  std::atomic<int> busy[RING];
  bool ok = false;
};

copy_pool& pool() {
  static const int n = [] {
    if (const char* e = getenv("BAGPU_COPY_THREADS")) return std::max(1, std::min(32, atoi(e)));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(2u, std::min(8u, hw ? hw / 2 : 4u));
  }();
  static copy_pool p(n);
  return p;
}

ring_state* ring_for(int device) {
  static std::mutex m;
  static ring_state* rings[64] = {};
  std::lock_guard<std::mutex> g(m);
  if (device < 0 || device >= 64) return nullptr;
  if (!rings[device]) {
    ring_state* r = new ring_state();
    for (auto& b : r->busy) b.store(0);
    if (cudaHostAlloc(reinterpret_cast<void**>(&r->base), CHUNK * RING, cudaHostAllocDefault) == cudaSuccess) {
      r->ok = true;
      for (auto& e : r->ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) r->ok = false;
    }
    if (!r->ok) cudaGetLastError();
    rings[device] = r;
  }
  return rings[device];
}

bool is_page_locked(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;  // registered or allocated by CUDA; plain malloc memory is "unregistered"
}

}  // namespace

// dst (host, any memory) <- src (device), `bytes`; ordered after the work already in h->stream; returns when dst is
// complete.  Calls on different handles of one device share the ring (a handle is not re-entrant; two handles used
// from two threads at once would interleave chunks safely because slots are claimed atomically).
int copy_to_host(ba_handle* h, void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return BA_OK;
  static const bool direct_only = getenv("BAGPU_DIRECT_D2H") != nullptr;
  ring_state* R = (direct_only || bytes < 4 * CHUNK || is_page_locked(dst)) ? nullptr : ring_for(h->device);
  if (!R || !R->ok) {
    BA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    BA_CUDA(cudaStreamSynchronize(h->stream));
    return BA_OK;
  }
  copy_pool& P = pool();
  size_t off = 0;
  int slot = 0;
  std::vector<int> used;
  while (off < bytes) {
    const size_t n = std::min(CHUNK, bytes - off);
    // claim a free slot (round robin; spin briefly: a slot frees every ~100 us)
    for (;;) {
      int expect = 0;
      if (R->busy[slot].compare_exchange_strong(expect, 1, std::memory_order_acquire)) break;
      slot = (slot + 1) % RING;
      if (slot == 0) std::this_thread::yield();
    }
    char* stage = R->base + (size_t)slot * CHUNK;
    BA_CUDA(cudaMemcpyAsync(stage, static_cast<const char*>(src) + off, n, cudaMemcpyDeviceToHost, h->stream));
    BA_CUDA(cudaEventRecord(R->ev[slot], h->stream));
    P.push(job{R->ev[slot], stage, static_cast<char*>(dst) + off, n, &R->busy[slot], h->device});
    used.push_back(slot);
    off += n;
    slot = (slot + 1) % RING;
  }
  // wait until every chunk of this call has been copied out
  for (int s : used)
    while (R->busy[s].load(std::memory_order_acquire) != 0) std::this_thread::yield();
  return BA_OK;
}

}  // namespace ba
