#include "host_util.h"
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <thread>

namespace fastcheb {

ErrState& err_state() {
  static thread_local ErrState st;
  return st;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err_state().msg = buf;
  return code;
}

namespace {
std::mutex g_mu;
std::map<int, DeviceInfo> g_devinfo;

struct WsBuf {
  void* p;
  size_t bytes;
  bool busy;
};
std::map<int, std::vector<WsBuf>> g_ws;

constexpr int kFlagSlots = 64;
struct FlagPool {
  long long* dev = nullptr;
  long long* host = nullptr;
  uint64_t used = 0;
  uint64_t dirty = ~uint64_t(0);  // slots whose device flag is not known to be kNoBad
};
std::map<int, FlagPool> g_flags;

__global__ void set_flag_kernel(long long* f, long long v) { *f = v; }
}  // namespace

std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

cudaError_t set_flag_async(long long* dev_flag, long long value, cudaStream_t st) {
  note_launch();
  set_flag_kernel<<<1, 1, 0, st>>>(dev_flag, value);
  return cudaGetLastError();
}

const DeviceInfo* device_info(int dev) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_devinfo.find(dev);
  if (it != g_devinfo.end()) return it->second.ok ? &it->second : nullptr;
  DeviceInfo di;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || dev < 0 || dev >= count) {
    fail(FASTCHEB_ECUDA, "no usable CUDA device %d (%s); libfastcheb_cuda has no CPU fallback", dev,
         e != cudaSuccess ? cudaGetErrorString(e) : "index out of range");
    cudaGetLastError();
    return nullptr;  // not cached: a later call may see a device
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    fail(FASTCHEB_ECUDA, "cudaGetDeviceProperties(%d) failed", dev);
    return nullptr;
  }
  di.sms = prop.multiProcessorCount;
  di.cc_major = prop.major;
  di.cc_minor = prop.minor;
  di.smem_optin = prop.sharedMemPerBlockOptin;
  di.ok = prop.major == 10;  // the fat binary holds sm_100a code only
  g_devinfo[dev] = di;
  if (!di.ok) {
    fail(FASTCHEB_ECUDA, "device %d is sm_%d%d; libfastcheb_cuda is built for sm_100a (B200) only", dev, prop.major,
         prop.minor);
    return nullptr;
  }
  return &g_devinfo[dev];
}

DeviceGuard::DeviceGuard(int dev) {
  if (cudaGetDevice(&prev_) != cudaSuccess) {
    ok_ = false;
    return;
  }
  if (prev_ != dev) {
    ok_ = cudaSetDevice(dev) == cudaSuccess;
    switched_ = ok_;
  }
}
DeviceGuard::~DeviceGuard() {
  if (switched_) cudaSetDevice(prev_);
}

void* ws_acquire(int dev, size_t bytes) {
  if (bytes == 0) bytes = 16;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& v = g_ws[dev];
    int best = -1;
    for (int i = 0; i < (int)v.size(); ++i)
      if (!v[i].busy && v[i].bytes >= bytes && (best < 0 || v[i].bytes < v[best].bytes)) best = i;
    if (best >= 0) {
      v[best].busy = true;
      return v[best].p;
    }
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    // drop idle cached buffers and retry once
    std::vector<void*> drop;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      auto& v = g_ws[dev];
      for (size_t i = 0; i < v.size();)
        if (!v[i].busy) {
          drop.push_back(v[i].p);
          v.erase(v.begin() + i);
        } else
          ++i;
    }
    for (void* d : drop) cudaFree(d);
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_ws[dev].push_back({p, bytes, true});
  return p;
}

void ws_release(int dev, void* p) {
  if (!p) return;
  void* to_free = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& v = g_ws[dev];
    size_t idle_bytes = 0;
    int idle = 0;
    for (auto& b : v)
      if (!b.busy) {
        ++idle;
        idle_bytes += b.bytes;
      }
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i].p == p) {
        // keep at most 8 idle buffers / 8 GiB cached
        if (idle >= 8 || idle_bytes + v[i].bytes > (size_t(8) << 30)) {
          to_free = p;
          v.erase(v.begin() + i);
        } else {
          v[i].busy = false;
        }
        break;
      }
  }
  if (to_free) cudaFree(to_free);
}

namespace {
std::map<int, std::vector<StreamPair>> g_idle_streams;
struct PinBuf {
  void* p;
  size_t bytes;
  bool busy;
};
std::vector<PinBuf> g_pin;

// a tiny persistent pool for par_memcpy; never destroyed (threads are detached: no static-destruction order issues)
struct CopyPool {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  int nthreads = 0;
  void start() {
    const unsigned hw = std::thread::hardware_concurrency();
    nthreads = (int)std::min<unsigned>(3, hw > 1 ? hw - 1 : 0);
    for (int i = 0; i < nthreads; ++i)
      std::thread([this] {
        for (;;) {
          std::function<void()> job;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [this] { return !q.empty(); });
            job = std::move(q.front());
            q.pop_front();
          }
          job();
        }
      }).detach();
  }
};
CopyPool* copy_pool() {
  static CopyPool* pool = [] {
    CopyPool* p = new CopyPool();
    p->start();
    return p;
  }();
  return pool;
}
}  // namespace

int streams_acquire(int dev, StreamPair* out) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto& v = g_idle_streams[dev];
    if (!v.empty()) {
      *out = v.back();
      v.pop_back();
      return FASTCHEB_OK;
    }
  }
  StreamPair sp;
  sp.dev = dev;
  for (int i = 0; i < 2; ++i)
    if (cudaStreamCreateWithFlags(&sp.s[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&sp.ev[i], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      for (int j = 0; j < 2; ++j) {
        if (sp.s[j]) cudaStreamDestroy(sp.s[j]);
        if (sp.ev[j]) cudaEventDestroy(sp.ev[j]);
      }
      return fail(FASTCHEB_ECUDA, "cudaStreamCreate failed");
    }
  *out = sp;
  return FASTCHEB_OK;
}

void streams_release(const StreamPair& sp) {
  if (sp.dev < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_idle_streams[sp.dev].push_back(sp);
}

void* pin_acquire(size_t bytes) {
  if (bytes == 0) bytes = 16;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pin.size(); ++i)
      if (!g_pin[i].busy && g_pin[i].bytes >= bytes && g_pin[i].bytes <= std::max<size_t>(bytes * 64, size_t(1) << 16) &&
          (best < 0 || g_pin[i].bytes < g_pin[best].bytes))
        best = i;  // (a small request does not take a staging buffer thousands of times its size)
    if (best >= 0) {
      g_pin[best].busy = true;
      return g_pin[best].p;
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_pin.push_back({p, bytes, true});
  return p;
}

void pin_release(void* p) {
  if (!p) return;
  void* to_free = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int idle = 0;
    for (auto& b : g_pin) idle += !b.busy;
    for (size_t i = 0; i < g_pin.size(); ++i)
      if (g_pin[i].p == p) {
        if (idle >= 8) {  // keep at most 8 idle staging buffers
          to_free = p;
          g_pin.erase(g_pin.begin() + i);
        } else {
          g_pin[i].busy = false;
        }
        break;
      }
  }
  if (to_free) cudaFreeHost(to_free);
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void par_memcpy(void* dst, const void* src, size_t bytes) {
  CopyPool* pool = bytes >= (size_t(4) << 20) ? copy_pool() : nullptr;
  const int parts = pool ? pool->nthreads + 1 : 1;
  if (parts <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t slice = ((bytes / parts) + 4095) & ~size_t(4095);
  std::mutex mu;
  std::condition_variable cv;
  int left = 0;
  for (int i = 1; i < parts; ++i) {
    const size_t off = slice * i;
    if (off >= bytes) break;
    const size_t len = std::min(slice, bytes - off);
    {
      std::lock_guard<std::mutex> lk(mu);
      ++left;
    }
    {
      std::lock_guard<std::mutex> lk(pool->mu);
      pool->q.push_back([=, &mu, &cv, &left] {
        memcpy((char*)dst + off, (const char*)src + off, len);
        std::lock_guard<std::mutex> lk2(mu);
        if (--left == 0) cv.notify_all();
      });
    }
    pool->cv.notify_one();
  }
  memcpy(dst, src, std::min(slice, bytes));
  std::unique_lock<std::mutex> lk(mu);
  cv.wait(lk, [&] { return left == 0; });
}

int flag_acquire(int dev, FlagSlot* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  FlagPool& fp = g_flags[dev];
  if (!fp.dev) {
    if (cudaMalloc(&fp.dev, kFlagSlots * sizeof(long long)) != cudaSuccess ||
        cudaHostAlloc(&fp.host, kFlagSlots * sizeof(long long), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return fail(FASTCHEB_ENOMEM, "cannot allocate status flags on device %d", dev);
    }
  }
  for (int i = 0; i < kFlagSlots; ++i)
    if (!(fp.used >> i & 1)) {
      fp.used |= uint64_t(1) << i;
      out->dev = fp.dev + i;
      out->host = fp.host + i;
      out->dev_id = dev;
      out->slot = i;
      out->dirty = (fp.dirty >> i & 1) != 0;
      out->clean = false;
      fp.dirty |= uint64_t(1) << i;  // until released clean
      return FASTCHEB_OK;
    }
  return fail(FASTCHEB_ENOMEM, "more than %d concurrent calls on device %d", kFlagSlots, dev);
}

void flag_release(const FlagSlot& s) {
  if (s.slot < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  FlagPool& fp = g_flags[s.dev_id];
  fp.used &= ~(uint64_t(1) << s.slot);
  if (s.clean) fp.dirty &= ~(uint64_t(1) << s.slot);
}

cudaError_t flag_prepare(FlagSlot& s, cudaStream_t st) {
  if (!s.dirty) return cudaSuccess;
  return set_flag_async(s.dev, kNoBad, st);
}


namespace {
struct LaunchKey {
  int dev;
  const void* fn;
  int threads;
  size_t smem;
  bool operator<(const LaunchKey& o) const {
    if (dev != o.dev) return dev < o.dev;
    if (fn != o.fn) return fn < o.fn;
    if (threads != o.threads) return threads < o.threads;
    return smem < o.smem;
  }
};
std::mutex g_launch_mu;
std::map<LaunchKey, int> g_launch_cache;  // threads == -1: attribute set; otherwise blocks per SM
}  // namespace

cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t st) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  {
    static std::mutex mu;
    static unsigned long long configured = 0;  // bit per device
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && !(configured >> dev & 1ull)) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = 1ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      configured |= 1ull << dev;
    }
  }
  return cudaMallocAsync(p, bytes ? bytes : 16, st);
}

void pool_free_all(void* const* ptrs, int n) {
  bool any = false;
  for (int i = 0; i < n; ++i) any = any || ptrs[i] != nullptr;
  if (!any) return;
  cudaDeviceSynchronize();
  for (int i = 0; i < n; ++i)
    if (ptrs[i]) cudaFreeAsync(ptrs[i], cudaStreamPerThread);
}

cudaError_t ensure_max_smem(const void* kernel) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const LaunchKey key{dev, kernel, -1, 0};
  {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    if (g_launch_cache.count(key)) return cudaSuccess;
  }
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptinMax);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_launch_mu);
  g_launch_cache[key] = 1;
  return cudaSuccess;
}

cudaError_t cached_occupancy(int* blocks_per_sm, const void* kernel, int threads, size_t smem) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const LaunchKey key{dev, kernel, threads, smem};
  {
    std::lock_guard<std::mutex> lk(g_launch_mu);
    auto it = g_launch_cache.find(key);
    if (it != g_launch_cache.end()) {
      *blocks_per_sm = it->second;
      return cudaSuccess;
    }
  }
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, threads, smem);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_launch_mu);
  g_launch_cache[key] = *blocks_per_sm;
  return cudaSuccess;
}

}  // namespace fastcheb
