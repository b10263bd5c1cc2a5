// Host-side plumbing shared by the C-ABI translation units: thread-local error state, device
// bookkeeping, a small device-buffer cache and the per-call status-flag pool.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/fastcheb.h"

namespace fastcheb {

struct ErrState {
  std::string msg;
  int64_t index = -1;
};
ErrState& err_state();

int fail(int code, const char* fmt, ...);

#define FC_CUDA(x)                                                                              \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess)                                                                      \
      return ::fastcheb::fail(e_ == cudaErrorMemoryAllocation ? FASTCHEB_ENOMEM : FASTCHEB_ECUDA, \
                              "%s failed: %s", #x, cudaGetErrorString(e_));                     \
  } while (0)
#define FC_TRY(x)                   \
  do {                              \
    int rc_ = (x);                  \
    if (rc_ != FASTCHEB_OK) return rc_; \
  } while (0)

struct DeviceInfo {
  bool ok = false;
  int sms = 0;
  int cc_major = 0, cc_minor = 0;
  size_t smem_optin = 0;
};
// cached per device; returns nullptr (and sets the error) if the device is unusable for this library
const DeviceInfo* device_info(int dev);

class DeviceGuard {
 public:
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
  bool ok() const { return ok_; }

 private:
  int prev_ = -1;
  bool switched_ = false, ok_ = true;
};

// grow-only cache of device scratch buffers (per device), so repeated host-memory calls do not pay
// cudaMalloc each time.
void* ws_acquire(int dev, size_t bytes);
void ws_release(int dev, void* p);

// Handle-lifetime device memory from the device's stream-ordered pool (cudaMallocAsync, release threshold 1 GiB):
// creating and destroying handles in a loop (roots, the adaptive chebinterp, chebinterp(vals) per call) then costs
// microseconds instead of the ~1 ms of a cudaMalloc / cudaFree pair.  pool_free_all synchronises the device first, like
// cudaFree does, because work on any stream may still read the buffers.
cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t st);
void pool_free_all(void* const* ptrs, int n);

// Host-staged calls (caller buffers in host memory) that are cut into chunks run on a cached pair of non-blocking streams
// with one event each, instead of creating and destroying streams per call.
struct StreamPair {
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int dev = -1;
};
int streams_acquire(int dev, StreamPair* out);
void streams_release(const StreamPair& sp);

// grow-only cache of PINNED host staging buffers: pageable caller memory (a Julia Vector, a numpy array) is copied
// through them by the calling thread so that the DMA transfers stay asynchronous and overlap the kernels.
void* pin_acquire(size_t bytes);
void pin_release(void* p);
// true when `p` is page-locked (cudaHostAlloc / cudaHostRegister) or managed memory, i.e. cudaMemcpyAsync from / to it
// does not block the host
bool is_pinned_host(const void* p);
// memcpy split over a small persistent pool of host threads (large pageable <-> pinned staging copies)
void par_memcpy(void* dst, const void* src, size_t bytes);

// one int64 device flag (+ pinned host mirror) per in-flight synchronous call
struct FlagSlot {
  long long* dev = nullptr;
  long long* host = nullptr;  // pinned
  int dev_id = -1;
  int slot = -1;
  bool dirty = true;   // at acquire: the device flag may hold something other than kNoBad
  bool clean = false;  // at release: the caller has seen kNoBad after its last kernel (flag_observe)
};
int flag_acquire(int dev, FlagSlot* out);
void flag_release(const FlagSlot& s);
// Slots go back to the pool CLEAN when the call that used them read kNoBad back after synchronising (the common case), so
// the next call skips the reset kernel: one launch less on every small synchronous call.
cudaError_t flag_prepare(FlagSlot& s, cudaStream_t st);  // reset the device flag to kNoBad unless the slot is clean
inline void flag_observe(FlagSlot& s) { s.clean = s.host && *s.host == 0x7fffffffffffffffLL; }  // after the final sync

constexpr long long kNoBad = 0x7fffffffffffffffLL;
cudaError_t set_flag_async(long long* dev_flag, long long value, cudaStream_t st);

// process-wide count of kernels launched by this library (bench.py reports it as gpu_launches)
void note_launch(int n = 1);
long long launch_count();

// Dynamic shared-memory ceiling requested for every kernel (the sm_100 opt-in maximum).  The attribute is per kernel
// and process-wide: setting it to the launch's own size would race between host threads that launch the same
// instantiation with different sizes, so every launch site sets this one value.
constexpr int kSmemOptinMax = 232448;

// Host-side launch set-up, paid once instead of on every call (each is several microseconds of driver time, which
// is what bounds small batches): the opt-in attribute once per (device, kernel), the occupancy once per
// (device, kernel, threads, dynamic shared memory).  Thread-safe.
cudaError_t ensure_max_smem(const void* kernel);
cudaError_t cached_occupancy(int* blocks_per_sm, const void* kernel, int threads, size_t smem);

inline size_t real_size(int dtype) { return dtype == FASTCHEB_F64 ? 8 : 4; }

}  // namespace fastcheb
