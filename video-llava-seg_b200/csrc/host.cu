#include "host.h"
#include <algorithm>

#include <cudaTypedefs.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

namespace vls {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("VLS_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;
  }
  return g_pdl == 1;
}
void pdl_set(bool on) { g_pdl = on ? 1 : 0; }

bool first_use_on_device(unsigned long long* flag_word) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;   // unknown device: just redo the setup
  const unsigned long long bit = 1ull << dev;
  const unsigned long long old = __atomic_fetch_or(flag_word, bit, __ATOMIC_ACQ_REL);
  return (old & bit) == 0;
}

struct ProfSlot {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  cudaEvent_t open = nullptr;
};
static bool g_prof = false;
static ProfSlot g_slots[PROF_SLOTS];
bool prof_enabled() { return g_prof; }
void prof_set(bool on) { g_prof = on; }
void prof_begin(int slot, cudaStream_t stream) {
  if (!g_prof) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  g_slots[slot].open = e;
}
void prof_end(int slot, cudaStream_t stream) {
  if (!g_prof || !g_slots[slot].open) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  g_slots[slot].ev.emplace_back(g_slots[slot].open, e);
  g_slots[slot].open = nullptr;
}
int prof_collect(int slot, int* count, double* total_ms) {
  *count = 0;
  *total_ms = 0.0;
  for (auto& pr : g_slots[slot].ev) {
    float ms = 0.f;
    if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
      *count += 1;
      *total_ms += ms;
    }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  g_slots[slot].ev.clear();
  return 0;
}

// Fork / join of independent kernel chains onto an internal side stream (events only: no host synchronisation, and
// the pattern is captured into CUDA graphs as parallel branches).  The per-frame path is a chain of small kernels
// that each fill a fraction of the 148 SMs, so independent sub-chains are run side by side.  VLS_NO_SIDE_STREAM=1
// keeps everything on the caller's stream.
struct Fork {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_step[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
bool overlap_enabled() {
  static const bool on = !(getenv("VLS_NO_SIDE_STREAM") && getenv("VLS_NO_SIDE_STREAM")[0] == '1');
  return on;
}
int fork_get(int idx, Fork** out) {
  // per host thread and per device: concurrent callers never share a side stream or its events
  static thread_local Fork forks[16][6];
  int dev = 0;
  VLS_CUDA(cudaGetDevice(&dev));
  VLS_REQUIRE(dev >= 0 && dev < 16, "device index %d out of range", dev);
  Fork& f = forks[dev][idx];
  if (!f.side) {
    // Fork 2 (self-attention value projection) sits on the frame's critical path: its stream gets the high priority the frame
    // graph's main chain is captured with (graphed.py), the other forks (0: memory K projections = 450 CTAs per layer, 1: decoder
    // helpers, 3: hole filling / output stage, 4: object-pointer MLP) the default, lowest one.  With every fork at the
    // default priority the value projection of layer 0 queued behind the K projection's CTAs and the first self-attention
    // started ~26 us late (warm timeline: 68 us into the frame instead of ~43); fork 1 at high priority let the image-side
    // GEMM take the SMs the decoder's 8-CTA cluster kernels need (decoder +16 us).
    int least = 0, greatest = 0;
    VLS_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const int prio = idx == 2 ? std::max(greatest, -1) : least;
    VLS_CUDA(cudaStreamCreateWithPriority(&f.side, cudaStreamNonBlocking, prio));
    VLS_CUDA(cudaEventCreateWithFlags(&f.ev_fork, cudaEventDisableTiming));
    VLS_CUDA(cudaEventCreateWithFlags(&f.ev_join, cudaEventDisableTiming));
  }
  *out = &f;
  return 0;
}
// returns the stream the forked chain must be launched on (the caller's own stream when overlap is disabled)
int fork_begin(int idx, cudaStream_t main, cudaStream_t* side) {
  *side = main;
  if (!overlap_enabled()) return 0;
  Fork* f;
  VLS_TRY(fork_get(idx, &f));
  VLS_CUDA(cudaEventRecord(f->ev_fork, main));
  VLS_CUDA(cudaStreamWaitEvent(f->side, f->ev_fork, 0));
  *side = f->side;
  return 0;
}
// milestones inside a forked chain: fork_mark records milestone k on the side stream, fork_wait makes `main` wait for it
// (the memory K projections of the four layers run as one forked chain, and layer l only needs the l-th of them)
int fork_mark(int idx, int k) {
  if (!overlap_enabled()) return 0;
  VLS_REQUIRE(k >= 0 && k < 8, "fork_mark: milestone %d out of range", k);
  Fork* f;
  VLS_TRY(fork_get(idx, &f));
  if (!f->ev_step[k]) VLS_CUDA(cudaEventCreateWithFlags(&f->ev_step[k], cudaEventDisableTiming));
  VLS_CUDA(cudaEventRecord(f->ev_step[k], f->side));
  return 0;
}
int fork_wait(int idx, int k, cudaStream_t main) {
  if (!overlap_enabled()) return 0;
  VLS_REQUIRE(k >= 0 && k < 8, "fork_wait: milestone %d out of range", k);
  Fork* f;
  VLS_TRY(fork_get(idx, &f));
  VLS_REQUIRE(f->ev_step[k] != nullptr, "fork_wait: milestone %d was never marked", k);
  VLS_CUDA(cudaStreamWaitEvent(main, f->ev_step[k], 0));
  return 0;
}
int fork_join(int idx, cudaStream_t main) {
  if (!overlap_enabled()) return 0;
  Fork* f;
  VLS_TRY(fork_get(idx, &f));
  VLS_CUDA(cudaEventRecord(f->ev_join, f->side));
  VLS_CUDA(cudaStreamWaitEvent(main, f->ev_join, 0));
  return 0;
}


static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                   uint64_t row_stride, uint64_t batch_stride, uint32_t box_rows) {
  std::call_once(g_encode_once, resolve_encode);
  VLS_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  VLS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  VLS_REQUIRE((row_stride * 2) % 16 == 0, "TMA row stride (%llu elements) must be a multiple of 8",
              (unsigned long long)row_stride);
  VLS_REQUIRE(batch <= 1 || (batch_stride * 2) % 16 == 0, "TMA batch stride must be a multiple of 8 elements");
  VLS_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  if (batch < 1) batch = 1;
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstr[2] = {row_stride * 2, (batch > 1 ? batch_stride : rows * row_stride) * 2};
  if (gstr[1] % 16 != 0) gstr[1] = (gstr[1] + 15) / 16 * 16;
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu batch=%llu)",
              (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch);
  return 0;
}

// f32 variant: box = (32 cols = 128 B, box_rows, 1), SWIZZLE_128B
int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t row_stride,
                  uint64_t batch_stride, uint32_t box_rows) {
  std::call_once(g_encode_once, resolve_encode);
  VLS_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  VLS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  VLS_REQUIRE((row_stride * 4) % 16 == 0 && (batch <= 1 || (batch_stride * 4) % 16 == 0), "TMA strides must be multiples of 16 bytes");
  VLS_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  if (batch < 1) batch = 1;
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstr[2] = {row_stride * 4, (batch > 1 ? batch_stride : rows * row_stride) * 4};
  cuuint32_t box[3] = {32, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed with CUresult %d", (int)r);
  return 0;
}

int make_tmap_f32_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_w,
                       uint32_t box_h) {
  std::call_once(g_encode_once, resolve_encode);
  VLS_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  VLS_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  VLS_REQUIRE(C >= 4 && C <= 256 && C % 4 == 0 && box_w >= 1 && box_w <= 256 && box_h >= 1 && box_h <= 256, "TMA box out of range");
  cuuint64_t gdim[4] = {C, W, H, B < 1 ? 1 : B};
  cuuint64_t gstr[3] = {C * 4, W * C * 4, H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)C, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32 nhwc) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace vls
