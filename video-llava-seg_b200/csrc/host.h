// Host-side plumbing shared by the C-ABI entry points: error reporting, CUDA checks, TMA tensor-map
// construction (driver entry point resolved at run time so the library links without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <utility>
#include <cstdio>

namespace vls {

void set_error(const char* fmt, ...);
const char* last_error();

#define VLS_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      vls::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

// Every kernel launch site reports itself (bench.py's `gpu_launches`) and checks the launch status.
void count_launches(int n);
long long launch_count();
#define VLS_POST_LAUNCH(n)               \
  do {                                   \
    vls::count_launches(n);              \
    VLS_CUDA(cudaGetLastError());        \
  } while (0)

// Programmatic dependent launch (PDL): every kernel of the library starts with griddepcontrol.wait (pdl_enter() in
// common.cuh), so a launch may be released as soon as all CTAs of the previous kernel in the stream have STARTED; the
// dependent grid's launch latency, CTA scheduling and parameter/descriptor fetch then overlap the previous grid's tail,
// and its CTAs block at the wait until that grid has completed and flushed.  The per-frame path is a chain of ~190
// mostly 3-25 us kernels.  Captured into CUDA graphs as programmatic edges.  MEASURED (B200, r1, whole frame as one graph
// replay): 1.7245 ms/frame with PDL vs 1.6988 ms without -- graph replay already hides the launch gaps, what remains
// is each small kernel's own ramp/drain -- so it is OFF by default; VLS_PDL=1 or vls_set_tuning("pdl", 1) enables it
// (useful for the eager, un-graphed path).
bool pdl_enabled();
void pdl_set(bool on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Fork / join of independent kernel chains onto internal side streams (events only, graph-capturable; per host thread
// and device).  fork_begin returns the stream the forked chain must be launched on (the caller's own stream when
// VLS_NO_SIDE_STREAM=1); fork_join makes `main` wait for everything enqueued on that side stream.  idx 0..3 select
// independent side streams.
int fork_begin(int idx, cudaStream_t main, cudaStream_t* side);
int fork_join(int idx, cudaStream_t main);
// milestones inside a forked chain: fork_mark(idx, k) records milestone k on the side stream, fork_wait makes `main` wait for it
int fork_mark(int idx, int k);
int fork_wait(int idx, int k, cudaStream_t main);

// One-time per-DEVICE initialisation (cudaFuncSetAttribute is a per-device setting): returns true the first time it is
// called with this flag word on the current device.  Thread-safe.
bool first_use_on_device(unsigned long long* flag_word);

// Optional per-kernel device timing (CUDA events on the launching stream), off by default.
// Used by bench.py to measure the dominant kernel's average launch duration live.
bool prof_enabled();
void prof_begin(int slot, cudaStream_t stream);
void prof_end(int slot, cudaStream_t stream);
enum { PROF_ATTN_CROSS = 0, PROF_ATTN_SELF = 1, PROF_SLOTS = 8 };
void prof_set(bool on);
int prof_collect(int slot, int* count, double* total_ms);

#define VLS_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      vls::set_error(__VA_ARGS__);  \
      return 1;                     \
    }                               \
  } while (0)

#define VLS_TRY(expr)        \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

// 3-D bf16 tensor map: dims (inner=cols, rows, batch), strides in ELEMENTS for rows/batch,
// box = (64 cols = 128 B, box_rows, 1), SWIZZLE_128B, zero fill out of bounds.
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                   uint64_t row_stride, uint64_t batch_stride, uint32_t box_rows);

int make_tmap_f32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t row_stride,
                  uint64_t batch_stride, uint32_t box_rows);

// 4-D f32 tensor map over a channel-last image batch [B][H][W][C] (C <= 256): box = (C, box_w, box_h, 1), no swizzle,
// zero fill out of bounds (coordinates may be negative: the halo of a convolution).
int make_tmap_f32_nhwc(CUtensorMap* out, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_w,
                       uint32_t box_h);

struct Workspace {  // bump allocator over a caller-owned device buffer
  char* base;
  size_t size, off;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  void* take(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    if (a + bytes > size) return nullptr;
    off = a + bytes;
    return base + a;
  }
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace vls
