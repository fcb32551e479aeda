// Host-side plumbing shared by the C-ABI entry points: error reporting, CUDA checks, TMA tensor-map
// construction (driver entry point resolved at run time so the library links without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
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
