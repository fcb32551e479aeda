#include "host.h"

#include <cudaTypedefs.h>

#include <cstring>
#include <mutex>

namespace vls {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

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

}  // namespace vls
