// Back-to-back FFN of a memory-attention layer in ONE kernel (memory_attention.py:95-98):
//
//     x[m][:] += relu(t[m][:] W1^T + b1) W2^T + b2        t = LN3(x) bf16 [M][256], W1 [2048][256], W2 [256][2048]
//
// r1 ran it as two GEMM launches (17 + 15 us at M = 4096): FFN-1 wrote the [M][2048] bf16 hidden tensor (16.8 MB) and
// FFN-2 read it back 4 x (one pass per 64-column output tile: 98 MB of L2->SM traffic).  Here the hidden activations
// never leave the SM:
//   * a CLUSTER of 4 CTAs owns one 128-row tile; CTA r owns hidden units [512 r, 512 r + 512), in 4 chunks of 128
//   * per chunk:  D1 = t_tile W1c^T (SS MMAs, M128 N128 K16 x 16) -> TMEM;  4 epilogue warps (thread = row) add b1, apply
//     ReLU, round to bf16 and write H back INTO TMEM over D1 (the P-over-S trick of the attention kernel);
//     Y += H W2c^T (TS MMAs, A = H in TMEM, M128 N256 K16 x 8).  D1/H is double-buffered, so GEMM-1 of chunk c+1 and
//     GEMM-2 of chunk c keep the tensor pipe busy while the epilogue warps work on chunk c+1
//   * weights stream through PANEL rings (W1: 6 x [128 x 64] = 16 KB slots, W2: 2 x [256 x 64] = 32 KB slots) with one
//     full/empty mbarrier pair per slot, released by a tcgen05.commit after the 4 MMAs that read the panel, so the
//     producer runs up to 1.5 chunks ahead without a second whole-chunk stage
//   * the four partial Y tiles (f32 [128][256] each) are reduced over DISTRIBUTED SHARED MEMORY in a fixed order (CTA r
//     sums columns [64 r, 64 r + 64) of all four, adds b2 and the residual and writes x): deterministic, no atomics,
//     no partials in global memory.
// TMEM: Y 256 columns | D1/H 2 x 128.  SMEM: t tile 64 KB + W1 ring 96 KB + W2 ring 64 KB = 224 KB (the Y dump for the
// cluster reduce re-uses the t tile + W1 ring once every MMA has completed).
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int C = 256;        // d_model
constexpr int FF = 2048;      // hidden
constexpr int BM = 128;       // rows per cluster
constexpr int CL = 4;         // CTAs per cluster = hidden quarters
constexpr int HPR = FF / CL;  // hidden units per CTA
constexpr int HC = 128;       // hidden chunk
constexpr int NCH = HPR / HC; // chunks per CTA
constexpr int W1_SLOTS = 6, W1_SLOT_BYTES = HC * 64 * 2;    // [128 hidden x 64 channels]
constexpr int W2_SLOTS = 2, W2_SLOT_BYTES = C * 64 * 2;     // [256 outputs x 64 hidden]
constexpr int A_BYTES = BM * C * 2;                         // 4 panels [128 rows x 64 channels]
constexpr int SMEM_BYTES = A_BYTES + W1_SLOTS * W1_SLOT_BYTES + W2_SLOTS * W2_SLOT_BYTES + 256 + 1024;   // 230 656 <= 232 448
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
constexpr int THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue (thread = row)
constexpr uint32_t TM_Y = 0, TM_D1 = 256;
static_assert(BM * C * 4 <= A_BYTES + W1_SLOTS * W1_SLOT_BYTES, "Y dump must fit into the t tile + W1 ring");

struct FfnParams {
  int M;                      // rows per batch element
  const float* b1;
  const float* b2;
  float* x;                   // f32 [B][M][256], updated in place
  long long x_bstride;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic smem base has the same offset in every CTA of the cluster, so this alignment is identical too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW1 = sA + A_BYTES;
  uint8_t* sW2 = sW1 + W1_SLOTS * W1_SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW2 + W2_SLOTS * W2_SLOT_BYTES);
  uint64_t* a_full = bars;                 // 1
  uint64_t* w1_full = bars + 1;            // W1_SLOTS
  uint64_t* w1_empty = w1_full + W1_SLOTS; // W1_SLOTS
  uint64_t* w2_full = w1_empty + W1_SLOTS; // W2_SLOTS
  uint64_t* w2_empty = w2_full + W2_SLOTS; // W2_SLOTS
  uint64_t* d1_full = w2_empty + W2_SLOTS; // 2
  uint64_t* h_ready = d1_full + 2;         // 2
  uint64_t* y_full = h_ready + 2;          // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * BM;
  const int bz = blockIdx.z;
  const int hbase = (int)rank * HPR;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < W1_SLOTS; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
    for (int s = 0; s < W2_SLOTS; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&d1_full[s], 1); mbar_init(&h_ready[s], 128); }
    mbar_init(y_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_enter();   // global memory from here on

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, A_BYTES);
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) tma_load_3d(sA + kp * (BM * 128), &tmA, a_full, kp * 64, m0, bz);
      // panel streams in consumption order: chunk 0 W1 panels, [chunk c+1 W1 panels, chunk c W2 panels] ...
      int i1 = 0, i2 = 0;
      auto load_w1 = [&](int c) {
#pragma unroll 1
        for (int kp = 0; kp < 4; ++kp, ++i1) {
          const int s = i1 % W1_SLOTS;
          mbar_wait(&w1_empty[s], ((i1 / W1_SLOTS) & 1) ^ 1);
          mbar_expect_tx(&w1_full[s], W1_SLOT_BYTES);
          tma_load_3d(sW1 + s * W1_SLOT_BYTES, &tmW1, &w1_full[s], kp * 64, hbase + c * HC, 0);
        }
      };
      auto load_w2 = [&](int c) {
#pragma unroll 1
        for (int hp = 0; hp < 2; ++hp, ++i2) {
          const int s = i2 % W2_SLOTS;
          mbar_wait(&w2_empty[s], ((i2 / W2_SLOTS) & 1) ^ 1);
          mbar_expect_tx(&w2_full[s], W2_SLOT_BYTES);
          tma_load_3d(sW2 + s * W2_SLOT_BYTES, &tmW2, &w2_full[s], hbase + c * HC + hp * 64, 0, 0);
        }
      };
      load_w1(0);
      for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) load_w1(c + 1);
        load_w2(c);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc_bf16(BM, HC);
      constexpr uint32_t idesc2 = make_idesc_bf16(BM, C);
      const uint32_t a_addr = smem_u32(sA);
      int i1 = 0, i2 = 0;
      auto gemm1 = [&](int c) {      // D1[c&1] = t_tile . W1c^T
        const uint32_t d = tmem + TM_D1 + uint32_t(c & 1) * HC;
#pragma unroll 1
        for (int kp = 0; kp < 4; ++kp, ++i1) {
          const int s = i1 % W1_SLOTS;
          mbar_wait(&w1_full[s], (i1 / W1_SLOTS) & 1);
          tc_fence_after();
          const uint64_t ad = make_desc_sw128(a_addr + kp * (BM * 128));
          const uint64_t bd = make_desc_sw128(smem_u32(sW1 + s * W1_SLOT_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(d, ad + 2 * kk, bd + 2 * kk, idesc1, (kp | kk) != 0 ? 1u : 0u);
          umma_commit(&w1_empty[s]);
        }
        umma_commit(&d1_full[c & 1]);
      };
      auto gemm2 = [&](int c) {      // Y += H[c&1] . W2c^T   (A = bf16 H in TMEM over D1[c&1])
        mbar_wait(&h_ready[c & 1], (c >> 1) & 1);
        const uint32_t a_h = tmem + TM_D1 + uint32_t(c & 1) * HC;
#pragma unroll 1
        for (int hp = 0; hp < 2; ++hp, ++i2) {
          const int s = i2 % W2_SLOTS;
          mbar_wait(&w2_full[s], (i2 / W2_SLOTS) & 1);
          tc_fence_after();
          const uint64_t bd = make_desc_sw128(smem_u32(sW2 + s * W2_SLOT_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + TM_Y, a_h + (hp * 4 + kk) * 8, bd + 2 * kk, idesc2, (c | hp | kk) != 0 ? 1u : 0u);
          umma_commit(&w2_empty[s]);
        }
      };
      mbar_wait(a_full, 0);
      tc_fence_after();
      gemm1(0);
      for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) gemm1(c + 1);   // tensor pipe works on chunk c+1 while the epilogue warps turn D1(c) into H(c)
        gemm2(c);
      }
      umma_commit(y_full);
    }
  } else {
    // ---- epilogue warps: thread = row (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    for (int c = 0; c < NCH; ++c) {
      const int b = c & 1;
      mbar_wait(&d1_full[b], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t base = tmem + lane_off + TM_D1 + uint32_t(b) * HC;
      const float* bias = p.b1 + hbase + c * HC;   // warp-uniform addresses: broadcast loads, L1-resident
#pragma unroll 1
      for (int k = 0; k < HC / 32; ++k) {   // reads columns [32k, 32k+32), writes packed bf16 to [16k, 16k+16): in place
        uint32_t r[32], h[16];
        tmem_ld32(base + k * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + k * 32 + 2 * i));
          h[i] = pack_bf16x2(fmaxf(__uint_as_float(r[2 * i]) + bb.x, 0.f), fmaxf(__uint_as_float(r[2 * i + 1]) + bb.y, 0.f));
        }
        tmem_st16(base + k * 16, h);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&h_ready[b]);
    }
    // ---- Y partial -> shared memory as [column group of 4][row] float4 (conflict-free for thread = row)
    mbar_wait(y_full, 0);        // every MMA has completed: the t tile and the W1 ring are dead
    tc_fence_after();
    float4* dump = reinterpret_cast<float4*>(smem);
#pragma unroll 1
    for (int k = 0; k < C / 32; ++k) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_off + TM_Y + k * 32, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dump[(k * 8 + i) * BM + rl] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                  __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
    }
  }
  tc_fence_before();
  cluster_sync_all();            // all four partial tiles are in shared memory
  if (warp >= 2) {
    // CTA r reduces output columns [64 r, 64 r + 64) over the four CTAs in rank order (deterministic), adds b2 and the
    // residual, and writes x.  thread = row: 16 float4 column groups.
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const int row = m0 + rl;
    const uint32_t my = smem_u32(smem);
    uint32_t peer[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) peer[r] = mapa_u32(my, (uint32_t)r);
    float* xr = p.x + (long long)bz * p.x_bstride + (long long)row * C + rank * 64;
    const float* b2 = p.b2 + rank * 64;
#pragma unroll 1
    for (int g = 0; g < 16; g += 4) {
      float4 acc[4], xin[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = uint32_t(((rank * 16 + g + i) * BM + rl) * 16);
        acc[i] = ld_dsmem_f4(peer[0] + off);
#pragma unroll
        for (int r = 1; r < CL; ++r) {
          const float4 v = ld_dsmem_f4(peer[r] + off);
          acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
        }
        if (row < p.M) xin[i] = *reinterpret_cast<const float4*>(xr + (g + i) * 4);
      }
      if (row < p.M) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = *reinterpret_cast<const float4*>(b2 + (g + i) * 4);
          *reinterpret_cast<float4*>(xr + (g + i) * 4) = make_float4(acc[i].x + bb.x + xin[i].x, acc[i].y + bb.y + xin[i].y,
                                                                     acc[i].z + bb.z + xin[i].z, acc[i].w + bb.w + xin[i].w);
        }
      }
    }
  }
  cluster_sync_all();            // peers may still be reading this CTA's partial tile
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

int g_ffn_fused = 1;   // memory attention: 1 = this kernel, 0 = two GEMM launches (vls_set_tuning "ffn_fused")

// x[b][m][:] += relu(t[b][m][:] W1^T + b1) W2^T + b2;  t bf16 [B][M][256] (row stride ldt), W1 bf16 [2048][256],
// W2 bf16 [256][2048], b1 f32 [2048], b2 f32 [256], x f32 [B][M][256] contiguous rows.
int launch_ffn_fused(const void* t, long long ldt, long long t_bstride, const void* w1, const float* b1, const void* w2,
                     const float* b2, float* x, long long x_bstride, int B, int M, cudaStream_t stream) {
  VLS_REQUIRE(t && w1 && b1 && w2 && b2 && x && B > 0 && M > 0, "ffn_fused: bad arguments");
  VLS_REQUIRE(ldt % 8 == 0, "ffn_fused: ldt must be a multiple of 8");
  CUtensorMap tmA, tmW1, tmW2;
  VLS_TRY(make_tmap_bf16(&tmA, t, C, M, B, ldt, t_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmW1, w1, C, FF, 1, C, (long long)FF * C, HC));
  VLS_TRY(make_tmap_bf16(&tmW2, w2, FF, C, 1, FF, (long long)FF * C, C));
  FfnParams p;
  p.M = M; p.b1 = b1; p.b2 = b2; p.x = x; p.x_bstride = x_bstride;
  static unsigned long long attr_set = 0;
  if (first_use_on_device(&attr_set))
    VLS_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  VLS_CUDA(launch_k(ffn_fused_kernel, dim3(CL, (M + BM - 1) / BM, B), dim3(THREADS), SMEM_BYTES, stream, tmA, tmW1, tmW2, p));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace vls
