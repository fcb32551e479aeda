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
//   * the four partial Y tiles (f32 [128][256] each) are reduced over DISTRIBUTED SHARED MEMORY: every CTA pushes the
//     64-column slice owned by CTA o into o's shared memory (st.shared::cluster), then CTA r sums its four slices in rank
//     order, adds b2 and the residual and writes x: deterministic, no atomics, no partials in global memory.
// Optional FRONT (the whole tail of a memory-attention layer in one launch, memory_attention.py:76-98): the CTA first
// computes  x_mid = x_in + ao (Wo Wv)^T + b  (the folded cross-attention output projection, K = 64, 4 MMAs into the Y
// columns) and t = LayerNorm3(x_mid) itself -- thread = row, so the statistics are thread-local -- and writes t as the
// bf16 A operand straight into shared memory in the 128B-swizzled K-major layout TMA would have produced; every CTA of
// the cluster does this redundantly (16 KB + 32 KB of operands), CTA r parks its 64-column slice of x_mid in x_out.
// Optional BACK: the LayerNorm that follows (next layer's norm1 or the final norm): every CTA holds a 64-column slice of
// the new rows, so the four exchange per-row (sum, sum of squares) over distributed shared memory and each normalises its
// slice.  A layer tail that was 6 launches (out-proj GEMM, LN, FFN-1, FFN-2, next LN + the gaps between) becomes one.
// TMEM: Y 256 columns | D1/H 2 x 128.  SMEM: t tile 64 KB + W1 ring 96 KB + W2 ring 64 KB = 224 KB (the Y dump for the
// cluster reduce re-uses the t tile + W1 ring once every MMA has completed).
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int C = 256;        // d_model
constexpr int BM = 128;       // rows per cluster
constexpr int CL = 4;         // CTAs per cluster = hidden quarters
constexpr int HC = 128;       // hidden chunk
// hidden width FF = CL * NCH * HC: NCH = 4 chunks per CTA for the memory-attention FFN (2048), 2 for the CXBlock of the
// memory encoder's fuser (256 -> 1024 -> 256 with GELU, memory_encoder.py:86-100)
constexpr int W1_SLOTS = 6, W1_SLOT_BYTES = HC * 64 * 2;    // [128 hidden x 64 channels]
constexpr int W2_SLOTS = 2, W2_SLOT_BYTES = C * 64 * 2;     // [256 outputs x 64 hidden]
constexpr int A_BYTES = BM * C * 2;                         // 4 panels [128 rows x 64 channels]
constexpr int SMEM_BYTES = A_BYTES + W1_SLOTS * W1_SLOT_BYTES + W2_SLOTS * W2_SLOT_BYTES + 256 + 1024;   // 230 656 <= 232 448
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
constexpr int THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue (thread = row)
constexpr uint32_t TM_Y = 0, TM_D1 = 256;
// landing zone of the cluster reduce in every CTA: [source rank][column group of 4][row] float4 with one float4 of padding
// per column group: the PUSH (thread = row, fixed column group) writes 512 contiguous bytes per warp -- remote stores
// that scatter 32 x 16 B cost 2.2 x as much (measured) -- and the (row, column group) reads of the reduce hit 8 distinct
// 16-byte bank groups per 8 lanes
constexpr int SL_PITCH = BM * 16 + 16;            // bytes per column group of a slice
constexpr int SL_BYTES = 16 * SL_PITCH;           // one source rank's 64-column slice
static_assert(CL * SL_BYTES <= A_BYTES + W1_SLOTS * W1_SLOT_BYTES, "reduce landing zone must fit into the t tile + W1 ring");
static_assert(BM * C * 4 <= A_BYTES + 4 * W1_SLOT_BYTES, "FRONT: the f32 x tile is staged in the t tile + 4 W1 slots");

struct FfnParams {
  int M;                      // rows per batch element
  const float* b1;
  const float* b2;
  const float* x_in;          // f32 [B][M][256] residual input
  float* x_out;               // f32 [B][M][256] result (== x_in unless FRONT, which needs a different buffer)
  long long x_bstride;
  // FRONT
  const float* b0;            // [256] bias of the folded output projection
  const float* ln_w; const float* ln_b; float ln_eps;       // LayerNorm3
  // BACK
  const float* ln2_w; const float* ln2_b; float ln2_eps;    // the LayerNorm applied to the result
  void* t_out; int t_out_bf16; long long t_out_st, t_out_sb;  // element (b, row, c) at b*sb + row*st + c
  long long* trace;           // optional dev trace: clock64 stamps of the first epilogue thread of CTA (0,0,0)
};

#define FFN_TRACE(slot)                                                                              \
  do {                                                                                               \
    if (p.trace && threadIdx.x == 64 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)       \
      p.trace[slot] = clock64();                                                                     \
  } while (0)

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

__device__ __forceinline__ void st_dsmem_f4(uint32_t cluster_addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(cluster_addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t cluster_addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}

// FRONT != 0: tmA = ao [B][M][64] (box 128 x 64), tmW0 = folded out-proj weight [256][64] (box 256 x 64, or 64 x 64 for
// FRONT == 2); else tmA = t [B][M][256].
// FRONT == 1: every CTA computes the full-width x_mid and LayerNorm3 itself (13.6 k cycles: 128 KB residual tile per CTA, two
//             passes over 256 TMEM columns per thread, thread = row global stores of the parked slice).
// FRONT == 2: the column-quarter pattern of mid_fused.cu: CTA r computes x_mid columns [64 r, 64 r + 64) (4 MMAs, N = 64), the
//             per-row statistics go over DSMEM, its normalised 64 columns are one K panel of the t operand and are
//             all-gathered with bulk shared -> shared::cluster copies; the new residual slice leaves through a TMA store.
//             The FRONT operands live in the (still idle) W2 ring, so the W1 ring streams from the first cycle.
template <int FRONT, bool BACK, int NCH, bool GELU>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXo, const FfnParams p) {
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
  uint64_t* g0_full = y_full + 1;          // FRONT: ao tile + folded out-proj weight have landed
  uint64_t* g0_done = g0_full + 1;         // FRONT: the out-proj MMAs have completed (x_mid partial in TMEM, W2 ring free)
  uint64_t* a_ready = g0_done + 1;         // FRONT: 128 epilogue threads have written t = LN(x_mid) into sA
  uint64_t* x_full = a_ready + 1;          // FRONT: the f32 residual tile has landed in its staging area (sA + 4 W1 slots)
  uint64_t* x_free = x_full + 1;           // FRONT: ... and has been consumed: the W1 ring (FRONT == 2: the W2 ring) may be filled
  uint64_t* t_full = x_free + 1;           // FRONT == 2: the three remote panels of the t tile have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = blockIdx.y * BM;
  const int bz = blockIdx.z;
  constexpr int HPR = NCH * HC;   // hidden units per CTA
  const int hbase = (int)rank * HPR;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < W1_SLOTS; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
    for (int s = 0; s < W2_SLOTS; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&d1_full[s], 1); mbar_init(&h_ready[s], 128); }
    mbar_init(y_full, 1);
    mbar_init(g0_full, 1);
    mbar_init(g0_done, 1);
    mbar_init(a_ready, 128);
    mbar_init(x_full, 1);
    mbar_init(x_free, 1);
    mbar_init(t_full, 1);
    if (FRONT == 2) mbar_expect_tx(t_full, (CL - 1) * BM * 128);   // armed before any peer can send
    fence_barrier_init();
    tma_prefetch_desc(&tmW0);
    if (FRONT) tma_prefetch_desc(&tmX);
    if (FRONT == 2) tma_prefetch_desc(&tmXo);
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
  FFN_TRACE(0);

  // FRONT == 2 has a cluster barrier in the middle of the epilogue warps' prologue: the producer / MMA warps arrive here and
  // consume the phase right before the next cluster barrier
  if (FRONT == 2 && warp < 2) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  if (warp == 0) {
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
      if (FRONT == 2) {  // W2 ring slot 0: [W0 slice 64 x 64 | ao tile 128 x 64 | statistics | parameters], slot 1: residual slice f32
        mbar_expect_tx(g0_full, 64 * 128 + BM * 128);
        tma_load_3d(sW2, &tmW0, g0_full, 0, (int)rank * 64, 0);
        tma_load_3d(sW2 + 64 * 128, &tmA, g0_full, 0, m0, bz);
        mbar_expect_tx(x_full, BM * 64 * 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) tma_load_3d(sW2 + W2_SLOT_BYTES + k * (BM * 128), &tmX, x_full, (int)rank * 64 + 32 * k, m0, bz);
      } else if (FRONT) {   // staged in the (still idle) W2 ring: slot 0 = folded out-proj weight [256 x 64], slot 1 = ao tile [128 x 64]
        mbar_expect_tx(g0_full, W2_SLOT_BYTES + BM * 128);
        tma_load_3d(sW2, &tmW0, g0_full, 0, 0, 0);
        tma_load_3d(sW2 + W2_SLOT_BYTES, &tmA, g0_full, 0, m0, bz);
        // residual tile x_in[m0 : m0+128][0:256] f32 = 8 boxes of [128 rows x 32 floats] (128B-swizzled): coalesced and
        // asynchronous (thread = row loads from global memory cost 12.6 k cycles in the first version of the prologue)
        mbar_expect_tx(x_full, BM * C * 4);
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) tma_load_3d(smem + kb * (BM * 128), &tmX, x_full, kb * 32, m0, bz);
      } else {
        mbar_expect_tx(a_full, A_BYTES);
#pragma unroll
        for (int kp = 0; kp < 4; ++kp) tma_load_3d(sA + kp * (BM * 128), &tmA, a_full, kp * 64, m0, bz);
      }
      // panel streams in consumption order: chunk 0 W1 panels, [chunk c+1 W1 panels, chunk c W2 panels] ...
      int i1 = 0, i2 = 0;
      if (FRONT == 1) mbar_wait(x_free, 0);
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
        if (FRONT == 1 && c == 0) mbar_wait(g0_done, 0);   // the out-proj operands have been consumed: the W2 ring is free
        if (FRONT == 2 && c == 0) mbar_wait(x_free, 0);    // ... and so have the statistics, parameters and the residual slice
        load_w2(c);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // elect.sync: a lane test makes the compiler wrap every TMA / tcgen05 instruction in an ELECT + BRA.U.ANY loop
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
      if (FRONT == 2) {
        mbar_wait(g0_full, 0);
        tc_fence_after();
        constexpr uint32_t idesc0 = make_idesc_bf16(BM, 64);
        const uint64_t ad = make_desc_sw128(smem_u32(sW2 + 64 * 128)), bd = make_desc_sw128(smem_u32(sW2));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + TM_Y, ad + 2 * kk, bd + 2 * kk, idesc0, kk != 0 ? 1u : 0u);
        umma_commit(g0_done);
        mbar_wait(a_ready, 0);     // own panel of t = LN(x_mid) written ...
        mbar_wait(t_full, 0);      // ... and the three remote ones have landed
      } else if (FRONT) {
        mbar_wait(g0_full, 0);
        tc_fence_after();
        const uint64_t ad = make_desc_sw128(smem_u32(sW2 + W2_SLOT_BYTES)), bd = make_desc_sw128(smem_u32(sW2));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + TM_Y, ad + 2 * kk, bd + 2 * kk, idesc2, kk != 0 ? 1u : 0u);
        umma_commit(g0_done);
        mbar_wait(a_ready, 0);     // t = LN(x_mid) is in sA (generic-proxy writes, fenced by the writers)
      } else {
        mbar_wait(a_full, 0);
      }
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
    if (FRONT == 2) {
      // ---- column quarter: x_mid[:, 64 r : 64 r + 64] = x_in + ao (Wo Wv)_r^T + b0
      float* prm = reinterpret_cast<float*>(sW2 + 64 * 128 + BM * 128 + CL * BM * 8);   // [b0 | ln_w | ln_b] slices
      float2* stats = reinterpret_cast<float2*>(sW2 + 64 * 128 + BM * 128);             // [source rank][row]
      uint8_t* sXq = sW2 + W2_SLOT_BYTES;                                                // residual slice: 2 boxes [128 x 32 f32]
      const int n0 = (int)rank * 64;
      {
        const int i = threadIdx.x - 64;   // 0..127: 192 floats as 96 float2
        if (i < 96) {
          const int which = i >> 5, c2 = (i & 31) * 2;
          const float* src = which == 0 ? p.b0 : which == 1 ? p.ln_w : p.ln_b;
          *reinterpret_cast<float2*>(&prm[which * 64 + c2]) = __ldg(reinterpret_cast<const float2*>(src + n0 + c2));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(g0_done, 0);
      mbar_wait(x_full, 0);
      tc_fence_after();
      FFN_TRACE(1);
      float xm[64];
      float sum = 0.f, ss = 0.f;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_off + TM_Y + k * 32, r);
        uint8_t* xs = sXq + k * (BM * 128) + rl * 128;
        float4 xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + ((i ^ (rl & 7)) << 4));
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = *reinterpret_cast<const float4*>(&prm[k * 32 + 4 * i]);
          float4 m;
          m.x = __uint_as_float(r[4 * i]) + bb.x + xv[i].x; m.y = __uint_as_float(r[4 * i + 1]) + bb.y + xv[i].y;
          m.z = __uint_as_float(r[4 * i + 2]) + bb.z + xv[i].z; m.w = __uint_as_float(r[4 * i + 3]) + bb.w + xv[i].w;
          sum += (m.x + m.y) + (m.z + m.w);
          ss += (m.x * m.x + m.y * m.y) + (m.z * m.z + m.w * m.w);
          xm[k * 32 + 4 * i] = m.x; xm[k * 32 + 4 * i + 1] = m.y; xm[k * 32 + 4 * i + 2] = m.z; xm[k * 32 + 4 * i + 3] = m.w;
          *reinterpret_cast<float4*>(xs + ((i ^ (rl & 7)) << 4)) = m;   // parked slice of x_mid: out through a TMA store
        }
      }
      fence_proxy_async();
      {
        const uint32_t dst = smem_u32(&stats[rank * BM + rl]);
#pragma unroll
        for (int r = 0; r < CL; ++r)
          asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(mapa_u32(dst, (uint32_t)r)), "f"(sum), "f"(ss) : "memory");
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
#pragma unroll
        for (int k = 0; k < 2; ++k) tma_store_3d(sXq + k * (BM * 128), &tmXo, n0 + 32 * k, m0, bz);
        tma_store_commit();
      }
      FFN_TRACE(2);
      tc_fence_before();
      cluster_sync_all();   // statistics of the four column quarters are here (warps 0 / 1 arrived at kernel start)
      {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int r = 0; r < CL; ++r) { const float2 q2 = stats[r * BM + rl]; s1 += q2.x; s2 += q2.y; }   // same order everywhere
        const float mean = s1 * (1.0f / C);
        const float rstd = rsqrtf(fmaxf(s2 * (1.0f / C) - mean * mean, 0.f) + p.ln_eps);
        uint8_t* prow = sA + rank * (BM * 128) + rl * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wa = *reinterpret_cast<const float4*>(&prm[64 + 8 * j]), wb = *reinterpret_cast<const float4*>(&prm[64 + 8 * j + 4]);
          const float4 ba = *reinterpret_cast<const float4*>(&prm[128 + 8 * j]), bb = *reinterpret_cast<const float4*>(&prm[128 + 8 * j + 4]);
          const uint32_t p0 = pack_bf16x2((xm[8 * j] - mean) * rstd * wa.x + ba.x, (xm[8 * j + 1] - mean) * rstd * wa.y + ba.y);
          const uint32_t p1 = pack_bf16x2((xm[8 * j + 2] - mean) * rstd * wa.z + ba.z, (xm[8 * j + 3] - mean) * rstd * wa.w + ba.w);
          const uint32_t p2 = pack_bf16x2((xm[8 * j + 4] - mean) * rstd * wb.x + bb.x, (xm[8 * j + 5] - mean) * rstd * wb.y + bb.y);
          const uint32_t p3 = pack_bf16x2((xm[8 * j + 6] - mean) * rstd * wb.z + bb.z, (xm[8 * j + 7] - mean) * rstd * wb.w + bb.w);
          *reinterpret_cast<uint4*>(prow + ((j ^ (rl & 7)) << 4)) = make_uint4(p0, p1, p2, p3);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(a_ready);
      asm volatile("bar.sync 1, 128;" ::: "memory");   // panel complete; statistics and parameters no longer needed
      if (threadIdx.x == 64) {
        const uint32_t src = smem_u32(sA + rank * (BM * 128)), bar = smem_u32(t_full);
#pragma unroll
        for (int d = 1; d < CL; ++d) {
          const uint32_t peer = (rank + d) % CL;
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(mapa_u32(src, peer)), "r"(src), "r"(BM * 128), "r"(mapa_u32(bar, peer)) : "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the parked slice has been written (the reduce re-reads it)
        mbar_arrive(x_free);                                         // the W2 ring is free
      }
      FFN_TRACE(3);
    } else if (FRONT) {
      // x_mid = x_in + ao (Wo Wv)^T + b0;  t = LN3(x_mid) -> sA (bf16, 128B-swizzled K-major panels)
      const int row = m0 + rl;
      const bool row_ok = row < p.M;
      float* xpark = p.x_out + (long long)bz * p.x_bstride + (long long)row * C;
      mbar_wait(g0_done, 0);
      mbar_wait(x_full, 0);
      tc_fence_after();
      FFN_TRACE(1);
      float sum = 0.f, ss = 0.f;
#pragma unroll 1
      for (int k = 0; k < C / 32; ++k) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_off + TM_Y + k * 32, r);
        const uint8_t* xrow = smem + k * (BM * 128) + rl * 128;   // staged residual: box k, this thread's row
        float4 xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(xrow + ((i ^ (rl & 7)) << 4));
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b0 + k * 32 + 4 * i));
          float4 m;
          m.x = __uint_as_float(r[4 * i]) + bb.x + xv[i].x; m.y = __uint_as_float(r[4 * i + 1]) + bb.y + xv[i].y;
          m.z = __uint_as_float(r[4 * i + 2]) + bb.z + xv[i].z; m.w = __uint_as_float(r[4 * i + 3]) + bb.w + xv[i].w;
          sum += (m.x + m.y) + (m.z + m.w);
          ss += (m.x * m.x + m.y * m.y) + (m.z * m.z + m.w * m.w);
          r[4 * i] = __float_as_uint(m.x); r[4 * i + 1] = __float_as_uint(m.y);
          r[4 * i + 2] = __float_as_uint(m.z); r[4 * i + 3] = __float_as_uint(m.w);
          if (row_ok && (k >> 1) == (int)rank) *reinterpret_cast<float4*>(xpark + k * 32 + 4 * i) = m;   // own slice of x_mid
        }
        tmem_st32(tmem + lane_off + TM_Y + k * 32, r);
      }
      tc_wait_st();
      // every thread is done with the staged residual: pass 2 overwrites the first 64 KB of it (sA), the producer the rest
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) mbar_arrive(x_free);
      FFN_TRACE(2);
      const float mean = sum * (1.0f / C);
      const float rstd = rsqrtf(fmaxf(ss * (1.0f / C) - mean * mean, 0.f) + p.ln_eps);
#pragma unroll 2
      for (int k = 0; k < C / 32; ++k) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_off + TM_Y + k * 32, r);
        float4 w4[8], b4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          w4[i] = __ldg(reinterpret_cast<const float4*>(p.ln_w + k * 32 + 4 * i));
          b4[i] = __ldg(reinterpret_cast<const float4*>(p.ln_b + k * 32 + 4 * i));
        }
        tc_wait_ld();
        uint8_t* prow = sA + (k >> 1) * (BM * 128) + rl * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 wa = w4[2 * j], wb = w4[2 * j + 1], ba = b4[2 * j], bbv = b4[2 * j + 1];
          const uint32_t p0 = pack_bf16x2((__uint_as_float(r[8 * j]) - mean) * rstd * wa.x + ba.x, (__uint_as_float(r[8 * j + 1]) - mean) * rstd * wa.y + ba.y);
          const uint32_t p1 = pack_bf16x2((__uint_as_float(r[8 * j + 2]) - mean) * rstd * wa.z + ba.z, (__uint_as_float(r[8 * j + 3]) - mean) * rstd * wa.w + ba.w);
          const uint32_t p2 = pack_bf16x2((__uint_as_float(r[8 * j + 4]) - mean) * rstd * wb.x + bbv.x, (__uint_as_float(r[8 * j + 5]) - mean) * rstd * wb.y + bbv.y);
          const uint32_t p3 = pack_bf16x2((__uint_as_float(r[8 * j + 6]) - mean) * rstd * wb.z + bbv.z, (__uint_as_float(r[8 * j + 7]) - mean) * rstd * wb.w + bbv.w);
          const int c16 = (k & 1) * 4 + j;   // 16-byte chunk inside the 128-byte panel row, XOR-swizzled by the row
          *reinterpret_cast<uint4*>(prow + ((c16 ^ (rl & 7)) << 4)) = make_uint4(p0, p1, p2, p3);
        }
      }
      fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      mbar_arrive(a_ready);
      FFN_TRACE(3);
    }
    for (int c = 0; c < NCH; ++c) {
      const int b = c & 1;
      mbar_wait(&d1_full[b], (c >> 1) & 1);
      tc_fence_after();
      FFN_TRACE(4 + c);
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
          const float h0 = __uint_as_float(r[2 * i]) + bb.x, h1 = __uint_as_float(r[2 * i + 1]) + bb.y;
          h[i] = GELU ? pack_bf16x2(gelu_erf(h0), gelu_erf(h1)) : pack_bf16x2(fmaxf(h0, 0.f), fmaxf(h1, 0.f));
        }
        tmem_st16(base + k * 16, h);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&h_ready[b]);
    }
    // ---- Y partial: PUSHED to the owners.  Columns [64 o, 64 o + 64) go into CTA o's shared memory, region [this rank], as
    //      [column group of 4][row] float4 (remote stores are fire-and-forget; the first version PULLED with 64 dependent
    //      distributed-shared-memory loads per thread, ~200 cycles each: 45 % of the kernel's stall samples in ncu)
    FFN_TRACE(8);
    mbar_wait(y_full, 0);        // every MMA of THIS CTA has completed ...
    tc_fence_after();
    FFN_TRACE(9);
  }
  if (FRONT == 2 && warp < 2) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  cluster_sync_all();            // ... and of every peer: their t tiles / W1 rings (the landing zone) are dead too
  FFN_TRACE(10);
  if (warp >= 2) {
    const int q = warp & 3;
    const int rl = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t my = smem_u32(smem);
#pragma unroll 1
    for (int k = 0; k < C / 32; ++k) {
      uint32_t r[32];
      tmem_ld32(tmem + lane_off + TM_Y + k * 32, r);
      const uint32_t dst = mapa_u32(my, (uint32_t)(k >> 1)) + uint32_t((int)rank * SL_BYTES + (k & 1) * 8 * SL_PITCH + rl * 16);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        st_dsmem_f4(dst + i * SL_PITCH, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                    __uint_as_float(r[4 * i + 3]));
    }
  }
  FFN_TRACE(11);
  // residual rows of the reduce below: requested BEFORE the barrier (the loads were batched four at a time behind it: four
  // L2 round trips of the 5.7 k-cycle reduce phase)
  float4 xin_all[16];
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    const int cg = et & 15, rsub = et >> 4;
    const float* xres = (FRONT ? p.x_out : p.x_in) + (long long)bz * p.x_bstride + rank * 64 + cg * 4;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int rl = rsub + 8 * j;
      xin_all[j] = (m0 + rl) < p.M ? *reinterpret_cast<const float4*>(xres + (long long)(m0 + rl) * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  cluster_sync_all();            // all four partial slices of this CTA's 64 columns are in ITS shared memory
  FFN_TRACE(12);
  // ---- reduce: thread t owns column group cg = t % 16 (4 columns) of rows t / 16 + 8 j: a half-warp reads / writes 256
  //      contiguous bytes of a row of x (thread = row made every global access a 16-byte gather at 1 KB stride)
  float4 yv[16];                 // BACK: the new values (rows t/16 + 8j, column group t%16)
  float2 stat[16];               // BACK: per-row partial (sum, sum of squares) of this CTA's 64 columns
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    const int cg = et & 15, rsub = et >> 4;
    const long long xbase = (long long)bz * p.x_bstride + rank * 64 + cg * 4;
    float* xr = p.x_out + xbase;
    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + rank * 64 + cg * 4));
#pragma unroll
    for (int j0 = 0; j0 < 16; j0 += 4) {
      float4 v[4][CL], xin[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = rsub + 8 * (j0 + i);
        xin[i] = xin_all[j0 + i];
#pragma unroll
        for (int r = 0; r < CL; ++r) v[i][r] = *reinterpret_cast<const float4*>(smem + r * SL_BYTES + cg * SL_PITCH + rl * 16);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = rsub + 8 * (j0 + i);
        float4 acc = v[i][0];                 // fixed rank order: bitwise deterministic
#pragma unroll
        for (int r = 1; r < CL; ++r) { acc.x += v[i][r].x; acc.y += v[i][r].y; acc.z += v[i][r].z; acc.w += v[i][r].w; }
        const float4 y = make_float4(acc.x + bb.x + xin[i].x, acc.y + bb.y + xin[i].y, acc.z + bb.z + xin[i].z, acc.w + bb.w + xin[i].w);
        if ((m0 + rl) < p.M) *reinterpret_cast<float4*>(xr + (long long)(m0 + rl) * C) = y;
        if (BACK) {
          yv[j0 + i] = y;
          float s1 = (y.x + y.y) + (y.z + y.w), s2 = (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {   // the 16 lanes that share a row
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          stat[j0 + i] = make_float2(s1, s2);
        }
      }
    }
    if (BACK) {
      // lanes cg = 0..3 push this CTA's per-row statistics to CTA cg: stats[source rank][row] in the dead W2 ring
      if (cg < CL) {
        const uint32_t dst = mapa_u32(smem_u32(sW2), (uint32_t)cg) + uint32_t((int)rank * BM * 8);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(dst + uint32_t((rsub + 8 * j) * 8)), "f"(stat[j].x), "f"(stat[j].y) : "memory");
      }
    }
  }
  FFN_TRACE(13);
  if (BACK) {
    cluster_sync_all();          // every CTA holds all four CTAs' per-row partial statistics
    if (warp >= 2) {
      const int et = threadIdx.x - 64;
      const int cg = et & 15, rsub = et >> 4;
      const float2* stats = reinterpret_cast<const float2*>(sW2);
      const float4 w = __ldg(reinterpret_cast<const float4*>(p.ln2_w + rank * 64 + cg * 4));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.ln2_b + rank * 64 + cg * 4));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int rl = rsub + 8 * j;
        float sum = 0.f, ss = 0.f;
#pragma unroll
        for (int r = 0; r < CL; ++r) { const float2 q2 = stats[r * BM + rl]; sum += q2.x; ss += q2.y; }   // same order everywhere
        const float mean = sum * (1.0f / C);
        const float rstd = rsqrtf(fmaxf(ss * (1.0f / C) - mean * mean, 0.f) + p.ln2_eps);
        if ((m0 + rl) < p.M) {
          const float o0 = (yv[j].x - mean) * rstd * w.x + bb.x, o1 = (yv[j].y - mean) * rstd * w.y + bb.y;
          const float o2 = (yv[j].z - mean) * rstd * w.z + bb.z, o3 = (yv[j].w - mean) * rstd * w.w + bb.w;
          const long long toff = (long long)bz * p.t_out_sb + (long long)(m0 + rl) * p.t_out_st + rank * 64 + cg * 4;
          if (p.t_out_bf16)
            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.t_out) + toff) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
          else
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.t_out) + toff) = make_float4(o0, o1, o2, o3);
        }
      }
    }
  }
  FFN_TRACE(14);
  cluster_sync_all();            // peers may still be reading this CTA's partial tile / statistics
  FFN_TRACE(15);
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

int g_ffn_fused = 1;   // memory attention: 1 = this kernel, 0 = two GEMM launches (vls_set_tuning "ffn_fused")
long long* g_ffn_trace = nullptr;   // dev-only: 16 int64 clock64 stamps (tools/trace_ffn.py)
int g_tail_quarter = 1;   // layer tail prologue: 1 = column quarters + all-gather (FRONT == 2), 0 = full width in every CTA
int g_tail_fused = 1;  // memory attention: 1 = out-proj + LN3 + FFN + next LN in one launch (needs ffn_fused), 0 = separate

namespace {

template <int FRONT, bool BACK, int NCH, bool GELU>
int launch_variant(const CUtensorMap& tmA, const CUtensorMap& tmW0, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                   const CUtensorMap& tmX, const CUtensorMap& tmXo, const FfnParams& p, int B, cudaStream_t stream) {
  static unsigned long long attr_set = 0;   // one flag word per instantiation
  auto kern = ffn_fused_kernel<FRONT, BACK, NCH, GELU>;
  if (first_use_on_device(&attr_set)) VLS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  VLS_CUDA(launch_k(kern, dim3(CL, (p.M + BM - 1) / BM, B), dim3(THREADS), SMEM_BYTES, stream, tmA, tmW0, tmW1, tmW2, tmX, tmXo, p));
  VLS_POST_LAUNCH(1);
  return 0;
}

}  // namespace

// x[b][m][:] += act(t[b][m][:] W1^T + b1) W2^T + b2;  t bf16 [B][M][256] (row stride ldt), W1 bf16 [ff][256],
// W2 bf16 [256][ff], b1 f32 [ff], b2 f32 [256], x f32 [B][M][256] contiguous rows.  (ff, act) = (2048, ReLU): the FFN of a
// memory-attention layer; (1024, GELU): the point-wise pair of a CXBlock of the memory encoder's fuser.
int launch_ffn_fused(const void* t, long long ldt, long long t_bstride, const void* w1, const float* b1, const void* w2,
                     const float* b2, float* x, long long x_bstride, int B, int M, cudaStream_t stream, int ff, int gelu) {
  VLS_REQUIRE(t && w1 && b1 && w2 && b2 && x && B > 0 && M > 0, "ffn_fused: bad arguments");
  VLS_REQUIRE(ldt % 8 == 0, "ffn_fused: ldt must be a multiple of 8");
  VLS_REQUIRE((ff == 2048 && !gelu) || (ff == 1024 && gelu), "ffn_fused: (hidden, activation) must be (2048, ReLU) or (1024, GELU)");
  CUtensorMap tmA, tmW1, tmW2;
  VLS_TRY(make_tmap_bf16(&tmA, t, C, M, B, ldt, t_bstride, BM));
  VLS_TRY(make_tmap_bf16(&tmW1, w1, C, ff, 1, C, (long long)ff * C, HC));
  VLS_TRY(make_tmap_bf16(&tmW2, w2, ff, C, 1, ff, (long long)ff * C, C));
  FfnParams p = {};
  p.M = M; p.b1 = b1; p.b2 = b2; p.x_in = x; p.x_out = x; p.x_bstride = x_bstride;
  p.trace = g_ffn_trace;
  if (gelu) return launch_variant<0, false, 2, true>(tmA, tmA, tmW1, tmW2, tmA, tmA, p, B, stream);
  return launch_variant<0, false, 4, false>(tmA, tmA, tmW1, tmW2, tmA, tmA, p, B, stream);
}

// The tail of a memory-attention layer in one launch (see the file header):
//   x_mid = x_in + ao W0^T + b0;  x_out = x_mid + FFN(LN3(x_mid));  t_out = LN_next(x_out)
// ao bf16 [B][M][64] (the cross-attention output over the 64-d memory), W0 bf16 [256][64] = Wo Wv.  x_in != x_out.
int launch_layer_tail(const LayerTailArgs& a, cudaStream_t stream) {
  VLS_REQUIRE(a.ao && a.w0 && a.b0 && a.ln_w && a.ln_b && a.w1 && a.b1 && a.w2 && a.b2 && a.x_in && a.x_out && a.ln2_w &&
              a.ln2_b && a.t_out && a.B > 0 && a.M > 0, "layer_tail: bad arguments");
  VLS_REQUIRE(a.x_in != a.x_out, "layer_tail: x_in and x_out must be different buffers");
  VLS_REQUIRE(a.t_out_st % 4 == 0 && a.t_out_sb % 4 == 0, "layer_tail: output strides must be multiples of 4");
  CUtensorMap tmA, tmW0, tmW1, tmW2, tmX, tmXo;
  const bool quarter = g_tail_quarter != 0;
  VLS_TRY(make_tmap_f32(&tmX, a.x_in, C, a.M, a.B, C, (long long)a.M * C, BM));
  VLS_TRY(make_tmap_f32(&tmXo, a.x_out, C, a.M, a.B, C, (long long)a.M * C, BM));
  VLS_TRY(make_tmap_bf16(&tmA, a.ao, 64, a.M, a.B, 64, (long long)a.M * 64, BM));
  VLS_TRY(make_tmap_bf16(&tmW0, a.w0, 64, C, 1, 64, (long long)C * 64, quarter ? 64 : C));
  constexpr int FF = 2048;
  VLS_TRY(make_tmap_bf16(&tmW1, a.w1, C, FF, 1, C, (long long)FF * C, HC));
  VLS_TRY(make_tmap_bf16(&tmW2, a.w2, FF, C, 1, FF, (long long)FF * C, C));
  FfnParams p = {};
  p.M = a.M; p.b1 = a.b1; p.b2 = a.b2; p.x_in = a.x_in; p.x_out = a.x_out; p.x_bstride = (long long)a.M * C;
  p.b0 = a.b0; p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_eps = a.ln_eps;
  p.ln2_w = a.ln2_w; p.ln2_b = a.ln2_b; p.ln2_eps = a.ln2_eps;
  p.t_out = a.t_out; p.t_out_bf16 = a.t_out_bf16; p.t_out_st = a.t_out_st; p.t_out_sb = a.t_out_sb;
  p.trace = g_ffn_trace;
  if (quarter) return launch_variant<2, true, 4, false>(tmA, tmW0, tmW1, tmW2, tmX, tmXo, p, a.B, stream);
  return launch_variant<1, true, 4, false>(tmA, tmW0, tmW1, tmW2, tmX, tmXo, p, a.B, stream);
}

}  // namespace vls
