// Connected components (8-connectivity) labelling + per-pixel component area, replacing
// sam2/csrc/connected_components.cu (6 launches per image in a host loop, :245-275) by ONE launch
// for the whole batch at the production shape [B,1,256,256], and the fused hole-filling of
// sam2/utils/misc.py:312-338.
//
// Result contract (bit-exact with the reference): union-find over 2x2 pixel blocks whose root is
// always the minimum index, so   label(pixel) = 1 + min over its component of ((r&~1)*W + (c&~1))
// and count(pixel) = component area; background pixels get 0 / 0.
//
// Small path (one CTA per image, (H/2)*(W/2) <= 16384 blocks): everything lives in shared memory.
//   A. each lane derives its block's 4-bit occupancy straight from global memory (16-bit loads)
//   B. warp-level run merge: a warp walks a row of blocks, __ballot_sync of "connected to the left
//      block" turns horizontal runs into star trees without a single atomic
//   C. vertical / diagonal unions (atomicMin union-find in smem), pruned when the previous lane of
//      the same run already linked to a horizontally connected upper block
//   D. path compression + area: __match_any_sync / __reduce_add_sync aggregate lanes sharing a
//      root so a large component costs one smem atomic per warp-row, not one per block
//   E. 128-bit stores of labels and areas (or, for hole filling, sparse in-place stores of 0.1)
// Larger images: 64 x 128 pixel tiles labelled in shared memory the same way, tile-border unions in global
// memory (the labels array is the forest, as in the reference), batched over N, four launches.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int CC_THREADS = 1024;
constexpr int CC_MAX_BLOCKS = 16384;

__device__ __forceinline__ int uf_find(const volatile int* s, int n) {
  int p = s[n];
  while (p != n) {
    n = p;
    p = s[n];
  }
  return n;
}

__device__ __forceinline__ void uf_union(int* s, int a, int b) {
  bool done;
  do {
    a = uf_find(s, a);
    b = uf_find(s, b);
    if (a < b) {
      const int old = atomicMin(s + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(s + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// occupancy bits: 0 = top-left, 1 = top-right, 2 = bottom-left, 3 = bottom-right
__device__ __forceinline__ bool conn_left(uint32_t me, uint32_t left) { return (me & 0x5u) && (left & 0xAu); }
__device__ __forceinline__ bool conn_up(uint32_t me, uint32_t up) { return (me & 0x3u) && (up & 0xCu); }
__device__ __forceinline__ bool conn_upleft(uint32_t me, uint32_t ul) { return (me & 0x1u) && (ul & 0x8u); }
__device__ __forceinline__ bool conn_upright(uint32_t me, uint32_t ur) { return (me & 0x2u) && (ur & 0x4u); }

template <bool FILL>
__device__ __forceinline__ uint32_t load_occ(const void* img, int H, int W, int by, int bx, float) {
  const int r = 2 * by, c = 2 * bx;
  if (FILL) {  // foreground of the hole search = (score <= 0)   (utils/misc.py:322)
    const float* f = reinterpret_cast<const float*>(img);
    const float2 t = *reinterpret_cast<const float2*>(f + (size_t)r * W + c);
    const float2 b = *reinterpret_cast<const float2*>(f + (size_t)(r + 1) * W + c);
    return (t.x <= 0.f ? 1u : 0u) | (t.y <= 0.f ? 2u : 0u) | (b.x <= 0.f ? 4u : 0u) | (b.y <= 0.f ? 8u : 0u);
  } else {
    const uint8_t* u = reinterpret_cast<const uint8_t*>(img);
    const uint16_t t = *reinterpret_cast<const uint16_t*>(u + (size_t)r * W + c);
    const uint16_t b = *reinterpret_cast<const uint16_t*>(u + (size_t)(r + 1) * W + c);
    return ((t & 0xFF) ? 1u : 0u) | ((t >> 8) ? 2u : 0u) | ((b & 0xFF) ? 4u : 0u) | ((b >> 8) ? 8u : 0u);
  }
}

// FILL=false: img uint8 -> labels/counts int32.   FILL=true: scores f32 updated in place.
template <bool FILL>
__global__ void __launch_bounds__(CC_THREADS, 1)
cc_small_kernel(const void* img_all, int H, int W, int32_t* labels_all, int32_t* counts_all, float* scores_all,
                int max_area, float fill_value) {
  extern __shared__ int cc_smem[];
  const int BH = H >> 1, BW = W >> 1, nb = BH * BW;
  int* lab = cc_smem;
  int* cnt = cc_smem + nb;
  uint8_t* occ = reinterpret_cast<uint8_t*>(cc_smem + 2 * nb);
  const size_t img_off = (size_t)blockIdx.x * H * W;
  const void* img = FILL ? static_cast<const void*>(scores_all + img_off)
                         : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img_all) + img_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = CC_THREADS / 32;
  const int chunks = (BW + 31) >> 5;

  // A. occupancy
  for (int bi = threadIdx.x; bi < nb; bi += CC_THREADS) {
    occ[bi] = (uint8_t)load_occ<FILL>(img, H, W, bi / BW, bi % BW, 0.f);
    cnt[bi] = 0;
  }
  __syncthreads();
  // B. horizontal runs -> star trees
  for (int t = warp; t < BH * chunks; t += nwarps) {
    const int by = t / chunks, c0 = (t % chunks) << 5, bx = c0 + lane;
    const bool in = bx < BW;
    const int bi = by * BW + bx;
    const uint32_t me = in ? occ[bi] : 0u;
    const uint32_t left = (in && bx > 0) ? occ[bi - 1] : 0u;
    const bool h = conn_left(me, left);
    const uint32_t hm = __ballot_sync(0xffffffffu, h);
    if (in) {
      const uint32_t stops = (~hm | 1u) & (0xffffffffu >> (31 - lane));
      lab[bi] = by * BW + c0 + (31 - __clz(stops));
    }
  }
  __syncthreads();
  // C. unions: chunk seams, up, up-left, up-right
  for (int t = warp; t < BH * chunks; t += nwarps) {
    const int by = t / chunks, c0 = (t % chunks) << 5, bx = c0 + lane;
    const bool in = bx < BW;
    const int bi = by * BW + bx;
    const uint32_t me = in ? occ[bi] : 0u;
    const uint32_t left = (in && bx > 0) ? occ[bi - 1] : 0u;
    const bool h = conn_left(me, left);
    uint32_t up = 0, ul = 0, ur = 0;
    if (in && by > 0) {
      up = occ[bi - BW];
      if (bx > 0) ul = occ[bi - BW - 1];
      if (bx + 1 < BW) ur = occ[bi - BW + 1];
    }
    const bool cu = conn_up(me, up);
    const bool cul = conn_upleft(me, ul) && !(cu && conn_left(up, ul));
    const bool cur = conn_upright(me, ur) && !(cu && conn_left(ur, up));
    const bool cu_prev = __shfl_up_sync(0xffffffffu, cu, 1);
    const bool cu_redundant = lane > 0 && h && cu_prev && conn_left(up, ul);
    if (lane == 0 && h) uf_union(lab, bi, bi - 1);
    if (cu && !cu_redundant) uf_union(lab, bi, bi - BW);
    if (cul) uf_union(lab, bi, bi - BW - 1);
    if (cur) uf_union(lab, bi, bi - BW + 1);
  }
  __syncthreads();
  // D. compression + areas
  for (int t = warp; t < BH * chunks; t += nwarps) {
    const int by = t / chunks, c0 = (t % chunks) << 5, bx = c0 + lane;
    const bool in = bx < BW;
    const int bi = by * BW + bx;
    const uint32_t me = in ? occ[bi] : 0u;
    int root = -1 - lane;  // unique negative key for empty lanes
    if (me) {
      root = uf_find(lab, bi);
      lab[bi] = root;  // benign race: concurrent finds still see a valid ancestor
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, root);
    const int area = __reduce_add_sync(peers, (int)__popc(me));
    if (me && lane == (__ffs(peers) - 1)) atomicAdd(cnt + root, area);
  }
  __syncthreads();
  // E. outputs
  if (FILL) {
    float* sc = scores_all + img_off;
    for (int bi = threadIdx.x; bi < nb; bi += CC_THREADS) {
      const uint32_t me = occ[bi];
      if (!me) continue;
      if (cnt[lab[bi]] > max_area) continue;
      const int r = 2 * (bi / BW), c = 2 * (bi % BW);
      if (me & 1u) sc[(size_t)r * W + c] = fill_value;
      if (me & 2u) sc[(size_t)r * W + c + 1] = fill_value;
      if (me & 4u) sc[(size_t)(r + 1) * W + c] = fill_value;
      if (me & 8u) sc[(size_t)(r + 1) * W + c + 1] = fill_value;
    }
  } else {
    int32_t* labels = labels_all + img_off;
    int32_t* counts = counts_all + img_off;
    if ((W & 3) == 0) {
      const int quads = (H * W) >> 2, qpr = W >> 2;
      for (int qd = threadIdx.x; qd < quads; qd += CC_THREADS) {
        const int r = qd / qpr, c = (qd % qpr) << 2;
        const int b0 = (r >> 1) * BW + (c >> 1);
        const int sh = (r & 1) << 1;
        const uint32_t o0 = occ[b0] >> sh, o1 = occ[b0 + 1] >> sh;
        int l0 = 0, l1 = 0, n0 = 0, n1 = 0;
        if (o0 & 3u) {
          const int rt = lab[b0];
          l0 = (rt / BW) * 2 * W + (rt % BW) * 2 + 1;
          n0 = cnt[rt];
        }
        if (o1 & 3u) {
          const int rt = lab[b0 + 1];
          l1 = (rt / BW) * 2 * W + (rt % BW) * 2 + 1;
          n1 = cnt[rt];
        }
        const int4 lv = make_int4((o0 & 1u) ? l0 : 0, (o0 & 2u) ? l0 : 0, (o1 & 1u) ? l1 : 0, (o1 & 2u) ? l1 : 0);
        const int4 cv = make_int4((o0 & 1u) ? n0 : 0, (o0 & 2u) ? n0 : 0, (o1 & 1u) ? n1 : 0, (o1 & 2u) ? n1 : 0);
        *reinterpret_cast<int4*>(labels + (size_t)r * W + c) = lv;
        *reinterpret_cast<int4*>(counts + (size_t)r * W + c) = cv;
      }
    } else {
      for (int px = threadIdx.x; px < H * W; px += CC_THREADS) {
        const int r = px / W, c = px % W;
        const int b0 = (r >> 1) * BW + (c >> 1);
        const bool fg = (occ[b0] >> (((r & 1) << 1) | (c & 1))) & 1u;
        int l = 0, n = 0;
        if (fg) {
          const int rt = lab[b0];
          l = (rt / BW) * 2 * W + (rt % BW) * 2 + 1;
          n = cnt[rt];
        }
        labels[px] = l;
        counts[px] = n;
      }
    }
  }
}

// ------------------------------------------------------------------ tiled path (images larger than 256 x 256)
// Tile = 64 x 128 pixels = 32 x 64 blocks, labelled entirely in shared memory with the same warp-level
// run merge as the small path; only tile-border blocks take part in global atomicMin unions.
//   cc_t_label   : occupancy (also cached as 1 byte / block for the later passes), local union-find,
//                  forest[block's top-left pixel] = GLOBAL pixel index of the local root
//   cc_t_border  : unions across tile borders (top row: up / up-left / up-right; left column: left / up-left;
//                  right column: up-right)
//   cc_t_count   : global find + path compression, area[root] += popcount (warp-aggregated atomics)
//   cc_t_final   : 8-byte stores of labels / areas (or sparse in-place fill of small holes)
constexpr int TBH = 32, TBW = 64, T_THREADS = 256;

template <bool FILL>
__global__ void __launch_bounds__(T_THREADS)
cc_t_label(const void* img_all, const float* scores_all, int H, int W, int32_t* forest_all, uint8_t* occ_all) {
  __shared__ int lab[TBH * TBW];
  __shared__ uint8_t occ[TBH * TBW];
  const int BH = H >> 1, BW = W >> 1;
  const size_t off = (size_t)blockIdx.z * H * W;
  const void* img = FILL ? static_cast<const void*>(scores_all + off)
                         : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img_all) + off);
  const int by0 = blockIdx.y * TBH, bx0 = blockIdx.x * TBW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = T_THREADS / 32;
  for (int i = threadIdx.x; i < TBH * TBW; i += T_THREADS) {
    const int by = by0 + i / TBW, bx = bx0 + i % TBW;
    const uint32_t o = (by < BH && bx < BW) ? load_occ<FILL>(img, H, W, by, bx, 0.f) : 0u;
    occ[i] = (uint8_t)o;
    if (by < BH && bx < BW) occ_all[(size_t)blockIdx.z * BH * BW + (size_t)by * BW + bx] = (uint8_t)o;
  }
  __syncthreads();
  constexpr int CH = TBW / 32;
  for (int t = warp; t < TBH * CH; t += nwarps) {
    const int ly = t / CH, c0 = (t % CH) << 5, lx = c0 + lane, li = ly * TBW + lx;
    const uint32_t me = occ[li], left = lx > 0 ? occ[li - 1] : 0u;
    const uint32_t hm = __ballot_sync(0xffffffffu, conn_left(me, left));
    const uint32_t stops = (~hm | 1u) & (0xffffffffu >> (31 - lane));
    lab[li] = ly * TBW + c0 + (31 - __clz(stops));
  }
  __syncthreads();
  for (int t = warp; t < TBH * CH; t += nwarps) {
    const int ly = t / CH, c0 = (t % CH) << 5, lx = c0 + lane, li = ly * TBW + lx;
    const uint32_t me = occ[li], left = lx > 0 ? occ[li - 1] : 0u;
    const bool h = conn_left(me, left);
    uint32_t up = 0, ul = 0, ur = 0;
    if (ly > 0) {
      up = occ[li - TBW];
      if (lx > 0) ul = occ[li - TBW - 1];
      if (lx + 1 < TBW) ur = occ[li - TBW + 1];
    }
    const bool cu = conn_up(me, up);
    const bool cul = conn_upleft(me, ul) && !(cu && conn_left(up, ul));
    const bool cur = conn_upright(me, ur) && !(cu && conn_left(ur, up));
    const bool cu_prev = __shfl_up_sync(0xffffffffu, cu, 1);
    const bool cu_redundant = lane > 0 && h && cu_prev && conn_left(up, ul);
    if (lane == 0 && h) uf_union(lab, li, li - 1);
    if (cu && !cu_redundant) uf_union(lab, li, li - TBW);
    if (cul) uf_union(lab, li, li - TBW - 1);
    if (cur) uf_union(lab, li, li - TBW + 1);
  }
  __syncthreads();
  int32_t* forest = forest_all + off;
  for (int i = threadIdx.x; i < TBH * TBW; i += T_THREADS) {
    const int by = by0 + i / TBW, bx = bx0 + i % TBW;
    if (by >= BH || bx >= BW) continue;
    const int r = occ[i] ? uf_find(lab, i) : i;
    forest[(2 * by) * W + 2 * bx] = (2 * (by0 + r / TBW)) * W + 2 * (bx0 + r % TBW);
  }
}

__global__ void cc_t_border(const uint8_t* __restrict__ occ_all, int H, int W, int32_t* forest_all) {
  const int BH = H >> 1, BW = W >> 1;
  const int tiles_x = (BW + TBW - 1) / TBW, tiles_y = (BH + TBH - 1) / TBH;
  const int per_tile = TBW + 2 * TBH;  // top row, left column, right column
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)tiles_x * tiles_y * per_tile) return;
  const int tile = (int)(id / per_tile), k = (int)(id % per_tile);
  const int ty = tile / tiles_x, tx = tile % tiles_x;
  int ly, lx, kind;  // kind 0: top row, 1: left column, 2: right column
  if (k < TBW) { ly = 0; lx = k; kind = 0; }
  else if (k < TBW + TBH) { ly = k - TBW; lx = 0; kind = 1; }
  else { ly = k - TBW - TBH; lx = TBW - 1; kind = 2; }
  const int by = ty * TBH + ly, bx = tx * TBW + lx;
  if (by >= BH || bx >= BW) return;
  const uint8_t* occ = occ_all + (size_t)blockIdx.z * BH * BW;
  int32_t* forest = forest_all + (size_t)blockIdx.z * H * W;
  const uint32_t me = occ[(size_t)by * BW + bx];
  if (!me) return;
  const int idx = 2 * by * W + 2 * bx;
  auto at = [&](int y, int x) -> uint32_t { return (y >= 0 && x >= 0 && x < BW) ? occ[(size_t)y * BW + x] : 0u; };
  if (kind == 0 && by > 0) {
    if (conn_up(me, at(by - 1, bx))) uf_union(forest, idx, idx - 2 * W);
    if (conn_upleft(me, at(by - 1, bx - 1))) uf_union(forest, idx, idx - 2 * W - 2);
    if (conn_upright(me, at(by - 1, bx + 1))) uf_union(forest, idx, idx - 2 * W + 2);
  } else if (kind == 1 && bx > 0) {
    if (conn_left(me, at(by, bx - 1))) uf_union(forest, idx, idx - 2);
    if (ly > 0 && conn_upleft(me, at(by - 1, bx - 1))) uf_union(forest, idx, idx - 2 * W - 2);
  } else if (kind == 2 && ly > 0) {
    if (conn_upright(me, at(by - 1, bx + 1))) uf_union(forest, idx, idx - 2 * W + 2);
  }
}

__global__ void cc_t_count(const uint8_t* __restrict__ occ_all, int H, int W, int32_t* forest_all, int32_t* area_all) {
  const int BW = W >> 1, BH = H >> 1;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  const bool in = bx < BW && by < BH;
  const size_t off = (size_t)blockIdx.z * H * W;
  const uint32_t me = in ? occ_all[(size_t)blockIdx.z * BH * BW + (size_t)by * BW + bx] : 0u;
  int root = -1 - (int)(threadIdx.x & 31);
  if (me) {
    const int idx = 2 * by * W + 2 * bx;
    root = uf_find(forest_all + off, idx);
    forest_all[off + idx] = root;
  }
  const uint32_t peers = __match_any_sync(0xffffffffu, root);
  const int area = __reduce_add_sync(peers, (int)__popc(me));
  if (me && (int)(threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(area_all + off + root, area);
}

template <bool FILL>
__global__ void cc_t_final(const uint8_t* __restrict__ occ_all, int H, int W, int32_t* labels_all,
                           const int32_t* __restrict__ area_all, int32_t* counts_all, float* scores_all, int max_area,
                           float fill_value) {
  const int BW = W >> 1, BH = H >> 1;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  if (bx >= BW || by >= BH) return;
  const size_t off = (size_t)blockIdx.z * H * W;
  const uint32_t me = occ_all[(size_t)blockIdx.z * BH * BW + (size_t)by * BW + bx];
  const int idx = 2 * by * W + 2 * bx;
  int root = 0, n = 0;
  if (me) {
    root = labels_all[off + idx];
    n = area_all[off + root];
  }
  if (FILL) {
    if (me && n <= max_area) {
      float* sc = scores_all + off;
      if (me & 1u) sc[idx] = fill_value;
      if (me & 2u) sc[idx + 1] = fill_value;
      if (me & 4u) sc[idx + W] = fill_value;
      if (me & 8u) sc[idx + W + 1] = fill_value;
    }
  } else {
    int32_t* L = labels_all + off;
    int32_t* C = counts_all + off;
    const int y = root + 1;
    *reinterpret_cast<int2*>(L + idx) = make_int2((me & 1u) ? y : 0, (me & 2u) ? y : 0);
    *reinterpret_cast<int2*>(L + idx + W) = make_int2((me & 4u) ? y : 0, (me & 8u) ? y : 0);
    *reinterpret_cast<int2*>(C + idx) = make_int2((me & 1u) ? n : 0, (me & 2u) ? n : 0);
    *reinterpret_cast<int2*>(C + idx + W) = make_int2((me & 4u) ? n : 0, (me & 8u) ? n : 0);
  }
}

bool small_ok(int h, int w) { return (h / 2) * (w / 2) <= CC_MAX_BLOCKS; }
size_t small_smem(int h, int w) {
  const size_t nb = (size_t)(h / 2) * (w / 2);
  return nb * 9;
}

template <bool FILL>
int run(const void* img, float* scores, int n, int h, int w, int32_t* labels, int32_t* counts, int max_area,
        float fill_value, void* ws, size_t ws_bytes, cudaStream_t stream) {
  VLS_REQUIRE(n >= 0 && h >= 0 && w >= 0, "cc: negative dimension");
  VLS_REQUIRE((h % 2) == 0, "height must be an even number");  // connected_components.cu:226
  VLS_REQUIRE((w % 2) == 0, "width must be an even number");   // connected_components.cu:227
  if (n == 0 || h == 0 || w == 0) return 0;
  if (small_ok(h, w)) {
    static bool attr[2] = {false, false};
    if (!attr[FILL]) {
      VLS_CUDA(cudaFuncSetAttribute(cc_small_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    CC_MAX_BLOCKS * 9));
      attr[FILL] = true;
    }
    cc_small_kernel<FILL><<<n, CC_THREADS, small_smem(h, w), stream>>>(img, h, w, labels, counts, scores, max_area,
                                                                        fill_value);
    VLS_POST_LAUNCH(1);
    return 0;
  }
  const size_t px = (size_t)n * h * w;
  const size_t blocks = px / 4;
  const size_t need = (FILL ? 2 * px * 4 : px * 4) + blocks;
  VLS_REQUIRE(ws != nullptr && ws_bytes >= need, "cc: workspace too small (%zu < %zu)", ws_bytes, need);
  int32_t* area = reinterpret_cast<int32_t*>(ws);
  int32_t* forest = FILL ? area + px : labels;
  uint8_t* occ = reinterpret_cast<uint8_t*>(area + (FILL ? 2 * px : px));
  VLS_CUDA(cudaMemsetAsync(area, 0, px * 4, stream));
  const int BH = h / 2, BW = w / 2;
  const int tiles_x = (BW + TBW - 1) / TBW, tiles_y = (BH + TBH - 1) / TBH;
  cc_t_label<FILL><<<dim3(tiles_x, tiles_y, n), T_THREADS, 0, stream>>>(img, scores, h, w, forest, occ);
  const long long border = (long long)tiles_x * tiles_y * (TBW + 2 * TBH);
  cc_t_border<<<dim3((unsigned)((border + 255) / 256), 1, n), 256, 0, stream>>>(occ, h, w, forest);
  dim3 blk(128, 1, 1), grd((BW + 127) / 128, BH, n);
  cc_t_count<<<grd, blk, 0, stream>>>(occ, h, w, forest, area);
  cc_t_final<FILL><<<grd, blk, 0, stream>>>(occ, h, w, forest, area, counts, scores, max_area, fill_value);
  VLS_POST_LAUNCH(4);
  return 0;
}

}  // namespace

size_t cc_workspace_bytes(int n, int h, int w, bool fill) {
  if (n <= 0 || h <= 0 || w <= 0 || small_ok(h, w)) return 0;
  return (size_t)n * h * w * 4 * (fill ? 2 : 1) + (size_t)n * h * w / 4;
}

int launch_cc_label(const uint8_t* img, int n, int h, int w, int32_t* labels, int32_t* counts, void* ws,
                    size_t ws_bytes, cudaStream_t stream) {
  VLS_REQUIRE(n == 0 || (img && labels && counts), "cc: null pointer");
  return run<false>(img, nullptr, n, h, w, labels, counts, 0, 0.f, ws, ws_bytes, stream);
}

int launch_fill_holes(float* scores, int n, int h, int w, int max_area, float fill_value, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  VLS_REQUIRE(n == 0 || scores, "fill_holes: null pointer");
  VLS_REQUIRE(max_area > 0, "max_area must be positive");  // utils/misc.py:318
  return run<true>(nullptr, scores, n, h, w, nullptr, nullptr, max_area, fill_value, ws, ws_bytes, stream);
}

}  // namespace vls
