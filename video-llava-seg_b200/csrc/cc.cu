// Connected components (8-connectivity) labelling + per-pixel component area, replacing
// sam2/csrc/connected_components.cu (6 launches per image in a host loop, :245-275) by ONE launch
// for the whole batch at the production shape [B,1,256,256], and the fused hole-filling of
// sam2/utils/misc.py:312-338.
//
// Result contract (bit-exact with the reference): union-find over 2x2 pixel blocks whose root is
// always the minimum index, so   label(pixel) = 1 + min over its component of ((r&~1)*W + (c&~1))
// and count(pixel) = component area; background pixels get 0 / 0.
//
// Small path (one CTA of 512 threads per image, (H/2)*(W/2) <= 16384 blocks, two CTAs per SM): everything lives in
// 5 bytes of shared memory per block (a union-find word that doubles as the area counter + an occupancy byte).
//   A. occupancy of every 2x2 block straight from global memory with 128-bit loads (16 pixels x 2 rows per thread)
//   B. region labelling: a warp walks its band of rows of one 32-block column strip top-down, keeping the previous
//      row's labels in registers; a horizontal run (found with __ballot_sync) inherits the smallest label of the upper
//      blocks it touches (warp reductions / segmented min-scan) or starts a new label.  No atomics unless a run joins
//      two differently named components.  Run areas are summed per label in registers and parked on the label's word.
//   C. seams between regions (band seams inside a strip, strip seams one lane per row) with lock-free atomicMin
//      min-root unions + path compression; area parked on a node travels with the link
//   D. the few blocks that ever started a label are flattened onto their roots (the only loop-y finds)
//   E. block -> label -> root -> area with plain loads; 128-bit streaming stores of labels and areas
//      (or, for hole filling, sparse in-place stores of 0.1)
// Larger images: 64 x 128 pixel tiles labelled in shared memory by the same region labeller; the global state is one
// forest word (parent << 4 | occupancy) and one area word per 2x2 block plus a list of the tile-local roots that touch
// a tile border; border unions (one warp per tile edge piece, pruned per run), area hand-over, output pass.
#include "common.cuh"
#include "kernels.h"

namespace vls {

namespace {

constexpr int CC_THREADS = 512;
constexpr int CC_MAX_BLOCKS = 16384;
constexpr int CC_IDX_BITS = 14;                       // block index < 16384; bits 14.. of a ROOT's word hold its area
constexpr int CC_IDX_MASK = (1 << CC_IDX_BITS) - 1;
constexpr int CC_NAME_CAP = 4096;                     // entries of the shared-memory list of NAME blocks (phase D)
constexpr uint32_t CC_NAME = 0x10u;                   // occupancy byte, bit 4: the block started a new label in phase B

__device__ __forceinline__ int uf_find(const volatile int* s, int n) {
  int p = s[n];
  while (p != n) {
    n = p;
    p = s[n];
  }
  return n;
}

// Every node on the path n -> ... -> r gets parent r.  r is an ancestor of n and parents are always smaller than
// their children, so the walk ends at r; atomicMin keeps "parent only ever moves to a smaller member of the set".
__device__ __forceinline__ void uf_compress(int* s, int n, int r) {
  while (n > r) {
    const int p = reinterpret_cast<const volatile int*>(s)[n];
    if (p == n) break;
    if (p > r) atomicMin(s + n, r);
    n = p;
  }
}

// min-root union (lock-free, atomicMin on roots) followed by path compression of both sides
__device__ __forceinline__ void uf_union(int* s, int a, int b) {
  int ra = uf_find(s, a), rb = uf_find(s, b);
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const int old = atomicMin(s + ra, rb);  // ra > rb
    if (old == ra) break;
    ra = uf_find(s, old);                   // ra had been linked meanwhile: carry on from its parent
    rb = uf_find(s, rb);
  }
  const int r = ra < rb ? ra : rb;
  uf_compress(s, a, r);
  uf_compress(s, b, r);
}

// ---- area-carrying variant for the shared-memory kernel: a word is parent | area << 14.  Areas are parked on NAME
// blocks (phase B) and travel with the links: atomicMin returns the old word, whose area part is re-added to the new
// parent.  Area left on a name that is no longer a root is collected when the names are flattened (phase D).
__device__ __forceinline__ int ufa_find(const volatile int* s, int n) {
  int p = s[n] & CC_IDX_MASK;
  while (p != n) {
    n = p;
    p = s[n] & CC_IDX_MASK;
  }
  return n;
}
__device__ __forceinline__ void ufa_move_area(int* s, int old_word, int to) {
  const uint32_t a = (uint32_t)old_word >> CC_IDX_BITS;
  if (a) atomicAdd(s + to, (int)(a << CC_IDX_BITS));
}
__device__ __forceinline__ void ufa_compress(int* s, int n, int r) {
  while (n > r) {
    const int p = reinterpret_cast<const volatile int*>(s)[n] & CC_IDX_MASK;
    if (p == n) break;
    if (p > r) ufa_move_area(s, atomicMin(s + n, r), r);
    n = p;
  }
}
__device__ __forceinline__ void ufa_union(int* s, int a, int b) {
  int ra = ufa_find(s, a), rb = ufa_find(s, b);
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const int old = atomicMin(s + ra, rb);  // ra > rb: the word becomes rb (any area bits made it larger than rb)
    ufa_move_area(s, old, rb);
    if ((old & CC_IDX_MASK) == ra) break;
    ra = ufa_find(s, old & CC_IDX_MASK);
    rb = ufa_find(s, rb);
  }
  const int r = ra < rb ? ra : rb;
  ufa_compress(s, a, r);
  ufa_compress(s, b, r);
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
    const float* t = reinterpret_cast<const float*>(img) + (size_t)r * W + c;   // scalar loads: any 4-byte aligned base
    const float* b = t + W;
    return (t[0] <= 0.f ? 1u : 0u) | (t[1] <= 0.f ? 2u : 0u) | (b[0] <= 0.f ? 4u : 0u) | (b[1] <= 0.f ? 8u : 0u);
  } else {
    const uint8_t* t = reinterpret_cast<const uint8_t*>(img) + (size_t)r * W + c;  // byte loads: any base address
    const uint8_t* b = t + W;
    return (t[0] ? 1u : 0u) | (t[1] ? 2u : 0u) | (b[0] ? 4u : 0u) | (b[1] ? 8u : 0u);
  }
}

// 4 top + 4 bottom uint8 pixels -> the occupancy bytes of 2 blocks (low 16 bits of the result)
__device__ __forceinline__ uint32_t occ2_from_u8(uint32_t top, uint32_t bot) {
  const uint32_t x = (__vcmpne4(top, 0u) & 0x01010101u) | ((__vcmpne4(bot, 0u) & 0x01010101u) << 2);
  const uint32_t a = (x & 0x5u) | ((x >> 7) & 0xAu);
  const uint32_t b = ((x >> 16) & 0x5u) | ((x >> 23) & 0xAu);
  return a | (b << 8);
}

// One warp-row = 32 consecutive blocks of one block row.  Neighbour occupancies come from two byte loads per lane
// plus shuffles; only the edge lanes read the adjacent chunk.
struct RowOcc {
  uint32_t me, left, up, ul, ur;
};
__device__ __forceinline__ RowOcc row_occ(const uint8_t* occ, int BW, int by, int bx, int bi, bool in, int lane) {
  RowOcc o;
  o.me = in ? occ[bi] : 0u;
  o.up = (in && by > 0) ? occ[bi - BW] : 0u;
  o.left = __shfl_up_sync(0xffffffffu, o.me, 1);
  o.ul = __shfl_up_sync(0xffffffffu, o.up, 1);
  o.ur = __shfl_down_sync(0xffffffffu, o.up, 1);
  if (lane == 0) {
    o.left = (in && bx > 0) ? occ[bi - 1] : 0u;
    o.ul = (in && bx > 0 && by > 0) ? occ[bi - BW - 1] : 0u;
  }
  if (lane == 31) o.ur = (in && bx + 1 < BW && by > 0) ? occ[bi - BW + 1] : 0u;
  return o;
}

// Phases B-D of the shared-memory labeller on a BH x BW block region held in `lab` / `occ` (row pitch BW), executed by
// the whole CTA (nthreads = blockDim.x, a multiple of 32).  `occ` must be complete (followed by __syncthreads) on
// entry; on return (after the caller's __syncthreads) every occupied block's word is a NAME block index, every name's
// word its root, and a root's word root | area << 14.
#ifdef CC_TRACE
__device__ long long g_cc_tr[8];
__device__ int g_cc_cnt[4];
#define CC_RMARK(i)                                            \
  do {                                                         \
    if (threadIdx.x == 0 && blockIdx.x == 0) g_cc_tr[i] = clock64(); \
  } while (0)
#else
#define CC_RMARK(i)
#endif
__device__ __forceinline__ void cc_label_region(int* lab, uint8_t* occ, int BH, int BW, uint16_t* name_list,
                                                int* name_count, int name_cap) {
  const int nthreads = blockDim.x;
  const int nb = BH * BW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nthreads >> 5;
  const int chunks = (BW + 31) >> 5;
  // warp-row tasks in column-strip-major order; each warp owns a contiguous band of rows of one strip and walks it
  // top-down, which (with path compression) keeps the union-find chains a few hops long
  const int ntask = BH * chunks, per = (ntask + nwarps - 1) / nwarps;
  const int t_begin = warp * per, t_end = min(ntask, t_begin + per);
  // B. region labelling.  A region = this warp's band of rows within one 32-block column strip, walked top-down with
  //    the previous row's labels kept in REGISTERS: a run takes the smallest label among the upper blocks it touches
  //    (segmented min-scan over the run's lanes) or becomes a new root; shared memory sees one store per block.
  //    Only a run that touches two differently named upper components costs a union-find operation.
  {
    constexpr int INF = 0x7fffffff;
    uint32_t up_me = 0u;
    int up_lab = INF, acc_name = -1, acc_sum = 0;
    int chunk = t_begin / BH, by = t_begin - chunk * BH;
    for (int t = t_begin; t < t_end; ++t, ++by) {
      if (by == BH) {
        by = 0;
        ++chunk;
      }
      if (by == 0) {
        up_me = 0u;
        up_lab = INF;
      }
      const int c0 = chunk << 5, bx = c0 + lane;
      const bool in = bx < BW;
      const int bi = by * BW + bx;
      const uint32_t me = in ? occ[bi] : 0u;
      uint32_t left = __shfl_up_sync(0xffffffffu, me, 1);
      uint32_t ulo = __shfl_up_sync(0xffffffffu, up_me, 1), uro = __shfl_down_sync(0xffffffffu, up_me, 1);
      const int ull = __shfl_up_sync(0xffffffffu, up_lab, 1), url = __shfl_down_sync(0xffffffffu, up_lab, 1);
      if (lane == 0) left = 0u, ulo = 0u;
      if (lane == 31) uro = 0u;
      const int ca = conn_up(me, up_me) ? up_lab : INF, cb = conn_upleft(me, ulo) ? ull : INF,
                cc = conn_upright(me, uro) ? url : INF;
      const int cand = min(ca, min(cb, cc));
      const uint32_t hm = __ballot_sync(0xffffffffu, conn_left(me, left));
      const uint32_t stops = ~hm | 1u;
      const int head = 31 - __clz(stops & (0xffffffffu >> (31 - lane)));
      const uint32_t above = stops & ~(0xffffffffu >> (31 - lane));
      // smallest candidate of the run.  Usually every candidate in the 32-block row carries the same name (one
      // component passes through): two warp reductions + a ballot.  Otherwise a segmented min-scan over the runs.
      const int mn = __reduce_min_sync(0xffffffffu, cand);
      const int mx = __reduce_max_sync(0xffffffffu, cand == INF ? -1 : cand);
      int runmin;
      if (mx == -1 || mx == mn) {
        const uint32_t run = (above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffffffffu) & ~((1u << head) - 1u);
        runmin = (__ballot_sync(0xffffffffu, cand != INF) & run) ? mn : INF;
      } else {
        int x = cand;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, x, d);
          if (lane - d >= head) x = min(x, v);
        }
        runmin = __shfl_sync(0xffffffffu, x, above ? __ffs(above) - 2 : 31);
      }
      int label = INF;
      if (me) {
        label = (runmin == INF) ? (by * BW + c0 + head) : runmin;
        lab[bi] = label;
        if (runmin == INF && head == lane) occ[bi] = (uint8_t)(me | CC_NAME);   // this block is a NAME: a tree node others point to
      }
      // area of every run goes to its label's word; the runs of this row that share the first run's label are summed
      // with one warp reduction and carried in registers from row to row (one atomic per label change, not per run)
      {
        const uint32_t run = (above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffffffffu) & ~((1u << lane) - 1u);
        const uint32_t b0 = __ballot_sync(0xffffffffu, me & 1u), b1 = __ballot_sync(0xffffffffu, me & 2u);
        const uint32_t b2 = __ballot_sync(0xffffffffu, me & 4u), b3 = __ballot_sync(0xffffffffu, me & 8u);
        const bool is_head = me && head == lane;
        const int area = __popc(b0 & run) + __popc(b1 & run) + __popc(b2 & run) + __popc(b3 & run);
        const uint32_t heads = __ballot_sync(0xffffffffu, is_head);
        if (heads) {
          const int n0 = __shfl_sync(0xffffffffu, label, __ffs(heads) - 1);
          const int total = __reduce_add_sync(0xffffffffu, (is_head && label == n0) ? area : 0);
          if (n0 == acc_name) {
            acc_sum += total;
          } else {
            __syncwarp();
            if (acc_sum && lane == 0) atomicAdd(lab + acc_name, acc_sum << CC_IDX_BITS);
            acc_name = n0;
            acc_sum = total;
          }
          if (is_head && label != n0) atomicAdd(lab + label, area << CC_IDX_BITS);
        }
      }
#ifdef CC_TRACE
      if (blockIdx.x == 0 && warp == 0) {
        const uint32_t un = __ballot_sync(0xffffffffu, (ca != INF && ca != runmin) || (cb != INF && cb != runmin && cb != ca) ||
                                                           (cc != INF && cc != runmin && cc != ca && cc != cb));
        if (lane == 0) {
          g_cc_cnt[0] += un != 0;
          g_cc_cnt[1] += __popc(un);
          g_cc_cnt[2] += !(mx == -1 || mx == mn);
          g_cc_cnt[3] += 1;
        }
      }
#endif
      // rare: a run joining differently named components (every candidate of every lane, not just the lane's minimum)
      if (ca != INF && ca != runmin) ufa_union(lab, ca, runmin);
      if (cb != INF && cb != runmin && cb != ca) ufa_union(lab, cb, runmin);
      if (cc != INF && cc != runmin && cc != ca && cc != cb) ufa_union(lab, cc, runmin);
      up_me = me;
      up_lab = label;
    }
    __syncwarp();
    if (acc_sum && lane == 0) atomicAdd(lab + acc_name, acc_sum << CC_IDX_BITS);
  }
  CC_RMARK(0);
  __syncthreads();
  CC_RMARK(1);
  // C. seams between regions (generic lock-free unions; every region is already labelled)
  //    C1: the first row of each band against the last row of the band above, inside the strip
  if (t_begin < t_end) {
    const int chunk = t_begin / BH, by = t_begin - chunk * BH;
    if (by > 0) {
      const int c0 = chunk << 5, bx = c0 + lane;
      const bool in = bx < BW;
      const int bi = by * BW + bx;
      const uint32_t me = in ? occ[bi] : 0u, up = in ? occ[bi - BW] : 0u;
      uint32_t left = __shfl_up_sync(0xffffffffu, me, 1);
      uint32_t ul = __shfl_up_sync(0xffffffffu, up, 1), ur = __shfl_down_sync(0xffffffffu, up, 1);
      if (lane == 0) left = 0u, ul = 0u;
      if (lane == 31) ur = 0u;
      const bool h = conn_left(me, left);
      const bool cu = conn_up(me, up);
      const bool cul = conn_upleft(me, ul) && !(cu && conn_left(up, ul));
      const bool cur = conn_upright(me, ur) && !(cu && conn_left(ur, up));
      const bool cu_prev = __shfl_up_sync(0xffffffffu, cu, 1);
      const bool cu_redundant = lane > 0 && h && cu_prev && conn_left(up, ul);
      if (cu && !cu_redundant) ufa_union(lab, bi, bi - BW);
      if (cul) ufa_union(lab, bi, bi - BW - 1);
      if (cur) ufa_union(lab, bi, bi - BW + 1);
    }
  }
  //    C2: strip seams, one lane per row: left edge block of a strip with its left / upper-left neighbour, and the right
  //    edge block of the strip before it with its upper-right neighbour
  {
    const int rg = (BH + 31) >> 5, ns = (chunks - 1) * rg;
    for (int u = warp; u < ns; u += nwarps) {
      const int sb = 1 + u / rg, by = ((u - (sb - 1) * rg) << 5) + lane;
      if (by < BH) {
        const int bi = by * BW + (sb << 5);
        const uint32_t me = occ[bi], lf = occ[bi - 1];
        if (conn_left(me, lf)) ufa_union(lab, bi, bi - 1);
        if (by > 0) {
          const uint32_t up = occ[bi - BW], ul = occ[bi - BW - 1];
          const bool across = conn_left(up, ul);
          if (conn_upleft(me, ul) && !(conn_up(me, up) && across)) ufa_union(lab, bi, bi - BW - 1);
          if (conn_upright(lf, up) && !(conn_up(lf, ul) && across)) ufa_union(lab, bi - 1, bi - BW);
        }
      }
    }
  }
  __syncthreads();
  CC_RMARK(2);
  // D. every parent pointer written so far targets a NAME block (a run that started a new label): flatten the names
  //    (the only loop-y finds left) and hand the area parked on a name to its root.  Afterwards block -> name -> root is
  //    two plain loads and a root's word is root | area << 14.  Names are a few % of the blocks, so they are first
  //    gathered into a list (128-bit scan of the occupancy bytes) and then handled one per thread with full warps;
  //    a scan that overflows the list flattens the surplus names in place.
  auto flatten = [&](int bi) {
    const int root = ufa_find(lab, bi);
    if (root != bi) {
      ufa_move_area(lab, lab[bi], root);
      lab[bi] = root;
    }
  };
  if (threadIdx.x == 0) *name_count = 0;
  __syncthreads();
  {
    auto push = [&](int bi) {
      const int slot = atomicAdd(name_count, 1);
      if (slot < name_cap) name_list[slot] = (uint16_t)bi;
      else flatten(bi);
    };
    const int nvec = (nb & 3) == 0 ? nb >> 4 : 0;     // the occupancy array starts 4*nb bytes into shared memory
    const uint4* occ4 = reinterpret_cast<const uint4*>(occ);
    for (int v = threadIdx.x; v < nvec; v += nthreads) {
      const uint4 o = occ4[v];
      const uint32_t wds[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        uint32_t m = wds[wi] & 0x10101010u;
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          push((v << 4) + (wi << 2) + (bit >> 3));
        }
      }
    }
    for (int bi = (nvec << 4) + threadIdx.x; bi < nb; bi += nthreads)
      if (occ[bi] & CC_NAME) push(bi);
  }
  __syncthreads();
  CC_RMARK(3);
  {
    const int cnt = min(*name_count, name_cap);
    for (int i = threadIdx.x; i < cnt; i += nthreads) flatten(name_list[i]);
  }
  CC_RMARK(4);
}

// FILL=false: img uint8 -> labels/counts int32.   FILL=true: scores f32 updated in place.
// Shared memory: one int32 per block (union-find parent; a root's word additionally carries area << 14 once the
// areas are accumulated) + one occupancy byte per block = 5 B/block, 80 KB at 256 x 256 -> two CTAs per SM, so one
// image's loads/stores overlap the other's union-find.
template <bool FILL>
__global__ void __launch_bounds__(CC_THREADS, 2)
cc_small_kernel(const void* img_all, int H, int W, int32_t* labels_all, int32_t* counts_all, float* scores_all,
                int max_area, float fill_value, int vec) {
  pdl_enter();
  extern __shared__ int cc_smem[];
  const int BH = H >> 1, BW = W >> 1, nb = BH * BW;
  int* lab = cc_smem;
  uint8_t* occ = reinterpret_cast<uint8_t*>(cc_smem + nb);
  uint16_t* name_list = reinterpret_cast<uint16_t*>(occ + ((nb + 15) & ~15));
  int* name_count = reinterpret_cast<int*>(name_list + CC_NAME_CAP);
  const size_t img_off = (size_t)blockIdx.x * H * W;
  const void* img = FILL ? static_cast<const void*>(scores_all + img_off)
                         : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img_all) + img_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = CC_THREADS / 32;

#ifdef CC_TRACE
  long long tr[6];
  tr[0] = clock64();
#define CC_MARK(i) tr[i] = clock64()
#else
#define CC_MARK(i)
#endif
  // A. occupancy
  if (vec) {
    if (FILL) {  // 2 x float4 -> 2 blocks
      const float* f = reinterpret_cast<const float*>(img);
      const int upr = W >> 2, units = BH * upr;
#pragma unroll 4
      for (int u = threadIdx.x; u < units; u += CC_THREADS) {
        const int by = u / upr, k = u - by * upr;
        const float4 t = *reinterpret_cast<const float4*>(f + (size_t)(2 * by) * W + 4 * k);
        const float4 b = *reinterpret_cast<const float4*>(f + (size_t)(2 * by + 1) * W + 4 * k);
        const uint32_t o0 = (t.x <= 0.f ? 1u : 0u) | (t.y <= 0.f ? 2u : 0u) | (b.x <= 0.f ? 4u : 0u) | (b.y <= 0.f ? 8u : 0u);
        const uint32_t o1 = (t.z <= 0.f ? 1u : 0u) | (t.w <= 0.f ? 2u : 0u) | (b.z <= 0.f ? 4u : 0u) | (b.w <= 0.f ? 8u : 0u);
        *reinterpret_cast<uint16_t*>(occ + by * BW + 2 * k) = (uint16_t)(o0 | (o1 << 8));
      }
    } else {     // 2 x 16 pixels -> 8 blocks
      const uint8_t* p = reinterpret_cast<const uint8_t*>(img);
      const int upr = W >> 4, units = BH * upr;
#pragma unroll 2
      for (int u = threadIdx.x; u < units; u += CC_THREADS) {
        const int by = u / upr, k = u - by * upr;
        const uint4 t = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by) * W + 16 * k);
        const uint4 b = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by + 1) * W + 16 * k);
        uint2 o;
        o.x = occ2_from_u8(t.x, b.x) | (occ2_from_u8(t.y, b.y) << 16);
        o.y = occ2_from_u8(t.z, b.z) | (occ2_from_u8(t.w, b.w) << 16);
        *reinterpret_cast<uint2*>(occ + by * BW + 8 * k) = o;
      }
    }
  } else {
    for (int bi = threadIdx.x; bi < nb; bi += CC_THREADS) occ[bi] = (uint8_t)load_occ<FILL>(img, H, W, bi / BW, bi % BW, 0.f);
  }
  __syncthreads();
  CC_MARK(1);
  cc_label_region(lab, occ, BH, BW, name_list, name_count, CC_NAME_CAP);
  __syncthreads();
  CC_MARK(4);
  // E. outputs
  if (FILL) {
    float* sc = scores_all + img_off;
    for (int bi = threadIdx.x; bi < nb; bi += CC_THREADS) {
      const uint32_t me = occ[bi];
      if (!me) continue;
      const int rt = lab[lab[bi] & CC_IDX_MASK] & CC_IDX_MASK;
      if ((int)((uint32_t)lab[rt] >> CC_IDX_BITS) > max_area) continue;
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
      // root block index -> label value needs root / BW: multiply-high by ceil(2^32 / BW) is exact for root < 2^14
      const uint32_t magic = BW > 1 ? (uint32_t)((0x100000000ull + BW - 1) / BW) : 0u;
      const int segs = (W + 127) >> 7;
#pragma unroll 2
      for (int task = warp; task < H * segs; task += nwarps) {
        const int r = task / segs, c = ((task - r * segs) << 7) + (lane << 2);
        if (c >= W) continue;
        const int b0 = (r >> 1) * BW + (c >> 1);
        const int sh = (r & 1) << 1;
        const uint32_t oo = *reinterpret_cast<const uint16_t*>(occ + b0);
        const uint32_t o0 = (oo & 0xFFu) >> sh, o1 = (oo >> 8) >> sh;
        int l0 = 0, l1 = 0, n0 = 0, n1 = 0;
        if (o0 & 3u) {
          const int rt = lab[lab[b0] & CC_IDX_MASK] & CC_IDX_MASK;
          const int q = BW > 1 ? (int)__umulhi((uint32_t)rt, magic) : rt;
          l0 = q * 2 * W + (rt - q * BW) * 2 + 1;
          n0 = (int)((uint32_t)lab[rt] >> CC_IDX_BITS);
        }
        if (o1 & 3u) {
          const int rt = lab[lab[b0 + 1] & CC_IDX_MASK] & CC_IDX_MASK;
          const int q = BW > 1 ? (int)__umulhi((uint32_t)rt, magic) : rt;
          l1 = q * 2 * W + (rt - q * BW) * 2 + 1;
          n1 = (int)((uint32_t)lab[rt] >> CC_IDX_BITS);
        }
        const int4 lv = make_int4((o0 & 1u) ? l0 : 0, (o0 & 2u) ? l0 : 0, (o1 & 1u) ? l1 : 0, (o1 & 2u) ? l1 : 0);
        const int4 cv = make_int4((o0 & 1u) ? n0 : 0, (o0 & 2u) ? n0 : 0, (o1 & 1u) ? n1 : 0, (o1 & 2u) ? n1 : 0);
        __stcs(reinterpret_cast<int4*>(labels + (size_t)r * W + c), lv);
        __stcs(reinterpret_cast<int4*>(counts + (size_t)r * W + c), cv);
      }
    } else {
      for (int px = threadIdx.x; px < H * W; px += CC_THREADS) {
        const int r = px / W, c = px % W;
        const int b0 = (r >> 1) * BW + (c >> 1);
        const bool fg = (occ[b0] >> (((r & 1) << 1) | (c & 1))) & 1u;
        int l = 0, n = 0;
        if (fg) {
          const int rt = lab[lab[b0] & CC_IDX_MASK] & CC_IDX_MASK;
          l = (rt / BW) * 2 * W + (rt % BW) * 2 + 1;
          n = (int)((uint32_t)lab[rt] >> CC_IDX_BITS);
        }
        labels[px] = l;
        counts[px] = n;
      }
    }
  }
#ifdef CC_TRACE
  __syncthreads();
  CC_MARK(5);
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    printf("cc_small<%d> cta %d: A %lld  B-D (region labeller) %lld  E %lld cycles\n", (int)FILL, blockIdx.x, tr[1] - tr[0],
           tr[4] - tr[1], tr[5] - tr[4]);
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("   cta 0 thread 0: B (own band) %lld  wait for the other bands %lld  C %lld  D scan %lld  D flatten %lld  names %d\n",
           g_cc_tr[0] - tr[1], g_cc_tr[1] - g_cc_tr[0], g_cc_tr[2] - g_cc_tr[1], g_cc_tr[3] - g_cc_tr[2],
           g_cc_tr[4] - g_cc_tr[3], *name_count);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    printf("   warp 0: %d rows, %d with a union (%d lanes), %d through the segmented scan\n", g_cc_cnt[3], g_cc_cnt[0], g_cc_cnt[1], g_cc_cnt[2]);
    g_cc_cnt[0] = g_cc_cnt[1] = g_cc_cnt[2] = g_cc_cnt[3] = 0;
  }
#endif
}

// ------------------------------------------------------------------ tiled path (images larger than 256 x 256)
// Tile = 64 x 128 pixels = 32 x 64 blocks, labelled in shared memory by the same region labeller as the small path.
// Global state is COMPACT (per 2x2 block, not per pixel):
//   forest[block] = parent block index << 4 | occupancy nibble      (1 B / pixel; doubles as the occupancy array)
//   area[block]   = pixel count, valid at tile-local roots only     (sparse writes, never cleared)
//   list          = the tile-local roots that touch their tile's border (the only ones a border union can demote)
//   cc_t_label  : occupancy -> shared-memory labelling -> forest words, root areas, open-root list
//   cc_t_border : lock-free min-root unions across tile borders on the forest words
//   cc_t_areas  : every listed root that is no longer a root hands its area to its new root
//   cc_t_final  : block -> root (usually one hop), 8-byte stores of labels / areas (or sparse fill of small holes)
// DRAM traffic ~11 B/pixel for 9 algorithmic (the first version scattered the forest into the labels array, kept separate
// occupancy and area arrays and cleared one of them: ~23 B/pixel).
constexpr int TBH = 32, TBW = 64, T_THREADS = 256, T_NAME_CAP = 512;
constexpr int T_BORDER = 2 * (TBH + TBW);   // open-root list entries per tile (one per border block at most)

__device__ __forceinline__ int gfind(const volatile uint32_t* f, int n) {
  uint32_t w = f[n];
  while ((int)(w >> 4) != n) {
    n = (int)(w >> 4);
    w = f[n];
  }
  return n;
}
// min-root union on words parent << 4 | occupancy: the nibble of a node never changes, so comparing whole words orders
// by parent
__device__ __forceinline__ void gunion(uint32_t* f, int a, int b) {
  int ra = gfind(f, a), rb = gfind(f, b);
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const uint32_t nib = reinterpret_cast<const volatile uint32_t*>(f)[ra] & 15u;
    const uint32_t old = atomicMin(f + ra, ((uint32_t)rb << 4) | nib);
    if ((int)(old >> 4) == ra) break;
    ra = gfind(f, (int)(old >> 4));
    rb = gfind(f, rb);
  }
  // path compression: a large component spans hundreds of tiles, and without it every border union walks the whole
  // chain of tile roots through L2 (the border kernel took 81 us for 1 M threads)
  const int r = ra < rb ? ra : rb;
  for (int side = 0; side < 2; ++side) {
    int n = side ? b : a;
    while (n > r) {
      const uint32_t w = reinterpret_cast<const volatile uint32_t*>(f)[n];
      const int p = (int)(w >> 4);
      if (p == n) break;
      if (p > r) atomicMin(f + n, ((uint32_t)r << 4) | (w & 15u));
      n = p;
    }
  }
}

template <bool FILL>
__global__ void __launch_bounds__(T_THREADS)
cc_t_label(const void* img_all, const float* scores_all, int H, int W, int vec, uint32_t* forest_all, int32_t* area_all,
           int* list_count, int2* list) {
  pdl_enter();
  __shared__ int lab[TBH * TBW];
  __shared__ __align__(16) uint8_t occ[TBH * TBW];
  __shared__ uint16_t names[T_NAME_CAP];
  __shared__ int name_count;
  const int BHg = H >> 1, BWg = W >> 1;
  const int z = blockIdx.z;
  const size_t off = (size_t)z * H * W;
  const void* img = FILL ? static_cast<const void*>(scores_all + off)
                         : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img_all) + off);
  const int by0 = blockIdx.y * TBH, bx0 = blockIdx.x * TBW;
  // A. occupancy of the tile (blocks outside the image are empty)
  if (vec && !FILL) {   // 2 x 16 pixels -> 8 blocks per thread: exactly one unit per thread
    const uint8_t* p = reinterpret_cast<const uint8_t*>(img);
    for (int u = threadIdx.x; u < TBH * (TBW / 8); u += T_THREADS) {
      const int ly = u / (TBW / 8), k = u % (TBW / 8);
      const int by = by0 + ly, bx = bx0 + 8 * k;
      uint2 o = make_uint2(0u, 0u);
      if (by < BHg && bx < BWg) {   // W % 16 == 0: a unit is inside the image or completely outside
        const uint4 t = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by) * W + 2 * bx);
        const uint4 b = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by + 1) * W + 2 * bx);
        o.x = occ2_from_u8(t.x, b.x) | (occ2_from_u8(t.y, b.y) << 16);
        o.y = occ2_from_u8(t.z, b.z) | (occ2_from_u8(t.w, b.w) << 16);
      }
      *reinterpret_cast<uint2*>(occ + ly * TBW + 8 * k) = o;
    }
  } else if (vec && FILL) {   // 2 x float4 -> 2 blocks
    const float* f = reinterpret_cast<const float*>(img);
#pragma unroll 4
    for (int u = threadIdx.x; u < TBH * (TBW / 2); u += T_THREADS) {
      const int ly = u / (TBW / 2), k = u % (TBW / 2);
      const int by = by0 + ly, bx = bx0 + 2 * k;
      uint32_t o = 0;
      if (by < BHg && bx < BWg) {   // W % 4 == 0
        const float4 t = *reinterpret_cast<const float4*>(f + (size_t)(2 * by) * W + 2 * bx);
        const float4 b = *reinterpret_cast<const float4*>(f + (size_t)(2 * by + 1) * W + 2 * bx);
        o = (t.x <= 0.f ? 1u : 0u) | (t.y <= 0.f ? 2u : 0u) | (b.x <= 0.f ? 4u : 0u) | (b.y <= 0.f ? 8u : 0u);
        o |= ((t.z <= 0.f ? 1u : 0u) | (t.w <= 0.f ? 2u : 0u) | (b.z <= 0.f ? 4u : 0u) | (b.w <= 0.f ? 8u : 0u)) << 8;
      }
      *reinterpret_cast<uint16_t*>(occ + ly * TBW + 2 * k) = (uint16_t)o;
    }
  } else {
    for (int i = threadIdx.x; i < TBH * TBW; i += T_THREADS) {
      const int by = by0 + i / TBW, bx = bx0 + i % TBW;
      occ[i] = (by < BHg && bx < BWg) ? (uint8_t)load_occ<FILL>(img, H, W, by, bx, 0.f) : (uint8_t)0;
    }
  }
  __syncthreads();
  cc_label_region(lab, occ, TBH, TBW, names, &name_count, T_NAME_CAP);
  __syncthreads();
  // roots that reach the tile border (bit 5 of the root's occupancy byte; every writer stores the same bit)
  for (int k = threadIdx.x; k < T_BORDER; k += T_THREADS) {
    int ly, lx;
    if (k < TBW) { ly = 0; lx = k; }
    else if (k < 2 * TBW) { ly = TBH - 1; lx = k - TBW; }
    else if (k < 2 * TBW + TBH) { ly = k - 2 * TBW; lx = 0; }
    else { ly = k - 2 * TBW - TBH; lx = TBW - 1; }
    const int i = ly * TBW + lx;
    if (occ[i] & 0xFu) {
      const int root = lab[lab[i] & CC_IDX_MASK] & CC_IDX_MASK;
      occ[root] = (uint8_t)(occ[root] | 0x20u);
    }
  }
  __syncthreads();
  uint32_t* forest = forest_all + (size_t)z * BHg * BWg;
  int32_t* area = area_all + (size_t)z * BHg * BWg;
  for (int i = threadIdx.x; i < TBH * TBW; i += T_THREADS) {
    const int ly = i / TBW, lx = i % TBW, by = by0 + ly, bx = bx0 + lx;
    if (by >= BHg || bx >= BWg) continue;
    const uint32_t o = occ[i];
    const int gb = by * BWg + bx;
    if (o & 0xFu) {
      const int root = lab[lab[i] & CC_IDX_MASK] & CC_IDX_MASK;
      const int groot = (by0 + root / TBW) * BWg + bx0 + root % TBW;
      forest[gb] = ((uint32_t)groot << 4) | (o & 0xFu);
      if (root == i) {
        area[gb] = (int)((uint32_t)lab[i] >> CC_IDX_BITS);
        if (o & 0x20u) list[atomicAdd(list_count, 1)] = make_int2(z, gb);
      }
    } else {
      forest[gb] = (uint32_t)gb << 4;
    }
  }
}

// One warp per tile edge piece: the two 32-block halves of the top row (lanes = consecutive blocks: coalesced loads,
// neighbours by shuffle, and the same pruning as inside a tile -- a run that crosses the border costs ONE union, not one
// per block), the left column and the right column (lanes = rows).
__global__ void cc_t_border(int H, int W, uint32_t* forest_all) {
  pdl_enter();
  const int BH = H >> 1, BW = W >> 1;
  const int tiles_x = (BW + TBW - 1) / TBW, tiles_y = (BH + TBH - 1) / TBH;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= tiles_x * tiles_y * 4) return;
  const int tile = gw >> 2, piece = gw & 3;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int by0 = ty * TBH, bx0 = tx * TBW;
  uint32_t* forest = forest_all + (size_t)blockIdx.z * BH * BW;
  auto nib = [&](int y, int x) -> uint32_t {
    return (y >= 0 && y < BH && x >= 0 && x < BW) ? (forest[y * BW + x] & 15u) : 0u;   // a word's nibble never changes
  };
  if (piece < 2) {   // top row, blocks bx0 + 32*piece + lane
    const int by = by0, bx = bx0 + 32 * piece + lane;
    if (by == 0 || by >= BH) return;
    const uint32_t me = nib(by, bx), up = nib(by - 1, bx);
    uint32_t left = __shfl_up_sync(0xffffffffu, me, 1), ul = __shfl_up_sync(0xffffffffu, up, 1);
    uint32_t ur = __shfl_down_sync(0xffffffffu, up, 1);
    if (lane == 0) {
      left = piece == 1 ? nib(by, bx - 1) : 0u;   // the block before the tile's first one belongs to another tile
      ul = nib(by - 1, bx - 1);
    }
    if (lane == 31) ur = nib(by - 1, bx + 1);
    const bool h = conn_left(me, left);          // same tile: already one component
    const bool cu = conn_up(me, up);
    const bool cul = conn_upleft(me, ul) && !(cu && conn_left(up, ul));
    const bool cur = conn_upright(me, ur) && !(cu && conn_left(ur, up));
    const bool cu_prev = __shfl_up_sync(0xffffffffu, cu, 1);
    const bool cu_redundant = lane > 0 && h && cu_prev && conn_left(up, ul);
    const int idx = by * BW + bx;
    if (cu && !cu_redundant) gunion(forest, idx, idx - BW);
    if (cul) gunion(forest, idx, idx - BW - 1);
    if (cur) gunion(forest, idx, idx - BW + 1);
  } else if (piece == 2) {   // left column, rows by0 + lane
    const int by = by0 + lane, bx = bx0;
    if (bx == 0 || by >= BH) return;
    const uint32_t me = nib(by, bx);
    if (!me) return;
    const uint32_t lf = nib(by, bx - 1);
    const int idx = by * BW + bx;
    if (conn_left(me, lf)) gunion(forest, idx, idx - 1);
    if (lane > 0) {          // the tile's first row is the top-row piece's business
      const uint32_t up = nib(by - 1, bx), ul = nib(by - 1, bx - 1);
      if (conn_upleft(me, ul) && !(conn_up(me, up) && conn_left(up, ul))) gunion(forest, idx, idx - BW - 1);
    }
  } else {   // right column
    const int by = by0 + lane, bx = bx0 + TBW - 1;
    if (lane == 0 || bx + 1 >= BW || by >= BH) return;
    const uint32_t me = nib(by, bx);
    if (!me) return;
    const uint32_t up = nib(by - 1, bx), ur = nib(by - 1, bx + 1);
    if (conn_upright(me, ur) && !(conn_up(me, up) && conn_left(ur, up))) gunion(forest, by * BW + bx, (by - 1) * BW + bx + 1);
  }
}

__global__ void cc_t_areas(int H, int W, uint32_t* forest_all, int32_t* area_all,
                           const int* __restrict__ list_count, const int2* __restrict__ list) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *list_count) return;
  const int2 e = list[i];
  const size_t nblk = (size_t)(H >> 1) * (W >> 1);
  uint32_t* forest = forest_all + (size_t)e.x * nblk;
  const int g = gfind(forest, e.y);
  if (g != e.y) {
    atomicAdd(area_all + (size_t)e.x * nblk + g, area_all[(size_t)e.x * nblk + e.y]);
    forest[e.y] = ((uint32_t)g << 4) | (forest[e.y] & 15u);   // flatten: cc_t_final then needs at most two hops
  }
}

template <bool FILL>
__global__ void cc_t_final(int H, int W, const uint32_t* __restrict__ forest_all, const int32_t* __restrict__ area_all,
                           int32_t* labels_all, int32_t* counts_all, float* scores_all, int max_area, float fill_value) {
  pdl_enter();
  const int BW = W >> 1, BH = H >> 1;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
  if (bx >= BW || by >= BH) return;
  const size_t nblk = (size_t)BH * BW;
  const uint32_t* forest = forest_all + (size_t)blockIdx.z * nblk;
  const uint32_t w0 = forest[by * BW + bx];
  const uint32_t me = w0 & 15u;
  const size_t off = (size_t)blockIdx.z * H * W;
  const int idx = 2 * by * W + 2 * bx;
  int y = 0, n = 0;
  if (me) {
    int root = (int)(w0 >> 4);
    uint32_t w1 = forest[root];
    while ((int)(w1 >> 4) != root) {
      root = (int)(w1 >> 4);
      w1 = forest[root];
    }
    n = area_all[(size_t)blockIdx.z * nblk + root];
    const int ry = root / BW;
    y = 2 * ry * W + 2 * (root - ry * BW) + 1;
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
    __stcs(reinterpret_cast<int2*>(L + idx), make_int2((me & 1u) ? y : 0, (me & 2u) ? y : 0));
    __stcs(reinterpret_cast<int2*>(L + idx + W), make_int2((me & 4u) ? y : 0, (me & 8u) ? y : 0));
    __stcs(reinterpret_cast<int2*>(C + idx), make_int2((me & 1u) ? n : 0, (me & 2u) ? n : 0));
    __stcs(reinterpret_cast<int2*>(C + idx + W), make_int2((me & 4u) ? n : 0, (me & 8u) ? n : 0));
  }
}

bool small_ok(int h, int w) { return (h / 2) * (w / 2) <= CC_MAX_BLOCKS; }
size_t small_smem(int h, int w) {
  const size_t nb = (size_t)(h / 2) * (w / 2);
  return nb * 4 + ((nb + 15) & ~(size_t)15) + CC_NAME_CAP * 2 + 16;
}

template <bool FILL>
int run(const void* img, float* scores, int n, int h, int w, int32_t* labels, int32_t* counts, int max_area,
        float fill_value, void* ws, size_t ws_bytes, cudaStream_t stream) {
  VLS_REQUIRE(n >= 0 && h >= 0 && w >= 0, "cc: negative dimension");
  VLS_REQUIRE((h % 2) == 0, "height must be an even number");  // connected_components.cu:226
  VLS_REQUIRE((w % 2) == 0, "width must be an even number");   // connected_components.cu:227
  if (n == 0 || h == 0 || w == 0) return 0;
  if (small_ok(h, w)) {
    static unsigned long long attr[2] = {0, 0};
    if (first_use_on_device(&attr[FILL])) {
      VLS_CUDA(cudaFuncSetAttribute(cc_small_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)small_smem(256, 256)));
    }
    // 128-bit loads need 16-byte aligned rows: W % 16 (uint8) / W % 4 (f32) and an aligned base
    const int vec = FILL ? ((w % 4) == 0 && ((uintptr_t)scores % 16) == 0)
                         : ((w % 16) == 0 && ((uintptr_t)img % 16) == 0);
    VLS_CUDA(launch_k(cc_small_kernel<FILL>, dim3(n), dim3(CC_THREADS), small_smem(h, w), stream, img, h, w, labels, counts, scores, max_area, fill_value, vec));
    VLS_POST_LAUNCH(1);
    return 0;
  }
  const size_t need = cc_workspace_bytes(n, h, w, FILL);
  VLS_REQUIRE(ws != nullptr && ws_bytes >= need, "cc: workspace too small (%zu < %zu)", ws_bytes, need);
  VLS_REQUIRE(((uintptr_t)ws % 16) == 0, "cc: workspace must be 16-byte aligned");
  const int BH = h / 2, BW = w / 2;
  const size_t nblk1 = (size_t)BH * BW;                       // blocks per image
  const int tiles_x = (BW + TBW - 1) / TBW, tiles_y = (BH + TBH - 1) / TBH;
  uint32_t* forest = reinterpret_cast<uint32_t*>(ws);
  int32_t* area = reinterpret_cast<int32_t*>(forest + nblk1 * n);
  int* list_count = reinterpret_cast<int*>(area + nblk1 * n);   // two counters (one per half), 16 bytes reserved
  int2* list = reinterpret_cast<int2*>(list_count + 4);
  const size_t list_per_image = (size_t)tiles_x * tiles_y * T_BORDER;
  VLS_CUDA(cudaMemsetAsync(list_count, 0, 16, stream));
  const int vec = FILL ? ((w % 4) == 0 && ((uintptr_t)scores % 16) == 0) : ((w % 16) == 0 && ((uintptr_t)img % 16) == 0);
  // The labelling kernel is issue bound and the final pass bandwidth bound, so a large batch is split in two halves
  // that run on two streams: one half's output pass overlaps the other half's labelling.
  const int halves = n >= 8 ? 2 : 1;
  for (int hf = halves - 1; hf >= 0; --hf) {   // the forked half first: the fork point precedes all of this call's work
    const int i0 = hf == 0 ? 0 : n / 2, cnt = halves == 1 ? n : (hf == 0 ? n / 2 : n - n / 2);
    cudaStream_t st = stream;
    if (hf == 1) VLS_TRY(fork_begin(3, stream, &st));
    const size_t px0 = (size_t)i0 * h * w;
    const void* img_h = FILL ? nullptr : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img) + px0);
    float* sc_h = FILL ? scores + px0 : nullptr;
    uint32_t* f_h = forest + nblk1 * i0;
    int32_t* a_h = area + nblk1 * i0;
    int2* l_h = list + list_per_image * i0;
    const long long list_cap = (long long)cnt * list_per_image;
    VLS_CUDA(launch_k(cc_t_label<FILL>, dim3(tiles_x, tiles_y, cnt), dim3(T_THREADS), 0, st, img_h, sc_h, h, w, vec, f_h, a_h,
                      list_count + hf, l_h));
    const long long border = (long long)tiles_x * tiles_y * 4 * 32;   // four warps per tile
    VLS_CUDA(launch_k(cc_t_border, dim3((unsigned)((border + 255) / 256), 1, cnt), dim3(256), 0, st, h, w, f_h));
    VLS_CUDA(launch_k(cc_t_areas, dim3((unsigned)((list_cap + 255) / 256)), dim3(256), 0, st, h, w, f_h, a_h, list_count + hf, l_h));
    dim3 blk(128, 1, 1), grd((BW + 127) / 128, BH, cnt);
    VLS_CUDA(launch_k(cc_t_final<FILL>, grd, blk, 0, st, h, w, f_h, a_h, FILL ? nullptr : labels + px0, FILL ? nullptr : counts + px0,
                      sc_h, max_area, fill_value));
    VLS_POST_LAUNCH(4);
  }
  if (halves == 2) VLS_TRY(fork_join(3, stream));
  return 0;
}

}  // namespace

// tiled path: forest + area words per 2x2 block, the open-root list and its counter
size_t cc_workspace_bytes(int n, int h, int w, bool fill) {
  (void)fill;
  if (n <= 0 || h <= 0 || w <= 0 || small_ok(h, w)) return 0;
  const size_t nblk = (size_t)n * (h / 2) * (w / 2);
  const size_t tiles = (size_t)n * ((w / 2 + TBW - 1) / TBW) * ((h / 2 + TBH - 1) / TBH);
  return nblk * 8 + 16 + tiles * T_BORDER * 8 + 256;
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
