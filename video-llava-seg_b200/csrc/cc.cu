// Connected components (8-connectivity) labelling + per-pixel component area, replacing
// sam2/csrc/connected_components.cu (6 launches per image in a host loop, :245-275) by ONE launch
// for the whole batch at the production shape [B,1,256,256], and the fused hole-filling of
// sam2/utils/misc.py:312-338.
//
// Result contract (bit-exact with the reference): union-find over 2x2 pixel blocks whose root is
// always the minimum index, so   label(pixel) = 1 + min over its component of ((r&~1)*W + (c&~1))
// and count(pixel) = component area; background pixels get 0 / 0.
//
// Small path (one CTA of 512 threads per image, (H/2)*(W/2) <= 16384 blocks, two CTAs per SM): everything lives in shared
// memory (a union-find word per block that doubles as the area counter, an occupancy byte per block, four 32-bit
// occupancy planes per chunk-row of 32 blocks).
//   A. occupancy of every 2x2 block straight from global memory with 128-bit loads (16 pixels x 2 rows per thread)
//   P. occupancy planes (one warp per chunk-row, four ballots); every occupied block's word := its run's head block
//   U. run-based unions, ONE THREAD per chunk-row: horizontal runs, the upper runs they touch and the seams to the
//      neighbouring chunks are bit operations on the planes; one lock-free atomicMin min-root union per (run, upper run)
//   F. every run head is re-parented to its root and adds its run's pixel count to the root's word
//   E. block -> head -> root -> area with plain loads; 128-bit streaming stores of labels and areas
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
constexpr int CC_MAX_CHUNK_ROWS = 1024;               // chunk-rows (32 blocks of one block row) per image on the small path

// Every node on the path n -> ... -> r gets parent r.  Only valid while every path stays inside its own set: links are made
// with compare-and-swap on ROOTS (uf_union).  The r1 labeller linked with atomicMin, which re-parents a node that has just
// stopped being a root (and repairs that by uniting its former parent as well); for a moment the path of a node then leads
// into a set it is not yet united with, a concurrent compression lowers the parents of THAT set's nodes and cuts real links
// (found with tools/debug_cc.py: one launch in ~50 split a component).
__device__ __forceinline__ void uf_compress(int* s, int n, int r) {
  while (n > r) {
    const int p = reinterpret_cast<const volatile int*>(s)[n];
    if (p == n) break;
    if (p > r) atomicMin(s + n, r);
    n = p;
  }
}

// min-root union (lock-free: a root is linked below the other set's root by compare-and-swap, so a word only changes from
// "root" to "child" once and paths never leave their set).  The two finds advance together (two independent loads per
// step); paths are compressed only when one of them was longer than a hop (a fresh head joining a flat tree needs nothing).
template <bool COMPRESS = true>
__device__ __forceinline__ void uf_union(int* s, int a, int b) {
  const volatile int* v = s;
  int ra = a, rb = b, pa = v[a], pb = v[b], hops = 0;
  while (pa != ra || pb != rb) {
    ra = pa;
    rb = pb;
    pa = v[ra];
    pb = v[rb];
    ++hops;
  }
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const int old = atomicCAS(s + ra, ra, rb);  // ra > rb
    if (old == ra) break;
    ra = old;                                   // ra had been linked meanwhile: carry on from its parent
    pa = v[ra];
    pb = v[rb];
    while (pa != ra || pb != rb) {
      ra = pa;
      rb = pb;
      pa = v[ra];
      pb = v[rb];
    }
    hops = 2;
  }
  if (COMPRESS && hops > 1) {
    const int r = ra < rb ? ra : rb;
    uf_compress(s, a, r);
    uf_compress(s, b, r);
  }
}

// find on words parent | area << 14 (a ROOT's word carries its component's area once the areas are accumulated)
__device__ __forceinline__ int ufa_find(const volatile int* s, int n) {
  int p = s[n] & CC_IDX_MASK;
  while (p != n) {
    n = p;
    p = s[n] & CC_IDX_MASK;
  }
  return n;
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

// ---- run-based region labeller (r2).
// A CHUNK-ROW = 32 consecutive blocks of one block row, held as four 32-bit occupancy planes (uint4: x = top-left, y =
// top-right, z = bottom-left, w = bottom-right pixels).  On planes, connectivity is a handful of bit operations for 32
// blocks at once, so ONE THREAD owns a chunk-row (the r1 labeller spent ~300 warp instructions on it: a warp walked its
// band of rows top-down with shuffles / ballots / warp reductions per row and was instruction-issue bound):
//   hm    = (x|z) & ((y|w) << 1)         block i joins block i-1;  run heads = occupied & ~hm
//   up    : (x|y) & (u.z|u.w)            block i touches the upper block i
//   up-l  : x & (u.w << 1)               ... the upper block i-1;    up-r: y & (u.z >> 1) the upper block i+1
// A run is named after its head block; only heads are union-find nodes.
//   U1  hooks: for every upper run a run touches (clz / ffs on the head masks, each upper run once) and for the seams to the
//       neighbouring chunks ONE atomicMin(word[head], other head) -- no find.  When the head already had another parent, the
//       pair (old parent, new one) still has to be united: the thread keeps it (registers) for phase U2.  Every row hooks at the same time, so a tall
//       component is now a parent chain as long as it is tall -- finds on it are what made the first version of this
//       labeller slow (45 k cycles of unions, 41-67 k of flattening) --
//   J   which pointer jumping (word[h] = word[word[h]], all heads in parallel, until nothing changes: log2(height) rounds)
//       collapses;
//   U2  the deferred pairs are united on the now flat trees (lock-free min-root unions);
//   F   every head is re-parented to its root and adds its run's pixel count to the root's word.
__device__ __forceinline__ uint32_t cc_hm(uint4 p) { return (p.x | p.z) & ((p.y | p.w) << 1); }
__device__ __forceinline__ uint32_t cc_heads(uint4 p) { return (p.x | p.y | p.z | p.w) & ~cc_hm(p); }
// the run that starts at head bit s: bits s .. e-1, e = first bit above s whose block does not join its left neighbour
__device__ __forceinline__ uint32_t cc_run_mask(uint32_t hm, int s) {
  const uint32_t stop = ~hm & ~((2u << s) - 1u);
  const uint32_t below_e = stop ? ((stop & (0u - stop)) - 1u) : 0xffffffffu;
  return below_e & ~((1u << s) - 1u);
}
constexpr int CC_PAIR_CAP = 2048;   // deferred unions (two 16-bit block indices per entry)

// head (bit index) of the run that contains the occupied block `bit`
__device__ __forceinline__ int cc_head_of(uint32_t heads, int bit) { return 31 - __clz(heads & (0xffffffffu >> (31 - bit))); }


#ifdef CC_TRACE
__device__ long long g_cc_tr[8];
#define CC_RMARK(i)                                                  \
  do {                                                               \
    if (threadIdx.x == 0 && blockIdx.x == 0) g_cc_tr[i] = clock64(); \
  } while (0)
#else
#define CC_RMARK(i)
#endif

// Occupancy planes of every chunk-row from the occupancy bytes (one warp per chunk-row, four ballots): the generic way;
// the 128-bit uint8 path of the small kernel writes the planes directly while it loads the image.
__device__ __forceinline__ void cc_planes_from_occ(const uint8_t* occ, int BH, int BW, uint4* planes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int chunks = (BW + 31) >> 5, nrows = BH * chunks;
  const uint32_t magic = chunks > 1 ? (uint32_t)((0x100000000ull + chunks - 1) / chunks) : 0u;
#pragma unroll 2
  for (int r = warp; r < nrows; r += nwarps) {
    const int by = chunks > 1 ? (int)__umulhi((uint32_t)r, magic) : r, ch = r - by * chunks;
    const int bx = (ch << 5) + lane;
    const uint32_t o = bx < BW ? occ[by * BW + bx] : 0u;
    uint4 p;
    p.x = __ballot_sync(0xffffffffu, o & 1u);
    p.y = __ballot_sync(0xffffffffu, o & 2u);
    p.z = __ballot_sync(0xffffffffu, o & 4u);
    p.w = __ballot_sync(0xffffffffu, o & 8u);
    if (lane == 0) planes[r] = p;
  }
}

// Labels the BH x BW block region described by `planes` (BH * ceil(BW / 32) chunk-row records, complete and followed by
// __syncthreads on entry), executed by the whole CTA (blockDim.x a multiple of 32).  `lab` has one word per block (row pitch
// BW); only the words of run heads are used.  `list` receives the block indices of all run heads (up to BH * BW
// entries: blocks that hold only left-column pixels are all heads): phases U1 / J / F run ONE HEAD PER THREAD over that list, so a chunk-row full of speckle (16 runs) does not hold up a
// warp whose other lanes own one run each.  `ctl` = two counters; `pairs` = CC_PAIR_CAP words.  On return (after the caller's __syncthreads): every head's
// word is its root (the smallest block index of the component) and a root's word is root | area << 14.  cc_root_of() maps
// a block to its root.
__device__ __forceinline__ void cc_label_region(int* lab, const uint4* planes, int BH, int BW, uint32_t* pairs, int* ctl,
                                                uint16_t* list) {
  const int nthreads = blockDim.x, lane = threadIdx.x & 31;
  const int chunks = (BW + 31) >> 5, nrows = BH * chunks, nb = BH * BW;
  const uint32_t magic = chunks > 1 ? (uint32_t)((0x100000000ull + chunks - 1) / chunks) : 0u;   // r / chunks, exact for r < 2^16
  const uint32_t bw_magic = BW > 1 ? (uint32_t)((0x100000000ull + BW - 1) / BW) : 0u;             // idx / BW, exact for idx < 2^14
  int* head_count = ctl;
  int* pair_count = ctl + 1;
  // I. every word its own root
  if ((nb & 3) == 0 && (reinterpret_cast<uintptr_t>(lab) & 15) == 0) {
    for (int i = threadIdx.x * 4; i < nb; i += nthreads * 4) *reinterpret_cast<int4*>(lab + i) = make_int4(i, i + 1, i + 2, i + 3);
  } else {
    for (int i = threadIdx.x; i < nb; i += nthreads) lab[i] = i;
  }
  if (threadIdx.x < 2) ctl[threadIdx.x] = 0;
  __syncthreads();
  // L. list of the run heads (thread per chunk-row; one atomic per warp)
  for (int r0 = threadIdx.x - lane; r0 < nrows; r0 += nthreads) {
    const int r = r0 + lane;
    uint32_t hh = 0u;
    int base = 0;
    if (r < nrows) {
      const int by = chunks > 1 ? (int)__umulhi((uint32_t)r, magic) : r, ch = r - by * chunks;
      hh = cc_heads(planes[r]);
      base = by * BW + (ch << 5);
    }
    const int cnt = __popc(hh);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    int wbase = 0;
    if (lane == 31 && incl) wbase = atomicAdd(head_count, incl);
    int pos = __shfl_sync(0xffffffffu, wbase, 31) + incl - cnt;
    while (hh) {
      list[pos++] = (uint16_t)(base + __ffs(hh) - 1);
      hh &= hh - 1;
    }
  }
  __syncthreads();
  CC_RMARK(0);
  const int nheads = *head_count;
  // a > b.  old == a: a was a root and now hangs below b.  Otherwise a already had the parent `old`: the smaller of the two
  // is its parent now and the pair (old, b) is still to be united (phase U2).
  auto hook = [&](int a, int b) {
    const int old = atomicMin(lab + a, b);
    if (old != a && old != b) {
      const int slot = atomicAdd(pair_count, 1);
      if (slot < CC_PAIR_CAP) pairs[slot] = (uint32_t)old | ((uint32_t)b << 16);
      else uf_union<false>(lab, old, b);   // list full: unite now (hooks still re-parent heads, so without compression)
    }
  };
  // U1. hooks of every run (all hooks of a head come from its one thread, in sequence)
  for (int e = threadIdx.x; e < nheads; e += nthreads) {
    const int idx = list[e];
    const int by = BW > 1 ? (int)__umulhi((uint32_t)idx, bw_magic) : idx, col = idx - by * BW;
    const int ch = col >> 5, s = col & 31, r = by * chunks + ch, base = idx - s;
    const uint4 m = planes[r];
    const uint32_t hm = cc_hm(m), run = cc_run_mask(hm, s);
    if (s == 0 && ch > 0 && ((m.x | m.z) & 1u)) {   // block 0 joins the last block of the chunk to the left
      const uint4 l = planes[r - 1];
      if ((l.y | l.w) >> 31) hook(idx, base - 32 + (31 - __clz(cc_heads(l))));
    }
    if (by == 0) continue;
    const uint4 u = planes[r - chunks];
    const int ubase = base - BW;
    uint32_t upm = (((m.x | m.y) & (u.z | u.w)) & run) | (((m.x & (u.w << 1)) & run) >> 1) | (((m.y & (u.z >> 1)) & run) << 1);
    if (upm) {                                      // upper blocks this run touches
      const uint32_t hmu = cc_hm(u), uheads = (u.x | u.y | u.z | u.w) & ~hmu;
      do {
        const int hj = cc_head_of(uheads, __ffs(upm) - 1);   // the upper run that contains the lowest of them
        upm &= ~cc_run_mask(hmu, hj);
        hook(idx, ubase + hj);
      } while (upm);
    }
    if (s == 0 && ch > 0 && (m.x & 1u)) {           // top-left pixel of block 0 against the bottom-right pixel of the upper-left chunk
      const uint4 ul = planes[r - chunks - 1];
      if (ul.w >> 31) hook(idx, ubase - 32 + (31 - __clz(cc_heads(ul))));
    }
    if ((run >> 31) && ch + 1 < chunks && (m.y >> 31)) {   // top-right pixel of block 31 against the bottom-left pixel of the upper-right chunk
      const uint4 ur = planes[r - chunks + 1];
      if (ur.z & 1u) hook(idx, ubase + 32);
    }
  }
  __syncthreads();
  CC_RMARK(1);
  // J. pointer jumping over the heads until every head points at a root of the hook forest.  A thread keeps the still
  //    unresolved ones of its list entries (e = thread + k * nthreads, k < 32) as a bit mask: after the first rounds only
  //    the heads of tall components are left.
  {
    const volatile int* v = lab;
    uint32_t pend = 0u;
    for (int k = 0, e = threadIdx.x; e < nheads && k < 32; ++k, e += nthreads) pend |= 1u << k;
    int changed;
#ifdef CC_TRACE
    int rounds = 0;
#endif
    do {
      changed = 0;
      uint32_t pp = pend;
      while (pp) {
        const int k = __ffs(pp) - 1;
        pp &= pp - 1;
        const int h = list[threadIdx.x + k * nthreads];
        const int p = v[h];
        const int gp = v[p];
        if (gp == p) {
          pend &= ~(1u << k);                      // the parent is a root (or h itself is): done
        } else {
          lab[h] = v[gp];                          // up to three levels per round (every ancestor is a valid parent)
          changed = 1;
        }
      }
      for (int e = threadIdx.x + 32 * nthreads; e < nheads; e += nthreads) {   // (more than 32 heads per thread: never on the paths in use)
        const int hh = list[e];
        const int pp = v[hh];
        const int gg = v[pp];
        if (gg != pp) {
          lab[hh] = v[gg];
          changed = 1;
        }
      }
#ifdef CC_TRACE
      ++rounds;
#endif
    } while (__syncthreads_or(changed));
#ifdef CC_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0) g_cc_tr[7] = rounds;
#endif
  }
  CC_RMARK(2);
  // U2. the deferred pairs, on the now flat trees
  {
    const int n = min(*pair_count, CC_PAIR_CAP);
    for (int i = threadIdx.x; i < n; i += nthreads) uf_union(lab, (int)(pairs[i] & 0xffffu), (int)(pairs[i] >> 16));
  }
  __syncthreads();
  CC_RMARK(3);
  // F. every head: parent -> root, and its run's pixel count onto the root's word (finds ignore the area bits; only roots
  //    receive area, only non-roots are re-parented, so the plain stores and the atomic adds never touch the same word).
  //    Plain fire-and-forget atomics: summing the areas per root inside the warp first (match.any + redux.add) made this phase
  //    twice as long (6.6 k vs 3.2 k cycles) -- same-address shared-memory atomics without a result are cheap.
  for (int e = threadIdx.x; e < nheads; e += nthreads) {
    const int idx = list[e];
    const int by = BW > 1 ? (int)__umulhi((uint32_t)idx, bw_magic) : idx, col = idx - by * BW;
    const uint4 m = planes[by * chunks + (col >> 5)];
    const uint32_t run = cc_run_mask(cc_hm(m), col & 31);
    const int area = __popc(m.x & run) + __popc(m.y & run) + __popc(m.z & run) + __popc(m.w & run);
    const int root = ufa_find(lab, idx);
    if (root != idx) lab[idx] = root;
    atomicAdd(lab + root, area << CC_IDX_BITS);
  }
  CC_RMARK(4);
}

// root of the component of the occupied block (by, bx) after cc_label_region (+ __syncthreads)
__device__ __forceinline__ int cc_root_of(const int* lab, const uint4* planes, int BW, int by, int bx) {
  const int chunks = (BW + 31) >> 5;
  const uint4 p = planes[by * chunks + (bx >> 5)];
  return lab[by * BW + (bx & ~31) + cc_head_of(cc_heads(p), bx & 31)] & CC_IDX_MASK;
}

// FILL=false: img uint8 -> labels/counts int32.   FILL=true: scores f32 updated in place.
// Shared memory: four plane words per chunk-row, one int32 per block (union-find parent of a run head; a root's word
// additionally carries area << 14), the deferred-pair list, and (generic paths only) one occupancy byte per block:
// 96 KB at 256 x 256 -> two CTAs per SM, so one image's loads / stores overlap the other's labelling.
template <bool FILL>
__global__ void __launch_bounds__(CC_THREADS, 2)
cc_small_kernel(const void* img_all, int H, int W, int32_t* labels_all, int32_t* counts_all, float* scores_all,
                int max_area, float fill_value, int vec) {
  pdl_enter();
  extern __shared__ __align__(16) int cc_smem[];
  const int BH = H >> 1, BW = W >> 1, nb = BH * BW;
  const int chunks = (BW + 31) >> 5, nrows = BH * chunks;
  uint4* planes = reinterpret_cast<uint4*>(cc_smem);              // [nrows] (first: 16-byte aligned for any nb)
  uint32_t* pairs = reinterpret_cast<uint32_t*>(cc_smem + 4 * nrows);
  int* ctl = cc_smem + 4 * nrows + CC_PAIR_CAP;                  // head / pair counters
  int* lab = ctl + 4;
  uint8_t* occ = reinterpret_cast<uint8_t*>(lab + nb);            // occupancy bytes of the generic paths, then the head list
  uint16_t* list = reinterpret_cast<uint16_t*>(occ);              // [nb]
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
  if (vec && !FILL) {
    // 2 x 16 pixels -> 8 blocks -> one byte of each of the chunk-row's four planes, written in place (no occupancy bytes,
    // no ballot pass).  Units beyond the image (the last chunk of a row when BW % 32 != 0) write zeros.
    const uint8_t* p = reinterpret_cast<const uint8_t*>(img);
    const int upr = chunks << 2, units = BH * upr;
    uint8_t* pb = reinterpret_cast<uint8_t*>(planes);
#pragma unroll 2
    for (int u = threadIdx.x; u < units; u += CC_THREADS) {
      const int by = u / upr, k = u - by * upr;
      uint32_t ox = 0u, oy = 0u;
      if (8 * k < BW) {   // W % 16 == 0: a unit is inside the image or completely outside
        const uint4 t = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by) * W + 16 * k);
        const uint4 b = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by + 1) * W + 16 * k);
        ox = occ2_from_u8(t.x, b.x) | (occ2_from_u8(t.y, b.y) << 16);   // one occupancy nibble per byte, blocks 0-3
        oy = occ2_from_u8(t.z, b.z) | (occ2_from_u8(t.w, b.w) << 16);   // blocks 4-7
      }
      uint8_t* dst = pb + (size_t)(by * chunks + (k >> 2)) * 16 + (k & 3);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // bit j of bytes 0..3 -> bits 24..27 (0x01020408 = 2^24 + 2^17 + 2^10 + 2^3: byte i lands on bit 24 + i, no carries)
        const uint32_t lo = (((ox >> j) & 0x01010101u) * 0x01020408u) >> 24;
        const uint32_t hi = (((oy >> j) & 0x01010101u) * 0x01020408u) >> 24;
        dst[4 * j] = (uint8_t)(lo | (hi << 4));
      }
    }
  } else {
    if (vec) {   // FILL: 2 x float4 -> 2 blocks
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
    } else {
      for (int bi = threadIdx.x; bi < nb; bi += CC_THREADS) occ[bi] = (uint8_t)load_occ<FILL>(img, H, W, bi / BW, bi % BW, 0.f);
    }
    __syncthreads();
    cc_planes_from_occ(occ, BH, BW, planes);
  }
  __syncthreads();
  CC_MARK(1);
  cc_label_region(lab, planes, BH, BW, pairs, ctl, list);
  __syncthreads();
  CC_MARK(4);
  // E. outputs
  if (FILL) {
    // one thread per chunk-row: a run whose component is small enough is written pixel by pixel (holes are rare and tiny)
    float* sc = scores_all + img_off;
    const uint32_t cmagic = chunks > 1 ? (uint32_t)((0x100000000ull + chunks - 1) / chunks) : 0u;
    for (int r = threadIdx.x; r < nrows; r += CC_THREADS) {
      const uint4 m = planes[r];
      const uint32_t occm = m.x | m.y | m.z | m.w;
      if (!occm) continue;
      const int by = chunks > 1 ? (int)__umulhi((uint32_t)r, cmagic) : r, ch = r - by * chunks;
      const uint32_t hm = cc_hm(m);
      const int base = by * BW + (ch << 5);
      uint32_t hh = occm & ~hm;
      while (hh) {
        const int s = __ffs(hh) - 1;
        hh &= hh - 1;
        const int rt = lab[base + s] & CC_IDX_MASK;
        if ((int)((uint32_t)lab[rt] >> CC_IDX_BITS) > max_area) continue;
        uint32_t run = cc_run_mask(hm, s);
        while (run) {
          const int i = __ffs(run) - 1;
          run &= run - 1;
          float* px = sc + (size_t)(2 * by) * W + 2 * ((ch << 5) + i);
          if ((m.x >> i) & 1u) px[0] = fill_value;
          if ((m.y >> i) & 1u) px[1] = fill_value;
          if ((m.z >> i) & 1u) px[W] = fill_value;
          if ((m.w >> i) & 1u) px[W + 1] = fill_value;
        }
      }
    }
  } else {
    int32_t* labels = labels_all + img_off;
    int32_t* counts = counts_all + img_off;
    // root block index -> label value needs root / BW: multiply-high by ceil(2^32 / BW) is exact for root < 2^14
    const uint32_t magic = BW > 1 ? (uint32_t)((0x100000000ull + BW - 1) / BW) : 0u;
    if ((W & 3) == 0) {
      // a thread writes its two blocks = 4 pixels of two pixel rows; block -> run head (bit tricks on the chunk-row's planes,
      // one broadcast load) -> root -> area is looked up once per block (the first version did it per pixel row, with a
      // runtime division per task: 48 % of the kernel's warp instructions were this loop)
      const int segs = (BW + 63) >> 6;
      for (int by = warp; by < BH; by += nwarps) {
#pragma unroll 2
        for (int sg = 0; sg < segs; ++sg) {
          const int bx0 = (sg << 6) + (lane << 1);   // W % 4 == 0: both blocks are inside the image or neither
          if (bx0 >= BW) continue;
          const uint4 p = planes[by * chunks + (bx0 >> 5)];
          const int i0 = bx0 & 31;
          const uint32_t hm = cc_hm(p), occm = p.x | p.y | p.z | p.w, heads = occm & ~hm;
          const int rowbase = by * BW + (bx0 & ~31);
          const bool o0 = (occm >> i0) & 1u, o1 = (occm >> (i0 + 1)) & 1u;
          int l0 = 0, l1 = 0, n0 = 0, n1 = 0;
          if (o0) {
            const int rt = lab[rowbase + cc_head_of(heads, i0)] & CC_IDX_MASK;
            const int q = BW > 1 ? (int)__umulhi((uint32_t)rt, magic) : rt;
            l0 = q * 2 * W + (rt - q * BW) * 2 + 1;
            n0 = (int)((uint32_t)lab[rt] >> CC_IDX_BITS);
          }
          if (o1) {
            if (o0 && ((hm >> (i0 + 1)) & 1u)) {   // the same run
              l1 = l0;
              n1 = n0;
            } else {
              const int rt = lab[rowbase + cc_head_of(heads, i0 + 1)] & CC_IDX_MASK;
              const int q = BW > 1 ? (int)__umulhi((uint32_t)rt, magic) : rt;
              l1 = q * 2 * W + (rt - q * BW) * 2 + 1;
              n1 = (int)((uint32_t)lab[rt] >> CC_IDX_BITS);
            }
          }
          const uint32_t x = p.x >> i0, y = p.y >> i0, z = p.z >> i0, w4 = p.w >> i0;
          const size_t off = (size_t)(2 * by) * W + 2 * bx0;
          __stcs(reinterpret_cast<int4*>(labels + off), make_int4((x & 1u) ? l0 : 0, (y & 1u) ? l0 : 0, (x & 2u) ? l1 : 0, (y & 2u) ? l1 : 0));
          __stcs(reinterpret_cast<int4*>(labels + off + W), make_int4((z & 1u) ? l0 : 0, (w4 & 1u) ? l0 : 0, (z & 2u) ? l1 : 0, (w4 & 2u) ? l1 : 0));
          __stcs(reinterpret_cast<int4*>(counts + off), make_int4((x & 1u) ? n0 : 0, (y & 1u) ? n0 : 0, (x & 2u) ? n1 : 0, (y & 2u) ? n1 : 0));
          __stcs(reinterpret_cast<int4*>(counts + off + W), make_int4((z & 1u) ? n0 : 0, (w4 & 1u) ? n0 : 0, (z & 2u) ? n1 : 0, (w4 & 2u) ? n1 : 0));
        }
      }
    } else {
      for (int px = threadIdx.x; px < H * W; px += CC_THREADS) {
        const int r = px / W, c = px % W;
        const int by = r >> 1, bx = c >> 1;
        const uint4 p = planes[by * chunks + (bx >> 5)];
        const uint32_t pl = (r & 1) ? ((c & 1) ? p.w : p.z) : ((c & 1) ? p.y : p.x);
        int l = 0, n = 0;
        if ((pl >> (bx & 31)) & 1u) {
          const int rt = cc_root_of(lab, planes, BW, by, bx);
          const int q = BW > 1 ? (int)__umulhi((uint32_t)rt, magic) : rt;
          l = q * 2 * W + (rt - q * BW) * 2 + 1;
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
    printf("cc_small<%d> cta %d: A %lld  region labeller %lld  E %lld cycles\n", (int)FILL, blockIdx.x, tr[1] - tr[0],
           tr[4] - tr[1], tr[5] - tr[4]);
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("   cta 0 thread 0: I + L (init, %d heads listed) %lld  U1 (hooks) %lld  J (pointer jumping) %lld  U2 (%d deferred pairs) %lld  F (flatten + areas) %lld cycles, %d jumping rounds\n",
           ctl[0], g_cc_tr[0] - tr[1], g_cc_tr[1] - g_cc_tr[0], g_cc_tr[2] - g_cc_tr[1], ctl[1], g_cc_tr[3] - g_cc_tr[2], g_cc_tr[4] - g_cc_tr[3], (int)g_cc_tr[7]);
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
constexpr int TBH = 32, TBW = 64, T_THREADS = 256;   // small tiles
constexpr int BTH = 128, BTW = 128;                   // big tiles (CC_THREADS threads)
constexpr int T_BORDER = 2 * (TBH + TBW);   // open-root list entries per small tile (one per border block at most)

__device__ __forceinline__ int gfind(const volatile uint32_t* f, int n) {
  uint32_t w = f[n];
  while ((int)(w >> 4) != n) {
    n = (int)(w >> 4);
    w = f[n];
  }
  return n;
}
// min-root union on words parent << 4 | occupancy (the nibble of a node never changes).  Roots are linked with
// compare-and-swap (see uf_union: linking with atomicMin lets paths leave their set for a moment, and the compression
// below would then cut links of the other set).
__device__ __forceinline__ void gunion(uint32_t* f, int a, int b) {
  int ra = gfind(f, a), rb = gfind(f, b);
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const uint32_t nib = reinterpret_cast<const volatile uint32_t*>(f)[ra] & 15u;
    const uint32_t old = atomicCAS(f + ra, ((uint32_t)ra << 4) | nib, ((uint32_t)rb << 4) | nib);
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

template <int TH, int TW>
constexpr size_t cc_tile_smem() {   // planes | deferred pairs | counters | words | occupancy bytes, then the head list
  return (size_t)TH * (TW / 32) * 16 + CC_PAIR_CAP * 4 + 16 + (size_t)TH * TW * 4 + (size_t)TH * TW * 2;
}

// Tile = TH x TW blocks (32 x 64 for images smaller than 256 pixels in a dimension, else 128 x 128 = the small path's
// 256 x 256 pixels: the labeller's phases are latency chains of about the same length whatever the tile size, and a tile's
// border work grows with its perimeter).
template <bool FILL, int TH, int TW, int THREADS>
__global__ void __launch_bounds__(THREADS)
cc_t_label(const void* img_all, const float* scores_all, int H, int W, int vec, uint32_t* forest_all, int32_t* area_all,
           int* list_count, int2* list) {
  pdl_enter();
  extern __shared__ __align__(16) int cc_tile_smem_base[];
  constexpr int CHT = TW / 32;
  uint4* planes = reinterpret_cast<uint4*>(cc_tile_smem_base);          // [TH * CHT]
  uint32_t* pairs = reinterpret_cast<uint32_t*>(cc_tile_smem_base + 4 * TH * CHT);
  int* ctl = cc_tile_smem_base + 4 * TH * CHT + CC_PAIR_CAP;
  int* lab = ctl + 4;                                                    // [TH * TW]
  uint8_t* occ = reinterpret_cast<uint8_t*>(lab + TH * TW);             // [TH * TW] bytes, then the head list (2 bytes per block)
  uint16_t* heads = reinterpret_cast<uint16_t*>(occ);
  const int BHg = H >> 1, BWg = W >> 1;
  const int z = blockIdx.z;
  const size_t off = (size_t)z * H * W;
  const void* img = FILL ? static_cast<const void*>(scores_all + off)
                         : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img_all) + off);
  const int by0 = blockIdx.y * TH, bx0 = blockIdx.x * TW;
  // A. occupancy of the tile (blocks outside the image are empty)
  if (vec && !FILL) {   // 2 x 16 pixels -> 8 blocks per thread
    const uint8_t* p = reinterpret_cast<const uint8_t*>(img);
#pragma unroll 2
    for (int u = threadIdx.x; u < TH * (TW / 8); u += THREADS) {
      const int ly = u / (TW / 8), k = u % (TW / 8);
      const int by = by0 + ly, bx = bx0 + 8 * k;
      uint2 o = make_uint2(0u, 0u);
      if (by < BHg && bx < BWg) {   // W % 16 == 0: a unit is inside the image or completely outside
        const uint4 t = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by) * W + 2 * bx);
        const uint4 b = *reinterpret_cast<const uint4*>(p + (size_t)(2 * by + 1) * W + 2 * bx);
        o.x = occ2_from_u8(t.x, b.x) | (occ2_from_u8(t.y, b.y) << 16);
        o.y = occ2_from_u8(t.z, b.z) | (occ2_from_u8(t.w, b.w) << 16);
      }
      *reinterpret_cast<uint2*>(occ + ly * TW + 8 * k) = o;
    }
  } else if (vec && FILL) {   // 2 x float4 -> 2 blocks
    const float* f = reinterpret_cast<const float*>(img);
#pragma unroll 4
    for (int u = threadIdx.x; u < TH * (TW / 2); u += THREADS) {
      const int ly = u / (TW / 2), k = u % (TW / 2);
      const int by = by0 + ly, bx = bx0 + 2 * k;
      uint32_t o = 0;
      if (by < BHg && bx < BWg) {   // W % 4 == 0
        const float4 t = *reinterpret_cast<const float4*>(f + (size_t)(2 * by) * W + 2 * bx);
        const float4 b = *reinterpret_cast<const float4*>(f + (size_t)(2 * by + 1) * W + 2 * bx);
        o = (t.x <= 0.f ? 1u : 0u) | (t.y <= 0.f ? 2u : 0u) | (b.x <= 0.f ? 4u : 0u) | (b.y <= 0.f ? 8u : 0u);
        o |= ((t.z <= 0.f ? 1u : 0u) | (t.w <= 0.f ? 2u : 0u) | (b.z <= 0.f ? 4u : 0u) | (b.w <= 0.f ? 8u : 0u)) << 8;
      }
      *reinterpret_cast<uint16_t*>(occ + ly * TW + 2 * k) = (uint16_t)o;
    }
  } else {
    for (int i = threadIdx.x; i < TH * TW; i += THREADS) {
      const int by = by0 + i / TW, bx = bx0 + i % TW;
      occ[i] = (by < BHg && bx < BWg) ? (uint8_t)load_occ<FILL>(img, H, W, by, bx, 0.f) : (uint8_t)0;
    }
  }
  __syncthreads();
  cc_planes_from_occ(occ, TH, TW, planes);
  __syncthreads();
  cc_label_region(lab, planes, TH, TW, pairs, ctl, heads);   // (the occupancy bytes are gone: their memory is the head list now)
  __syncthreads();
  // roots that reach the tile border: the sign bit of the root's word (area needs 17 bits above the 14 index bits)
  constexpr int NBORDER = 2 * (TH + TW);
  for (int k = threadIdx.x; k < NBORDER; k += THREADS) {
    int ly, lx;
    if (k < TW) { ly = 0; lx = k; }
    else if (k < 2 * TW) { ly = TH - 1; lx = k - TW; }
    else if (k < 2 * TW + TH) { ly = k - 2 * TW; lx = 0; }
    else { ly = k - 2 * TW - TH; lx = TW - 1; }
    const uint4 p = planes[ly * CHT + (lx >> 5)];
    if (((p.x | p.y | p.z | p.w) >> (lx & 31)) & 1u) atomicOr(lab + cc_root_of(lab, planes, TW, ly, lx), (int)0x80000000);
  }
  __syncthreads();
  uint32_t* forest = forest_all + (size_t)z * BHg * BWg;
  int32_t* area = area_all + (size_t)z * BHg * BWg;
  for (int i = threadIdx.x; i < TH * TW; i += THREADS) {
    const int ly = i / TW, lx = i % TW, by = by0 + ly, bx = bx0 + lx;
    if (by >= BHg || bx >= BWg) continue;
    const uint4 p = planes[ly * CHT + (lx >> 5)];
    const int bit = lx & 31;
    const uint32_t o = ((p.x >> bit) & 1u) | (((p.y >> bit) & 1u) << 1) | (((p.z >> bit) & 1u) << 2) | (((p.w >> bit) & 1u) << 3);
    const int gb = by * BWg + bx;
    if (o) {
      const int root = lab[ly * TW + (lx & ~31) + cc_head_of(cc_heads(p), bit)] & CC_IDX_MASK;
      const int groot = (by0 + root / TW) * BWg + bx0 + root % TW;
      forest[gb] = ((uint32_t)groot << 4) | o;
      if (root == i) {
        const int word = lab[i];
        area[gb] = (int)(((uint32_t)word & 0x7fffffffu) >> CC_IDX_BITS);
        if (word < 0) list[atomicAdd(list_count, 1)] = make_int2(z, gb);
      }
    } else {
      forest[gb] = (uint32_t)gb << 4;
    }
  }
}

// One warp per tile edge piece: the 32-block pieces of the top row (lanes = consecutive blocks: coalesced loads,
// neighbours by shuffle, and the same pruning as inside a tile -- a run that crosses the border costs ONE union, not one
// per block), of the left column and of the right column (lanes = rows).
template <int TH, int TW>
__global__ void cc_t_border(int H, int W, uint32_t* forest_all) {
  pdl_enter();
  constexpr int PT = TW / 32, PC = TH / 32, PIECES = PT + 2 * PC;
  const int BH = H >> 1, BW = W >> 1;
  const int tiles_x = (BW + TW - 1) / TW, tiles_y = (BH + TH - 1) / TH;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= tiles_x * tiles_y * PIECES) return;
  const int tile = gw / PIECES, piece = gw - tile * PIECES;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int by0 = ty * TH, bx0 = tx * TW;
  uint32_t* forest = forest_all + (size_t)blockIdx.z * BH * BW;
  auto nib = [&](int y, int x) -> uint32_t {
    return (y >= 0 && y < BH && x >= 0 && x < BW) ? (forest[y * BW + x] & 15u) : 0u;   // a word's nibble never changes
  };
  if (piece < PT) {   // top row, blocks bx0 + 32*piece + lane
    const int by = by0, bx = bx0 + 32 * piece + lane;
    if (by == 0 || by >= BH) return;
    const uint32_t me = nib(by, bx), up = nib(by - 1, bx);
    uint32_t left = __shfl_up_sync(0xffffffffu, me, 1), ul = __shfl_up_sync(0xffffffffu, up, 1);
    uint32_t ur = __shfl_down_sync(0xffffffffu, up, 1);
    if (lane == 0) {
      left = piece > 0 ? nib(by, bx - 1) : 0u;   // the block before the tile's first one belongs to another tile
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
  } else if (piece < PT + PC) {   // left column, rows by0 + 32*(piece - PT) + lane
    const int ly = 32 * (piece - PT) + lane, by = by0 + ly, bx = bx0;
    if (bx == 0 || by >= BH) return;
    const uint32_t me = nib(by, bx);
    if (!me) return;
    const uint32_t lf = nib(by, bx - 1);
    const int idx = by * BW + bx;
    if (conn_left(me, lf)) gunion(forest, idx, idx - 1);
    if (ly > 0) {            // the tile's first row is the top-row piece's business
      const uint32_t up = nib(by - 1, bx), ul = nib(by - 1, bx - 1);
      if (conn_upleft(me, ul) && !(conn_up(me, up) && conn_left(up, ul))) gunion(forest, idx, idx - BW - 1);
    }
  } else {   // right column
    const int ly = 32 * (piece - PT - PC) + lane, by = by0 + ly, bx = bx0 + TW - 1;
    if (ly == 0 || bx + 1 >= BW || by >= BH) return;
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

size_t chunk_rows(int h, int w) { return (size_t)(h / 2) * ((w / 2 + 31) / 32); }
bool small_ok(int h, int w) { return (h / 2) * (w / 2) <= CC_MAX_BLOCKS && chunk_rows(h, w) <= CC_MAX_CHUNK_ROWS; }
size_t small_smem(int h, int w) {
  const size_t nb = (size_t)(h / 2) * (w / 2);
  // planes | deferred pairs | counters | words | occupancy bytes, re-used as the list of run heads (16 bits per block)
  return chunk_rows(h, w) * 16 + CC_PAIR_CAP * 4 + 16 + nb * 4 + ((nb * 2 + 15) & ~(size_t)15) + 16;
}
constexpr size_t CC_SMALL_SMEM_MAX = (size_t)CC_MAX_CHUNK_ROWS * 16 + CC_PAIR_CAP * 4 + 32 + (size_t)CC_MAX_BLOCKS * 6 + 16;

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
      VLS_CUDA(cudaFuncSetAttribute(cc_small_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC_SMALL_SMEM_MAX));
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
  // 64 x 128-pixel tiles.  256 x 256-pixel ones (the kernels are templates on the tile size) were measured on 64 x 1024^2:
  // border unions 39 -> 24 us per half, but labelling 79 -> 96 us (512 CTAs per half on 296 slots = 1.7 waves of a CTA that
  // takes 90 k cycles; noise input 630 vs 400 us in total), so they stay off.
  const bool big = false;
  const int th = big ? BTH : TBH, tw = big ? BTW : TBW;
  const int tiles_x = (BW + tw - 1) / tw, tiles_y = (BH + th - 1) / th;
  uint32_t* forest = reinterpret_cast<uint32_t*>(ws);
  int32_t* area = reinterpret_cast<int32_t*>(forest + nblk1 * n);
  int* list_count = reinterpret_cast<int*>(area + nblk1 * n);   // one counter per slice, 16 bytes reserved
  int2* list = reinterpret_cast<int2*>(list_count + 4);
  const size_t list_per_image = (size_t)tiles_x * tiles_y * 2 * (th + tw);   // one entry per border block of a tile at most
  VLS_CUDA(cudaMemsetAsync(list_count, 0, 16, stream));
  const int vec = FILL ? ((w % 4) == 0 && ((uintptr_t)scores % 16) == 0) : ((w % 16) == 0 && ((uintptr_t)img % 16) == 0);
  static unsigned long long attr[2] = {0, 0};
  if (first_use_on_device(&attr[FILL])) {
    VLS_CUDA(cudaFuncSetAttribute(cc_t_label<FILL, BTH, BTW, CC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)cc_tile_smem<BTH, BTW>()));
  }
  // The labelling kernel is latency / issue bound and the final pass store bound, so a large batch is cut into slices
  // of work (two from 8 images on); odd slices run on a forked stream, so one slice's output pass
  // (store bound) overlaps the next slice's labelling (latency / issue bound).
  const int slices = n >= 8 ? 2 : 1;   // (4 slices measured slower: 366 vs 330 us for 64 x 1024^2; the smaller grids lose more than the overlap gains)
  cudaStream_t side = stream;
  if (slices > 1) VLS_TRY(fork_begin(3, stream, &side));   // the fork point precedes all of this call's work
  for (int sl = 0; sl < slices; ++sl) {
    const int i0 = (int)((long long)n * sl / slices), cnt = (int)((long long)n * (sl + 1) / slices) - i0;
    cudaStream_t st = (sl & 1) ? side : stream;
    const size_t px0 = (size_t)i0 * h * w;
    const void* img_h = FILL ? nullptr : static_cast<const void*>(reinterpret_cast<const uint8_t*>(img) + px0);
    float* sc_h = FILL ? scores + px0 : nullptr;
    uint32_t* f_h = forest + nblk1 * i0;
    int32_t* a_h = area + nblk1 * i0;
    int2* l_h = list + list_per_image * i0;
    const long long list_cap = (long long)cnt * list_per_image;
    const long long border = (long long)tiles_x * tiles_y * (tw / 32 + 2 * (th / 32)) * 32;   // one warp per edge piece
    if (big) {
      VLS_CUDA(launch_k(cc_t_label<FILL, BTH, BTW, CC_THREADS>, dim3(tiles_x, tiles_y, cnt), dim3(CC_THREADS), cc_tile_smem<BTH, BTW>(),
                        st, img_h, sc_h, h, w, vec, f_h, a_h, list_count + sl, l_h));
      VLS_CUDA(launch_k(cc_t_border<BTH, BTW>, dim3((unsigned)((border + 255) / 256), 1, cnt), dim3(256), 0, st, h, w, f_h));
    } else {
      VLS_CUDA(launch_k(cc_t_label<FILL, TBH, TBW, T_THREADS>, dim3(tiles_x, tiles_y, cnt), dim3(T_THREADS), cc_tile_smem<TBH, TBW>(),
                        st, img_h, sc_h, h, w, vec, f_h, a_h, list_count + sl, l_h));
      VLS_CUDA(launch_k(cc_t_border<TBH, TBW>, dim3((unsigned)((border + 255) / 256), 1, cnt), dim3(256), 0, st, h, w, f_h));
    }
    VLS_CUDA(launch_k(cc_t_areas, dim3((unsigned)((list_cap + 255) / 256)), dim3(256), 0, st, h, w, f_h, a_h, list_count + sl, l_h));
    dim3 blk(128, 1, 1), grd((BW + 127) / 128, BH, cnt);
    VLS_CUDA(launch_k(cc_t_final<FILL>, grd, blk, 0, st, h, w, f_h, a_h, FILL ? nullptr : labels + px0, FILL ? nullptr : counts + px0,
                      sc_h, max_area, fill_value));
    VLS_POST_LAUNCH(4);
  }
  if (slices > 1) VLS_TRY(fork_join(3, stream));
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
