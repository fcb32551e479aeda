/* TEST INFRASTRUCTURE ONLY -- plain-C sequential restatement of the reference's connected
 * components labelling, sam2/csrc/connected_components.cu (paths relative to /root/reference).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 *
 * It walks the same six stages as the reference host loop (:245-275), one image at a time,
 * with the same 2x2-block union-find:
 *   init_labeling (:62-70)   label[idx] = idx for every block's top-left pixel
 *   merge (:72-118)          neighbourhood bitmask P, union with up-left/up/up-right/left blocks
 *   compression (:120-127)   full path compression
 *   final_labeling (:129-168) label = root+1 on foreground pixels of the block, 0 elsewhere
 *   init_counting (:170-187) count_init[label-1] += 1 per foreground pixel
 *   final_counting (:189-209) count_final[idx] = count_init[label-1]
 * The GPU kernel's atomicMin union (:42-60) always links the larger root under the smaller,
 * so the final root of a component is its minimum block index regardless of thread order;
 * a sequential min-root union reproduces it exactly.
 *
 * Pinning: there are no golden vectors for this path in the reference.
 * This file is cross-checked against a scipy closed form (oracle/sam2_path.py:cc_label) on CPU
 * and, on the GPU box, against the reference .cu itself compiled into oracle/_ref/ (when built).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int32_t find_root(const int32_t *s, int32_t n) {
  while (s[n] != n) n = s[n];
  return n;
}

static void unite(int32_t *s, int32_t a, int32_t b) {
  a = find_root(s, a);
  b = find_root(s, b);
  if (a < b) s[b] = a;
  else if (b < a) s[a] = b;
}

/* img: [n,h,w] uint8 (non-zero = foreground); labels, counts: [n,h,w] int32 (outputs).
 * Returns 0, or 1 if h or w is odd (the reference asserts evenness, :226-227). */
int cc_oracle_label(const uint8_t *img_all, int n, int h, int w, int32_t *labels_all, int32_t *counts_all) {
  if ((h & 1) || (w & 1) || n < 0) return 1;
  const int64_t hw = (int64_t)h * w;
  int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(hw > 0 ? hw : 1));
  if (!cnt) return 2;
  for (int i = 0; i < n; ++i) {
    const uint8_t *img = img_all + (int64_t)i * hw;
    int32_t *label = labels_all + (int64_t)i * hw;
    int32_t *count = counts_all + (int64_t)i * hw;
    memset(label, 0, sizeof(int32_t) * (size_t)hw);
    memset(cnt, 0, sizeof(int32_t) * (size_t)hw);
    for (int r = 0; r < h; r += 2)
      for (int c = 0; c < w; c += 2) label[r * w + c] = r * w + c;
    for (int r = 0; r < h; r += 2)
      for (int c = 0; c < w; c += 2) {
        const int idx = r * w + c;
        uint32_t P = 0;
        if (img[idx]) P |= 0x777;
        if (r + 1 < h && img[idx + w]) P |= 0x777 << 4;
        if (c + 1 < w && img[idx + 1]) P |= 0x777 << 1;
        if (c == 0) P &= 0xEEEE;
        if (c + 1 >= w) P &= 0x3333;
        else if (c + 2 >= w) P &= 0x7777;
        if (r == 0) P &= 0xFFF0;
        if (r + 1 >= h) P &= 0xFF;
        if (P > 0) {
          if (((P >> 0) & 1) && img[idx - w - 1]) unite(label, idx, idx - 2 * w - 2);
          if ((((P >> 1) & 1) && img[idx - w]) || (((P >> 2) & 1) && img[idx - w + 1]))
            unite(label, idx, idx - 2 * w);
          if (((P >> 3) & 1) && img[idx + 2 - w]) unite(label, idx, idx - 2 * w + 2);
          if ((((P >> 4) & 1) && img[idx - 1]) || (((P >> 8) & 1) && img[idx + w - 1]))
            unite(label, idx, idx - 2);
        }
      }
    for (int r = 0; r < h; r += 2)
      for (int c = 0; c < w; c += 2) {
        const int idx = r * w + c;
        label[idx] = find_root(label, idx);
      }
    for (int r = 0; r < h; r += 2)
      for (int c = 0; c < w; c += 2) {
        const int idx = r * w + c;
        const int32_t y = label[idx] + 1;
        label[idx] = img[idx] ? y : 0;
        label[idx + 1] = img[idx + 1] ? y : 0;
        label[idx + w] = img[idx + w] ? y : 0;
        label[idx + w + 1] = img[idx + w + 1] ? y : 0;
      }
    /* NOTE: final_labeling overwrites label[] of a block's top-left pixel before later blocks
     * are processed; in the reference that is a separate launch after compression, and every
     * block reads only its own top-left entry (:140), so sequential order is equivalent. */
    for (int64_t p = 0; p < hw; ++p)
      if (label[p] > 0) cnt[label[p] - 1] += 1;
    for (int64_t p = 0; p < hw; ++p) count[p] = label[p] > 0 ? cnt[label[p] - 1] : 0;
  }
  free(cnt);
  return 0;
}
