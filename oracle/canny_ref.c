/* Plain-C restatement of cv2.Canny(img_u8_hwc, low, high) (aperture 3, L1 gradient)
 * as the reference calls it at all_utils/utils.py:81-85 (via generate_canny,
 * all_utils/utils.py:87-109; run_aug/run_aug.py:436-437).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA path and
 * the "port" CPU baseline in bench.py.  Parity PINNED: tests/test_canny_oracle.py checks
 * it bit-for-bit against the committed golden vectors, which were produced by the
 * reference's own generate_canny (tests/golden/make_canny_golden.py).
 *
 * Algorithm = published OpenCV 4.x modules/imgproc/src/canny.cpp (third-party,
 * opencv-python 4.8.0.74 pinned in environment.yml:24); see oracle/canny_np.py header.
 * Scalar, single-threaded: cpu_baseline.cores == 1 per call (callers may fan out images).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TG22 13573

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* img: [h][w][c] u8; out: [h][w] u8 (0/255).  Returns 0, or -1 on allocation failure. */
int oracle_canny_u8(const uint8_t *img, int h, int w, int c, int low, int high, uint8_t *out) {
  if (low > high) { int t = low; low = high; high = t; }
  size_t n = (size_t)h * w;
  if (n == 0) return 0;
  int16_t *dx = (int16_t *)malloc(n * sizeof(int16_t));
  int16_t *dy = (int16_t *)malloc(n * sizeof(int16_t));
  int32_t *mag = (int32_t *)calloc((size_t)(h + 2) * (w + 2), sizeof(int32_t)); /* zero border */
  uint8_t *lab = (uint8_t *)malloc(n);
  int32_t *stack = (int32_t *)malloc(n * sizeof(int32_t));
  if (!dx || !dy || !mag || !lab || !stack) { free(dx); free(dy); free(mag); free(lab); free(stack); return -1; }
  const int ms = w + 2;
  /* 1-2: Sobel with replicated border, winning channel = first with the largest |dx|+|dy| */
  for (int y = 0; y < h; ++y) {
    int y0 = clampi(y - 1, 0, h - 1), y2 = clampi(y + 1, 0, h - 1);
    for (int x = 0; x < w; ++x) {
      int x0 = clampi(x - 1, 0, w - 1), x2 = clampi(x + 1, 0, w - 1);
      int best = -1, bdx = 0, bdy = 0;
      for (int k = 0; k < c; ++k) {
#define P(yy, xx) ((int)img[((size_t)(yy) * w + (xx)) * c + k])
        int gx = (P(y0, x2) + 2 * P(y, x2) + P(y2, x2)) - (P(y0, x0) + 2 * P(y, x0) + P(y2, x0));
        int gy = (P(y2, x0) + 2 * P(y2, x) + P(y2, x2)) - (P(y0, x0) + 2 * P(y0, x) + P(y0, x2));
#undef P
        int m = abs(gx) + abs(gy);
        if (m > best) { best = m; bdx = gx; bdy = gy; }
      }
      dx[(size_t)y * w + x] = (int16_t)bdx;
      dy[(size_t)y * w + x] = (int16_t)bdy;
      mag[(size_t)(y + 1) * ms + (x + 1)] = best;
    }
  }
  /* 3-4: non-maximum suppression + double threshold */
  int sp = 0;
  for (int y = 0; y < h; ++y) {
    const int32_t *mc = mag + (size_t)(y + 1) * ms + 1;
    const int32_t *mu = mc - ms, *md = mc + ms;
    for (int x = 0; x < w; ++x) {
      int m = mc[x];
      uint8_t l = 0;
      if (m > low) {
        int xs = dx[(size_t)y * w + x], ys = dy[(size_t)y * w + x];
        int64_t ax = abs(xs), ay = (int64_t)abs(ys) << 15;
        int64_t tg22x = ax * TG22;
        int keep;
        if (ay < tg22x) {
          keep = (m > mc[x - 1]) && (m >= mc[x + 1]);
        } else {
          int64_t tg67x = tg22x + (ax << 16);
          if (ay > tg67x) {
            keep = (m > mu[x]) && (m >= md[x]);
          } else {
            int s = ((xs ^ ys) < 0) ? -1 : 1;
            keep = (m > mu[x - s]) && (m > md[x + s]);
          }
        }
        if (keep) l = (m > high) ? 2 : 1;
      }
      lab[(size_t)y * w + x] = l;
      if (l == 2) stack[sp++] = (int32_t)((size_t)y * w + x);
    }
  }
  /* 5: 8-connected hysteresis */
  while (sp > 0) {
    int32_t p = stack[--sp];
    int py = p / w, px = p % w;
    for (int oy = -1; oy <= 1; ++oy) {
      int yy = py + oy;
      if (yy < 0 || yy >= h) continue;
      for (int ox = -1; ox <= 1; ++ox) {
        int xx = px + ox;
        if (xx < 0 || xx >= w) continue;
        size_t q = (size_t)yy * w + xx;
        if (lab[q] == 1) { lab[q] = 2; stack[sp++] = (int32_t)q; }
      }
    }
  }
  /* 6 */
  for (size_t i = 0; i < n; ++i) out[i] = (lab[i] == 2) ? 255 : 0;
  free(dx); free(dy); free(mag); free(lab); free(stack);
  return 0;
}

/* Batched helper: imgs [n][h][w][c] -> out [n][h][w]. */
int oracle_canny_u8_batch(const uint8_t *imgs, int n, int h, int w, int c, int low, int high, uint8_t *out) {
  for (int i = 0; i < n; ++i) {
    int rc = oracle_canny_u8(imgs + (size_t)i * h * w * c, h, w, c, low, high, out + (size_t)i * h * w);
    if (rc) return rc;
  }
  return 0;
}
