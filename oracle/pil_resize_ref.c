/* Plain-C restatement of Pillow's 8-bit antialiased Image.resize (ImagingResample),
 * the arithmetic behind the filter's preprocessing transforms:
 *   all_utils/dataset_utils.py:78-85   Resize((256,256)) [PIL bilinear] -> CenterCrop(224)
 *   openai-clip clip/clip.py _transform  Resize(224, bicubic) -> CenterCrop(224)   [third-party]
 * called at all_utils/utils.py:360 and :404 (via get_semantic_filtering :169-177).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity PINNED: bit-for-bit against
 * PIL.Image.resize (Pillow is installed here; tests/test_resize_oracle.py) and against
 * the committed golden vectors.
 *
 * Published algorithm (Pillow src/libImaging/Resample.c; third-party, unpinned transitive
 * dependency, 12.2 installed): separable, horizontal pass then vertical pass, uint8
 * intermediate; per output index: centre = (i+0.5)*scale, support = base*max(scale,1),
 * window [ (int)(centre-support+.5), (int)(centre+support+.5) ) clipped to the axis;
 * weights filter((x-centre+.5)/max(scale,1)) normalised in double, quantised to
 * (int)(+-0.5 + w * 2^22); accumulate int32 from 2^21, >> 22, clamp to [0,255].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

static double bicubic_filter(double x) {
#define a -0.5
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
#undef a
}

/* Returns ksize; *bounds_out = [out][2] (xmin, count), *kk_out = [out][ksize] int32 coefficients. */
int oracle_pil_coeffs(int in_size, int out_size, int filt, int **bounds_out, int32_t **kk_out) {
  double (*filter)(double) = filt ? bicubic_filter : bilinear_filter;
  double fsupport = filt ? 2.0 : 1.0;
  double scale, filterscale, support;
  filterscale = scale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  support = fsupport * filterscale;
  int ksize = (int)ceil(support) * 2 + 1;
  double *prekk = (double *)malloc(sizeof(double) * out_size * ksize);
  int *bounds = (int *)malloc(sizeof(int) * out_size * 2);
  int32_t *kk = (int32_t *)malloc(sizeof(int32_t) * out_size * ksize);
  if (!prekk || !bounds || !kk) { free(prekk); free(bounds); free(kk); return -1; }
  for (int xx = 0; xx < out_size; xx++) {
    double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double *k = &prekk[xx * ksize];
    int x;
    for (x = 0; x < xmax; x++) {
      double w = filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; x++) {
      if (ww != 0.0) k[x] /= ww;
    }
    for (; x < ksize; x++) k[x] = 0;
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  for (int i = 0; i < out_size * ksize; i++) {
    if (prekk[i] < 0) kk[i] = (int32_t)(-0.5 + prekk[i] * (1 << PRECISION_BITS));
    else kk[i] = (int32_t)(0.5 + prekk[i] * (1 << PRECISION_BITS));
  }
  free(prekk);
  *bounds_out = bounds;
  *kk_out = kk;
  return ksize;
}

static inline uint8_t clip8(int32_t v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* img [h][w][c] u8 -> out [oh][ow][c] u8.  filt: 0 bilinear, 1 bicubic. */
int oracle_pil_resize_u8(const uint8_t *img, int h, int w, int c, uint8_t *out, int oh, int ow, int filt) {
  int *bx = 0, *by = 0;
  int32_t *kx = 0, *ky = 0;
  int ksx = oracle_pil_coeffs(w, ow, filt, &bx, &kx);
  int ksy = oracle_pil_coeffs(h, oh, filt, &by, &ky);
  uint8_t *tmp = (uint8_t *)malloc((size_t)h * ow * c);
  if (ksx < 0 || ksy < 0 || !tmp) { free(bx); free(by); free(kx); free(ky); free(tmp); return -1; }
  /* Pillow skips a pass whose size is unchanged; the pass below is then an exact identity
   * anyway (single coefficient 2^22), so results agree. */
  for (int y = 0; y < h; y++)
    for (int xx = 0; xx < ow; xx++) {
      int xmin = bx[xx * 2], cnt = bx[xx * 2 + 1];
      const int32_t *k = &kx[xx * ksx];
      for (int ch = 0; ch < c; ch++) {
        int32_t ss = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < cnt; x++) ss += (int32_t)img[((size_t)y * w + x + xmin) * c + ch] * k[x];
        tmp[((size_t)y * ow + xx) * c + ch] = clip8(ss);
      }
    }
  for (int yy = 0; yy < oh; yy++) {
    int ymin = by[yy * 2], cnt = by[yy * 2 + 1];
    const int32_t *k = &ky[yy * ksy];
    for (int xx = 0; xx < ow; xx++)
      for (int ch = 0; ch < c; ch++) {
        int32_t ss = 1 << (PRECISION_BITS - 1);
        for (int y = 0; y < cnt; y++) ss += (int32_t)tmp[((size_t)(y + ymin) * ow + xx) * c + ch] * k[y];
        out[((size_t)yy * ow + xx) * c + ch] = clip8(ss);
      }
  }
  free(bx); free(by); free(kx); free(ky); free(tmp);
  return 0;
}
