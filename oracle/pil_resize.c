/*
 * CPU oracle (TEST INFRASTRUCTURE ONLY) for the image pre-processing of the reference's input pipeline:
 *
 *     transforms.Resize((256, 256))  ->  transforms.ToTensor()  ->  transforms.Normalize(mean, std)
 *     (train.py:92-102, inference.py:28-31; applied per image by data/dataset.py:21-27)
 *
 * On a PIL image torchvision's Resize is `img.resize((w, h), Image.BILINEAR)`, so the arithmetic lives in the
 * third-party dependency Pillow (requirements.txt:3 "pillow", unpinned; 12.2.0 in this image), which is not vendored
 * in the reference.  This file restates Pillow's published two-pass 8-bit resampling algorithm for the bilinear
 * (triangle) filter: per-axis coefficient windows computed in double precision, normalised, converted to 22-bit
 * fixed point; horizontal pass to a uint8 intermediate, then vertical pass, each with round-half-up and clipping.
 * Pinned by tests/test_oracle_resize.py against Pillow itself (bit-exact) on random images and sizes.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file; the product never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

static uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  if (v < 0) return 0;
  if (v > 255) return 255;
  return (uint8_t)v;
}

/* Coefficient windows of one axis: bounds[2*i] = first input index, bounds[2*i+1] = window length, kk[i*ksize + x]
 * fixed-point weights.  Returns ksize; *bounds_out / *kk_out are malloc'ed. */
int fnst_oracle_resize_coeffs(int in_size, int out_size, int** bounds_out, int** kk_out) {
  const double support0 = 1.0; /* bilinear */
  double scale, filterscale, support;
  filterscale = scale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  support = support0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  int* bounds = (int*)malloc(sizeof(int) * 2 * out_size);
  int* kk = (int*)malloc(sizeof(int) * (size_t)out_size * ksize);
  double* k = (double*)malloc(sizeof(double) * ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = bilinear_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    for (x = 0; x < ksize; ++x) {
      if (k[x] < 0) kk[xx * ksize + x] = (int)(-0.5 + k[x] * (1 << PRECISION_BITS));
      else kk[xx * ksize + x] = (int)(0.5 + k[x] * (1 << PRECISION_BITS));
    }
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  free(k);
  *bounds_out = bounds;
  *kk_out = kk;
  return ksize;
}

/* in: HWC uint8 RGB (in_pitch bytes per row); out: HWC uint8 RGB, contiguous. */
void fnst_oracle_resize_bilinear_u8(const uint8_t* in, int in_h, int in_w, int64_t in_pitch, uint8_t* out, int out_h, int out_w) {
  const int need_h = out_w != in_w, need_v = out_h != in_h;
  const uint8_t* src = in;
  int64_t src_pitch = in_pitch;
  uint8_t* tmp = NULL;
  int *bh = NULL, *kh = NULL, *bv = NULL, *kv = NULL;
  int ksh = 0, ksv = 0;
  if (need_h) ksh = fnst_oracle_resize_coeffs(in_w, out_w, &bh, &kh);
  if (need_v) ksv = fnst_oracle_resize_coeffs(in_h, out_h, &bv, &kv);
  if (need_h) {
    /* Pillow resamples only the rows the vertical pass will read; the values of those rows are the same either way */
    uint8_t* dst = need_v ? (tmp = (uint8_t*)malloc((size_t)in_h * out_w * 3)) : out;
    for (int y = 0; y < in_h; ++y)
      for (int xx = 0; xx < out_w; ++xx) {
        const int xmin = bh[xx * 2], xmax = bh[xx * 2 + 1];
        const int* k = &kh[xx * ksh];
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < xmax; ++x) {
          const uint8_t* p = src + y * src_pitch + (int64_t)(x + xmin) * 3;
          s0 += p[0] * k[x]; s1 += p[1] * k[x]; s2 += p[2] * k[x];
        }
        uint8_t* q = dst + ((int64_t)y * out_w + xx) * 3;
        q[0] = clip8(s0); q[1] = clip8(s1); q[2] = clip8(s2);
      }
    src = dst;
    src_pitch = (int64_t)out_w * 3;
  }
  if (need_v) {
    for (int yy = 0; yy < out_h; ++yy) {
      const int ymin = bv[yy * 2], ymax = bv[yy * 2 + 1];
      const int* k = &kv[yy * ksv];
      for (int xx = 0; xx < out_w; ++xx) {
        int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < ymax; ++y) {
          const uint8_t* p = src + (y + ymin) * src_pitch + (int64_t)xx * 3;
          s0 += p[0] * k[y]; s1 += p[1] * k[y]; s2 += p[2] * k[y];
        }
        uint8_t* q = out + ((int64_t)yy * out_w + xx) * 3;
        q[0] = clip8(s0); q[1] = clip8(s1); q[2] = clip8(s2);
      }
    }
  } else if (!need_h) {
    for (int y = 0; y < in_h; ++y) memcpy(out + (int64_t)y * out_w * 3, in + y * in_pitch, (size_t)out_w * 3);
  }
  free(tmp); free(bh); free(kh); free(bv); free(kv);
}

/* ToTensor + Normalize (train.py:92-102): out[c][y][x] = ((u8 / 255.0f) - mean[c]) / std[c], float32 arithmetic in that order. */
void fnst_oracle_to_tensor(const uint8_t* hwc, int h, int w, const float* mean3, const float* std3, float* out_chw) {
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < h * w; ++i) {
      float v = (float)hwc[i * 3 + c] / 255.0f;
      if (mean3 && std3) v = (v - mean3[c]) / std3[c];
      out_chw[(size_t)c * h * w + i] = v;
    }
}
