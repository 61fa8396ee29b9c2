// Stand-alone GPU parity check of fnst_resize_to_tensor (no Python, starts in a second): the device kernel against the
// CPU oracle oracle/pil_resize.c (itself pinned bit-exact to Pillow by tests/test_oracle_resize.py) -- uint8 output must
// be identical, float output must be bit-identical to the oracle's ToTensor / Normalize arithmetic.
// Build + run: make -C tests/native && tests/native/bin/resize_selftest
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/fnst.h"

extern "C" void fnst_oracle_resize_bilinear_u8(const uint8_t* in, int in_h, int in_w, int64_t in_pitch, uint8_t* out, int out_h, int out_w);
extern "C" void fnst_oracle_to_tensor(const uint8_t* hwc, int h, int w, const float* mean3, const float* std3, float* out_chw);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main() {
  const int cases[][4] = {{444, 444, 256, 256}, {609, 800, 256, 256}, {650, 650, 256, 256}, {1080, 1920, 256, 256}, {256, 256, 256, 256},
                          {256, 300, 256, 256}, {300, 256, 256, 256}, {100, 120, 256, 256}, {17, 23, 256, 256}, {3000, 4000, 256, 256},
                          {8000, 31, 256, 256}, {1, 1, 256, 256}, {500, 333, 224, 320}, {64, 64, 1080, 1920}, {777, 1234, 40, 48}};
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  int failures = 0;
  for (int staged = 0; staged <= 1; ++staged) {
  fnst_set_tuning("resize_staged", staged);
  printf("---- resize_staged = %d ----\n", staged);
  uint32_t seed = 12345u;
  for (const auto& c : cases) {
    const int ih = c[0], iw = c[1], oh = c[2], ow = c[3];
    const int64_t pitch = (int64_t)iw * 3 + 5;                        // padded rows: exercises in_pitch_bytes
    std::vector<uint8_t> img((size_t)ih * pitch);
    for (auto& b : img) { seed = seed * 1664525u + 1013904223u; b = (uint8_t)(seed >> 24); }
    std::vector<uint8_t> ref_u8((size_t)oh * ow * 3), got_u8(ref_u8.size());
    std::vector<float> ref_f((size_t)oh * ow * 3), got_f(ref_f.size()), ref_f0(ref_f.size()), got_f0(ref_f.size());
    fnst_oracle_resize_bilinear_u8(img.data(), ih, iw, pitch, ref_u8.data(), oh, ow);
    fnst_oracle_to_tensor(ref_u8.data(), oh, ow, mean, stdv, ref_f.data());
    fnst_oracle_to_tensor(ref_u8.data(), oh, ow, nullptr, nullptr, ref_f0.data());
    uint8_t *d_img, *d_u8; float *d_f, *d_f0;
    CK(cudaMalloc(&d_img, img.size())); CK(cudaMalloc(&d_u8, got_u8.size())); CK(cudaMalloc(&d_f, got_f.size() * 4)); CK(cudaMalloc(&d_f0, got_f.size() * 4));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    int rc = fnst_resize_to_tensor(d_img, ih, iw, pitch, oh, ow, d_f, d_u8, mean, stdv, 0, nullptr);
    if (rc == 0) rc = fnst_resize_to_tensor(d_img, ih, iw, pitch, oh, ow, d_f0, nullptr, nullptr, nullptr, 0, nullptr);
    if (rc != 0) { printf("%dx%d -> %dx%d: rc=%d %s\n", ih, iw, oh, ow, rc, fnst_last_error()); ++failures; continue; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got_u8.data(), d_u8, got_u8.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got_f.data(), d_f, got_f.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got_f0.data(), d_f0, got_f0.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad_u8 = 0, bad_f = 0, bad_f0 = 0;
    for (size_t i = 0; i < ref_u8.size(); ++i) bad_u8 += ref_u8[i] != got_u8[i];
    bad_f = memcmp(ref_f.data(), got_f.data(), ref_f.size() * 4) != 0;
    bad_f0 = memcmp(ref_f0.data(), got_f0.data(), ref_f0.size() * 4) != 0;
    printf("%5dx%-5d -> %4dx%-4d  u8 mismatches %zu  float(normalised) %s  float(plain) %s\n", ih, iw, oh, ow, bad_u8,
           bad_f ? "DIFFERS" : "bit-exact", bad_f0 ? "DIFFERS" : "bit-exact");
    failures += (bad_u8 != 0) + bad_f + bad_f0;
    cudaFree(d_img); cudaFree(d_u8); cudaFree(d_f); cudaFree(d_f0);
  }
  }
  // timing of the training-pipeline case (decoded 1080p frame -> normalised 256x256 tensor), device time per image
  {
    const int ih = 1080, iw = 1920, n = 64;
    uint8_t* d_img; float* d_f;
    CK(cudaMalloc(&d_img, (size_t)n * ih * iw * 3)); CK(cudaMalloc(&d_f, (size_t)n * 3 * 256 * 256 * 4));
    CK(cudaMemset(d_img, 77, (size_t)n * ih * iw * 3));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int staged = 0; staged <= 1; ++staged) {
    fnst_set_tuning("resize_staged", staged);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      for (int i = 0; i < n; ++i)
        fnst_resize_to_tensor(d_img + (size_t)i * ih * iw * 3, ih, iw, (int64_t)iw * 3, 256, 256, d_f + (size_t)i * 3 * 256 * 256, nullptr, mean, stdv, 0, nullptr);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)ih * iw * 3 + 3.0 * 256 * 256 * 4;
    printf("timing (resize_staged = %d): 1080x1920 -> 256x256 normalised tensor: %.2f us/image, %.0f GB/s of algorithmic bytes (%d images, 398 MB > L2)\n",
           staged, 1e3 * ms / n, bytes * n / (ms * 1e-3) / 1e9, n);
    }
  }
  printf(failures ? "FAILED (%d)\n" : "ALL OK\n", failures);
  return failures ? 1 : 0;
}
