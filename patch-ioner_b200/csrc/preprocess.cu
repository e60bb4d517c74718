// preprocess.cu -- image preprocessing of the reference on the device:
//   T.Resize(resize_dim, BICUBIC) [-> T.CenterCrop(crop_dim)] -> T.ToTensor() -> T.Normalize(mean, std)     (src/model.py:347-357)
// torchvision resizes PIL images with Pillow's ImagingResample (8 bits per channel): two separable passes (horizontal, then
// vertical), antialiased bicubic (a = -0.5, support 2 x max(scale, 1)), coefficients normalised per output pixel and converted
// to 22-bit fixed point, every pass rounded and clipped to uint8.  The host builds the coefficient tables exactly as Pillow
// does (patch-ioner_b200/preprocess.py); these kernels do the integer arithmetic, so the resized bytes are bit-identical to
// Pillow's, and the float tail (u / 255, then (v - mean) / std in fp32) is the same two roundings as torchvision's.
// HBM-bound byte work: a thread per output pixel, 3 channels, taps from a table; rows outside the crop are never computed.
#include "common.cuh"

namespace pio {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= PRECISION_BITS;  // arithmetic shift, like Pillow's lookup table index
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[b][y - row_first][xx - crop_left][c] = clip8(2^21 + sum_x img[b][y][xmin + x][c] * kx[xx][x])
__global__ void __launch_bounds__(256) resample_h_kernel(const unsigned char* __restrict__ img, int H, int W,
                                                         const int* __restrict__ kx, const int* __restrict__ bx, int ksize,
                                                         int crop_left, int cw, int row_first, int rows,
                                                         unsigned char* __restrict__ tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over rows * cw
  const int b = blockIdx.y;
  if (i >= rows * cw) return;
  const int y = row_first + i / cw, xo = i % cw, xx = crop_left + xo;
  const int xmin = bx[2 * xx], n = bx[2 * xx + 1];
  const int* k = kx + (long long)xx * ksize;
  const unsigned char* p = img + ((long long)b * H + y) * W * 3 + (long long)xmin * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int kv = __ldg(k + x);
    s0 += p[3 * x] * kv;
    s1 += p[3 * x + 1] * kv;
    s2 += p[3 * x + 2] * kv;
  }
  unsigned char* o = tmp + (((long long)b * rows + (y - row_first)) * cw + xo) * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// out[b][c][yo][xo] = (clip8(2^21 + sum_y tmp[..]) / 255 - mean[c]) / std[c]
__global__ void __launch_bounds__(256) resample_v_normalize_kernel(const unsigned char* __restrict__ tmp, int rows, int cw,
                                                                   const int* __restrict__ ky, const int* __restrict__ by, int ksize,
                                                                   int crop_top, int ch, int row_first, float m0, float m1, float m2,
                                                                   float d0, float d1, float d2, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over ch * cw
  const int b = blockIdx.y;
  if (i >= ch * cw) return;
  const int yo = i / cw, xo = i % cw, yy = crop_top + yo;
  const int ymin = by[2 * yy], n = by[2 * yy + 1];
  const int* k = ky + (long long)yy * ksize;
  const unsigned char* p = tmp + (((long long)b * rows + (ymin - row_first)) * cw + xo) * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const int kv = __ldg(k + y);
    const unsigned char* q = p + (long long)y * cw * 3;
    s0 += q[0] * kv;
    s1 += q[1] * kv;
    s2 += q[2] * kv;
  }
  const long long plane = (long long)ch * cw;
  float* o = out + (long long)b * 3 * plane + (long long)yo * cw + xo;
  o[0] = ((float)clip8(s0) / 255.0f - m0) / d0;           // ToTensor: u8 -> float / 255; Normalize: (v - mean) / std
  o[plane] = ((float)clip8(s1) / 255.0f - m1) / d1;
  o[2 * plane] = ((float)clip8(s2) / 255.0f - m2) / d2;
}

}  // namespace
}  // namespace pio

extern "C" {

size_t pio_preprocess_workspace_bytes(int B, int rows, int crop_w) { return (size_t)B * rows * crop_w * 3 + 256; }

int pio_preprocess(const unsigned char* imgs, int B, int H, int W, const int* kx, const int* bx, int ksize_x, const int* ky,
                   const int* by, int ksize_y, int crop_left, int crop_top, int crop_w, int crop_h, int row_first, int rows,
                   const float* mean3, const float* std3, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pio;
  if (B == 0) return PIO_OK;
  PIO_CHECK(imgs && kx && bx && ky && by && mean3 && std3 && out && workspace, "preprocess: null argument");
  PIO_CHECK(workspace_bytes >= pio_preprocess_workspace_bytes(B, rows, crop_w), "preprocess: workspace too small");
  PIO_CHECK(rows > 0 && row_first >= 0 && row_first + rows <= H && crop_w > 0 && crop_h > 0, "preprocess: bad row / crop window");
  cudaStream_t st = as_stream(stream);
  unsigned char* tmp = (unsigned char*)workspace;
  resample_h_kernel<<<dim3(cdiv((long long)rows * crop_w, 256), B), 256, 0, st>>>(imgs, H, W, kx, bx, ksize_x, crop_left, crop_w,
                                                                                  row_first, rows, tmp);
  PIO_LAUNCHED();
  resample_v_normalize_kernel<<<dim3(cdiv((long long)crop_h * crop_w, 256), B), 256, 0, st>>>(
      tmp, rows, crop_w, ky, by, ksize_y, crop_top, crop_h, row_first, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out);
  PIO_LAUNCHED();
  return PIO_OK;
}
}
