// postprocess.cuh — the reference's Postprocessor (reference include/Postprocessor.hpp:29-197) as
// kernels: emissive extract (:131-156), separable Gaussian blur (:64-128), add (:158-175) and the
// exposure tone map (:182-207).  SURVEY.md §8 f-3.
//
// Every read of the reference goes through Texture::getRGBat(clamp(0, 0.999, x / w), clamp(0, 0.999, y / h))
// (Texture.hpp:18-39), which is NOT the identity on pixel coordinates:
//   * u == 0 takes the `else` branch (u = 1 - (|u| - (int)|u|) = 1), so column 0 reads index y*w + w (the
//     first texel of the NEXT row) and row 0 reads index h*w + x, which the bounds clamp turns into the
//     LAST texel of the image; taps that fall above the image or left of it behave the same way;
//   * (int)(fl(x / w) * w) is x - 1 for the x whose quotient rounds down;
//   * u, v > 0.999 clamp (the last columns / rows of frames wider than 1000 pixels repeat).
// pp_index() below is that function, literally, in exact fp32 (no contraction), so the kernels gather
// exactly the texels the reference gathers.  All arithmetic that reaches the output is written with
// round-to-nearest intrinsics in the reference's operation order: results are bit-identical except where
// libm enters (expf in the tone map: evaluated in double and rounded once, which equals a correctly
// rounded expf; the Gaussian weights are evaluated on the host with the reference's own expression).
//
// These are HBM-streaming kernels (12 B read + 12 B written per pixel and stage; the 10 blur taps of a
// pixel hit L1/L2); grids are a multiple of the SM count with a grid-stride loop.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tutu {

constexpr int kPostMaxKernel = 64;

struct PostWeights {       // by value in the kernel-parameter bank
  float g[kPostMaxKernel]; // gaussian(start + i, stddev), Postprocessor.hpp:77-79
  float sum;               // kernelSum accumulated in tap order
  int n;                   // kernelSize
  int start;               // (int)(-kernelSize * 0.5)
};

// std::max(lo, std::min(hi, v)) (global.hpp:52-55); v is never NaN here (x / w of integers)
__device__ __forceinline__ float pp_clamp(float lo, float hi, float v) {
  const float m = hi < v ? hi : v;  // std::min(hi, v)
  return lo < m ? m : lo;           // std::max(lo, m)
}

// Texture::getRGBat's texel index for the (already clamped) coordinates u, v
__device__ __forceinline__ int pp_texel(float u, float v, int width, int height) {
  if (u > 0.f)
    u = __fsub_rn(u, (float)(int)u);
  else
    u = __fsub_rn(1.f, __fsub_rn(fabsf(u), (float)(int)fabsf(u)));
  if (v > 0.f)
    v = __fsub_rn(v, (float)(int)v);
  else
    v = __fsub_rn(1.f, __fsub_rn(fabsf(v), (float)(int)fabsf(v)));
  const int x = (int)__fmul_rn(u, (float)width);
  const int y = (int)__fmul_rn(v, (float)height);
  long long index = (long long)y * width + x;
  const long long size = (long long)width * height;
  if (index < 0) index = 0;
  if (index >= size) index = size - 1;
  return (int)index;
}

// the texel the reference reads for pixel column x / tap row y (either may lie outside the image)
__device__ __forceinline__ int pp_index(int x, int y, int width, int height) {
  const float U = __fdiv_rn((float)x, (float)width);
  const float V = __fdiv_rn((float)y, (float)height);
  return pp_texel(pp_clamp(0.f, 0.999f, U), pp_clamp(0.f, 0.999f, V), width, height);
}

struct PostRgb {
  float x, y, z;
};
__device__ __forceinline__ PostRgb pp_load(const float* __restrict__ img, int texel) {
  const float* p = img + 3 * (size_t)texel;
  return PostRgb{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}
__device__ __forceinline__ void pp_store(float* __restrict__ img, size_t pixel, PostRgb c) {
  float* p = img + 3 * pixel;
  p[0] = c.x, p[1] = c.y, p[2] = c.z;
}

// getEmmisiveTexture: pixels brighter than |rgb| > threshold are rescaled so that their largest channel
// becomes `strength` (rescale(), global.hpp:66-68: targetMin + (targetMax - targetMin) * (in - 0) / (mx - 0)).
__device__ __forceinline__ PostRgb pp_emissive(PostRgb col, float threshold, float strength) {
  const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(col.x, col.x), __fmul_rn(col.y, col.y)), __fmul_rn(col.z, col.z)));
  PostRgb out{0.f, 0.f, 0.f};
  if (norm > threshold) {
    float mx = col.x > col.y ? col.x : col.y;
    mx = mx > col.z ? mx : col.z;
    const float range = __fsub_rn(strength, 0.f), den = __fsub_rn(mx, 0.f);
    out.x = __fadd_rn(0.f, __fdiv_rn(__fmul_rn(range, __fsub_rn(col.x, 0.f)), den));
    out.y = __fadd_rn(0.f, __fdiv_rn(__fmul_rn(range, __fsub_rn(col.y, 0.f)), den));
    out.z = __fadd_rn(0.f, __fdiv_rn(__fmul_rn(range, __fsub_rn(col.z, 0.f)), den));
  }
  return out;
}

// 1 - exp(-c * EXPOSURE): expf evaluated in double and rounded once (= a correctly rounded expf)
__device__ __forceinline__ float pp_tonemap(float c, float exposure) {
  const float a = __fmul_rn(-c, exposure);
  return __fsub_rn(1.f, (float)exp((double)a));
}

__global__ void __launch_bounds__(256)
pp_extract(const float* __restrict__ src, int width, int height, float threshold, float strength, float* __restrict__ dst) {
  const size_t n = (size_t)width * height;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % (size_t)width), y = (int)(i / (size_t)width);
    pp_store(dst, i, pp_emissive(pp_load(src, pp_index(x, y, width, height)), threshold, strength));
  }
}

// one pass of getGaussianBlurTexture: VERTICAL taps walk rows (Postprocessor.hpp:82-101), else columns (:103-121)
template <bool VERTICAL>
__global__ void __launch_bounds__(256)
pp_blur(const float* __restrict__ src, int width, int height, const __grid_constant__ PostWeights w, float* __restrict__ dst) {
  const size_t n = (size_t)width * height;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % (size_t)width), y = (int)(i / (size_t)width);
    PostRgb col{0.f, 0.f, 0.f};
    for (int k = 0; k < w.n; ++k) {
      const int tx = VERTICAL ? x : x + k + w.start, ty = VERTICAL ? y + k + w.start : y;
      const PostRgb t = pp_load(src, pp_index(tx, ty, width, height));
      const float g = w.g[k];
      col.x = __fadd_rn(col.x, __fmul_rn(t.x, g));
      col.y = __fadd_rn(col.y, __fmul_rn(t.y, g));
      col.z = __fadd_rn(col.z, __fmul_rn(t.z, g));
    }
    pp_store(dst, i, PostRgb{__fdiv_rn(col.x, w.sum), __fdiv_rn(col.y, w.sum), __fdiv_rn(col.z, w.sum)});
  }
}

// add (pixelwise, no texture fetch) and, fused behind it when TONEMAP, getHDRtexture (which DOES fetch through
// pp_index).  bloom == nullptr: tone map of `src` alone.
template <bool TONEMAP>
__global__ void __launch_bounds__(256)
pp_combine(const float* __restrict__ src, const float* __restrict__ bloom, int width, int height, float exposure,
           float* __restrict__ dst) {
  const size_t n = (size_t)width * height;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = i;
    if (TONEMAP) t = (size_t)pp_index((int)(i % (size_t)width), (int)(i / (size_t)width), width, height);
    PostRgb c = pp_load(src, (int)t);
    if (bloom) {
      const PostRgb b = pp_load(bloom, (int)t);
      c.x = __fadd_rn(c.x, b.x), c.y = __fadd_rn(c.y, b.y), c.z = __fadd_rn(c.z, b.z);
    }
    if (TONEMAP) c = PostRgb{pp_tonemap(c.x, exposure), pp_tonemap(c.y, exposure), pp_tonemap(c.z, exposure)};
    pp_store(dst, i, c);
  }
}

}  // namespace tutu
