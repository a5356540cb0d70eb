// fp32 NCHW kernels behind the per-layer plugin path (Layer::Forward_gpu called one layer at a
// time, e.g. Net::ForwardFromTo or a user driving a single layer).  Each replaces the reference
// kernel named beside it; they are simple grid-stride HBM kernels -- the throughput path is the
// fused plan, these exist so every layer of the boundary has a CUDA implementation and no CPU one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dc {

#define DC_GRID_STRIDE(i, n)                                                                         \
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < (n);         \
       i += static_cast<long long>(gridDim.x) * blockDim.x)

// BatchNormLayer::Forward_gpu inference branch (batch_norm_layer.cu:22-89): (x - mean[c]) / std[c]
__global__ void bn_nchw_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ sd,
                               long long total, int c, int hw, float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) {
    const int ch = static_cast<int>((i / hw) % c);
    y[i] = (x[i] - __ldg(mean + ch)) / __ldg(sd + ch);
  }
}
// ScaleBiasForward (scale_layer.cu:19-27): x * gamma[c] + beta[c]
__global__ void scale_nchw_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, long long total, int c, int hw, float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) {
    const int ch = static_cast<int>((i / hw) % c);
    const float v = x[i] * __ldg(gamma + ch);
    y[i] = beta ? v + __ldg(beta + ch) : v;
  }
}
// ReLUForward (relu_layer.cu:8-14)
__global__ void relu_kernel(const float* __restrict__ x, long long total, float slope, float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) { const float v = x[i]; y[i] = v > 0.f ? v : v * slope; }
}
// SigmoidForward (sigmoid_layer.cu:8-13)
__global__ void sigmoid_kernel(const float* __restrict__ x, long long total, float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) { y[i] = 1.f / (1.f + expf(-x[i])); }
}
// Eltwise SUM (eltwise_layer.cu:47-53: memset + axpy per bottom): y = ca*a + cb*b
__global__ void axpby_kernel(const float* __restrict__ a, float ca, const float* __restrict__ b, float cb, long long total,
                             float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) { y[i] = (0.f + ca * a[i]) + cb * b[i]; }
}
// Crop copy_kernel (crop_layer.cu:9-38)
__global__ void crop_nchw_kernel(const float* __restrict__ x, int h, int w, int off_h, int off_w, int ho, int wo,
                                 long long total, float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) {
    const int ox = static_cast<int>(i % wo);
    long long r = i / wo;
    const int oy = static_cast<int>(r % ho);
    const long long nc = r / ho;
    y[i] = x[(nc * h + oy + off_h) * w + ox + off_w];
  }
}
// MaxPoolForward (pooling_layer.cu:10-47)
__global__ void maxpool_nchw_kernel(const float* __restrict__ x, int h, int w, int kh, int kw, int sh, int sw, int ph,
                                    int pw, int ho, int wo, long long total, float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) {
    const int ox = static_cast<int>(i % wo);
    long long r = i / wo;
    const int oy = static_cast<int>(r % ho);
    const long long nc = r / ho;
    int y0 = oy * sh - ph, x0 = ox * sw - pw;
    const int y1 = min(y0 + kh, h), x1 = min(x0 + kw, w);
    y0 = max(y0, 0);
    x0 = max(x0, 0);
    float best = -3.402823466e+38f;
    const float* src = x + nc * h * w;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) best = fmaxf(best, src[yy * w + xx]);
    y[i] = best;
  }
}
// Generic direct convolution, any kernel/stride/pad/dilation, group 1: the per-layer path for
// geometries the tcgen05 kernel does not take (cin % 64 != 0, stride > 1, > 9 taps).
// Same sum as im2col_gpu + SGEMM (im2col.cu:8-62): zero outside the image.
__global__ void conv_direct_nchw_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                                        const float* __restrict__ bias, int cin, int h, int w, int cout, int kh, int kw,
                                        int sh, int sw, int ph, int pw, int dh, int dw, int ho, int wo, long long total,
                                        float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) {
    const int ox = static_cast<int>(i % wo);
    long long r = i / wo;
    const int oy = static_cast<int>(r % ho);
    r /= ho;
    const int co = static_cast<int>(r % cout);
    const long long n = r / cout;
    float acc = 0.f;
    for (int ci = 0; ci < cin; ++ci) {
      const float* xs = x + (n * cin + ci) * h * w;
      const float* ws = wgt + (static_cast<long long>(co) * cin + ci) * kh * kw;
      for (int p = 0; p < kh; ++p) {
        const int iy = oy * sh - ph + p * dh;
        if (iy < 0 || iy >= h) continue;
        for (int q = 0; q < kw; ++q) {
          const int ix = ox * sw - pw + q * dw;
          if (ix < 0 || ix >= w) continue;
          acc = fmaf(xs[iy * w + ix], __ldg(ws + p * kw + q), acc);
        }
      }
    }
    y[i] = bias ? acc + __ldg(bias + co) : acc;
  }
}
// Generic transposed convolution, gather form like col2im_gpu_kernel (im2col.cu:246-285):
// y[n,co,oy,ox] = b[co] + sum_{ci,p,q : oy = iy*s - pad + p*d} x[n,ci,iy,ix] * W[ci,co,p,q]
__global__ void deconv_direct_nchw_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                                          const float* __restrict__ bias, int cin, int h, int w, int cout, int kh, int kw,
                                          int sh, int sw, int ph, int pw, int dh, int dw, int ho, int wo, long long total,
                                          float* __restrict__ y) {
  DC_GRID_STRIDE(i, total) {
    const int ox = static_cast<int>(i % wo);
    long long r = i / wo;
    const int oy = static_cast<int>(r % ho);
    r /= ho;
    const int co = static_cast<int>(r % cout);
    const long long n = r / cout;
    float acc = 0.f;
    for (int p = 0; p < kh; ++p) {
      const int ty = oy + ph - p * dh;
      if (ty < 0 || ty % sh) continue;
      const int iy = ty / sh;
      if (iy >= h) continue;
      for (int q = 0; q < kw; ++q) {
        const int tx = ox + pw - q * dw;
        if (tx < 0 || tx % sw) continue;
        const int ix = tx / sw;
        if (ix >= w) continue;
        for (int ci = 0; ci < cin; ++ci)
          acc = fmaf(x[((n * cin + ci) * h + iy) * w + ix], __ldg(wgt + ((static_cast<long long>(ci) * cout + co) * kh + p) * kw + q), acc);
      }
    }
    y[i] = bias ? acc + __ldg(bias + co) : acc;
  }
}

}  // namespace dc
