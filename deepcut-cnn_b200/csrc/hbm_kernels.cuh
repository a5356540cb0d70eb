// HBM-bound kernels of the DeeperCut forward path (sm_100a): stem convolution, max-pool,
// stride-2 subsample, head finish (col2im + crop + add + sigmoid) and layout converters.
// All activations between kernels are NHWC "split fp16" (plane 0 = hi, plane 1 = lo).
#pragma once
#include "dc_ptx.cuh"

namespace dc {

// ---------------------------------------------------------------------------------------
// conv1: 7x7 / stride 2 / pad 3, 3 -> 64 channels, fp32 NCHW in (the `data` blob as Caffe
// holds it), fused BatchNorm+Scale+ReLU, split-fp16 NHWC out.
// Replaces ConvolutionLayer::Forward_gpu (conv_layer.cu:8-24: im2col K=147 + SGEMM M=64) +
// bn_conv1/scale_conv1/conv1_relu.  K=147 is too ragged for a TMA/UMMA tile and the layer is
// 0.8 % of the net's FLOPs, so this is an fp32 FFMA direct convolution: a CTA stages the
// input patch of an 8x32 output tile and the whole 147x64 filter bank in shared memory; each
// thread owns one output pixel x 64 channels (weights are warp-broadcast LDS.128).
// ---------------------------------------------------------------------------------------
constexpr int kC1TileH = 8, kC1TileW = 32, kC1K = 7, kC1Cout = 64;
constexpr int kC1PatchH = kC1TileH * 2 + 5, kC1PatchW = kC1TileW * 2 + 5;   // 21 x 69
constexpr int kC1PatchWPad = kC1PatchW + 2;                                 // 71: odd stride, no 2-way conflicts
constexpr int kC1SmemFloats = 147 * 64 + 3 * kC1PatchH * kC1PatchWPad;

__global__ void __launch_bounds__(256, 2) conv1_7x7s2_kernel(const float* __restrict__ x, const float* __restrict__ wk,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, __half* __restrict__ out,
                                                          long long out_plane, int N, int H, int W, int Ho, int Wo) {
  extern __shared__ float c1s[];
  float* sw = c1s;                    // [147][64]  (k = (ci*7+p)*7+q, co)
  float* sp = c1s + 147 * 64;         // [3][21][71]
  const int tiles_x = (Wo + kC1TileW - 1) / kC1TileW;
  const int tiles_y = (Ho + kC1TileH - 1) / kC1TileH;
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int oy0 = ty * kC1TileH, ox0 = tx * kC1TileW;
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;

  for (int i = threadIdx.x; i < 147 * 64 / 4; i += blockDim.x)
    reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(wk) + i);
  for (int i = threadIdx.x; i < 3 * kC1PatchH * kC1PatchW; i += blockDim.x) {
    const int c = i / (kC1PatchH * kC1PatchW);
    const int r = (i / kC1PatchW) % kC1PatchH;
    const int col = i % kC1PatchW;
    const int iy = iy0 + r, ix = ix0 + col;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((static_cast<long long>(n) * 3 + c) * H + iy) * W + ix);
    sp[(c * kC1PatchH + r) * kC1PatchWPad + col] = v;
  }
  __syncthreads();

  const int ly = threadIdx.x / kC1TileW, lx = threadIdx.x % kC1TileW;
  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;
  for (int c = 0; c < 3; ++c) {
    for (int p = 0; p < 7; ++p) {
      const float* prow = sp + (c * kC1PatchH + ly * 2 + p) * kC1PatchWPad + lx * 2;
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float xv = prow[q];
        const float4* w4 = reinterpret_cast<const float4*>(sw + ((c * 7 + p) * 7 + q) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 wv = w4[j];
          acc[4 * j + 0] = fmaf(xv, wv.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(xv, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(xv, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(xv, wv.w, acc[4 * j + 3]);
        }
      }
    }
  }
  const int oy = oy0 + ly, ox = ox0 + lx;
  if (oy < Ho && ox < Wo) {
    const long long off = ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * 64;
    __half* oh = out + off;
    __half* ol = out + out_plane + off;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      uint4 h4, l4;
      __half2* hh = reinterpret_cast<__half2*>(&h4);
      __half2* ll = reinterpret_cast<__half2*>(&l4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c0 = g * 8 + e * 2;
        float a = fmaxf(fmaf(acc[c0], __ldg(scale + c0), __ldg(shift + c0)), 0.f);
        float b = fmaxf(fmaf(acc[c0 + 1], __ldg(scale + c0 + 1), __ldg(shift + c0 + 1)), 0.f);
        __half ah, al, bh, bl;
        split_f16(a, ah, al);
        split_f16(b, bh, bl);
        hh[e] = __halves2half2(ah, bh);
        ll[e] = __halves2half2(al, bl);
      }
      reinterpret_cast<uint4*>(oh)[g] = h4;
      reinterpret_cast<uint4*>(ol)[g] = l4;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Stem on tensor cores.  A 7x7 / stride-2 / pad-3 convolution over 3 channels is a 4x4 / stride-1
// convolution over the 2x2 space-to-depth image (12 channels: c' = (py*2 + px)*3 + ci, padded to 16):
// input row 2*oy - 3 + P = 2*(oy + t) + py with t = floor((P-3)/2) in {-2..1}.  With 16 channels per
// pixel, FOUR horizontally adjacent pixels are 64 contiguous fp16 = one 128-byte K-chunk, so a tensor
// map whose W stride is one pixel (32 B) but whose inner extent is four pixels (overlapping windows)
// feeds conv_igemm directly: 4 vertical taps x K = 64.  This kernel writes that image, split fp16,
// with 2 zero columns on the left and 1 on the right so every window is in bounds:
//   out[plane][n][Y][X + 2][c'],  Y < ceil(H/2), X in [-2, ceil(W/2) + 1).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_s2d_kernel(const float* __restrict__ x, __half* __restrict__ out,
                                                       long long out_plane, int N, int H, int W, int H2, int W2) {
  const int Wp = W2 + 3;
  const long long total = static_cast<long long>(N) * H2 * Wp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xp = static_cast<int>(i % Wp);
    long long r = i / Wp;
    const int Y = static_cast<int>(r % H2);
    const int n = static_cast<int>(r / H2);
    const int X = xp - 2;
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.f;
    if (X >= 0 && X < W2) {
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const int iy = 2 * Y + py;
        if (iy >= H) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float* row = x + ((static_cast<long long>(n) * 3 + ci) * H + iy) * W;
          const int ix = 2 * X;
          v[(py * 2 + 0) * 3 + ci] = __ldg(row + ix);
          if (ix + 1 < W) v[(py * 2 + 1) * 3 + ci] = __ldg(row + ix + 1);
        }
      }
    }
    uint4 h4[2], l4[2];
    __half2* hh = reinterpret_cast<__half2*>(h4);
    __half2* ll = reinterpret_cast<__half2*>(l4);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const __half2 h2 = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      const float2 hf = __half22float2(h2);
      hh[e] = h2;
      ll[e] = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
    }
    uint4* oh = reinterpret_cast<uint4*>(out + i * 16);
    uint4* ol = reinterpret_cast<uint4*>(out + out_plane + i * 16);
    oh[0] = h4[0]; oh[1] = h4[1];
    ol[0] = l4[0]; ol[1] = l4[1];
  }
}

// ---------------------------------------------------------------------------------------
// MAX pooling k x k / stride s, pad 0, Caffe ceil-mode output size, windows clipped at the
// bottom/right edge (PoolingLayer::Forward_gpu pooling_layer.cu:10-47).  One thread = one
// output pixel x 8 channels (128-bit loads of hi and lo).  The winner's (hi, lo) pair is
// copied verbatim, so the result is bit-identical to pooling the joined values.
// ---------------------------------------------------------------------------------------
// KT > 0: compile-time window (the net's 3x3) -- all KT*KT*2 128-bit loads are issued before the first compare, so a
// thread keeps 18 loads in flight instead of 2 (the runtime-k loop is latency-bound at ~50 % of the copy bandwidth).
template <int KT>
__global__ void __launch_bounds__(256) maxpool_split_kernel(const __half* __restrict__ in, long long in_plane,
                                                            __half* __restrict__ out, long long out_plane, int N,
                                                            int H, int W, int C, int Ho, int Wo, int k, int s) {
  const int cg = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * cg;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % cg);
    long long pidx = i / cg;
    const int ox = static_cast<int>(pidx % Wo);
    pidx /= Wo;
    const int oy = static_cast<int>(pidx % Ho);
    const int n = static_cast<int>(pidx / Ho);
    const int y0 = oy * s, x0 = ox * s;
    float best[8];
    __half bh[8], bl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -3.402823466e+38f; bh[e] = __float2half(0.f); bl[e] = __float2half(0.f); }
    auto consider = [&](const uint4& h4, const uint4& l4) {
      const __half* hh = reinterpret_cast<const __half*>(&h4);
      const __half* ll = reinterpret_cast<const __half*>(&l4);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float v = join_f16(hh[e], ll[e]);
        if (v > best[e]) { best[e] = v; bh[e] = hh[e]; bl[e] = ll[e]; }
      }
    };
    if constexpr (KT > 0) {
      constexpr int kWin = KT > 0 ? KT * KT : 1;
      uint4 vh[kWin], vl[kWin];
#pragma unroll
      for (int dy = 0; dy < KT; ++dy)
#pragma unroll
        for (int dx = 0; dx < KT; ++dx) {
          // windows are clipped at the bottom/right edge: clamp the address, skip the compare below
          const int y = min(y0 + dy, H - 1), xx = min(x0 + dx, W - 1);
          const long long off = ((static_cast<long long>(n) * H + y) * W + xx) * C + g * 8;
          vh[dy * KT + dx] = __ldg(reinterpret_cast<const uint4*>(in + off));
          vl[dy * KT + dx] = __ldg(reinterpret_cast<const uint4*>(in + in_plane + off));
        }
#pragma unroll
      for (int dy = 0; dy < KT; ++dy)
#pragma unroll
        for (int dx = 0; dx < KT; ++dx)
          if (y0 + dy < H && x0 + dx < W) consider(vh[dy * KT + dx], vl[dy * KT + dx]);
    } else {
      const int y1 = min(y0 + k, H), x1 = min(x0 + k, W);
      for (int y = y0; y < y1; ++y)
        for (int xx = x0; xx < x1; ++xx) {
          const long long off = ((static_cast<long long>(n) * H + y) * W + xx) * C + g * 8;
          consider(__ldg(reinterpret_cast<const uint4*>(in + off)), __ldg(reinterpret_cast<const uint4*>(in + in_plane + off)));
        }
    }
    const long long ooff = ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + g * 8;
    *reinterpret_cast<uint4*>(out + ooff) = *reinterpret_cast<const uint4*>(bh);
    *reinterpret_cast<uint4*>(out + out_plane + ooff) = *reinterpret_cast<const uint4*>(bl);
  }
}

// ---------------------------------------------------------------------------------------
// Stride-s spatial subsample (both planes): out[n,y,x,:] = in[n,s*y,s*x,:].  Feeds the four
// stride-2 1x1 convolutions (res3a/res4a branch1 + branch2a), which the reference runs through
// im2col because stride != 1 (base_conv_layer.cpp:109-116).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) subsample_split_kernel(const __half* __restrict__ in, long long in_plane,
                                                              __half* __restrict__ out, long long out_plane, int N,
                                                              int H, int W, int C, int Ho, int Wo, int s) {
  const int cg = C / 8;
  const long long total = static_cast<long long>(N) * Ho * Wo * cg * 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % cg);
    long long pidx = i / cg;
    const int ox = static_cast<int>(pidx % Wo);
    pidx /= Wo;
    const int oy = static_cast<int>(pidx % Ho);
    pidx /= Ho;
    const int n = static_cast<int>(pidx % N);
    const int plane = static_cast<int>(pidx / N);
    const long long src = plane * in_plane + ((static_cast<long long>(n) * H + oy * s) * W + ox * s) * C + g * 8;
    const long long dst = plane * out_plane + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * C + g * 8;
    *reinterpret_cast<uint4*>(out + dst) = __ldg(reinterpret_cast<const uint4*>(in + src));
  }
}

// ---------------------------------------------------------------------------------------
// Head finish.  Inputs are the two fp32 GEMM results of the merged heads, channel-major (the
// layout of Caffe's own col buffer, base_conv_layer.cpp:358-365):
//   col  [rows = co*9 + p*3 + q][ldcol >= N*h*w]   col(r, pixel(n,i,j)) = sum_ci res5c[n,ci,i,j] * Wd[ci,co,p,q]
//   skip [rows = co][ldskip >= N*Ho*Wo]            1x1 heads on res3b7 (+ both biases, folded into the GEMM shift)
// Output (fp32 NCHW, the layout Caffe exposes):  out[n,co,y,x] =
//   skip + sum_{p,q : (y-p),(x-q) even, in range} col((co,p,q), (n,(y-p)/2,(x-q)/2))  [-> sigmoid]
// i.e. DeconvolutionLayer col2im (im2col.cu:246-305) + Crop to Ho x Wo at offset 0
// (crop_layer.cu:9-38) + Eltwise SUM (eltwise_layer.cu:47-53) + Sigmoid (sigmoid_layer.cu:8-24).
// One thread per INPUT cell (i, j) of the h x w grid = the 2 x 2 output pixels (2i..2i+1, 2j..2j+1) it feeds: the nine
// col rows are read once each at (i, j) / (i-1, j) / (i, j-1) / (i-1, j-1) -- unit-stride along j across the warp -- and
// the two output rows leave as float2 (a warp writes 256 contiguous bytes per row).  Per output pixel the taps are
// summed in (p, q) order, then added to the skip value, like the one-thread-per-pixel formulation it replaces.
// ---------------------------------------------------------------------------------------
// grid = (N * Cout, ceil(ceil(Ho/2) * ceil(Wo/2) / 256)); VEC: Wo even and 8-byte-aligned rows (float2 path).
template <bool VEC>
__global__ void __launch_bounds__(256) head_finish_kernel(const float* __restrict__ col, long long ldcol, int col_row0,
                                                          const float* __restrict__ skip, long long ldskip, int skip_row0,
                                                          float* __restrict__ out, int N, int Cout, int h, int w,
                                                          int Ho, int Wo, int do_sigmoid) {
  const int n = blockIdx.x / Cout, co = blockIdx.x % Cout;
  const int plane = Ho * Wo;
  const int ch = (Ho + 1) >> 1, cw = (Wo + 1) >> 1;
  const int cell = blockIdx.y * 256 + threadIdx.x;
  if (cell >= ch * cw) return;
  const int i = cell / cw, j = cell - i * cw;
  const float* crow = col + (static_cast<long long>(col_row0) + co * 9) * ldcol + static_cast<long long>(n) * h * w;
  const float* srow = skip + (static_cast<long long>(skip_row0) + co) * ldskip + static_cast<long long>(n) * plane;
  float* orow = out + (static_cast<long long>(n) * Cout + co) * plane;
  auto tap = [&](int pq, int ii, int jj) -> float {
    return (ii >= 0 && ii < h && jj >= 0 && jj < w) ? __ldg(crow + pq * ldcol + ii * w + jj) : 0.f;
  };
  // (p, q) order within each output pixel
  const float u00 = ((tap(0, i, j) + tap(2, i, j - 1)) + tap(6, i - 1, j)) + tap(8, i - 1, j - 1);   // (2i,   2j)
  const float u01 = tap(1, i, j) + tap(7, i - 1, j);                                                  // (2i,   2j+1)
  const float u10 = tap(3, i, j) + tap(5, i, j - 1);                                                  // (2i+1, 2j)
  const float u11 = tap(4, i, j);                                                                     // (2i+1, 2j+1)
  const int y0 = 2 * i, x0 = 2 * j;
  const bool x1ok = x0 + 1 < Wo, y1ok = y0 + 1 < Ho;
  auto fin = [&](float s_, float u) -> float {
    float v = s_ + u;
    if (do_sigmoid) v = 1.f / (1.f + expf(-v));
    return v;
  };
  if (VEC) {      // Wo even => x1ok always
    const float2 s0 = __ldg(reinterpret_cast<const float2*>(srow + y0 * Wo + x0));
    *reinterpret_cast<float2*>(orow + y0 * Wo + x0) = make_float2(fin(s0.x, u00), fin(s0.y, u01));
    if (y1ok) {
      const float2 s1 = __ldg(reinterpret_cast<const float2*>(srow + (y0 + 1) * Wo + x0));
      *reinterpret_cast<float2*>(orow + (y0 + 1) * Wo + x0) = make_float2(fin(s1.x, u10), fin(s1.y, u11));
    }
  } else {
    orow[y0 * Wo + x0] = fin(__ldg(srow + y0 * Wo + x0), u00);
    if (x1ok) orow[y0 * Wo + x0 + 1] = fin(__ldg(srow + y0 * Wo + x0 + 1), u01);
    if (y1ok) {
      orow[(y0 + 1) * Wo + x0] = fin(__ldg(srow + (y0 + 1) * Wo + x0), u10);
      if (x1ok) orow[(y0 + 1) * Wo + x0 + 1] = fin(__ldg(srow + (y0 + 1) * Wo + x0 + 1), u11);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Pose read-out on the device (the demo's _pose_from_mats, python/pose/estimate_pose.py:131-143):
// per image and joint j, the arg-max of prob[n, j] (first maximum in row-major order, like np.argmax),
// then  y = (my*stride + stride/2 + loc[n, 2j+1, my, mx]*s) / scale,  x likewise with loc[n, 2j].
// out[n][5][J] = {x, y, confidence, loc[2j+1]*s/scale, loc[2j]*s/scale}  (the demo's row order).
// One CTA per (n, joint); block-wide arg-max by warp shuffles, ties broken towards the lower index.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pose_from_maps_kernel(const float* __restrict__ prob, const float* __restrict__ loc,
                                                             int J, int H, int W, float stride, float locref_scale,
                                                             float scale, float* __restrict__ out) {
  const int n = blockIdx.x / J, j = blockIdx.x % J;
  const int plane = H * W;
  const float* pm = prob + (static_cast<long long>(n) * J + j) * plane;
  float best = -3.402823466e+38f;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < plane; i += blockDim.x) {
    const float v = __ldg(pm + i);
    if (v > best) { best = v; besti = i; }        // strictly greater: keeps the first maximum of this thread's stride
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k)
      if (sv[k] > best || (sv[k] == best && si[k] < besti)) { best = sv[k]; besti = si[k]; }
    const int my = besti / W, mx = besti % W;
    const float* lp = loc + (static_cast<long long>(n) * 2 * J + 2 * j) * plane + besti;
    const float off_x = __ldg(lp) * locref_scale, off_y = __ldg(lp + plane) * locref_scale;
    float* o = out + static_cast<long long>(n) * 5 * J + j;
    o[0 * J] = (mx * stride + 0.5f * stride + off_x) / scale;
    o[1 * J] = (my * stride + 0.5f * stride + off_y) / scale;
    o[2 * J] = best;
    o[3 * J] = off_y / scale;
    o[4 * J] = off_x / scale;
  }
}

// ---------------------------------------------------------------------------------------
// Layout converters (blob materialisation / test harness): fp32 NCHW <-> split-fp16 NHWC.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_split_kernel(const float* __restrict__ in, __half* __restrict__ out,
                                                            long long out_plane, int N, int C, int H, int W) {
  const long long total = static_cast<long long>(N) * C * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const float v = __ldg(in + ((static_cast<long long>(n) * C + c) * H + y) * W + x);
    __half hi, lo;
    split_f16(v, hi, lo);
    out[i] = hi;
    out[out_plane + i] = lo;
  }
}

__global__ void __launch_bounds__(256) split_to_nchw_kernel(const __half* __restrict__ in, long long in_plane,
                                                            float* __restrict__ out, int N, int C, int H, int W) {
  const long long total = static_cast<long long>(N) * C * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    long long r = i / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int c = static_cast<int>(r % C);
    const int n = static_cast<int>(r / C);
    const long long src = ((static_cast<long long>(n) * H + y) * W + x) * C + c;
    out[i] = join_f16(in[src], in[in_plane + src]);
  }
}

// ---------------------------------------------------------------------------------------
// Demo pre-processing on the device (python/pose/estimate_pose.py:83-105): uint8 HWC image -> edge-replicated 64 px
// below / right -> PIL bilinear rescale (Pillow Resample.c, 8 bits per channel: separable triangle filter, 22-bit
// fixed-point coefficients, uint8 rounding after EACH pass, horizontal first) -> minus the per-channel mean -> top-left
// crop (zero fill beyond the rescaled image) as fp32 CHW.  Byte/integer work, HBM-bound and tiny next to the net
// (a 720p frame is 2.8 MB in, 11 MB out): one thread per output pixel, x fastest, so both passes read and write
// coalesced rows.  bounds[i] = {first source index, tap count}, kk[i*ksize + k] the integer coefficients
// (host: build_resample_table in dc_abi.cu).
// ---------------------------------------------------------------------------------------
constexpr int kResampleBits = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) {
  v >>= kResampleBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass over the virtual padded image: mid[y][xo][c], y < rows, xo < cols
__global__ void __launch_bounds__(256) preprocess_hpass_kernel(const unsigned char* __restrict__ img, int h, int w, int rows, int cols,
                                                               const int2* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                               unsigned char* __restrict__ mid) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xo >= cols || y >= rows) return;
  const unsigned char* row = img + static_cast<size_t>(min(y, h - 1)) * w * 3;
  const int2 b = bounds[xo];
  int a0 = 1 << (kResampleBits - 1), a1 = a0, a2 = a0;
  for (int k = 0; k < b.y; ++k) {
    const int c = kk[xo * ksize + k];
    const unsigned char* px = row + min(b.x + k, w - 1) * 3;
    a0 += px[0] * c;
    a1 += px[1] * c;
    a2 += px[2] * c;
  }
  unsigned char* o = mid + (static_cast<size_t>(y) * cols + xo) * 3;
  o[0] = static_cast<unsigned char>(clip8(a0));
  o[1] = static_cast<unsigned char>(clip8(a1));
  o[2] = static_cast<unsigned char>(clip8(a2));
}

// vertical pass (or none) + mean subtraction + crop / zero fill: out[c][yo][xo] fp32.  HP: the source is the horizontal
// pass's output `mid` [.][cols][3]; otherwise the virtual padded image itself.  VP: apply the vertical taps.
template <bool HP, bool VP>
__global__ void __launch_bounds__(256) preprocess_finish_kernel(const unsigned char* __restrict__ src, int h, int w, int cols, int valid_h,
                                                                int valid_w, const int2* __restrict__ bounds, const int* __restrict__ kk,
                                                                int ksize, float m0, float m1, float m2, float* __restrict__ out, int out_h,
                                                                int out_w) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  const int yo = blockIdx.y;
  if (xo >= out_w || yo >= out_h) return;
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (yo < valid_h && xo < valid_w) {
    auto pixel = [&](int y) -> const unsigned char* {
      if (HP) return src + (static_cast<size_t>(y) * cols + xo) * 3;
      return src + (static_cast<size_t>(min(y, h - 1)) * w + min(xo, w - 1)) * 3;
    };
    int a0, a1, a2;
    if (VP) {
      const int2 b = bounds[yo];
      a0 = a1 = a2 = 1 << (kResampleBits - 1);
      for (int k = 0; k < b.y; ++k) {
        const int c = kk[yo * ksize + k];
        const unsigned char* px = pixel(b.x + k);
        a0 += px[0] * c;
        a1 += px[1] * c;
        a2 += px[2] * c;
      }
      a0 = clip8(a0);
      a1 = clip8(a1);
      a2 = clip8(a2);
    } else {
      const unsigned char* px = pixel(yo);
      a0 = px[0];
      a1 = px[1];
      a2 = px[2];
    }
    v0 = static_cast<float>(a0) - m0;
    v1 = static_cast<float>(a1) - m1;
    v2 = static_cast<float>(a2) - m2;
  }
  const size_t plane = static_cast<size_t>(out_h) * out_w;
  const size_t o = static_cast<size_t>(yo) * out_w + xo;
  out[o] = v0;
  out[plane + o] = v1;
  out[2 * plane + o] = v2;
}

// ---------------------------------------------------------------------------------------
// Batch of uint8 images [n][h][w][3] (what a decoder hands over; 3 B/pixel over PCIe / NVLink instead of the 12 B/pixel of the
// float net input) -> the `data` blob, fp32 [n][3][h][w], minus the per-channel mean (estimate_pose.py:25,99).
// Byte work, HBM-bound: each thread converts 4 adjacent pixels = 12 bytes in (three aligned 32-bit loads when w % 4 == 0),
// three coalesced float4 stores out.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) images_u8_to_blob_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                                                 long long quads, int hw, float m0, float m1, float m2) {
  // quads = n * h * w / 4 (w % 4 == 0); a quad never straddles images because h*w % 4 == 0
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < quads; q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long px = q * 4;
    const long long n = px / hw;
    const int o = static_cast<int>(px - n * hw);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(img + px * 3);
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    // bytes: c0 c1 c2 | c0 c1 c2 | c0 c1 c2 | c0 c1 c2
    const float4 p0 = make_float4((w0 & 0xFF) - m0, ((w0 >> 24) & 0xFF) - m0, ((w1 >> 16) & 0xFF) - m0, ((w2 >> 8) & 0xFF) - m0);
    const float4 p1 = make_float4(((w0 >> 8) & 0xFF) - m1, (w1 & 0xFF) - m1, ((w1 >> 24) & 0xFF) - m1, ((w2 >> 16) & 0xFF) - m1);
    const float4 p2 = make_float4(((w0 >> 16) & 0xFF) - m2, ((w1 >> 8) & 0xFF) - m2, (w2 & 0xFF) - m2, ((w2 >> 24) & 0xFF) - m2);
    float* dst = out + n * 3 * static_cast<long long>(hw) + o;
    *reinterpret_cast<float4*>(dst) = p0;
    *reinterpret_cast<float4*>(dst + hw) = p1;
    *reinterpret_cast<float4*>(dst + 2 * static_cast<long long>(hw)) = p2;
  }
}
__global__ void __launch_bounds__(256) images_u8_to_blob_scalar_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                                                        long long pixels, int hw, float m0, float m1, float m2) {
  for (long long px = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; px < pixels; px += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = px / hw;
    const long long o = px - n * hw;
    const unsigned char* s = img + px * 3;
    float* dst = out + n * 3 * static_cast<long long>(hw) + o;
    dst[0] = static_cast<float>(s[0]) - m0;
    dst[hw] = static_cast<float>(s[1]) - m1;
    dst[2 * static_cast<long long>(hw)] = static_cast<float>(s[2]) - m2;
  }
}

}  // namespace dc
