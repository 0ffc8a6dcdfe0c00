// Fused PONO + SPADE modulation + LeakyReLU + reflection pad (HBM bound).
//
//   y = pad_reflect( lrelu( (x - mean_c) / sqrt(var_c + eps) * (1 + gamma) + beta ) )
//
// replaces, per SPADE layer, reference normalization.py:63-68 (PositionalNorm2d,
// unbiased variance over channels per pixel), :149 (x_hat * (1 + gamma) + beta),
// architecture.py:94-95 (leaky_relu 0.2) and the nn.ReflectionPad2d in front of
// the following 3x3 conv (architecture.py:31,73-74): ~11 elementwise launches
// and as many full-activation HBM round trips in forward, ~20 in backward,
// become one kernel each way.
//
// Layout NCHW fp32.  One thread per (padded) output pixel; the channel loop
// strides by H*W so a warp reads 32 consecutive pixels of one channel
// (coalesced).  gamma/beta come as one tensor gb [B, 2C, H, W] (the gamma and
// beta convs share their input and run as one conv with concatenated filters).
// Algorithmic bytes per element of x: forward 4 (x) + 8 (gamma,beta) + 4 (y)
// = 16 B; backward 4 (dy) + 4 (x) + 8 (gb) + 4 (dx) + 8 (dgb) = 28 B.
#include "corr_kernels.h"
#include "tmap.h"

namespace cocos {

namespace {

__device__ __forceinline__ int reflect(int i, int n) {
  // index into [0, n) of padded coordinate i in [-pad, n + pad)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Thread mapping: blockDim = (32 pixels, S channel slices).  A warp = 32 consecutive (padded) pixels of one
// slice, so every global access is a coalesced 128 B row; thread (x, y) walks channels y, y+S, y+2S, ...; the
// per-pixel sums are reduced across the S slices through shared memory.  S = 8 for wide maps, 32 for the
// 8x8 .. 32x32 maps with 512-1024 channels (a thread-per-pixel mapping leaves those layers with a few hundred
// threads on the whole GPU).
template <int S>
__device__ __forceinline__ float slice_reduce(float v, float (*red)[32]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < S; ++i) t += red[i][threadIdx.x];
  __syncthreads();
  return t;
}

template <int S>
__global__ void __launch_bounds__(32 * S)
spade_mod_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gb, float* __restrict__ y,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int H, int W, int pad,
                     float slope, float eps) {
  __shared__ float red[S][32];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.y;
  const int op = blockIdx.x * 32 + threadIdx.x;  // padded output pixel
  const bool live = op < Hp * Wp;
  const int opc = live ? op : 0;
  const int ho = opc / Wp, wo = opc - ho * Wp;
  const int hs = reflect(ho - pad, H), ws = reflect(wo - pad, W);
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t pix = static_cast<size_t>(hs) * W + ws;
  const float* xp = x + static_cast<size_t>(b) * C * hw + pix;
  const float* gp = gb + static_cast<size_t>(b) * 2 * C * hw + pix;
  float sum = 0.f;
  for (int c = threadIdx.y; c < C; c += S) sum += xp[c * hw];
  const float mean = slice_reduce<S>(sum, red) / C;
  float ss = 0.f;
  for (int c = threadIdx.y; c < C; c += S) {
    const float d = xp[c * hw] - mean;
    ss = fmaf(d, d, ss);
  }
  const float rstd = rsqrtf(slice_reduce<S>(ss, red) / (C - 1) + eps);
  if (!live) return;
  if (threadIdx.y == 0 && ho - pad == hs && wo - pad == ws) {  // interior pixel: owns the saved statistics
    mean_out[static_cast<size_t>(b) * hw + pix] = mean;
    rstd_out[static_cast<size_t>(b) * hw + pix] = rstd;
  }
  const size_t hwp = static_cast<size_t>(Hp) * Wp;
  float* yp = y + static_cast<size_t>(b) * C * hwp + op;
  for (int c = threadIdx.y; c < C; c += S) {
    const float xh = (xp[c * hw] - mean) * rstd;
    float z = fmaf(xh, 1.0f + gp[c * hw], gp[(C + c) * hw]);
    z = z > 0.f ? z : z * slope;
    yp[c * hwp] = z;
  }
}

// gradient of the padded output folded back onto source pixel (h, w)
__device__ __forceinline__ float folded_dy(const float* __restrict__ dyc, int h, int w, int H, int W, int pad,
                                           int Wp) {
  // padded coordinates whose reflection source is h: h+pad, and the mirror images
  int hc[3], wc[3], nh = 0, nw = 0;
  hc[nh++] = h + pad;
  if (h >= 1 && h <= pad) hc[nh++] = pad - h;
  if (h <= H - 2 && h >= H - 1 - pad) hc[nh++] = 2 * (H - 1) - h + pad;
  wc[nw++] = w + pad;
  if (w >= 1 && w <= pad) wc[nw++] = pad - w;
  if (w <= W - 2 && w >= W - 1 - pad) wc[nw++] = 2 * (W - 1) - w + pad;
  float acc = 0.f;
  for (int i = 0; i < nh; ++i)
    for (int j = 0; j < nw; ++j) acc += dyc[static_cast<size_t>(hc[i]) * Wp + wc[j]];
  return acc;
}

template <int S>
__global__ void __launch_bounds__(32 * S)
spade_mod_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gb,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, float* __restrict__ dx,
                     float* __restrict__ dgb, int C, int H, int W, int pad, float slope) {
  __shared__ float red[S][32];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.y;
  const size_t hw = static_cast<size_t>(H) * W;
  const int pixi = blockIdx.x * 32 + threadIdx.x;
  const bool live = pixi < static_cast<int>(hw);
  const int pix = live ? pixi : 0;
  const int h = pix / W, w = pix - h * W;
  const size_t hwp = static_cast<size_t>(Hp) * Wp;
  const float* xp = x + static_cast<size_t>(b) * C * hw + pix;
  const float* gp = gb + static_cast<size_t>(b) * 2 * C * hw + pix;
  const float* dyb = dy + static_cast<size_t>(b) * C * hwp;
  float* dxp = dx + static_cast<size_t>(b) * C * hw + pix;
  float* dgp = dgb + static_cast<size_t>(b) * 2 * C * hw + pix;
  const float mean = mean_in[static_cast<size_t>(b) * hw + pix];
  const float rstd = rstd_in[static_cast<size_t>(b) * hw + pix];
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.y; c < C; c += S) {
    const float xh = (xp[c * hw] - mean) * rstd;
    const float g1 = 1.0f + gp[c * hw];
    const float z = fmaf(xh, g1, gp[(C + c) * hw]);
    float dz = pad ? folded_dy(dyb + c * hwp, h, w, H, W, pad, Wp) : dyb[c * hwp + pix];
    dz = z > 0.f ? dz : dz * slope;
    const float dxh = dz * g1;
    if (live) {
      dgp[c * hw] = dz * xh;   // d gamma
      dgp[(C + c) * hw] = dz;  // d beta
      dxp[c * hw] = dxh;       // stash; finished in the second pass (same thread re-reads it)
    }
    s1 += dxh;
    s2 = fmaf(dxh, xh, s2);
  }
  const float m1 = slice_reduce<S>(s1, red) / C;
  const float m2 = slice_reduce<S>(s2, red) / (C - 1);
  if (!live) return;
  for (int c = threadIdx.y; c < C; c += S) {
    const float xh = (xp[c * hw] - mean) * rstd;
    dxp[c * hw] = rstd * (dxp[c * hw] - m1 - xh * m2);
  }
}

// ------------------------------------------------------------------ NHWC (channels_last) variants
// One warp per pixel: the channel vector of a pixel is contiguous, lanes stride over it with float4 accesses and
// the per-pixel sums are warp-shuffle reductions.  C % 4 == 0.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
spade_mod_fwd_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ gb, float* __restrict__ y,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out, int C, int H, int W, int pad,
                          float slope, float eps) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int op = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (op >= Hp * Wp) return;  // whole warp
  const int ho = op / Wp, wo = op - ho * Wp;
  const int hs = reflect(ho - pad, H), ws = reflect(wo - pad, W);
  const size_t spix = (static_cast<size_t>(b) * H + hs) * W + ws;
  const float4* xp = reinterpret_cast<const float4*>(x + spix * C);
  const float4* gp = reinterpret_cast<const float4*>(gb + spix * 2 * C);
  float4* yp = reinterpret_cast<float4*>(y + ((static_cast<size_t>(b) * Hp + ho) * Wp + wo) * C);
  const int n4 = C >> 2;
  float sum = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xp[i];
    sum += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(sum) / C;
  float ss = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xp[i];
    const float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
    ss += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) / (C - 1) + eps);
  if (lane == 0 && ho - pad == hs && wo - pad == ws) {
    mean_out[spix] = mean;
    rstd_out[spix] = rstd;
  }
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xp[i], g = gp[i], be = gp[n4 + i];
    float4 z;
    z.x = fmaf((v.x - mean) * rstd, 1.0f + g.x, be.x);
    z.y = fmaf((v.y - mean) * rstd, 1.0f + g.y, be.y);
    z.z = fmaf((v.z - mean) * rstd, 1.0f + g.z, be.z);
    z.w = fmaf((v.w - mean) * rstd, 1.0f + g.w, be.w);
    z.x = z.x > 0.f ? z.x : z.x * slope;
    z.y = z.y > 0.f ? z.y : z.y * slope;
    z.z = z.z > 0.f ? z.z : z.z * slope;
    z.w = z.w > 0.f ? z.w : z.w * slope;
    yp[i] = z;
  }
}

__global__ void __launch_bounds__(256)
spade_mod_bwd_nhwc_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gb,
                          const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                          float* __restrict__ dx, float* __restrict__ dgb, int C, int H, int W, int pad, float slope) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= H * W) return;
  const int h = pix / W, w = pix - h * W;
  const size_t spix = static_cast<size_t>(b) * H * W + pix;
  const int n4 = C >> 2;
  const float4* xp = reinterpret_cast<const float4*>(x + spix * C);
  const float4* gp = reinterpret_cast<const float4*>(gb + spix * 2 * C);
  float4* dxp = reinterpret_cast<float4*>(dx + spix * C);
  float4* dgp = reinterpret_cast<float4*>(dgb + spix * 2 * C);
  // padded positions whose reflection source is (h, w)
  int hc[3], wc[3], nh = 0, nw = 0;
  hc[nh++] = h + pad;
  if (h >= 1 && h <= pad) hc[nh++] = pad - h;
  if (h <= H - 2 && h >= H - 1 - pad) hc[nh++] = 2 * (H - 1) - h + pad;
  wc[nw++] = w + pad;
  if (w >= 1 && w <= pad) wc[nw++] = pad - w;
  if (w <= W - 2 && w >= W - 1 - pad) wc[nw++] = 2 * (W - 1) - w + pad;
  const float mean = mean_in[spix], rstd = rstd_in[spix];
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < n4; i += 32) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < nh; ++a)
      for (int c = 0; c < nw; ++c) {
        const float4 t =
            reinterpret_cast<const float4*>(dy + ((static_cast<size_t>(b) * Hp + hc[a]) * Wp + wc[c]) * C)[i];
        d.x += t.x; d.y += t.y; d.z += t.z; d.w += t.w;
      }
    const float4 v = xp[i], g = gp[i], be = gp[n4 + i];
    float4 xh, dg, dxh;
    xh.x = (v.x - mean) * rstd; xh.y = (v.y - mean) * rstd; xh.z = (v.z - mean) * rstd; xh.w = (v.w - mean) * rstd;
    d.x = fmaf(xh.x, 1.0f + g.x, be.x) > 0.f ? d.x : d.x * slope;
    d.y = fmaf(xh.y, 1.0f + g.y, be.y) > 0.f ? d.y : d.y * slope;
    d.z = fmaf(xh.z, 1.0f + g.z, be.z) > 0.f ? d.z : d.z * slope;
    d.w = fmaf(xh.w, 1.0f + g.w, be.w) > 0.f ? d.w : d.w * slope;
    dg.x = d.x * xh.x; dg.y = d.y * xh.y; dg.z = d.z * xh.z; dg.w = d.w * xh.w;
    dxh.x = d.x * (1.0f + g.x); dxh.y = d.y * (1.0f + g.y); dxh.z = d.z * (1.0f + g.z); dxh.w = d.w * (1.0f + g.w);
    dgp[i] = dg;       // d gamma
    dgp[n4 + i] = d;   // d beta
    dxp[i] = dxh;      // stash
    s1 += (dxh.x + dxh.y) + (dxh.z + dxh.w);
    s2 += (dxh.x * xh.x + dxh.y * xh.y) + (dxh.z * xh.z + dxh.w * xh.w);
  }
  const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / (C - 1);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xp[i];
    float4 t = dxp[i];
    t.x = rstd * (t.x - m1 - (v.x - mean) * rstd * m2);
    t.y = rstd * (t.y - m1 - (v.y - mean) * rstd * m2);
    t.z = rstd * (t.z - m1 - (v.z - mean) * rstd * m2);
    t.w = rstd * (t.w - m1 - (v.w - mean) * rstd * m2);
    dxp[i] = t;
  }
}

}  // namespace

int spade_mod_fwd_launch(const float* x, const float* gb, float* y, float* mean, float* rstd, int B, int C, int H,
                         int W, int pad, float slope, float eps, int nhwc, cudaStream_t stream) {
  if (B <= 0 || C < 2 || H <= pad || W <= pad || pad < 0 || (nhwc && (C % 4))) {
    set_error("spade_mod_fwd: bad shape (B=%d C=%d H=%d W=%d pad=%d nhwc=%d)", B, C, H, W, pad, nhwc);
    return -1;
  }
  const int npix = (H + 2 * pad) * (W + 2 * pad);
  if (nhwc) {
    spade_mod_fwd_nhwc_kernel<<<dim3((npix + 7) / 8, B), 256, 0, stream>>>(x, gb, y, mean, rstd, C, H, W, pad, slope,
                                                                           eps);
    COCOS_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const dim3 grid((npix + 31) / 32, B);
  if (C >= 256 && npix <= 64 * 64)
    spade_mod_fwd_kernel<32><<<grid, dim3(32, 32), 0, stream>>>(x, gb, y, mean, rstd, C, H, W, pad, slope, eps);
  else
    spade_mod_fwd_kernel<8><<<grid, dim3(32, 8), 0, stream>>>(x, gb, y, mean, rstd, C, H, W, pad, slope, eps);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int spade_mod_bwd_launch(const float* dy, const float* x, const float* gb, const float* mean, const float* rstd,
                         float* dx, float* dgb, int B, int C, int H, int W, int pad, float slope, int nhwc,
                         cudaStream_t stream) {
  if (B <= 0 || C < 2 || H <= pad || W <= pad || pad < 0 || (nhwc && (C % 4))) {
    set_error("spade_mod_bwd: bad shape (B=%d C=%d H=%d W=%d pad=%d nhwc=%d)", B, C, H, W, pad, nhwc);
    return -1;
  }
  if (nhwc) {
    spade_mod_bwd_nhwc_kernel<<<dim3((H * W + 7) / 8, B), 256, 0, stream>>>(dy, x, gb, mean, rstd, dx, dgb, C, H, W,
                                                                            pad, slope);
    COCOS_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  const dim3 grid((H * W + 31) / 32, B);
  if (C >= 256 && H * W <= 64 * 64)
    spade_mod_bwd_kernel<32><<<grid, dim3(32, 32), 0, stream>>>(dy, x, gb, mean, rstd, dx, dgb, C, H, W, pad, slope);
  else
    spade_mod_bwd_kernel<8><<<grid, dim3(32, 8), 0, stream>>>(dy, x, gb, mean, rstd, dx, dgb, C, H, W, pad, slope);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
