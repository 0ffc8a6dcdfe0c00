// Fused operand prologue of the correspondence (HBM bound), PONO_C variant:
//   unfold(mk x mk, zero padded) -> subtract the mean over the K = C*mk*mk vector -> divide by (L2 norm + eps)
//   -> fp16, position-major [B, N, K]
// i.e. reference correspondence.py:273-281 (theta) / 283-289 (phi) for `--PONO_C` plus the operand rounding, in
// two launches instead of unfold + mean + sub + norm + div + permute over a [B, K, N] fp32 tensor (K = 2304:
// 37.7 MB per image per operand, several passes).  The K axis is written tap-major (k = tap*C + c): both operands
// use the same order, the dot products are unchanged.
//   step 1: [B,C,N] fp32 -> [B,N,C] fp32 (coalesced tiled transpose) so a position's channel vector is contiguous;
//   step 2: one warp per output position gathers its mk*mk neighbour vectors (L2 resident), reduces, writes K fp16.
// Algorithmic bytes per position: 4*C read + 2*K written (+ the 8*C transpose round trip).
#include "corr_kernels.h"
#include "tmap.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace cocos {

namespace {

__global__ void __launch_bounds__(256)
transpose_cn_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int N) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + static_cast<size_t>(b) * C * N;
  float* d = dst + static_cast<size_t>(b) * C * N;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, n = n0 + tx;
    tile[ty + i][tx] = (c < C && n < N) ? s[static_cast<size_t>(c) * N + n] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int n = n0 + ty + i, c = c0 + tx;
    if (n < N && c < C) d[static_cast<size_t>(n) * C + c] = tile[tx][ty + i];
  }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// xt: [B, N=h*w, C] fp32; out: [B, N, C*mk*mk] fp16 (tap-major); C % 4 == 0
template <int MK>
__global__ void __launch_bounds__(256)
norm_pack_kernel(const float* __restrict__ xt, __half* __restrict__ out, float* __restrict__ mean_out,
                 float* __restrict__ inv_out, int C, int h, int w, float eps) {
  constexpr int TAPS = MK * MK;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= h * w) return;
  const int i = n / w, j = n - i * w;
  const int K = C * TAPS, n4 = C >> 2;
  const float* base = xt + static_cast<size_t>(b) * h * w * C;
  const float4* nb[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    const int ii = i + t / MK - MK / 2, jj = j + t % MK - MK / 2;
    nb[t] = (ii >= 0 && ii < h && jj >= 0 && jj < w)
                ? reinterpret_cast<const float4*>(base + (static_cast<size_t>(ii) * w + jj) * C)
                : nullptr;
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < TAPS; ++t)
    if (nb[t])
      for (int q = lane; q < n4; q += 32) {
        const float4 v = nb[t][q];
        s += (v.x + v.y) + (v.z + v.w);
      }
  const float mean = warp_sum_f(s) / K;  // zero-padded taps count in K (F.unfold pads before the mean)
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    if (nb[t]) {
      for (int q = lane; q < n4; q += 32) {
        const float4 v = nb[t][q];
        const float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
        ss += (a * a + bb * bb) + (c * c + d * d);
      }
    } else {
      ss += (lane < 1 ? 1.f : 0.f) * C * mean * mean;  // C entries equal to (0 - mean)
    }
  }
  const float inv = 1.0f / (sqrtf(warp_sum_f(ss)) + eps);
  if (mean_out && lane == 0) {  // saved for the backward (norm_pack_bwd_*)
    mean_out[static_cast<size_t>(b) * h * w + n] = mean;
    inv_out[static_cast<size_t>(b) * h * w + n] = inv;
  }
  __half* o = out + (static_cast<size_t>(b) * h * w + n) * K;
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    for (int q = lane; q < n4; q += 32) {
      float4 v = nb[t] ? nb[t][q] : make_float4(0.f, 0.f, 0.f, 0.f);
      const __half2 lo = __floats2half2_rn((v.x - mean) * inv, (v.y - mean) * inv);
      const __half2 hi = __floats2half2_rn((v.z - mean) * inv, (v.w - mean) * inv);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(o + t * C + 4 * q) = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// With g = dL/d(fhat) [B, K, N] fp32 (k = tap*C + c, what the dS GEMM of the correspondence backward emits), per
// position n:  a_n = <fhat_n, g_n>,  s_n = sum_k g_nk,  and
//   dL/df_nk = (g_nk - fhat_nk * a_n - s_n / K) * inv_n          (normalise + centre; sum_k fhat_nk = 0)
//   dL/dx[c, i', j'] = sum over taps t with n' = (i', j') - offset(t) inside the map of dL/df_{n', (t, c)}
// where fhat_{n',(t,c)} = (x[c, i', j'] - mean_n') * inv_n' only needs x at the OUTPUT pixel: the fold (col2im) is a
// 9-term gather and the [B, K, N] unfolded tensors of the reference (correspondence.py:273-289) never exist.
//
// pass 1: a_n, s_n.  Block = 32 consecutive positions (lanes) x 8 warps splitting K; coalesced reads of g.
template <int MK>
__global__ void __launch_bounds__(256)
norm_pack_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                           const float* __restrict__ inv, float* __restrict__ a_out, float* __restrict__ s_out, int C,
                           int h, int w) {
  constexpr int TAPS = MK * MK;
  __shared__ float red[2][8][32];
  const int N = h * w, b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool live = n < N;
  const int nc = live ? n : 0;
  const int i = nc / w, j = nc - i * w;
  const float mu = mean[static_cast<size_t>(b) * N + nc], iv = inv[static_cast<size_t>(b) * N + nc];
  const float* gb = g + static_cast<size_t>(b) * C * TAPS * N + nc;
  const float* xb = x + static_cast<size_t>(b) * C * N;
  float a = 0.f, s = 0.f;
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    const int ii = i + t / MK - MK / 2, jj = j + t % MK - MK / 2;
    const bool in = ii >= 0 && ii < h && jj >= 0 && jj < w;
    const float* xp = xb + (in ? ii * w + jj : 0);
    for (int c = wid; c < C; c += 8) {
      const float gv = live ? gb[static_cast<size_t>(t * C + c) * N] : 0.f;
      const float xv = in ? xp[static_cast<size_t>(c) * N] : 0.f;
      a = fmaf((xv - mu) * iv, gv, a);
      s += gv;
    }
  }
  red[0][wid][lane] = a;
  red[1][wid][lane] = s;
  __syncthreads();
  if (wid == 0 && live) {
    float ta = 0.f, ts = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { ta += red[0][k][lane]; ts += red[1][k][lane]; }
    a_out[static_cast<size_t>(b) * N + n] = ta;
    s_out[static_cast<size_t>(b) * N + n] = ts;
  }
}

// pass 2: one thread per (c, i', j'), consecutive threads along j' (coalesced g and dx).
template <int MK>
__global__ void __launch_bounds__(256)
norm_pack_bwd_fold_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                          const float* __restrict__ inv, const float* __restrict__ a_in, const float* __restrict__ s_in,
                          float* __restrict__ dx, int C, int h, int w) {
  constexpr int TAPS = MK * MK;
  const int N = h * w, b = blockIdx.z, c = blockIdx.y;
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const int i = n / w, j = n - i * w;
  const size_t bo = static_cast<size_t>(b) * N;
  const float xv = x[(static_cast<size_t>(b) * C + c) * N + n];
  const float invK = 1.0f / static_cast<float>(C * TAPS);
  const float* gb = g + static_cast<size_t>(b) * C * TAPS * N;
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    // the position whose tap t reads pixel (i, j)
    const int ii = i - (t / MK - MK / 2), jj = j - (t % MK - MK / 2);
    if (ii < 0 || ii >= h || jj < 0 || jj >= w) continue;
    const int np = ii * w + jj;
    const float iv = inv[bo + np];
    const float fh = (xv - mean[bo + np]) * iv;
    acc += (gb[static_cast<size_t>(t * C + c) * N + np] - fh * a_in[bo + np] - s_in[bo + np] * invK) * iv;
  }
  dx[(static_cast<size_t>(b) * C + c) * N + n] = acc;
}

// [B, N, K] fp16 -> [B, K, N] bf16: the channel-major GEMM operands of the correspondence backward from the packed
// forward operand (64 x 64 smem-tiled transpose, 2 elements per thread access)
__global__ void __launch_bounds__(256)
transpose_f16_bf16_kernel(const __half* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int K) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z, n0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  const __half* s = src + static_cast<size_t>(b) * N * K;
  __nv_bfloat16* d = dst + static_cast<size_t>(b) * N * K;
#pragma unroll
  for (int r = 0; r < 64; r += 4) {
    const int n = n0 + ty + r, k = k0 + tx;
    tile[ty + r][tx] = (n < N && k < K) ? __half2float(s[static_cast<size_t>(n) * K + k]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 64; r += 4) {
    const int k = k0 + ty + r, n = n0 + tx;
    if (k < K && n < N) d[static_cast<size_t>(k) * N + n] = __float2bfloat16_rn(tile[tx][ty + r]);
  }
}

}  // namespace

int norm_pack_bwd_launch(const float* g, const float* x, const float* mean, const float* inv, float* a_ws, float* s_ws,
                         float* dx, int B, int C, int h, int w, int mk, cudaStream_t stream) {
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || (mk != 1 && mk != 3)) {
    set_error("norm_pack_bwd: bad arguments (B=%d C=%d h=%d w=%d mk=%d)", B, C, h, w, mk);
    return -1;
  }
  const int N = h * w;
  const dim3 g1((N + 31) / 32, B), g2((N + 255) / 256, C, B);
  if (mk == 1) {
    norm_pack_bwd_stats_kernel<1><<<g1, 256, 0, stream>>>(g, x, mean, inv, a_ws, s_ws, C, h, w);
    norm_pack_bwd_fold_kernel<1><<<g2, 256, 0, stream>>>(g, x, mean, inv, a_ws, s_ws, dx, C, h, w);
  } else {
    norm_pack_bwd_stats_kernel<3><<<g1, 256, 0, stream>>>(g, x, mean, inv, a_ws, s_ws, C, h, w);
    norm_pack_bwd_fold_kernel<3><<<g2, 256, 0, stream>>>(g, x, mean, inv, a_ws, s_ws, dx, C, h, w);
  }
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int transpose_f16_bf16_launch(const void* src, void* dst, int B, int N, int K, cudaStream_t stream) {
  if (B <= 0 || N <= 0 || K <= 0) {
    set_error("transpose_f16_bf16: bad arguments (B=%d N=%d K=%d)", B, N, K);
    return -1;
  }
  transpose_f16_bf16_kernel<<<dim3((K + 63) / 64, (N + 63) / 64, B), 256, 0, stream>>>(
      static_cast<const __half*>(src), static_cast<__nv_bfloat16*>(dst), N, K);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int norm_pack_launch(const float* x, float* xt_workspace, void* out, float* mean_out, float* inv_out, int B, int C, int h,
                     int w, int mk, float eps, cudaStream_t stream) {
  if (B <= 0 || C <= 0 || (C % 4) || h <= 0 || w <= 0 || (mk != 1 && mk != 3) || ((C * mk * mk) % 64) ||
      ((mean_out == nullptr) != (inv_out == nullptr))) {
    set_error("norm_pack: need C %% 4 == 0, mk in {1,3}, C*mk*mk %% 64 == 0 (B=%d C=%d h=%d w=%d mk=%d)", B, C, h, w, mk);
    return -1;
  }
  const int N = h * w;
  transpose_cn_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B), 256, 0, stream>>>(x, xt_workspace, C, N);
  COCOS_CUDA_CHECK(cudaGetLastError());
  const dim3 grid((N + 7) / 8, B);
  if (mk == 1)
    norm_pack_kernel<1><<<grid, 256, 0, stream>>>(xt_workspace, static_cast<__half*>(out), mean_out, inv_out, C, h, w,
                                                  eps);
  else
    norm_pack_kernel<3><<<grid, 256, 0, stream>>>(xt_workspace, static_cast<__half*>(out), mean_out, inv_out, C, h, w,
                                                  eps);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
