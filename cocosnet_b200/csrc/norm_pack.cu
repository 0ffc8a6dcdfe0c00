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
norm_pack_kernel(const float* __restrict__ xt, __half* __restrict__ out, int C, int h, int w, float eps) {
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

}  // namespace

int norm_pack_launch(const float* x, float* xt_workspace, void* out, int B, int C, int h, int w, int mk, float eps,
                     cudaStream_t stream) {
  if (B <= 0 || C <= 0 || (C % 4) || h <= 0 || w <= 0 || (mk != 1 && mk != 3) || ((C * mk * mk) % 64)) {
    set_error("norm_pack: need C %% 4 == 0, mk in {1,3}, C*mk*mk %% 64 == 0 (B=%d C=%d h=%d w=%d mk=%d)", B, C, h, w, mk);
    return -1;
  }
  const int N = h * w;
  transpose_cn_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B), 256, 0, stream>>>(x, xt_workspace, C, N);
  COCOS_CUDA_CHECK(cudaGetLastError());
  const dim3 grid((N + 7) / 8, B);
  if (mk == 1)
    norm_pack_kernel<1><<<grid, 256, 0, stream>>>(xt_workspace, static_cast<__half*>(out), C, h, w, eps);
  else
    norm_pack_kernel<3><<<grid, 256, 0, stream>>>(xt_workspace, static_cast<__half*>(out), C, h, w, eps);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
