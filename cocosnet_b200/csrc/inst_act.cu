// Fused InstanceNorm2d(affine=False) + LeakyReLU, forward and backward (HBM bound).
//
//   y = lrelu_slope( (x - mean_hw) / sqrt(var_hw + eps) )      per (batch, channel) plane, biased variance
//
// replaces the nn.InstanceNorm2d -> nn.LeakyReLU(0.2) pairs of the domain adaptor
// (reference generator.py:104-113,141-145 with normalization.py:52-53) and of the
// PatchGAN discriminator (discriminator.py:92-115): batch_norm_collect_statistics +
// transform_input + leaky_relu (and their three backward kernels) become one kernel
// each way.  One CTA per plane, float4 accesses, two reductions.  slope = 1 gives the
// plain instance norm.  Algorithmic bytes per element: forward 4 + 4, backward 4 + 4 + 4.
#include "corr_kernels.h"
#include "tmap.h"

namespace cocos {

namespace {

constexpr int NT = 512;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < NT / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(NT)
inst_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ mean_out,
                    float* __restrict__ rstd_out, int HW, float slope, float eps) {
  __shared__ float red[32];
  const size_t plane = blockIdx.x;
  const float* xp = x + plane * HW;
  float* yp = y + plane * HW;
  const bool vec = (HW & 3) == 0;
  float s = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    for (int i = threadIdx.x; i < (HW >> 2); i += NT) {
      const float4 v = x4[i];
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += NT) s += xp[i];
  }
  const float mean = block_sum(s, red) / HW;
  float ss = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    for (int i = threadIdx.x; i < (HW >> 2); i += NT) {
      const float4 v = x4[i];
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += NT) {
      const float d = xp[i] - mean;
      ss = fmaf(d, d, ss);
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / HW + eps);
  if (threadIdx.x == 0) {
    mean_out[plane] = mean;
    rstd_out[plane] = rstd;
  }
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    float4* y4 = reinterpret_cast<float4*>(yp);
    for (int i = threadIdx.x; i < (HW >> 2); i += NT) {
      float4 v = x4[i];
      v.x = (v.x - mean) * rstd; v.y = (v.y - mean) * rstd; v.z = (v.z - mean) * rstd; v.w = (v.w - mean) * rstd;
      v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
      v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
      y4[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < HW; i += NT) {
      float v = (xp[i] - mean) * rstd;
      yp[i] = v > 0.f ? v : v * slope;
    }
  }
}

__global__ void __launch_bounds__(NT)
inst_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
                    const float* __restrict__ rstd_in, float* __restrict__ dx, int HW, float slope) {
  __shared__ float red[32];
  const size_t plane = blockIdx.x;
  const float* xp = x + plane * HW;
  const float* dyp = dy + plane * HW;
  float* dxp = dx + plane * HW;
  const float mean = mean_in[plane], rstd = rstd_in[plane];
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < HW; i += NT) {
    const float xh = (xp[i] - mean) * rstd;
    const float dz = xh > 0.f ? dyp[i] : dyp[i] * slope;
    s1 += dz;
    s2 = fmaf(dz, xh, s2);
  }
  const float m1 = block_sum(s1, red) / HW;
  const float m2 = block_sum(s2, red) / HW;
  for (int i = threadIdx.x; i < HW; i += NT) {
    const float xh = (xp[i] - mean) * rstd;
    const float dz = xh > 0.f ? dyp[i] : dyp[i] * slope;
    dxp[i] = rstd * (dz - m1 - xh * m2);
  }
}

}  // namespace

int inst_act_fwd_launch(const float* x, float* y, float* mean, float* rstd, int planes, int HW, float slope, float eps,
                        cudaStream_t stream) {
  if (planes <= 0 || HW <= 0) {
    set_error("inst_act_fwd: bad shape (planes=%d HW=%d)", planes, HW);
    return -1;
  }
  inst_act_fwd_kernel<<<planes, NT, 0, stream>>>(x, y, mean, rstd, HW, slope, eps);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int inst_act_bwd_launch(const float* dy, const float* x, const float* mean, const float* rstd, float* dx, int planes,
                        int HW, float slope, cudaStream_t stream) {
  if (planes <= 0 || HW <= 0) {
    set_error("inst_act_bwd: bad shape (planes=%d HW=%d)", planes, HW);
    return -1;
  }
  inst_act_bwd_kernel<<<planes, NT, 0, stream>>>(dy, x, mean, rstd, dx, HW, slope);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
