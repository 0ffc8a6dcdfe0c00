// K2n: persistent TMA-fed tcgen05 implicit-GEMM "tap convolution" over 16-bit NHWC activations, sm_100a.
//
//   Y[b, h, w, n] = epilogue( sum_{g < ngroups} sum_{c < 64*kchunks}
//                             X[b, s*h + dh[g], s*w + dw[g], coff[g] + c] * Wt[n, (g*kchunks)*64 + c] )
//
// One kernel for every convolution of the SPADE generator / domain adaptor / residual blocks / PatchGAN / VGG19
// (reference architecture.py:31-33,73-74; normalization.py:112-120; correspondence.py:17-22,79-146;
// discriminator.py:92-115; generator.py:104-113), forward AND backward-data:
//   * forward KSxKS stride s: one group per filter tap, dh = r - pad (+ the producer's materialised halo);
//   * backward-data of a stride-1 conv: x = dY, flipped taps, Wt = W^T;
//   * backward-data of a stride-2 conv: one launch per input-pixel parity class (a stride-1 tap conv over dY with the
//     taps of matching parity), scattered to the class's pixels through (y_sh, y_sw, y_oh, y_ow);
//   * 2-term fp16 split operands ([hi | lo] along channels): three groups per tap (hi*Whi, lo*Whi, hi*Wlo) - the
//     host lays the weight rows out in the same order; `coff` selects the hi or lo half of x.
// Operands:
//   * A tile = ONE 5-D TMA box {64 ch, TW, 1, TH, TB} of the activation tensor seen as
//     [B, Hin/s, s (row parity), Win/s, s (col parity) * Ca]: its TB*TH*TW rows of 128 B land pixel after pixel with
//     the 128B swizzle, i.e. directly in the K-major operand layout of tcgen05.mma (no im2col, no strided gather:
//     a stride-2 tap is a plain tile of the parity-split view).  Out-of-image taps and channels >= Ca are TMA
//     zero fill (zero padding for free; reflection halos are materialised by the producing epilogue).
//   * B tile = weights [Cout, ngroups*kchunks*64] K-major, box {64, BN}.
//   * accumulators: TWO TMEM buffers of BN fp32 columns; the epilogue of tile i overlaps the main loop of tile i+1.
// Epilogue (4 warps, one thread per output pixel): + bias, + residual, activation, then store as fp16 / bf16 / fp32
// NHWC (optionally with a 1-pixel halo filled by reflection, optionally as a 2-term [hi | lo] split) or fp32 NCHW.
// warp 4: TMA producer, warp 5: MMA issuer (UMMA M=128, N=BN), warps 0-3: epilogue.  Grid = min(#tiles, #SMs).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/cocos_b200.h"
#include "corr_kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace cocos {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int ATOM_BYTES = 128 * BK * 2;
constexpr int NUM_THREADS = 192;
constexpr int MAXG = COCOS_TAPCONV_MAX_GROUPS;

struct TapParams {
  int B, H, W, Cout;
  int TB, TH, TW, tiles_b, tiles_h, tiles_w, tiles_n, num_tiles;
  int a_stride, Ca, bf16;
  int ngroups, kchunks;
  int8_t dh[MAXG], dw[MAXG];
  int16_t coff[MAXG];
  // epilogue
  const float* scale;
  const float* bias;
  const void* res;
  int res_kind, res_Cs;
  int act;
  float slope;
  void* y;
  int y_kind, y_H, y_W, y_Cs, y_coff, y_lo_off, y_pad, y_reflect;
  int y_sh, y_sw, y_oh, y_ow;
  // SPADE modulation epilogue (mod_W > 0): an N tile holds [gamma of mod_W channels | beta of the same channels]
  int mod_W;
  const void* mod_x; int mod_x_kind, mod_x_Cs;
  const float* mod_mean; const float* mod_rstd;
  void* gb; int gb_kind, gb_Cs;
};

struct TapBars {
  uint64_t full[6];
  uint64_t empty[6];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ int floordiv2(int v) { return v >> 1; }  // arithmetic shift: floor for negatives

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : v * slope;
  if (act == 3) return tanhf(v);
  return v;
}

__device__ __forceinline__ uint32_t pack2(float a, float b, bool bf16) {
  if (bf16) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
  }
  const __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

__device__ __forceinline__ float lo_of(float v) { return v - __half2float(__float2half_rn(v)); }

// 32 consecutive channels of one pixel -> 16-bit NHWC at `dst` (16-byte aligned when `vec`), n_ok valid channels
__device__ __forceinline__ void store16(uint16_t* dst, const float* v, int n_ok, bool vec, bool bf16) {
  if (vec && n_ok == 32) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d4[q] = make_uint4(pack2(v[8 * q], v[8 * q + 1], bf16), pack2(v[8 * q + 2], v[8 * q + 3], bf16),
                         pack2(v[8 * q + 4], v[8 * q + 5], bf16), pack2(v[8 * q + 6], v[8 * q + 7], bf16));
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n_ok)
        dst[i] = bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v[i])) : __half_as_ushort(__float2half_rn(v[i]));
  }
}

__device__ __forceinline__ void store32(float* dst, const float* v, int n_ok, bool vec) {
  if (vec && n_ok == 32) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int q = 0; q < 8; ++q) d4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < n_ok) dst[i] = v[i];
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tapconv_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ TapParams p) {
  constexpr int STAGES = (BN == 128) ? 6 : 4;
  constexpr int STAGE_BYTES = ATOM_BYTES + (BN / 128) * ATOM_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
  TapBars* bars = reinterpret_cast<TapBars*>(smem_gen + STAGES * STAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int iters = p.ngroups * p.kchunks;
  const uint32_t a_bytes = static_cast<uint32_t>(p.TB * p.TH * p.TW) * 128u;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->acc_full[i]), 1);
      mbar_init(smem_u32(&bars->acc_empty[i]), 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&bars->tmem_base), 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int t = tile;
        const int n_i = t % p.tiles_n; t /= p.tiles_n;
        const int tw_i = t % p.tiles_w; t /= p.tiles_w;
        const int th_i = t % p.tiles_h;
        const int tb_i = t / p.tiles_h;
        const int h0 = th_i * p.TH, w0 = tw_i * p.TW, b0 = tb_i * p.TB, n0 = n_i * BN;
        for (int g = 0; g < p.ngroups; ++g) {
          int dh = p.dh[g], dw = p.dw[g], hpar = 0, cbase = p.coff[g];
          if (p.a_stride == 2) {
            hpar = dh & 1;
            dh = floordiv2(dh);
            cbase += (dw & 1) * p.Ca;
            dw = floordiv2(dw);
          }
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
            const uint32_t full = smem_u32(&bars->full[st]);
            mbar_expect_tx(full, a_bytes + (BN / 128) * ATOM_BYTES);
            tma_load_5d(smem0 + st * STAGE_BYTES, &tm_x, full, cbase + kc * BK, w0 + dw, hpar, h0 + dh, b0);
            tma_load_3d(smem0 + st * STAGE_BYTES + ATOM_BYTES, &tm_w, full, (g * p.kchunks + kc) * BK, n0, 0);
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(BM, BN, p.bf16 != 0);
    uint32_t st = 0, ph = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      mbar_wait(smem_u32(&bars->acc_empty[buf]), ((local >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * BN;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(smem_u32(&bars->full[st]), ph);
        tc_fence_after();
        if (leader) {
          const uint32_t a_addr = smem0 + st * STAGE_BYTES;
          const uint64_t da = make_desc_k_sw128(a_addr), db = make_desc_k_sw128(a_addr + ATOM_BYTES);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4)
            umma_f16(acc, desc_advance_k16(da, s4), desc_advance_k16(db, s4), idesc, (it | s4) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&bars->empty[st]));
          if (it == iters - 1) umma_commit(smem_u32(&bars->acc_full[buf]));
        }
        __syncwarp();
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (one thread per output pixel)
    const int m = tid;
    const int rows_per_img = p.TH * p.TW;
    const int tb = m / rows_per_img, rem = m - tb * rows_per_img;
    const int th = rem / p.TW, tw = rem - th * p.TW;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const int yHp = p.y_H + 2 * p.y_pad, yWp = p.y_W + 2 * p.y_pad;
    const bool out16 = p.y_kind == 1 || p.y_kind == 2;
    const bool obf = p.y_kind == 2;
    const float scale = p.scale ? __ldg(p.scale) : 1.0f;
    int local = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
      int t = tile;
      const int n_i = t % p.tiles_n; t /= p.tiles_n;
      const int tw_i = t % p.tiles_w; t /= p.tiles_w;
      const int th_i = t % p.tiles_h;
      const int tb_i = t / p.tiles_h;
      const int b = tb_i * p.TB + tb, h = th_i * p.TH + th, w = tw_i * p.TW + tw;
      const int n0 = n_i * BN;
      const bool ok = (tb < p.TB) && b < p.B && h < p.H && w < p.W;
      const int yh = h * p.y_sh + p.y_oh, yw = w * p.y_sw + p.y_ow;
      // store targets: the pixel itself, plus its mirror images in the reflection halo
      int rt[3], ct[3], nr = 0, nc = 0;
      rt[nr++] = yh + p.y_pad;
      ct[nc++] = yw + p.y_pad;
      if (p.y_reflect && p.y_pad == 1) {
        if (yh == 1) rt[nr++] = 0;
        if (yh == p.y_H - 2) rt[nr++] = p.y_H + 1;
        if (yw == 1) ct[nc++] = 0;
        if (yw == p.y_W - 2) ct[nc++] = p.y_W + 1;
      }
      const int buf = local & 1;
      mbar_wait(smem_u32(&bars->acc_full[buf]), (local >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * BN + lane_sel;
      if (p.mod_W) {
        // ---- SPADE: y = reflect_pad(lrelu(PONO(x) * (1 + gamma) + beta)) straight from the gamma / beta accumulators
        // (normalization.py:132-149 + architecture.py:73-74); the raw [gamma | beta] pair is stored too when the
        // backward will need it.  Tile n_i covers channels [n_i * mod_W, (n_i + 1) * mod_W) of x / y.
        const size_t spix = (static_cast<size_t>(b) * p.y_H + yh) * p.y_W + yw;
        float mean = 0.f, rstd = 0.f;
        if (ok) { mean = __ldg(p.mod_mean + spix); rstd = __ldg(p.mod_rstd + spix); }
#pragma unroll 1
        for (int c = 0; c < p.mod_W / 32; ++c) {
          uint32_t rg[32], rb[32];
          tmem_ld32(acc + c * 32, rg);
          tmem_ld32(acc + p.mod_W + c * 32, rb);
          tmem_wait_ld();
          if (!ok) continue;
          const int ch = n_i * p.mod_W + c * 32;        // channel of x / y
          const int ng = n0 + c * 32, nbt = ng + p.mod_W;  // columns (= rows of the interleaved weight / bias)
          float v[32], g[32], be[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            g[i] = __uint_as_float(rg[i]) * scale;
            be[i] = __uint_as_float(rb[i]) * scale;
          }
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              g[i] += __ldg(p.bias + ng + i);
              be[i] += __ldg(p.bias + nbt + i);
            }
          }
          if (p.gb) {
            const size_t go = spix * p.gb_Cs;
            if (p.gb_kind == 3) {
              store32(static_cast<float*>(p.gb) + go + ng, g, 32, true);
              store32(static_cast<float*>(p.gb) + go + nbt, be, 32, true);
            } else {
              store16(static_cast<uint16_t*>(p.gb) + go + ng, g, 32, true, false);
              store16(static_cast<uint16_t*>(p.gb) + go + nbt, be, 32, true, false);
            }
          }
          if (p.mod_x_kind == 3) {
            const float4* xp = reinterpret_cast<const float4*>(static_cast<const float*>(p.mod_x) + spix * p.mod_x_Cs + ch);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 t = xp[q];
              v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
          } else {
            const uint4* xp = reinterpret_cast<const uint4*>(static_cast<const __half*>(p.mod_x) + spix * p.mod_x_Cs + ch);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 t = xp[q];
              const __half2* h2 = reinterpret_cast<const __half2*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                v[8 * q + 2 * e] = f.x; v[8 * q + 2 * e + 1] = f.y;
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float z = fmaf((v[i] - mean) * rstd, 1.0f + g[i], be[i]);
            v[i] = z > 0.f ? z : z * p.slope;
          }
          float lo[32];
          if (p.y_lo_off) {
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = lo_of(v[i]);
          }
          for (int a = 0; a < nr; ++a)
            for (int e = 0; e < nc; ++e) {
              const size_t po = ((static_cast<size_t>(b) * yHp + rt[a]) * yWp + ct[e]) * p.y_Cs + ch;
              store16(static_cast<uint16_t*>(p.y) + po, v, 32, true, false);
              if (p.y_lo_off) store16(static_cast<uint16_t*>(p.y) + po + p.y_lo_off, lo, 32, true, false);
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[buf]));
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nb = n0 + c * 32;
        if (nb >= p.Cout) break;  // uniform
        uint32_t r[32];
        tmem_ld32(acc + c * 32, r);
        tmem_wait_ld();
        if (!ok) continue;
        const int n_ok = min(32, p.Cout - nb);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * scale;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < n_ok) v[i] += __ldg(p.bias + nb + i);
        }
        if (p.res) {
          const size_t ro = ((static_cast<size_t>(b) * p.y_H + yh) * p.y_W + yw) * p.res_Cs + nb;
          if (p.res_kind == 3) {
            const float* rp = static_cast<const float*>(p.res) + ro;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < n_ok) v[i] += rp[i];
          } else {
            const __half* rp = static_cast<const __half*>(p.res) + ro;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < n_ok) v[i] += __half2float(rp[i]);
          }
        }
        if (p.act) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], p.act, p.slope);
        }
        if (p.y_kind == 0) {  // fp32 NCHW
          const size_t hw = static_cast<size_t>(p.y_H) * p.y_W;
          float* yb = static_cast<float*>(p.y) + (static_cast<size_t>(b) * p.y_Cs + p.y_coff + nb) * hw +
                      static_cast<size_t>(yh) * p.y_W + yw;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < n_ok) yb[static_cast<size_t>(i) * hw] = v[i];
        } else {
          const int ch = p.y_coff + nb;
          const bool vec = ((ch | p.y_Cs | p.y_lo_off) & 7) == 0;
          float lo[32];
          if (p.y_lo_off) {
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = lo_of(v[i]);
          }
          for (int a = 0; a < nr; ++a)
            for (int e = 0; e < nc; ++e) {
              const size_t po = ((static_cast<size_t>(b) * yHp + rt[a]) * yWp + ct[e]) * p.y_Cs + ch;
              if (out16) {
                store16(static_cast<uint16_t*>(p.y) + po, v, n_ok, vec, obf);
                if (p.y_lo_off) store16(static_cast<uint16_t*>(p.y) + po + p.y_lo_off, lo, n_ok, vec, false);
              } else {
                store32(static_cast<float*>(p.y) + po, v, n_ok, vec && ((ch | p.y_Cs) & 3) == 0);
              }
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[buf]));
    }
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 2 * BN);
  }
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace

// Tile shape of the M dimension: TB images x TH rows x TW columns <= 128 pixels, minimising the number of tiles.
void tapconv_tile_shape(int B, int H, int W, int* TB, int* TH, int* TW) {
  if (H * W <= 64) {  // tiny maps (8x8): several images per tile
    *TW = W; *TH = H;
    int tb = 128 / (H * W);
    if (tb > B) tb = B;
    *TB = tb;
    return;
  }
  *TB = 1;
  long long best = -1;
  for (int tw = (W < 128 ? W : 128); tw >= 1; --tw) {
    int th = 128 / tw;
    if (th > H) th = H;
    const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    if (best < 0 || tiles < best) { best = tiles; *TW = tw; *TH = th; }
  }
}

int tapconv_launch(const cocos_tapconv_desc* d, cudaStream_t stream) {
  if (!d || !d->x || !d->w || !d->y) {
    set_error("tapconv: null pointer argument");
    return -1;
  }
  const int s = d->a_stride;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Hin <= 0 || d->Win <= 0 || d->Ca <= 0 || (d->Ca % 8) || d->Cout <= 0 ||
      (s != 1 && s != 2) || (s == 2 && ((d->Hin | d->Win) & 1)) || d->ngroups <= 0 || d->ngroups > MAXG ||
      d->kchunks <= 0 || d->y_kind < 0 || d->y_kind > 3 || d->y_pad < 0 || d->y_pad > 1 || d->y_Cs <= 0 ||
      d->y_sh <= 0 || d->y_sw <= 0 || d->act < 0 || d->act > 3 || (d->res && d->res_kind != 1 && d->res_kind != 3) ||
      (d->res && (d->y_sh != 1 || d->y_sw != 1 || d->y_oh || d->y_ow)) || (d->y_lo_off && d->y_kind != 1) ||
      (d->y_kind == 0 && (d->y_pad || d->y_lo_off))) {
    set_error("tapconv: bad descriptor (B=%d H=%d W=%d Hin=%d Win=%d Ca=%d Cout=%d stride=%d groups=%d kchunks=%d "
              "y_kind=%d y_pad=%d)", d->B, d->H, d->W, d->Hin, d->Win, d->Ca, d->Cout, s, d->ngroups, d->kchunks,
              d->y_kind, d->y_pad);
    return -1;
  }
  TapParams p;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
  tapconv_tile_shape(d->B, d->H, d->W, &p.TB, &p.TH, &p.TW);
  p.tiles_b = (d->B + p.TB - 1) / p.TB;
  p.tiles_h = (d->H + p.TH - 1) / p.TH;
  p.tiles_w = (d->W + p.TW - 1) / p.TW;
  const int BN = d->mod_W ? 2 * d->mod_W : (d->Cout > 128 ? 256 : 128);
  p.tiles_n = (d->Cout + BN - 1) / BN;
  p.num_tiles = p.tiles_b * p.tiles_h * p.tiles_w * p.tiles_n;
  p.a_stride = s; p.Ca = d->Ca; p.bf16 = d->bf16;
  p.ngroups = d->ngroups; p.kchunks = d->kchunks;
  for (int g = 0; g < d->ngroups; ++g) {
    p.dh[g] = d->dh[g]; p.dw[g] = d->dw[g]; p.coff[g] = d->coff[g];
    if (d->coff[g] < 0 || (d->coff[g] % 8)) {
      set_error("tapconv: channel offset of group %d (%d) must be a non-negative multiple of 8", g, d->coff[g]);
      return -1;
    }
  }
  p.scale = d->scale; p.bias = d->bias; p.res = d->res; p.res_kind = d->res_kind; p.res_Cs = d->res_Cs;
  p.act = d->act; p.slope = d->slope;
  p.y = d->y; p.y_kind = d->y_kind; p.y_H = d->y_H; p.y_W = d->y_W; p.y_Cs = d->y_Cs; p.y_coff = d->y_coff;
  p.y_lo_off = d->y_lo_off; p.y_pad = d->y_pad; p.y_reflect = d->y_reflect;
  p.y_sh = d->y_sh; p.y_sw = d->y_sw; p.y_oh = d->y_oh; p.y_ow = d->y_ow;
  p.mod_W = d->mod_W; p.mod_x = d->mod_x; p.mod_x_kind = d->mod_x_kind; p.mod_x_Cs = d->mod_x_Cs;
  p.mod_mean = d->mod_mean; p.mod_rstd = d->mod_rstd; p.gb = d->gb; p.gb_kind = d->gb_kind; p.gb_Cs = d->gb_Cs;
  if (d->mod_W) {
    const int C = d->Cout / 2;
    if ((d->mod_W != 64 && d->mod_W != 128) || (d->Cout % (2 * d->mod_W)) || (d->mod_W == 64 && C != 64) || !d->mod_x ||
        !d->mod_mean || !d->mod_rstd || (d->mod_x_kind != 1 && d->mod_x_kind != 3) || (d->mod_x_Cs % 8) || d->y_kind != 1 ||
        (d->y_Cs % 8) || (d->y_lo_off % 8) || d->y_coff || d->res || d->act != 2 || d->y_sh != 1 || d->y_sw != 1 ||
        d->y_oh || d->y_ow || d->H != d->y_H || d->W != d->y_W ||
        (d->gb && ((d->gb_kind != 1 && d->gb_kind != 3) || (d->gb_Cs % 8) || d->gb_Cs < d->Cout))) {
      set_error("tapconv: bad SPADE-epilogue descriptor (Cout=%d mod_W=%d x_kind=%d y_kind=%d act=%d)", d->Cout, d->mod_W,
                d->mod_x_kind, d->y_kind, d->act);
      return -1;
    }
  }
  if ((d->H - 1) * d->y_sh + d->y_oh >= d->y_H || (d->W - 1) * d->y_sw + d->y_ow >= d->y_W) {
    set_error("tapconv: output pixels fall outside y (%dx%d into %dx%d)", d->H, d->W, d->y_H, d->y_W);
    return -1;
  }
  CUtensorMap tm_x, tm_w;
  int rc;
  {
    // [B, Hin/s, s, Win/s, s*Ca]: innermost = (column parity, channel)
    const uint64_t Ca = (uint64_t)d->Ca, Win = (uint64_t)d->Win, Hin = (uint64_t)d->Hin;
    const uint64_t dims[5] = {Ca * s, Win / s, (uint64_t)s, Hin / s, (uint64_t)d->B};
    const uint64_t pitches[4] = {Ca * s * 2, Win * Ca * 2, Win * Ca * s * 2, Hin * Win * Ca * 2};
    const uint32_t box[5] = {BK, (uint32_t)p.TW, 1, (uint32_t)p.TH, (uint32_t)p.TB};
    if ((rc = make_tmap_16_5d(&tm_x, d->x, dims, pitches, box))) return rc;
  }
  const uint64_t Kt = (uint64_t)d->ngroups * d->kchunks * BK;
  const uint64_t wrows = (uint64_t)(d->w_rows >= d->Cout ? d->w_rows : d->Cout);
  if ((rc = make_tmap_f16_3d(&tm_w, d->w, Kt, wrows, 1, Kt * 2, wrows * Kt * 2, BK, BN, 1))) return rc;
  const int smem_bytes = 1024 + 196608 + 256;
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (BN == 256) {
    COCOS_CUDA_CHECK(cudaFuncSetAttribute(tapconv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    tapconv_kernel<256><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_x, tm_w, p);
  } else {
    COCOS_CUDA_CHECK(cudaFuncSetAttribute(tapconv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    tapconv_kernel<128><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_x, tm_w, p);
  }
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
