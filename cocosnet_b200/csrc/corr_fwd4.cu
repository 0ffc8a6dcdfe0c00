// K1 v4: the BASELINE shape (exemplar warp with Cv <= 4 value channels, e.g. the
// avg-pooled RGB exemplar of correspondence.py:313-318).
//
// With only 3 value channels the second product P V is 3 FMAs per exponential:
// cheaper on the CUDA cores, fused into the exp loop, than as a tensor-core
// pass (P -> fp16 -> TMEM/smem -> MMA padded to 16 columns -> commit ->
// barrier).  That removes P, O and every PV barrier from tensor memory and lets
// the 256-key S tile (UMMA N = 256, the shape that sustains ~75% of the 1-CTA
// tensor rate) be DOUBLE buffered (2 x 256 = all 512 TMEM columns): the S stream
// of tile t+1 runs under the softmax of tile t with no bubble.  P and V stay
// fp32 (no fp16 rounding of the probabilities: kernel-only error ~1e-6).
//
//   warps 0-3 : softmax+warp WG0 (keys [0,128) of every 256-key tile)   warp 8 : TMA Q / K
//   warps 4-7 : softmax+warp WG1 (keys [128,256))                       warp 9 : TMA V (fp32 rows)
//   warp 10   : issuer of S = Q K^T (N = 256), TMEM owner               warp 11: idle (warpgroup padding)
#include "corr_kernels.h"
#include "ptx.cuh"
#include "tmap.h"

#include <stdlib.h>

namespace cocos {

namespace {

constexpr int BM = 128, BN2 = 256, BK = 64;
constexpr int ATOM_BYTES = 128 * BK * 2;
constexpr int KSTAGE_BYTES = 2 * ATOM_BYTES;
constexpr int NUM_THREADS = 384;
constexpr int MAX_KSTAGES = 6;
constexpr int MAX_VSTAGES = 3;
constexpr int MAXC = 4;

struct Fwd4Params {
  int B, Nq, Nk, Kd, Cv;
  int kc_count, n_tiles, q_resident, ns_k, ns_v;
  float scale, scale_log2;
  float* out;
  float* lse;
};

struct Bars4 {
  uint64_t q_full;
  uint64_t k_full[MAX_KSTAGES];
  uint64_t k_empty[MAX_KSTAGES];
  uint64_t v_full[MAX_VSTAGES];
  uint64_t v_empty[MAX_VSTAGES];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float merge[128][MAXC + 2];  // WG1 -> WG0: m, l, o[0..3]
};

template <int CV>
__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_fwd4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Fwd4Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BM, bidx = blockIdx.y;

  const uint32_t stage_bytes = p.q_resident ? KSTAGE_BYTES : (ATOM_BYTES + KSTAGE_BYTES);
  constexpr uint32_t v_stage_bytes = CV * BN2 * 4;  // CV rows of 256 fp32
  const uint32_t q_smem = smem0;
  const uint32_t k_ring = q_smem + (p.q_resident ? p.kc_count * ATOM_BYTES : 0);
  const uint32_t v_ring = k_ring + p.ns_k * stage_bytes;
  const uint32_t bar_off = v_ring + p.ns_v * v_stage_bytes - smem0;
  Bars4* bars = reinterpret_cast<Bars4*>(smem_gen + bar_off);
  const float* v_gen = reinterpret_cast<const float*>(smem_gen + (v_ring - smem0));

  if (tid == 0) {
    mbar_init(smem_u32(&bars->q_full), 1);
    for (int i = 0; i < p.ns_k; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
    }
    for (int i = 0; i < p.ns_v; ++i) {
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 256);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_empty[i]), 256);
    }
    fence_mbar_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
  }
  if (warp == 9 && lane == 0) tma_prefetch_desc(&tm_v);
  if (warp == 10) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const int T = p.n_tiles;

  if (warp >= 8) {
    setmaxnreg_dec<48>();
    const bool leader = elect_one();
    if (warp == 8) {
      if (leader) {
        if (p.q_resident) {
          mbar_expect_tx(smem_u32(&bars->q_full), p.kc_count * ATOM_BYTES);
          for (int kc = 0; kc < p.kc_count; ++kc)
            tma_load_3d(q_smem + kc * ATOM_BYTES, &tm_q, smem_u32(&bars->q_full), kc * BK, q0, bidx);
        }
        uint32_t ks = 0, kph = 0;
        for (int j = 0; j < T; ++j) {
          for (int kc = 0; kc < p.kc_count; ++kc) {
            mbar_wait(smem_u32(&bars->k_empty[ks]), kph ^ 1);
            const uint32_t full = smem_u32(&bars->k_full[ks]);
            mbar_expect_tx(full, stage_bytes);
            uint32_t dst = k_ring + ks * stage_bytes;
            if (!p.q_resident) {
              tma_load_3d(dst, &tm_q, full, kc * BK, q0, bidx);
              dst += ATOM_BYTES;
            }
            tma_load_3d(dst, &tm_k, full, kc * BK, j * BN2, bidx);
            if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
          }
        }
      }
    } else if (warp == 9) {
      if (leader) {
        uint32_t vs = 0, vph = 0;
        for (int j = 0; j < T; ++j) {
          mbar_wait(smem_u32(&bars->v_empty[vs]), vph ^ 1);
          const uint32_t vfull = smem_u32(&bars->v_full[vs]);
          mbar_expect_tx(vfull, v_stage_bytes);
          tma_load_3d(v_ring + vs * v_stage_bytes, &tm_v, vfull, j * BN2, 0, bidx);  // box [256 keys x CV rows]
          if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == 10) {
      const uint32_t idesc_s = make_idesc_f16(BM, BN2);
      uint32_t ks = 0, kph = 0;
      if (p.q_resident) {
        mbar_wait(smem_u32(&bars->q_full), 0);
        tc_fence_after();
      }
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        if (t >= 2) mbar_wait(smem_u32(&bars->s_empty[b]), ((t >> 1) - 1) & 1);
        tc_fence_after();
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          tc_fence_after();
          if (leader) {
            const uint32_t st = k_ring + ks * stage_bytes;
            const uint64_t da = make_desc_k_sw128(p.q_resident ? (q_smem + kc * ATOM_BYTES) : st);
            const uint64_t db = make_desc_k_sw128(p.q_resident ? st : (st + ATOM_BYTES));
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
              umma_f16(tmem + b * BN2, desc_advance_k16(da, s4), desc_advance_k16(db, s4), idesc_s,
                       (kc | s4) != 0 ? 1u : 0u);
            umma_commit(smem_u32(&bars->k_empty[ks]));
            if (kc == p.kc_count - 1) umma_commit(smem_u32(&bars->s_full[b]));
          }
          __syncwarp();
          if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------- softmax + warp warpgroups (warps 0..7)
    setmaxnreg_inc<224>();
    const int g = warp >> 2;
    const int row = tid & 127;
    const int q = q0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float c2 = p.scale_log2;
    float m = -INFINITY, l = 0.f;
    float o[CV];
#pragma unroll
    for (int c = 0; c < CV; ++c) o[c] = 0.f;
    uint32_t vs = 0, vph = 0;

    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      mbar_wait(smem_u32(&bars->s_full[b]), (t >> 1) & 1);
      tc_fence_after();
      float s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_ld32(tmem + b * BN2 + g * 128 + lane_sel + c * 32, reinterpret_cast<uint32_t*>(&s[c * 32]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_empty[b]));

      const int valid = p.Nk - (t * BN2 + g * 128);
      mbar_wait(smem_u32(&bars->v_full[vs]), vph);
      if (valid > 0) {  // warp-uniform
        if (valid < 128) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= valid) s[c] = -INFINITY;
        }
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = fmaxf(s[i], s[i + 8]);
#pragma unroll
        for (int c = 16; c < 128; c += 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) mx[i] = fmaxf(mx[i], s[c + i]);
        }
        const float tmax = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                                 fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        const float m_new = fmaxf(m, tmax);       // finite: at least one valid key
        const float alpha = ex2((m - m_new) * c2);  // exact online softmax; 0 on the first tile (m = -inf)
        m = m_new;
        l *= alpha;
#pragma unroll
        for (int c = 0; c < CV; ++c) o[c] *= alpha;
        const float mc = m * c2;
        const float* vt = v_gen + vs * (CV * BN2) + g * 128;
        float sum = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 32; ++k4) {
          float pr[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            pr[i] = ex2(fmaf(s[4 * k4 + i], c2, -mc));
            sum += pr[i];
          }
#pragma unroll
          for (int c = 0; c < CV; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(vt + c * BN2 + 4 * k4);  // warp-wide broadcast
            o[c] = fmaf(pr[0], v.x, fmaf(pr[1], v.y, fmaf(pr[2], v.z, fmaf(pr[3], v.w, o[c]))));
          }
        }
        l += sum;
      }
      mbar_arrive(smem_u32(&bars->v_empty[vs]));
      if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
    }

    // ---- merge the two column halves; WG0 writes the result
    if (g == 1) {
      bars->merge[row][0] = m;
      bars->merge[row][1] = l;
#pragma unroll
      for (int c = 0; c < CV; ++c) bars->merge[row][2 + c] = o[c];
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (g == 0 && q < p.Nq) {
      const float m1 = bars->merge[row][0], l1 = bars->merge[row][1];
      const float mm = fmaxf(m, m1);
      const float a0 = ex2((m - mm) * c2);
      const float a1 = (m1 == -INFINITY) ? 0.f : ex2((m1 - mm) * c2);
      const float lt = l * a0 + l1 * a1;
      const float inv_l = 1.0f / lt;
#pragma unroll
      for (int c = 0; c < CV; ++c) {
        const float o1 = (m1 == -INFINITY) ? 0.f : bars->merge[row][2 + c];
        p.out[(static_cast<size_t>(bidx) * p.Cv + c) * p.Nq + q] = (o[c] * a0 + o1 * a1) * inv_l;
      }
      if (p.lse != nullptr) p.lse[static_cast<size_t>(bidx) * p.Nq + q] = (mm * c2 + log2f(lt)) * 0.6931471805599453f;
    }
  }

  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int CV>
int launch4(const CUtensorMap& tm_q, const CUtensorMap& tm_k, const CUtensorMap& tm_v, const Fwd4Params& p,
            int smem_bytes, cudaStream_t stream) {
  COCOS_CUDA_CHECK(
      cudaFuncSetAttribute(corr_fwd4_kernel<CV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid((p.Nq + BM - 1) / BM, p.B);
  corr_fwd4_kernel<CV><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

// v: fp32 [B, Cv, Nk] channel-major, the reference's own layout (no packing).  Returns 1 if this variant does not apply.
int corr_warp_fwd4_launch(const void* q, const void* k, const float* v, float* out, float* lse, int B, int Nq, int Nk,
                          int Kd, int Cv, float scale, cudaStream_t stream) {
  if (Cv < 1 || Cv > MAXC || (Nk % 4) != 0 || (reinterpret_cast<uintptr_t>(v) & 15)) return 1;
  Fwd4Params p;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.Kd = Kd; p.Cv = Cv;
  p.kc_count = Kd / BK;
  p.n_tiles = (Nk + BN2 - 1) / BN2;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse;
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(Bars4)) - 64;
  const int v_stage = Cv * BN2 * 4;
  p.q_resident = (Kd <= 256) ? 1 : 0;
  const int q_bytes = p.q_resident ? p.kc_count * ATOM_BYTES : 0;
  const int stage = p.q_resident ? KSTAGE_BYTES : (ATOM_BYTES + KSTAGE_BYTES);
  p.ns_v = 3;
  p.ns_k = (budget - q_bytes - p.ns_v * v_stage) / stage;
  if (p.ns_k > MAX_KSTAGES) p.ns_k = MAX_KSTAGES;
  if (p.ns_k < 2) return 1;
  const int smem_bytes = 1024 + q_bytes + p.ns_k * stage + p.ns_v * v_stage + sizeof(Bars4) + 64;

  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  if ((rc = make_tmap_f16_3d(&tm_q, q, Kd, Nq, B, (uint64_t)Kd * 2, (uint64_t)Nq * Kd * 2, BK, BM, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_k, k, Kd, Nk, B, (uint64_t)Kd * 2, (uint64_t)Nk * Kd * 2, BK, BN2, 1))) return rc;
  if ((rc = make_tmap_f32_3d_plain(&tm_v, v, Nk, Cv, B, (uint64_t)Nk * 4, (uint64_t)Cv * Nk * 4, BN2, Cv, 1)))
    return rc;
  switch (Cv) {
    case 1: return launch4<1>(tm_q, tm_k, tm_v, p, smem_bytes, stream);
    case 2: return launch4<2>(tm_q, tm_k, tm_v, p, smem_bytes, stream);
    case 3: return launch4<3>(tm_q, tm_k, tm_v, p, smem_bytes, stream);
    default: return launch4<4>(tm_q, tm_k, tm_v, p, smem_bytes, stream);
  }
}

}  // namespace cocos
