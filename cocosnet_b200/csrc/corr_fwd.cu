// K1: fused correlation -> softmax -> warp for sm_100a.
//
//   O[b,c,i] = sum_j softmax_j(scale * <Q[b,i,:], K[b,j,:]>) * V[b,j,c]
//
// Replaces reference correspondence.py:291 (matmul theta^T phi), :304
// (/temperature), :307 (row softmax) and :318/:334 (matmul with the exemplar /
// its mask) without ever writing the HWxHW matrix to HBM.
//
// One CTA owns 128 query rows and streams 128-key tiles:
//   warp 4  : TMA producer  (Q once or chunk-streamed, K chunks, V tiles)
//   warp 5  : tcgen05.mma issuer (S = Q K^T into a double-buffered TMEM tile,
//             O += P V into a TMEM accumulator), TMEM owner
//   warps 0-3: online softmax, one thread per query row (TMEM lane): S from
//             TMEM -> registers, exp2, P as fp16 into 128B-swizzled smem (the A
//             operand of the second MMA), lazy rescale of O, final epilogue.
#include "corr_kernels.h"
#include "ptx.cuh"
#include "tmap.h"

#include <stdlib.h>

namespace cocos {

namespace {

constexpr int BM = 128;  // queries per CTA
constexpr int BN = 128;  // keys per tile
constexpr int BK = 64;   // fp16 elements per 128B swizzle atom row
constexpr int ATOM_BYTES = 128 * BK * 2;  // 16 KiB: 128 rows x 64 fp16
constexpr int NUM_THREADS = 192;
constexpr int MAX_KSTAGES = 8;
constexpr int MAX_VSTAGES = 2;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units (P stays <= 2^8)

struct FwdParams {
  int B, Nq, Nk, Kd, Cv, Cvp;
  int kc_count;    // Kd / 64
  int n_tiles;     // ceil(Nk / 128)
  int q_resident;  // Q tile kept in smem for the whole CTA
  int ns_k, ns_v;
  float scale, scale_log2;
  float* out;   // [B, Cv, Nq]
  float* lse;   // [B, Nq] (natural log) or null
  float* corr;  // [B, Nq, Nk] scaled logits or null (return_corr / debug)
};

struct Barriers {
  uint64_t q_full;
  uint64_t k_full[MAX_KSTAGES];
  uint64_t k_empty[MAX_KSTAGES];
  uint64_t v_full[MAX_VSTAGES];
  uint64_t v_empty[MAX_VSTAGES];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t p_full;
  uint64_t pv_done;
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * BM;
  const int bidx = blockIdx.y;

  const uint32_t stage_bytes = p.q_resident ? ATOM_BYTES : 2 * ATOM_BYTES;
  const uint32_t v_stage_bytes = static_cast<uint32_t>(p.Cvp) * 256u;  // 2 atoms of [Cvp x 64] fp16
  const uint32_t q_smem = smem0;
  const uint32_t k_ring = q_smem + (p.q_resident ? p.kc_count * ATOM_BYTES : 0);
  const uint32_t v_ring = k_ring + p.ns_k * stage_bytes;
  const uint32_t p_smem = v_ring + p.ns_v * v_stage_bytes;
  const uint32_t bar_off = p_smem + 2 * ATOM_BYTES - smem0;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_gen + bar_off);

  if (tid == 0) {
    mbar_init(smem_u32(&bars->q_full), 1);
    for (int i = 0; i < p.ns_k; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
    }
    for (int i = 0; i < p.ns_v; ++i) {
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_empty[i]), 128);
    }
    mbar_init(smem_u32(&bars->p_full), 128);
    mbar_init(smem_u32(&bars->pv_done), 1);
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tmem_o = tmem + 256;  // S0: [0,128) S1: [128,256) O: [256, 256+Cvp)
  const int T = p.n_tiles;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (p.q_resident) {
        mbar_expect_tx(smem_u32(&bars->q_full), p.kc_count * ATOM_BYTES);
        for (int kc = 0; kc < p.kc_count; ++kc)
          tma_load_3d(q_smem + kc * ATOM_BYTES, &tm_q, smem_u32(&bars->q_full), kc * BK, q0, bidx);
      }
      uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
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
          tma_load_3d(dst, &tm_k, full, kc * BK, j * BN, bidx);
          if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
        }
        mbar_wait(smem_u32(&bars->v_empty[vs]), vph ^ 1);
        const uint32_t vfull = smem_u32(&bars->v_full[vs]);
        mbar_expect_tx(vfull, v_stage_bytes);
        const uint32_t vdst = v_ring + vs * v_stage_bytes;
        tma_load_3d(vdst, &tm_v, vfull, j * BN, 0, bidx);
        tma_load_3d(vdst + v_stage_bytes / 2, &tm_v, vfull, j * BN + BK, 0, bidx);
        if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(BM, BN);
      const uint32_t idesc_pv = make_idesc_f16(BM, p.Cvp);
      uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
      if (p.q_resident) {
        mbar_wait(smem_u32(&bars->q_full), 0);
        tc_fence_after();
      }
      auto issue_s = [&](int t) {
        const int b = t & 1;
        if (t >= 2) {
          mbar_wait(smem_u32(&bars->s_empty[b]), ((t >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem + b * BN;
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          tc_fence_after();
          const uint32_t st = k_ring + ks * stage_bytes;
          const uint32_t a_addr = p.q_resident ? (q_smem + kc * ATOM_BYTES) : st;
          const uint32_t b_addr = p.q_resident ? st : (st + ATOM_BYTES);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            umma_f16(d_tmem, make_desc_k_sw128(a_addr + s4 * 32), make_desc_k_sw128(b_addr + s4 * 32), idesc_s,
                     (kc | s4) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bars->k_empty[ks]));
          if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
        }
        umma_commit(smem_u32(&bars->s_full[b]));
      };
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        mbar_wait(smem_u32(&bars->p_full), j & 1);
        mbar_wait(smem_u32(&bars->v_full[vs]), vph);
        tc_fence_after();
        const uint32_t vb = v_ring + vs * v_stage_bytes;
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const uint32_t a_addr = p_smem + (st >> 2) * ATOM_BYTES + (st & 3) * 32;
          const uint32_t b_addr = vb + (st >> 2) * (v_stage_bytes / 2) + (st & 3) * 32;
          umma_f16(tmem_o, make_desc_k_sw128(a_addr), make_desc_k_sw128(b_addr), idesc_pv,
                   (j | st) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars->v_empty[vs]));
        umma_commit(smem_u32(&bars->pv_done));
        if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------- softmax warps (0..3)
    const int row = tid;  // TMEM lane == query row within the tile
    const int q = q0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const float c2 = p.scale_log2;
    float m = -INFINITY, l = 0.f;
    const uint32_t row_off = p_smem + row * 128;
    const uint32_t sw = static_cast<uint32_t>(row & 7);

    for (int j = 0; j < T; ++j) {
      const int b = j & 1;
      mbar_wait(smem_u32(&bars->s_full[b]), (j >> 1) & 1);
      tc_fence_after();
      float s[BN];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_ld32(tmem + lane_sel + b * BN + c * 32, reinterpret_cast<uint32_t*>(&s[c * 32]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_empty[b]));

      if (p.corr != nullptr && q < p.Nq) {
        float* dst = p.corr + (static_cast<size_t>(bidx) * p.Nq + q) * p.Nk + j * BN;
        const int valid = p.Nk - j * BN;
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (c < valid) dst[c] = s[c] * p.scale;
      }
      if (j == T - 1 && (p.Nk & (BN - 1)) != 0) {
        const int valid = p.Nk - j * BN;
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      float tmax = s[0];
#pragma unroll
      for (int c = 1; c < BN; ++c) tmax = fmaxf(tmax, s[c]);
      const float m_new = fmaxf(m, tmax);
      if (j == 0) {
        m = m_new;
      } else {
        const bool resc = (m_new - m) * c2 > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, resc)) {
          const float alpha = resc ? ex2((m - m_new) * c2) : 1.0f;
          if (resc) m = m_new;
          l *= alpha;
          mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);
          tc_fence_after();
          for (int cc = 0; cc < p.Cvp; cc += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_sel + cc, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_o + lane_sel + cc, o);
          }
          tmem_wait_st();
          tc_fence_before();
        }
      }
      const float mc = m * c2;
      float sum = 0.f;
      uint32_t pk[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        const float p0 = ex2(fmaf(s[2 * i], c2, -mc));
        const float p1 = ex2(fmaf(s[2 * i + 1], c2, -mc));
        sum += p0 + p1;
        pk[i] = pack_h2(p0, p1);
      }
      l += sum;
      if (j > 0) mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1);  // P buffer free again
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const uint32_t addr = row_off + (c >> 3) * ATOM_BYTES + (((c & 7) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                     "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                     : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bars->p_full));
    }

    // epilogue: O / l -> out[b, c, q]
    mbar_wait(smem_u32(&bars->pv_done), (T - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    for (int cc = 0; cc < p.Cvp; cc += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_o + lane_sel + cc, o);
      tmem_wait_ld();
      if (q < p.Nq) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = cc + i;
          if (c < p.Cv) p.out[(static_cast<size_t>(bidx) * p.Cv + c) * p.Nq + q] = __uint_as_float(o[i]) * inv_l;
        }
      }
    }
    if (p.lse != nullptr && q < p.Nq)
      p.lse[static_cast<size_t>(bidx) * p.Nq + q] = (m * c2 + log2f(l)) * 0.6931471805599453f;
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int corr_warp_fwd_launch(const void* q, const void* k, const void* vt, const float* v32, float* out, float* lse,
                         float* corr, int B, int Nq, int Nk, int Kd, int Cv, int Cvp, int Nkp, float scale,
                         cudaStream_t stream) {
  if (B <= 0 || Nq <= 0 || Nk <= 0) {
    set_error("corr_warp_fwd: empty problem (B=%d Nq=%d Nk=%d)", B, Nq, Nk);
    return -1;
  }
  if (Kd <= 0 || (Kd % BK) != 0) {
    set_error("corr_warp_fwd: Kd=%d must be a positive multiple of 64 (pad in the pack kernel)", Kd);
    return -1;
  }
  static const int variant = [] {
    // default: best kernel that applies; COCOS_K1_VARIANT=1..4 caps the choice (A/B timing)
    const char* e = getenv("COCOS_K1_VARIANT");
    return e ? atoi(e) : 4;
  }();
  if (variant >= 4 && v32 != nullptr && corr == nullptr && Cv >= 1 && Cv <= 4) {
    const int rc4 = corr_warp_fwd4_launch(q, k, v32, out, lse, B, Nq, Nk, Kd, Cv, scale, stream);
    if (rc4 != 1) return rc4;
  }
  if (vt == nullptr) {
    set_error("corr_warp_fwd: packed fp16 values (vt) are required for this shape (Cv=%d)", Cv);
    return -1;
  }
  if (Cv <= 0 || Cvp < Cv || (Cvp % 16) != 0 || Cvp > 256) {
    set_error("corr_warp_fwd: need 0 < Cv <= Cvp <= 256, Cvp %% 16 == 0 (Cv=%d Cvp=%d)", Cv, Cvp);
    return -1;
  }
  if (Nkp < Nk || (Nkp % 8) != 0) {
    set_error("corr_warp_fwd: V row pitch Nkp=%d must be >= Nk=%d and a multiple of 8", Nkp, Nk);
    return -1;
  }
  {
    // 256-key tiles (v3) when Cvp <= 64, else the two-warpgroup 128-key pipeline (v2), else the
    // single-warpgroup kernel below
    if (variant >= 3 && corr == nullptr && Cvp <= 64) {
      const int rc3 = corr_warp_fwd3_launch(q, k, vt, out, lse, B, Nq, Nk, Kd, Cv, Cvp, Nkp, scale, stream);
      if (rc3 != 1) return rc3;
    }
    if (variant >= 2 && corr == nullptr && Cvp <= 128) {
      const int rc2 = corr_warp_fwd2_launch(q, k, vt, out, lse, B, Nq, Nk, Kd, Cv, Cvp, Nkp, scale, stream);
      if (rc2 != 1) return rc2;
    }
  }
  FwdParams p;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.Kd = Kd; p.Cv = Cv; p.Cvp = Cvp;
  p.kc_count = Kd / BK;
  p.n_tiles = (Nk + BN - 1) / BN;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse; p.corr = corr;

  // shared memory plan (<= 227 KiB): [Q] [K ring] [V ring] [P] [barriers]
  const int budget = 227 * 1024 - 1024 /*alignment slack*/ - 512 /*barriers*/;
  const int p_bytes = 2 * ATOM_BYTES;
  const int v_stage = Cvp * 256;
  p.q_resident = (Kd <= 256) ? 1 : 0;
  const int q_bytes = p.q_resident ? p.kc_count * ATOM_BYTES : 0;
  const int stage = p.q_resident ? ATOM_BYTES : 2 * ATOM_BYTES;
  p.ns_v = 2;
  int rem = budget - p_bytes - q_bytes - p.ns_v * v_stage;
  if (rem / stage < 3) {
    p.ns_v = 1;
    rem = budget - p_bytes - q_bytes - v_stage;
  }
  p.ns_k = rem / stage;
  if (p.ns_k > MAX_KSTAGES) p.ns_k = MAX_KSTAGES;
  if (p.ns_k < 2) {
    set_error("corr_warp_fwd: shared memory plan failed (Kd=%d Cvp=%d)", Kd, Cvp);
    return -1;
  }
  const int smem_bytes = 1024 + q_bytes + p.ns_k * stage + p.ns_v * v_stage + p_bytes + 512;

  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  if ((rc = make_tmap_f16_3d(&tm_q, q, Kd, Nq, B, (uint64_t)Kd * 2, (uint64_t)Nq * Kd * 2, BK, BM, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_k, k, Kd, Nk, B, (uint64_t)Kd * 2, (uint64_t)Nk * Kd * 2, BK, BN, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_v, vt, Nk, Cvp, B, (uint64_t)Nkp * 2, (uint64_t)Cvp * Nkp * 2, BK, Cvp, 1)))
    return rc;

  COCOS_CUDA_CHECK(cudaFuncSetAttribute(corr_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid((Nq + BM - 1) / BM, B);
  corr_fwd_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
