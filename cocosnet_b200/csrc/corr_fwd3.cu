// K1 v3: fused correlation -> softmax -> warp, 256-key S tiles (UMMA N = 256).
//
// Measured on B200 (profiles/r01_k1_timeline.md, tools/bench_gemm.py): a 1-CTA
// tcgen05.mma of shape M128 x N128 x K16 sustains only ~55% of the tensor peak
// (~116-160 clk per instruction instead of 64), while M128 x N256 x K16 reaches
// ~76% in the same pipeline.  So S = Q K^T is issued as N = 256 instructions over
// a 256-key tile; the two softmax warpgroups split the tile by COLUMNS
// (warpgroup g owns keys [g*128, g*128+128) of every tile), each with its own
// running (max, sum), its own P (kept in TENSOR memory, fed to the second MMA as
// the A operand) and its own O accumulator; the two partial states merge once at
// the end.  With P in TMEM the shared memory holds only Q, a 4-deep ring of
// 32 KiB K chunks and the V tiles.
//
//   warps 0-3 : softmax WG0 (left half of every key tile)    warp 8 : TMA Q / K
//   warps 4-7 : softmax WG1 (right half)                     warp 9 : TMA V
//   warp 10   : issuer of S = Q K^T (N=256), TMEM owner      warp 11: issuer of O_g += P_g V_g
// TMEM columns: S [0,256)  P0 [256,320)  P1 [320,384)  O0 [384,448)  O1 [448,512); Cvp <= 64.
#include "corr_kernels.h"
#include "ptx.cuh"
#include "tmap.h"

#include <stdlib.h>

namespace cocos {

namespace {

#define TRACE(role, tile, ev)                                                                      \
  do {                                                                                             \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (tile) < 64 && (threadIdx.x & 31) == 0)   \
      p.trace[((role) * 64 + (tile)) * 8 + (ev)] = clock64();                                      \
  } while (0)

constexpr int BM = 128;   // queries per CTA
constexpr int BN2 = 256;  // keys per S tile
constexpr int BK = 64;
constexpr int ATOM_BYTES = 128 * BK * 2;    // 16 KiB: 128 rows x 64 fp16
constexpr int KSTAGE_BYTES = 2 * ATOM_BYTES;  // 256 keys x 64 fp16
constexpr int NUM_THREADS = 384;
constexpr int MAX_KSTAGES = 6;
constexpr int MAX_VSTAGES = 3;
constexpr float RESCALE_THRESHOLD = 8.0f;

struct Fwd3Params {
  int B, Nq, Nk, Kd, Cv, Cvp;
  int kc_count, n_tiles, q_resident, ns_k, ns_v;
  float scale, scale_log2;
  float* out;
  float* lse;
  unsigned long long* trace;  // COCOS_K1_DBG=512: clock64 stamps of CTA (0,0)
};

struct Bars3 {
  uint64_t q_full;
  uint64_t k_full[MAX_KSTAGES];
  uint64_t k_empty[MAX_KSTAGES];
  uint64_t v_full[MAX_VSTAGES];
  uint64_t v_empty[MAX_VSTAGES];
  uint64_t s_full;
  uint64_t s_empty;
  uint64_t p_full[2];
  uint64_t pv_done[2];
  uint32_t tmem_base;
  uint32_t pad;
  float merge_m[128];
  float merge_l[128];
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_fwd3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Fwd3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BM, bidx = blockIdx.y;

  // streamed-Q mode (Kd > 256): a stage holds [Q chunk | K chunk]
  const uint32_t stage_bytes = p.q_resident ? KSTAGE_BYTES : (ATOM_BYTES + KSTAGE_BYTES);
  const uint32_t v_stage_bytes = static_cast<uint32_t>(p.Cvp) * 512u;  // 4 atoms of [Cvp x 64 keys]
  const uint32_t q_smem = smem0;
  const uint32_t k_ring = q_smem + (p.q_resident ? p.kc_count * ATOM_BYTES : 0);
  const uint32_t v_ring = k_ring + p.ns_k * stage_bytes;
  const uint32_t bar_off = v_ring + p.ns_v * v_stage_bytes - smem0;
  Bars3* bars = reinterpret_cast<Bars3*>(smem_gen + bar_off);
  float* merge_o = reinterpret_cast<float*>(smem_gen + (k_ring - smem0));  // K ring is free after the last MMA

  if (tid == 0) {
    mbar_init(smem_u32(&bars->q_full), 1);
    for (int i = 0; i < p.ns_k; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
    }
    for (int i = 0; i < p.ns_v; ++i) {
      mbar_init(smem_u32(&bars->v_full[i]), 1);
      mbar_init(smem_u32(&bars->v_empty[i]), 2);  // both PV streams (g = 0, 1) release a V tile
    }
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->s_empty), 256);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->p_full[i]), 128);
      mbar_init(smem_u32(&bars->pv_done[i]), 1);
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
  const int T = p.n_tiles;  // 256-key tiles

  if (warp >= 8) {
    setmaxnreg_dec<48>();
    const bool leader = elect_one();
    if (warp == 8) {
      // ------------------------------------------------------ TMA producer: Q, K
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
            tma_load_3d(dst, &tm_k, full, kc * BK, j * BN2, bidx);  // one 256-row box
            if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
          }
        }
      }
    } else if (warp == 9) {
      // --------------------------------------------------------- TMA producer: V
      if (leader) {
        uint32_t vs = 0, vph = 0;
        for (int j = 0; j < T; ++j) {
          mbar_wait(smem_u32(&bars->v_empty[vs]), vph ^ 1);
          const uint32_t vfull = smem_u32(&bars->v_full[vs]);
          mbar_expect_tx(vfull, v_stage_bytes);
          const uint32_t vdst = v_ring + vs * v_stage_bytes;
#pragma unroll
          for (int a = 0; a < 4; ++a) tma_load_3d(vdst + a * (v_stage_bytes / 4), &tm_v, vfull, j * BN2 + a * BK, 0, bidx);
          if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == 10) {
      // ------------------------------------------- MMA issuer: S = Q K^T, N = 256
      const uint32_t idesc_s = make_idesc_f16(BM, BN2);
      uint32_t ks = 0, kph = 0;
      if (p.q_resident) {
        mbar_wait(smem_u32(&bars->q_full), 0);
        tc_fence_after();
      }
      for (int t = 0; t < T; ++t) {
        TRACE(4, t, 0);
        if (t >= 1) mbar_wait(smem_u32(&bars->s_empty), (t - 1) & 1);
        tc_fence_after();
        TRACE(4, t, 1);
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          tc_fence_after();
          if (kc < 4) TRACE(4, t, 2 + kc);
          if (leader) {
            const uint32_t st = k_ring + ks * stage_bytes;
            const uint64_t da = make_desc_k_sw128(p.q_resident ? (q_smem + kc * ATOM_BYTES) : st);
            const uint64_t db = make_desc_k_sw128(p.q_resident ? st : (st + ATOM_BYTES));
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4)
              umma_f16(tmem, desc_advance_k16(da, s4), desc_advance_k16(db, s4), idesc_s, (kc | s4) != 0 ? 1u : 0u);
            umma_commit(smem_u32(&bars->k_empty[ks]));
            if (kc == p.kc_count - 1) umma_commit(smem_u32(&bars->s_full));
          }
          __syncwarp();
          if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
        }
        TRACE(4, t, 6);
      }
    } else {
      // ---------------------------------- MMA issuer: O_g += P_g V_g (A = P from TMEM)
      const uint32_t idesc_pv = make_idesc_f16(BM, p.Cvp);
      uint32_t vs = 0, vph = 0;
      for (int t = 0; t < T; ++t) {
        mbar_wait(smem_u32(&bars->v_full[vs]), vph);
        const uint32_t vb = v_ring + vs * v_stage_bytes;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          TRACE(5, t, g * 3);
          mbar_wait(smem_u32(&bars->p_full[g]), t & 1);
          tc_fence_after();
          TRACE(5, t, g * 3 + 1);
          if (leader) {
#pragma unroll
            for (int at = 0; at < 2; ++at) {
              const uint64_t db = make_desc_k_sw128(vb + (g * 2 + at) * (v_stage_bytes / 4));
#pragma unroll
              for (int s4 = 0; s4 < 4; ++s4)
                umma_f16_ts(tmem + 384 + g * 64, tmem + 256 + g * 64 + (at * 4 + s4) * 8, desc_advance_k16(db, s4),
                            idesc_pv, (t | at | s4) != 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&bars->v_empty[vs]));
            umma_commit(smem_u32(&bars->pv_done[g]));
          }
          __syncwarp();
          TRACE(5, t, g * 3 + 2);
        }
        if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------- softmax warpgroups (warps 0..7)
    setmaxnreg_inc<224>();
    const int g = warp >> 2;    // which half of every 256-key tile
    const int row = tid & 127;  // TMEM lane == query row
    const int q = q0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem + g * 128 + lane_sel;
    const uint32_t tmem_p = tmem + 256 + g * 64 + lane_sel;
    const uint32_t tmem_o = tmem + 384 + g * 64 + lane_sel;
    const float c2 = p.scale_log2;
    float m = -INFINITY, l = 0.f;
    bool seen = false;  // at least one valid key processed

    for (int t = 0; t < T; ++t) {
      if ((warp & 3) == 0) TRACE(g, t, 0);
      mbar_wait(smem_u32(&bars->s_full), t & 1);
      tc_fence_after();
      if ((warp & 3) == 0) TRACE(g, t, 1);
      float s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + c * 32, reinterpret_cast<uint32_t*>(&s[c * 32]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_empty));
      if ((warp & 3) == 0) TRACE(g, t, 2);

      const int valid = p.Nk - (t * BN2 + g * 128);  // keys of my half that exist (<= 0: none, last tile only)
      if (valid < 128) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      // row max: 8 independent chains (a single dependent FMNMX chain costs ~1.6k clk per tile)
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
      const float m_new = fmaxf(m, tmax);
      if (!seen) {
        m = m_new;  // stays -inf while nothing valid has been seen
      } else {
        const bool resc = (m_new - m) * c2 > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, resc)) {
          const float alpha = resc ? ex2((m - m_new) * c2) : 1.0f;
          if (resc) m = m_new;
          l *= alpha;
          mbar_wait(smem_u32(&bars->pv_done[g]), (t - 1) & 1);
          tc_fence_after();
          for (int cc = 0; cc < p.Cvp; cc += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + cc, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_o + cc, o);
          }
          tmem_wait_st();
          tc_fence_before();
        }
      }
      seen = seen || (valid > 0);
      if ((warp & 3) == 0) TRACE(g, t, 3);
      const float mc = seen ? m * c2 : 0.f;  // all-masked half: ex2(-inf - 0) = 0, no NaN
      float sum = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float p0 = ex2(fmaf(s[2 * i], c2, -mc));
        const float p1 = ex2(fmaf(s[2 * i + 1], c2, -mc));
        sum += p0 + p1;
        pk[i] = pack_h2(p0, p1);
      }
      l += sum;
      if ((warp & 3) == 0) TRACE(g, t, 4);
      if (t > 0) mbar_wait(smem_u32(&bars->pv_done[g]), (t - 1) & 1);  // my P columns are free again
      tc_fence_after();
      if ((warp & 3) == 0) TRACE(g, t, 5);
      tmem_st32(tmem_p, pk);
      tmem_st32(tmem_p + 32, pk + 32);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->p_full[g]));
      if ((warp & 3) == 0) TRACE(g, t, 6);
    }

    // ---- merge the two column halves; WG0 writes the result
    mbar_wait(smem_u32(&bars->pv_done[g]), (T - 1) & 1);
    tc_fence_after();
    asm volatile("bar.sync 2, 256;" ::: "memory");  // every MMA has completed: the K ring is reusable
    if (g == 1) {
      bars->merge_m[row] = seen ? m : -INFINITY;
      bars->merge_l[row] = l;
      for (int cc = 0; cc < p.Cvp; cc += 16) {
        uint32_t o[16];
        tmem_ld16(tmem_o + cc, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) merge_o[(cc + i) * 128 + row] = __uint_as_float(o[i]);
      }
      tc_fence_before();
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (g == 0) {
      const float m1 = bars->merge_m[row], l1 = bars->merge_l[row];
      const float mm = fmaxf(m, m1);
      const float a0 = ex2((m - mm) * c2);
      const float a1 = (m1 == -INFINITY) ? 0.f : ex2((m1 - mm) * c2);
      const float lt = l * a0 + l1 * a1;
      const float inv_l = 1.0f / lt;
      for (int cc = 0; cc < p.Cvp; cc += 16) {
        uint32_t o[16];
        tmem_ld16(tmem_o + cc, o);
        tmem_wait_ld();
        if (q < p.Nq) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cc + i;
            if (c < p.Cv) {
              const float o1 = (m1 == -INFINITY) ? 0.f : merge_o[c * 128 + row];
              p.out[(static_cast<size_t>(bidx) * p.Cv + c) * p.Nq + q] = (__uint_as_float(o[i]) * a0 + o1 * a1) * inv_l;
            }
          }
        }
      }
      if (p.lse != nullptr && q < p.Nq)
        p.lse[static_cast<size_t>(bidx) * p.Nq + q] = (mm * c2 + log2f(lt)) * 0.6931471805599453f;
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// returns 1 when this variant does not apply (caller falls back)
int corr_warp_fwd3_launch(const void* q, const void* k, const void* vt, float* out, float* lse, int B, int Nq,
                          int Nk, int Kd, int Cv, int Cvp, int Nkp, float scale, cudaStream_t stream) {
  if (Cvp > 64) return 1;
  Fwd3Params p;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.Kd = Kd; p.Cv = Cv; p.Cvp = Cvp;
  p.kc_count = Kd / BK;
  p.n_tiles = (Nk + BN2 - 1) / BN2;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse;
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(Bars3)) - 64;
  const int v_stage = Cvp * 512;
  p.q_resident = (Kd <= 256) ? 1 : 0;
  const int q_bytes = p.q_resident ? p.kc_count * ATOM_BYTES : 0;
  const int stage = p.q_resident ? KSTAGE_BYTES : (ATOM_BYTES + KSTAGE_BYTES);
  p.ns_v = 2;
  int rem = budget - q_bytes - p.ns_v * v_stage;
  p.ns_k = rem / stage;
  if (p.ns_k > MAX_KSTAGES) p.ns_k = MAX_KSTAGES;
  if (p.ns_k < 2 || p.ns_k * stage < Cvp * 128 * 4) return 1;
  const int smem_bytes = 1024 + q_bytes + p.ns_k * stage + p.ns_v * v_stage + sizeof(Bars3) + 64;

  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  if ((rc = make_tmap_f16_3d(&tm_q, q, Kd, Nq, B, (uint64_t)Kd * 2, (uint64_t)Nq * Kd * 2, BK, BM, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_k, k, Kd, Nk, B, (uint64_t)Kd * 2, (uint64_t)Nk * Kd * 2, BK, BN2, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_v, vt, Nk, Cvp, B, (uint64_t)Nkp * 2, (uint64_t)Cvp * Nkp * 2, BK, Cvp, 1)))
    return rc;
  COCOS_CUDA_CHECK(
      cudaFuncSetAttribute(corr_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid((Nq + BM - 1) / BM, B);
  p.trace = nullptr;
  const char* dbg = getenv("COCOS_K1_DBG");
  if (dbg && (atoi(dbg) & 512)) {  // debug only: timeline of CTA (0,0); synchronises and prints to stderr
    const size_t nb = 6 * 64 * 8 * sizeof(unsigned long long);
    COCOS_CUDA_CHECK(cudaMalloc(&p.trace, nb));
    COCOS_CUDA_CHECK(cudaMemset(p.trace, 0, nb));
    corr_fwd3_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
    COCOS_CUDA_CHECK(cudaDeviceSynchronize());
    static unsigned long long h[6 * 64 * 8];
    COCOS_CUDA_CHECK(cudaMemcpy(h, p.trace, nb, cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < 6 * 64 * 8; ++i)
      if (h[i] && h[i] < t0) t0 = h[i];
    const char* names[6] = {"WG0", "WG1", "", "", "MMA_S", "MMA_PV"};
    for (int t = 0; t < 16 && t < p.n_tiles; ++t)
      for (int r = 0; r < 6; ++r) {
        if (!names[r][0] || !h[(r * 64 + t) * 8]) continue;
        fprintf(stderr, "trace tile %2d %-6s:", t, names[r]);
        for (int e = 0; e < 7; ++e)
          fprintf(stderr, " %7lld", (long long)(h[(r * 64 + t) * 8 + e] ? (long long)(h[(r * 64 + t) * 8 + e] - t0) : -1));
        fprintf(stderr, "\n");
      }
    return 0;
  }
  corr_fwd3_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
