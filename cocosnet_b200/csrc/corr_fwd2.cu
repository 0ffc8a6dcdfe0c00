// K1 v2: fused correlation -> softmax -> warp with TWO softmax warpgroups.
//
// Same math and operand layouts as corr_fwd.cu; the difference is the
// pipeline.  The exp throughput (MUFU, 16/clk/SM) is co-critical with the
// tensor core at C=256, and one warpgroup cannot both keep MUFU busy and hide
// the TMEM->register->smem latency chain.  Here warpgroup g handles key tiles
// t = g (mod 2) with its own S buffer, P buffer, running (max, sum) and its
// own O accumulator in TMEM, so while one group is in its exp loop the other
// is loading/stashing, and the MMA warp alternates S_t / PV_{t-1}.  The two
// partial (m, l, O) are merged once at the end through shared memory.
//
//   warps 0-3 : softmax WG0 (even key tiles)     warp 8 : TMA producer for Q / K chunks
//   warps 4-7 : softmax WG1 (odd key tiles)      warp 9 : TMA producer for V tiles
//   warp 10   : issuer of the S = Q K^T MMAs, TMEM owner
//   warp 11   : issuer of the O += P V MMAs
// Two issuers so that the S stream never stalls behind a P tile that is still
// being exponentiated; descriptors are advanced with one add per MMA and the
// issue loops are warp-converged (elect.sync) -- the issue path, not the tensor
// pipe, was the limiter of the first version (profiles/r01_k1_fwd_v1_ncu.md).
// TMEM columns: S0 [0,128) S1 [128,256) O0 [256,256+Cvp) O1 [384,384+Cvp), Cvp <= 128; with P in tensor memory
// (Cvp <= 64): S0 S1 | P0 [256,320) P1 [320,384) | O0 [384,448) O1 [448,512) and no P buffers in shared memory,
// which leaves room for a K ring deep enough to cover the TMA latency (profiles/r01_k1_timeline.md).
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

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int ATOM_BYTES = 128 * BK * 2;
constexpr int NUM_THREADS = 384;
constexpr int MAX_KSTAGES = 10;
constexpr int MAX_VSTAGES = 4;
constexpr float RESCALE_THRESHOLD = 8.0f;

struct Fwd2Params {
  int B, Nq, Nk, Kd, Cv, Cvp;
  int kc_count, n_tiles, q_resident, ns_k, ns_v;
  int q_tmem;  // Q tile kept in TENSOR memory (A operand of the S MMA from TMEM): Kd <= 256, Cvp <= 64, P in smem
  const __half* q_ptr;
  int p_tmem;  // P kept in tensor memory (A operand of the PV MMA from TMEM); needs Cvp <= 64
  float scale, scale_log2;
  float* out;
  float* lse;
  unsigned long long* trace;  // dbg & 512: per-role clock64 stamps of CTA (0,0), [role][tile][8]
  int dbg;  // COCOS_K1_DBG bitmask, timing experiments only: 1 no ex2, 2 no P store, 4 no PV mma, 8 no S mma
};

struct Bars2 {
  uint64_t q_full;
  uint64_t k_full[MAX_KSTAGES];
  uint64_t k_empty[MAX_KSTAGES];
  uint64_t v_full[MAX_VSTAGES];
  uint64_t v_empty[MAX_VSTAGES];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t p_full[2];
  uint64_t pv_done[2];
  uint32_t tmem_base;
  uint32_t pad;
  float merge_m[128];
  float merge_l[128];
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Fwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BM, bidx = blockIdx.y;

  const uint32_t stage_bytes = p.q_resident ? ATOM_BYTES : 2 * ATOM_BYTES;
  const uint32_t v_stage_bytes = static_cast<uint32_t>(p.Cvp) * 256u;
  const uint32_t q_smem = smem0;
  const uint32_t k_ring = q_smem + ((p.q_resident && !p.q_tmem) ? p.kc_count * ATOM_BYTES : 0);
  const uint32_t v_ring = k_ring + p.ns_k * stage_bytes;
  const uint32_t p_smem = v_ring + p.ns_v * v_stage_bytes;  // P0 | P1, 2 atoms each (absent with p_tmem)
  const uint32_t bar_off = p_smem + (p.p_tmem ? 0 : 4 * ATOM_BYTES) - smem0;
  Bars2* bars = reinterpret_cast<Bars2*>(smem_gen + bar_off);
  // merge scratch, reused after the last MMA: the P buffers, or the K ring when P lives in TMEM
  float* merge_o = reinterpret_cast<float*>(smem_gen + ((p.p_tmem ? k_ring : p_smem) - smem0));
  // TMEM column map.  q_tmem: Q [0,128) S0 [128,256) S1 [256,384) O0 [384,448) O1 [448,512)
  const uint32_t s_col = p.q_tmem ? 128u : 0u;
  const uint32_t o_col = (p.p_tmem || p.q_tmem) ? 384u : 256u, o_stride = (p.p_tmem || p.q_tmem) ? 64u : 128u;
  const uint32_t pt_col = 256u;

  if (tid == 0) {
    mbar_init(smem_u32(&bars->q_full), p.q_tmem ? 128 : 1);
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
  const int T = p.n_tiles;
  // Rotate the key-tile order per CTA: the CTAs of one image would otherwise all stream the same K tile at the
  // same moment (same L2 lines requested 32x concurrently).  Online softmax is order independent.
  const int t_shift = (p.dbg & 64) ? static_cast<int>((static_cast<long long>(blockIdx.x) * T) / gridDim.x) : 0;

  if (warp >= 8) {
    setmaxnreg_dec<48>();
    const bool leader = elect_one();
    if (warp == 8) {
      // ------------------------------------------------------ TMA producer: Q, K
      if (leader) {
        if (p.q_resident && !p.q_tmem) {
          mbar_expect_tx(smem_u32(&bars->q_full), p.kc_count * ATOM_BYTES);
          for (int kc = 0; kc < p.kc_count; ++kc)
            tma_load_3d(q_smem + kc * ATOM_BYTES, &tm_q, smem_u32(&bars->q_full), kc * BK, q0, bidx);
        }
        uint32_t ks = 0, kph = 0;
        for (int j = 0; j < T; ++j) {
          const int jj = (j + t_shift) % T;
          for (int kc = 0; kc < p.kc_count; ++kc) {
            mbar_wait(smem_u32(&bars->k_empty[ks]), kph ^ 1);
            const uint32_t full = smem_u32(&bars->k_full[ks]);
            if (p.dbg & 128) {  // timing experiment: no TMA traffic at all
              mbar_arrive(full);
            } else {
              mbar_expect_tx(full, stage_bytes);
              uint32_t dst = k_ring + ks * stage_bytes;
              if (!p.q_resident) {
                tma_load_3d(dst, &tm_q, full, kc * BK, q0, bidx);
                dst += ATOM_BYTES;
              }
              tma_load_3d(dst, &tm_k, full, kc * BK, jj * BN, bidx);
            }
            if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
          }
        }
      }
    } else if (warp == 9) {
      // --------------------------------------------------------- TMA producer: V
      if (leader) {
        uint32_t vs = 0, vph = 0;
        for (int j = 0; j < T; ++j) {
          const int jj = (j + t_shift) % T;
          mbar_wait(smem_u32(&bars->v_empty[vs]), vph ^ 1);
          const uint32_t vfull = smem_u32(&bars->v_full[vs]);
          mbar_expect_tx(vfull, v_stage_bytes);
          const uint32_t vdst = v_ring + vs * v_stage_bytes;
          tma_load_3d(vdst, &tm_v, vfull, jj * BN, 0, bidx);
          tma_load_3d(vdst + v_stage_bytes / 2, &tm_v, vfull, jj * BN + BK, 0, bidx);
          if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
        }
      }
    } else if (warp == 10) {
      // ------------------------------------------------ MMA issuer: S = Q K^T
      const uint32_t idesc_s = make_idesc_f16(BM, BN);
      uint32_t ks = 0, kph = 0;
      if (p.q_resident) {
        mbar_wait(smem_u32(&bars->q_full), 0);
        tc_fence_after();
      }
      for (int t = 0; t < T; ++t) {
        const int g = t & 1, u = t >> 1;
        TRACE(4, t, 0);
        if (u >= 1) mbar_wait(smem_u32(&bars->s_empty[g]), (u - 1) & 1);
        tc_fence_after();
        TRACE(4, t, 1);
        const uint32_t d_tmem = tmem + s_col + g * BN;
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          tc_fence_after();
          if (kc < 4) TRACE(4, t, 2 + kc);
          if (leader) {
            const uint32_t st = k_ring + ks * stage_bytes;
            const uint64_t da = make_desc_k_sw128(p.q_resident ? (q_smem + kc * ATOM_BYTES) : st);
            const uint64_t db = make_desc_k_sw128(p.q_resident ? st : (st + ATOM_BYTES));
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              if ((p.dbg & 8) && t >= 2) continue;
              if (p.q_tmem)  // A = Q from tensor memory: 8 columns (16 fp16) per K step, 32 per 64-wide chunk
                umma_f16_ts(d_tmem, tmem + kc * 32 + s4 * 8, desc_advance_k16(db, s4), idesc_s,
                            (kc | s4) != 0 ? 1u : 0u);
              else
                umma_f16(d_tmem, desc_advance_k16(da, s4), desc_advance_k16(db, s4), idesc_s,
                         (kc | s4) != 0 ? 1u : 0u);
            }
            umma_commit(smem_u32(&bars->k_empty[ks]));
            if (kc == p.kc_count - 1) umma_commit(smem_u32(&bars->s_full[g]));
          }
          __syncwarp();
          if (++ks == static_cast<uint32_t>(p.ns_k)) { ks = 0; kph ^= 1; }
        }
        TRACE(4, t, 6);
      }
    } else {
      // ------------------------------------------------ MMA issuer: O += P V
      const uint32_t idesc_pv = make_idesc_f16(BM, p.Cvp);
      uint32_t vs = 0, vph = 0;
      for (int t = 0; t < T; ++t) {
        const int g = t & 1, u = t >> 1;
        TRACE(5, t, 0);
        mbar_wait(smem_u32(&bars->p_full[g]), u & 1);
        TRACE(5, t, 1);
        mbar_wait(smem_u32(&bars->v_full[vs]), vph);
        tc_fence_after();
        TRACE(5, t, 2);
        if (leader) {
          const uint32_t vb = v_ring + vs * v_stage_bytes;
          const uint32_t pb = p_smem + g * 2 * ATOM_BYTES;
#pragma unroll
          for (int at = 0; at < 2; ++at) {
            const uint64_t da = make_desc_k_sw128(pb + at * ATOM_BYTES);
            const uint64_t db = make_desc_k_sw128(vb + at * (v_stage_bytes / 2));
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              if ((p.dbg & 4) && t >= 2) continue;
              const uint32_t acc = (u | at | s4) != 0 ? 1u : 0u;
              if (p.p_tmem)  // A = P from tensor memory: 8 columns (16 fp16) per K step
                umma_f16_ts(tmem + o_col + g * o_stride, tmem + pt_col + g * 64 + (at * 4 + s4) * 8,
                            desc_advance_k16(db, s4), idesc_pv, acc);
              else
                umma_f16(tmem + o_col + g * o_stride, desc_advance_k16(da, s4), desc_advance_k16(db, s4), idesc_pv,
                         acc);
            }
          }
          umma_commit(smem_u32(&bars->v_empty[vs]));
          umma_commit(smem_u32(&bars->pv_done[g]));
        }
        __syncwarp();
        TRACE(5, t, 3);
        if (++vs == static_cast<uint32_t>(p.ns_v)) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------- softmax warpgroups (warps 0..7)
    setmaxnreg_inc<208>();
    const int g = warp >> 2;          // warpgroup: even / odd key tiles
    const int row = tid & 127;        // TMEM lane == query row
    const int q = q0 + row;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem + s_col + g * BN;
    const uint32_t tmem_o = tmem + o_col + g * o_stride;
    const float c2 = p.scale_log2;
    if (p.q_tmem && g == 0) {
      // my query row straight from global memory into tensor memory (lane = row, 2 fp16 per 32-bit column)
      const uint4* src = reinterpret_cast<const uint4*>(p.q_ptr + (static_cast<size_t>(bidx) * p.Nq + q) * p.Kd);
      for (int cc = 0; cc < p.Kd / 2; cc += 32) {  // 32 columns = 64 fp16 = 8 x 16 B
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (q < p.Nq) v = __ldg(src + (cc / 4) + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        tmem_st32(tmem + lane_sel + cc, r);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->q_full));
    }
    float m = -INFINITY, l = 0.f;
    const uint32_t row_off = p_smem + g * 2 * ATOM_BYTES + row * 128;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    int n_mine = 0;

    for (int t = g; t < T; t += 2) {
      const int u = t >> 1;
      ++n_mine;
      if ((warp & 3) == 0) TRACE(g, t, 0);
      mbar_wait(smem_u32(&bars->s_full[g]), u & 1);
      tc_fence_after();
      if ((warp & 3) == 0) TRACE(g, t, 1);
      if (p.dbg & 16) {  // timing experiment: barrier skeleton only
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->s_empty[g]));
        if (u > 0) mbar_wait(smem_u32(&bars->pv_done[g]), (u - 1) & 1);
        mbar_arrive(smem_u32(&bars->p_full[g]));
        continue;
      }
      float s[BN];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + lane_sel + c * 32, reinterpret_cast<uint32_t*>(&s[c * 32]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->s_empty[g]));
      if ((warp & 3) == 0) TRACE(g, t, 2);

      const int tt = (t + t_shift) % T;  // the key tile this step actually holds
      if (tt == T - 1 && (p.Nk & (BN - 1)) != 0) {
        const int valid = p.Nk - tt * BN;
#pragma unroll
        for (int c = 0; c < BN; ++c)
          if (c >= valid) s[c] = -INFINITY;
      }
      float tmax = s[0];
#pragma unroll
      for (int c = 1; c < BN; ++c) tmax = fmaxf(tmax, s[c]);
      const float m_new = fmaxf(m, tmax);
      if (u == 0) {
        m = m_new;
      } else {
        const bool resc = (m_new - m) * c2 > RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, resc)) {
          const float alpha = resc ? ex2((m - m_new) * c2) : 1.0f;
          if (resc) m = m_new;
          l *= alpha;
          mbar_wait(smem_u32(&bars->pv_done[g]), (u - 1) & 1);
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
      if ((warp & 3) == 0) TRACE(g, t, 3);
      const float mc = m * c2;
      float sum = 0.f;
      uint32_t pk[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        float p0 = fmaf(s[2 * i], c2, -mc), p1 = fmaf(s[2 * i + 1], c2, -mc);
        if (!(p.dbg & 1)) {
          p0 = ex2(p0);
          p1 = ex2(p1);
        }
        sum += p0 + p1;
        pk[i] = pack_h2(p0, p1);
      }
      l += sum;
      if ((warp & 3) == 0) TRACE(g, t, 4);
      if (u > 0) mbar_wait(smem_u32(&bars->pv_done[g]), (u - 1) & 1);  // my P buffer is free again
      if ((warp & 3) == 0) TRACE(g, t, 5);
      if (p.p_tmem) {
        tmem_st32(tmem + pt_col + g * 64 + lane_sel, pk);
        tmem_st32(tmem + pt_col + g * 64 + 32 + lane_sel, pk + 32);
        tmem_wait_st();
        tc_fence_before();
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (p.dbg & 2) break;
          const uint32_t addr = row_off + (c >> 3) * ATOM_BYTES + (((c & 7) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                       "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
      }
      mbar_arrive(smem_u32(&bars->p_full[g]));
      if ((warp & 3) == 0) TRACE(g, t, 6);
    }

    // ---- merge the two partial softmax states; WG0 writes the result
    if (n_mine > 0) {
      mbar_wait(smem_u32(&bars->pv_done[g]), (n_mine - 1) & 1);
      tc_fence_after();
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");  // every MMA has completed: P buffers are reusable
    if (g == 1) {
      bars->merge_m[row] = m;
      bars->merge_l[row] = l;
      for (int cc = 0; cc < p.Cvp; cc += 16) {
        uint32_t o[16];
        if (n_mine > 0) {
          tmem_ld16(tmem_o + lane_sel + cc, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = 0u;
        }
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
        tmem_ld16(tmem_o + lane_sel + cc, o);
        tmem_wait_ld();
        if (q < p.Nq) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cc + i;
            if (c < p.Cv) {
              const float v = (__uint_as_float(o[i]) * a0 + merge_o[c * 128 + row] * a1) * inv_l;
              p.out[(static_cast<size_t>(bidx) * p.Cv + c) * p.Nq + q] = v;
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

int corr_warp_fwd2_launch(const void* q, const void* k, const void* vt, float* out, float* lse, int B, int Nq,
                          int Nk, int Kd, int Cv, int Cvp, int Nkp, float scale, cudaStream_t stream) {
  Fwd2Params p;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.Kd = Kd; p.Cv = Cv; p.Cvp = Cvp;
  p.kc_count = Kd / BK;
  p.n_tiles = (Nk + BN - 1) / BN;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.lse = lse;
  {
    const char* e = getenv("COCOS_K1_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  {
    const char* e = getenv("COCOS_K1_PTMEM");
    p.p_tmem = (Cvp <= 64) && (e ? atoi(e) != 0 : false);
    const char* e2 = getenv("COCOS_K1_QTMEM");
    p.q_tmem = (!p.p_tmem) && (Cvp <= 64) && (Kd <= 256) && (e2 ? atoi(e2) != 0 : true);
    p.q_ptr = static_cast<const __half*>(q);
  }
  const int budget = 227 * 1024 - 1024 - static_cast<int>(sizeof(Bars2)) - 64;
  const int p_bytes = p.p_tmem ? 0 : 4 * ATOM_BYTES;
  const int v_stage = Cvp * 256;
  p.q_resident = (Kd <= 256) ? 1 : 0;
  const int q_bytes = (p.q_resident && !p.q_tmem) ? p.kc_count * ATOM_BYTES : 0;
  const int stage = p.q_resident ? ATOM_BYTES : 2 * ATOM_BYTES;
  p.ns_v = (v_stage <= 8192) ? 4 : 2;
  int rem = budget - p_bytes - q_bytes - p.ns_v * v_stage;
  if (rem / stage < 3) {
    p.ns_v = 1;
    rem = budget - p_bytes - q_bytes - v_stage;
  }
  p.ns_k = rem / stage;
  if (p.ns_k > MAX_KSTAGES) p.ns_k = MAX_KSTAGES;
  if ((p.dbg & 32) && p.ns_k > 2) p.ns_k = 2;
  if (p.ns_k < 2) return 1;  // caller falls back to the single-warpgroup kernel
  const int smem_bytes = 1024 + q_bytes + p.ns_k * stage + p.ns_v * v_stage + p_bytes + sizeof(Bars2) + 64;

  CUtensorMap tm_q, tm_k, tm_v;
  int rc;
  if ((rc = make_tmap_f16_3d(&tm_q, q, Kd, Nq, B, (uint64_t)Kd * 2, (uint64_t)Nq * Kd * 2, BK, BM, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_k, k, Kd, Nk, B, (uint64_t)Kd * 2, (uint64_t)Nk * Kd * 2, BK, BN, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_v, vt, Nk, Cvp, B, (uint64_t)Nkp * 2, (uint64_t)Cvp * Nkp * 2, BK, Cvp, 1)))
    return rc;
  COCOS_CUDA_CHECK(
      cudaFuncSetAttribute(corr_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid((Nq + BM - 1) / BM, B);
  p.trace = nullptr;
  if (p.dbg & 512) {  // debug only: timeline of CTA (0,0); synchronises and prints
    const size_t nb = 6 * 64 * 8 * sizeof(unsigned long long);
    COCOS_CUDA_CHECK(cudaMalloc(&p.trace, nb));
    COCOS_CUDA_CHECK(cudaMemset(p.trace, 0, nb));
    corr_fwd2_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
    COCOS_CUDA_CHECK(cudaDeviceSynchronize());
    static unsigned long long h[6 * 64 * 8];
    COCOS_CUDA_CHECK(cudaMemcpy(h, p.trace, nb, cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < 6 * 64 * 8; ++i)
      if (h[i] && h[i] < t0) t0 = h[i];
    const char* names[6] = {"WG0", "WG1", "", "", "MMA_S", "MMA_PV"};
    for (int t = 0; t < 12 && t < p.n_tiles; ++t)
      for (int r = 0; r < 6; ++r) {
        if (!names[r][0] || !h[(r * 64 + t) * 8]) continue;
        fprintf(stderr, "trace tile %2d %-6s:", t, names[r]);
        for (int e = 0; e < 7; ++e) fprintf(stderr, " %7lld", (long long)(h[(r * 64 + t) * 8 + e] ? h[(r * 64 + t) * 8 + e] - t0 : -1));
        fprintf(stderr, "\n");
      }
    return 0;
  }
  corr_fwd2_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_v, p);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
