// K1 backward, stage A: recompute the correlation tile by tile and emit the
// softmax-Jacobian product
//     dS[i,j] = P[i,j] * (dP[i,j] - D[i]) * scale,   P = exp(scale*S - lse),
//     dP = dO V^T,  D[i] = sum_c dO[c,i] O[c,i]
// as bf16 (dO enters row-scaled in fp16, the scale is divided back out) in BOTH
// orientations -- dS [B,Nq,Nkp] and dS^T [B,Nk,Nqp] -- plus optionally P^T.
// Two plain tcgen05 GEMMs (gemm.cu) then give dQ^T = K^T-major x dS and
// dK^T = Q^T-major x dS^T; a third gives dV from P^T.  This is what autograd
// does through reference correspondence.py:291-318, without keeping the three
// fp32 N x N tensors alive.
//
// CTA = 128 query rows, streams 128-key tiles.  warp 4: TMA, warp 5: MMA
// (S and dP, both double-buffered in TMEM: 4 x 128 columns), warps 0-3: one
// thread per query row.
#include "corr_kernels.h"
#include "ptx.cuh"
#include "tmap.h"

#include <cuda_bf16.h>

namespace cocos {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int ATOM_BYTES = 128 * BK * 2;
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 4;
constexpr int TPITCH = 136;                     // staging tile pitch (fp16 elements)
constexpr int TILE_BYTES = 128 * TPITCH * 2;    // 34816

struct BwdParams {
  int B, Nq, Nk, Kd, Cv, Cvk;
  int kc_count, vc_count, n_tiles, ns;
  int Nkp, Nqp;
  float scale, scale_log2;
  const __half* do16;  // [B, Nq, Cvk] row-scaled dO (the dP operand)
  const float* rscale; // [B, Nq] the row scale r_i applied to dO
  const float* out;    // [B, Cv, Nq]
  const float* lse;    // [B, Nq]
  uint16_t* ds;        // bf16 [B, Nq, Nkp]
  uint16_t* dst;       // bf16 [B, Nk, Nqp]
  uint16_t* pt;        // bf16 [B, Nk, Nqp] or null
};

struct BwdBars {
  uint64_t do_full;
  uint64_t k_full[MAX_STAGES];
  uint64_t k_empty[MAX_STAGES];
  uint64_t v_full;
  uint64_t v_empty;
  uint64_t sdp_full[2];
  uint64_t sdp_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void wg_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// copy a staged 128x128 fp16 tile (pitch TPITCH) to global rows of pitch `gpitch`
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void copy_tile_out(const uint8_t* tile, uint16_t* gbase, int gpitch, int valid_rows,
                                              int valid_cols, int tid) {
  const int chunk = tid & 15;
  if (chunk * 8 >= valid_cols) return;
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const int r = (tid >> 4) + it * 8;
    if (r < valid_rows) {
      const uint4 v = *reinterpret_cast<const uint4*>(tile + (r * TPITCH + chunk * 8) * 2);
      *reinterpret_cast<uint4*>(gbase + static_cast<size_t>(r) * gpitch + chunk * 8) = v;
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_bwd_ds_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_v,
                   const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BM, bidx = blockIdx.y;
  const int T = p.n_tiles;

  const uint32_t stage_bytes = 2 * ATOM_BYTES;                 // [Q chunk | K chunk]
  const uint32_t op_bytes = p.vc_count * ATOM_BYTES;           // dO tile / V tile
  const uint32_t k_ring = smem0;
  const uint32_t do_smem = k_ring + p.ns * stage_bytes;
  const uint32_t v_smem = do_smem + op_bytes;
  const uint32_t stg_off = (v_smem + op_bytes) - smem0;        // staging tiles A, B
  uint8_t* tile = smem_gen + stg_off;  // one staging tile, reused for dS, dS^T (and P^T)
  BwdBars* bars = reinterpret_cast<BwdBars*>(tile + TILE_BYTES);

  if (tid == 0) {
    mbar_init(smem_u32(&bars->do_full), 1);
    for (int i = 0; i < p.ns; ++i) {
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->v_full), 1);
    mbar_init(smem_u32(&bars->v_empty), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->sdp_full[i]), 1);
      mbar_init(smem_u32(&bars->sdp_empty[i]), 128);
    }
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;  // S0 [0,128) S1 [128,256) dP0 [256,384) dP1 [384,512)

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bars->do_full), op_bytes);
      for (int vc = 0; vc < p.vc_count; ++vc)
        tma_load_3d(do_smem + vc * ATOM_BYTES, &tm_do, smem_u32(&bars->do_full), vc * BK, q0, bidx);
      uint32_t ks = 0, kph = 0;
      for (int j = 0; j < T; ++j) {
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(smem_u32(&bars->k_empty[ks]), kph ^ 1);
          const uint32_t full = smem_u32(&bars->k_full[ks]);
          mbar_expect_tx(full, stage_bytes);
          tma_load_3d(k_ring + ks * stage_bytes, &tm_q, full, kc * BK, q0, bidx);
          tma_load_3d(k_ring + ks * stage_bytes + ATOM_BYTES, &tm_k, full, kc * BK, j * BN, bidx);
          if (++ks == static_cast<uint32_t>(p.ns)) { ks = 0; kph ^= 1; }
        }
        mbar_wait(smem_u32(&bars->v_empty), (j & 1) ^ 1);
        mbar_expect_tx(smem_u32(&bars->v_full), op_bytes);
        for (int vc = 0; vc < p.vc_count; ++vc)
          tma_load_3d(v_smem + vc * ATOM_BYTES, &tm_v, smem_u32(&bars->v_full), vc * BK, j * BN, bidx);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BM, BN);
      uint32_t ks = 0, kph = 0;
      mbar_wait(smem_u32(&bars->do_full), 0);
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        if (t >= 2) mbar_wait(smem_u32(&bars->sdp_empty[b]), ((t >> 1) - 1) & 1);
        tc_fence_after();
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(smem_u32(&bars->k_full[ks]), kph);
          tc_fence_after();
          const uint32_t a_addr = k_ring + ks * stage_bytes;
          const uint32_t b_addr = a_addr + ATOM_BYTES;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4)
            umma_f16(tmem + b * BN, make_desc_k_sw128(a_addr + s4 * 32), make_desc_k_sw128(b_addr + s4 * 32), idesc,
                     (kc | s4) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&bars->k_empty[ks]));
          if (++ks == static_cast<uint32_t>(p.ns)) { ks = 0; kph ^= 1; }
        }
        mbar_wait(smem_u32(&bars->v_full), t & 1);
        tc_fence_after();
        for (int vc = 0; vc < p.vc_count; ++vc) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4)
            umma_f16(tmem + 256 + b * BN, make_desc_k_sw128(do_smem + vc * ATOM_BYTES + s4 * 32),
                     make_desc_k_sw128(v_smem + vc * ATOM_BYTES + s4 * 32), idesc, (vc | s4) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars->v_empty));
        umma_commit(smem_u32(&bars->sdp_full[b]));
      }
    }
    __syncwarp();
  } else {
    const int row = tid;
    const int q = q0 + row;
    const bool row_ok = q < p.Nq;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const float c2 = p.scale_log2;
    // D'[i] = sum_c dO'[i,c] * O[c,i] with the SAME fp16-rounded, row-scaled dO' the dP MMA consumes, so
    // that sum_j P[i,j] (dP'[i,j] - D'[i]) cancels to rounding noise instead of to an fp16-vs-fp32 bias.
    float lse2 = INFINITY, dsum = 0.f, gs = 0.f;
    if (row_ok) {
      lse2 = p.lse[static_cast<size_t>(bidx) * p.Nq + q] * 1.4426950408889634f;
      const __half* drow = p.do16 + (static_cast<size_t>(bidx) * p.Nq + q) * p.Cvk;
      for (int c = 0; c < p.Cv; ++c)
        dsum = fmaf(__half2float(drow[c]), p.out[(static_cast<size_t>(bidx) * p.Cv + c) * p.Nq + q], dsum);
      gs = p.scale / p.rscale[static_cast<size_t>(bidx) * p.Nq + q];  // undo the row scale: true dS in bf16
    }
    const int valid_rows = min(BM, p.Nq - q0);

    for (int j = 0; j < T; ++j) {
      const int b = j & 1;
      const int valid_cols = min(BN, p.Nk - j * BN);
      mbar_wait(smem_u32(&bars->sdp_full[b]), (j >> 1) & 1);
      tc_fence_after();
      float s[BN];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_ld32(tmem + lane_sel + b * BN + c * 32, reinterpret_cast<uint32_t*>(&s[c * 32]));
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < BN; ++c) {
        const float pv = ex2(fmaf(s[c], c2, -lse2));
        s[c] = (c < valid_cols) ? pv : 0.f;
      }
      uint32_t dsp[BN / 2];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_sel + 256 + b * BN + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float d0 = s[c * 32 + i] * (__uint_as_float(r[i]) - dsum) * gs;
          const float d1 = s[c * 32 + i + 1] * (__uint_as_float(r[i + 1]) - dsum) * gs;
          dsp[(c * 32 + i) >> 1] = pack_bf2(d0, d1);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->sdp_empty[b]));

      // stage dS row-major, then transposed, (then P^T) through one smem tile
#pragma unroll
      for (int c = 0; c < 16; ++c)
        *reinterpret_cast<uint4*>(tile + (row * TPITCH + c * 8) * 2) =
            make_uint4(dsp[4 * c], dsp[4 * c + 1], dsp[4 * c + 2], dsp[4 * c + 3]);
      wg_sync();
      copy_tile_out(tile, p.ds + (static_cast<size_t>(bidx) * p.Nq + q0) * p.Nkp + j * BN, p.Nkp, valid_rows,
                    valid_cols, tid);
      wg_sync();
#pragma unroll
      for (int c = 0; c < BN / 2; ++c) {
        *reinterpret_cast<uint16_t*>(tile + ((2 * c) * TPITCH + row) * 2) = static_cast<uint16_t>(dsp[c] & 0xFFFF);
        *reinterpret_cast<uint16_t*>(tile + ((2 * c + 1) * TPITCH + row) * 2) = static_cast<uint16_t>(dsp[c] >> 16);
      }
      wg_sync();
      copy_tile_out(tile, p.dst + (static_cast<size_t>(bidx) * p.Nk + j * BN) * p.Nqp + q0, p.Nqp, valid_cols,
                    valid_rows, tid);
      wg_sync();
      if (p.pt != nullptr) {
#pragma unroll
        for (int c = 0; c < BN / 2; ++c) {
          const uint32_t h = pack_bf2(s[2 * c], s[2 * c + 1]);
          *reinterpret_cast<uint16_t*>(tile + ((2 * c) * TPITCH + row) * 2) = static_cast<uint16_t>(h & 0xFFFF);
          *reinterpret_cast<uint16_t*>(tile + ((2 * c + 1) * TPITCH + row) * 2) = static_cast<uint16_t>(h >> 16);
        }
        wg_sync();
        copy_tile_out(tile, p.pt + (static_cast<size_t>(bidx) * p.Nk + j * BN) * p.Nqp + q0, p.Nqp, valid_cols,
                      valid_rows, tid);
        wg_sync();
      }
    }
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int corr_bwd_ds_launch(const void* q, const void* k, const void* do16, const void* v16, const float* rscale,
                       const float* out, const float* lse, void* ds, void* dst, void* pt, int B, int Nq, int Nk,
                       int Kd, int Cv, int Cvk, int Nkp, int Nqp, float scale, cudaStream_t stream) {
  if (B <= 0 || Nq <= 0 || Nk <= 0 || Kd <= 0 || (Kd % BK) || Cvk <= 0 || (Cvk % BK) || Cv <= 0 || Cv > Cvk) {
    set_error("corr_bwd_ds: bad shape (B=%d Nq=%d Nk=%d Kd=%d Cv=%d Cvk=%d)", B, Nq, Nk, Kd, Cv, Cvk);
    return -1;
  }
  if (Nkp < Nk || Nqp < Nq || (Nkp % 8) || (Nqp % 8)) {
    set_error("corr_bwd_ds: pitches must be >= extent and multiples of 8 (Nkp=%d Nqp=%d)", Nkp, Nqp);
    return -1;
  }
  BwdParams p;
  p.B = B; p.Nq = Nq; p.Nk = Nk; p.Kd = Kd; p.Cv = Cv; p.Cvk = Cvk;
  p.kc_count = Kd / BK; p.vc_count = Cvk / BK; p.n_tiles = (Nk + BN - 1) / BN;
  p.Nkp = Nkp; p.Nqp = Nqp;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.do16 = static_cast<const __half*>(do16); p.rscale = rscale; p.out = out; p.lse = lse;
  p.ds = static_cast<uint16_t*>(ds); p.dst = static_cast<uint16_t*>(dst); p.pt = static_cast<uint16_t*>(pt);
  const int budget = 227 * 1024 - 1024 - 512;
  const int fixed = 2 * p.vc_count * ATOM_BYTES + TILE_BYTES;
  p.ns = (budget - fixed) / (2 * ATOM_BYTES);
  if (p.ns > MAX_STAGES) p.ns = MAX_STAGES;
  if (p.ns < 2) {
    set_error("corr_bwd_ds: shared memory plan failed (Cvk=%d)", Cvk);
    return -1;
  }
  const int smem_bytes = 1024 + p.ns * 2 * ATOM_BYTES + fixed + 512;
  CUtensorMap tm_q, tm_k, tm_do, tm_v;
  int rc;
  if ((rc = make_tmap_f16_3d(&tm_q, q, Kd, Nq, B, (uint64_t)Kd * 2, (uint64_t)Nq * Kd * 2, BK, BM, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_k, k, Kd, Nk, B, (uint64_t)Kd * 2, (uint64_t)Nk * Kd * 2, BK, BN, 1))) return rc;
  if ((rc = make_tmap_f16_3d(&tm_do, do16, Cvk, Nq, B, (uint64_t)Cvk * 2, (uint64_t)Nq * Cvk * 2, BK, BM, 1)))
    return rc;
  if ((rc = make_tmap_f16_3d(&tm_v, v16, Cvk, Nk, B, (uint64_t)Cvk * 2, (uint64_t)Nk * Cvk * 2, BK, BN, 1)))
    return rc;
  COCOS_CUDA_CHECK(
      cudaFuncSetAttribute(corr_bwd_ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid((Nq + BM - 1) / BM, B);
  corr_bwd_ds_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_q, tm_k, tm_do, tm_v, p);
  COCOS_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace cocos
