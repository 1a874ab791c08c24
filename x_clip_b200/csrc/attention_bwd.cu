// Fused multi-head attention backward (dim_head = 64, n <= 320), sm_100a tcgen05.
//
// Backward of the reference's Attention core (x_clip/x_clip.py:217-244) as produced by
// autograd over einsum/softmax/einsum; scores are recomputed on-chip from Q,K and the saved
// per-row log-sum-exp, never read from HBM.
//
// One CTA owns one (batch, head).  Outer loop over key tiles j (128 keys), inner loop over
// query tiles i (128 queries).  Per (j,i) "pair", five tcgen05 MMAs, all fed from the natural
// TMA boxes [128 tokens x 64] of Q, K, V, dO (K-major or MN-major descriptors pick the
// orientation - nothing is transposed in memory):
//   S   = Q_i K_j^T          (queries on TMEM lanes)        [128 x 128]
//   dP  = dO_i V_j^T                                        [128 x 128]
//   threads: P = exp2(S*c - lse_i), dS = P * (dP - delta_i) * scale  -> bf16 in smem
//   dV_j += P^T  dO_i        (A = P  as MN-major, B = dO_i as MN-major)   [128 x 64]
//   dK_j += dS^T Q_i         (A = dS as MN-major, B = Q_i  as MN-major)   [128 x 64]
//   dQ_i  = dS   K_j         (A = dS as K-major,  B = K_j  as MN-major)   [128 x 64]
// dK_j/dV_j accumulate in TMEM across i.  dQ_i accumulates across the (<= 3) key tiles through
// an fp32 workspace that the same thread re-reads (same CTA, same row: no atomics, L2 resident).
//
// 8 compute warps (a query row is shared by two threads, 64 key columns each) + 1 control warp.
// The control thread runs a software pipeline over the CTA's linear stream of pairs: K/V and
// Q/dO tiles are double-buffered and prefetched by TMA, and S/dP of the next pair are issued
// right behind the three gradient MMAs of the current one.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace xclip {

constexpr int kBwdComputeWarps = 8;
constexpr int kBwdThreads = (kBwdComputeWarps + 1) * 32;
constexpr int kBT = 128;
constexpr int kBDh = 64;
constexpr int kBBox = kBT * kBDh * 2;  // 16 KiB

struct AttnBwdParams {
  int B, H, n;
  float scale, scale_log2;
  const uint8_t* mask;   // [B,n] or null
  const float* lse;      // [B,H,n] base-2
  const float* delta;    // [B,H,n] rowsum(dO * O)
  bf16* dqkv;            // [B*n, ld] q | k | v gradients
  long long ld;
  float* dq_ws;          // fp32 [B*n, H*64] or null when n <= 128
  int nt;                // == n (kept as a separate name: tokens covered by the tiles)
};

__device__ __forceinline__ float bwd_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bwd_sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                           uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ float4 bwd_lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void bwd_sts_f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]; one warp per (token, 4 heads at a time)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ d_o,
                  long long lddo, float* __restrict__ delta, int B, int n, int H) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long tokens = (long long)B * n;
  const int groups = (H + 3) / 4;  // 4 heads (256 elements) per warp pass, 8 lanes per head
  for (long long w = warp; w < tokens * groups; w += nwarps) {
    const long long tok = w / groups;
    const int hg = (int)(w % groups);
    const int h = hg * 4 + (lane >> 3);
    float s = 0.f;
    if (h < H) {
      const int col = h * kBDh + (lane & 7) * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + tok * ldo + col);
      const uint4 c = *reinterpret_cast<const uint4*>(d_o + tok * lddo + col);
      const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16x2(aa[i]), y = unpack_bf16x2(cc[i]);
        s += x.x * y.x + x.y * y.y;
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (h < H && (lane & 7) == 0) {
      const long long b = tok / n, i = tok % n;
      delta[(b * H + h) * n + i] = s;
    }
  }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sKV = smem;                  // 2 buffers x (K box, V box)
  uint8_t* sQdO = sKV + 4 * kBBox;      // 2 buffers x (Q box, dO box)
  uint8_t* sP = sQdO + 4 * kBBox;       // 2 blocks of [128 x 64] bf16
  uint8_t* sdS = sP + 2 * kBBox;        // 2 blocks
  uint8_t* tail = sdS + 2 * kBBox;
  uint64_t* kv_bar = reinterpret_cast<uint64_t*>(tail);  // [2]
  uint64_t* qdo_bar = kv_bar + 2;                        // [2]
  uint64_t* s_bar = kv_bar + 4;
  uint64_t* pds_bar = kv_bar + 5;
  uint64_t* g_bar = kv_bar + 6;
  uint64_t* dq_bar = kv_bar + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_bar + 8);
  const uint32_t sMul = smem_u32(tail + 128);  // [128] f32: scale*log2e for attendable keys, else 0
  const uint32_t sAdd = sMul + 128 * 4;        // [128] f32: 0 or -inf (masked / beyond n)
  const uint32_t sLse = sAdd + 128 * 4;        // [384] f32: log-sum-exp of every query of this (b,h)
  const uint32_t sDelta = sLse + 384 * 4;      // [384] f32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool is_control = warp == kBwdComputeWarps;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("xclip attn_bwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(&kv_bar[0], 1);
    mbar_init(&kv_bar[1], 1);
    mbar_init(&qdo_bar[0], 1);
    mbar_init(&qdo_bar[1], 1);
    mbar_init(s_bar, 1);
    mbar_init(pds_bar, kBwdComputeWarps);
    mbar_init(g_bar, 1);
    mbar_init(dq_bar, 1);
    fence_barrier_init();
  }
  if (is_control) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_do);
    }
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256,
                 tdK = tmem_base + 320, tdQ = tmem_base + 384;

  const int ntiles = (p.nt + kBT - 1) / kBT;
  const int pairs_per_bh = ntiles * ntiles;
  const int inner = p.H * kBDh;
  const int num_bh = p.B * p.H;
  // this CTA's work: bh = blockIdx.x + k*gridDim.x, k = 0..my_bh-1; linear pair index pc
  const int my_bh = (num_bh - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int my_pairs = my_bh * pairs_per_bh;

  if (is_control) {
    // ===================== control warp =====================
    if (XCLIP_ONE_LANE(lane)) {
      // Coordinates of a pair are tracked incrementally (no divisions on this single-thread
      // critical path): `cur` is the pair whose gradient MMAs are issued, `nxt` the one whose
      // tiles are prefetched / whose S,dP are issued, `kvn` the next key step to prefetch.
      struct Coord { int bh, j, i; };
      auto advance = [&](Coord& c) {
        if (++c.i == ntiles) { c.i = 0; if (++c.j == ntiles) { c.j = 0; c.bh += gridDim.x; } }
      };
      auto load_kv = [&](const Coord& c, int kc) {
        const int b = c.bh / p.H, h = c.bh - b * p.H;
        uint8_t* dst = sKV + (kc & 1) * 2 * kBBox;
        mbar_arrive_expect_tx(&kv_bar[kc & 1], 2 * kBBox);
        tma_load_3d(dst, &tm_qkv, &kv_bar[kc & 1], inner + h * kBDh, c.j * kBT, b);
        tma_load_3d(dst + kBBox, &tm_qkv, &kv_bar[kc & 1], 2 * inner + h * kBDh, c.j * kBT, b);
      };
      auto load_qdo = [&](const Coord& c, int pc) {
        const int b = c.bh / p.H, h = c.bh - b * p.H;
        uint8_t* dst = sQdO + (pc & 1) * 2 * kBBox;
        mbar_arrive_expect_tx(&qdo_bar[pc & 1], 2 * kBBox);
        tma_load_3d(dst, &tm_qkv, &qdo_bar[pc & 1], h * kBDh, c.i * kBT, b);
        tma_load_3d(dst + kBBox, &tm_do, &qdo_bar[pc & 1], h * kBDh, c.i * kBT, b);
      };
      const uint32_t sKV_a = smem_u32(sKV), sQdO_a = smem_u32(sQdO);
      // descriptors whose operand never moves: P^T / dS^T (MN-major A) and dS (K-major A)
      const uint64_t desc_pT = make_smem_desc(smem_u32(sP), kBBox, 1024);
      const uint64_t desc_dsT = make_smem_desc(smem_u32(sdS), kBBox, 1024);
      const uint64_t desc_dsK = make_smem_desc(smem_u32(sdS), 0, 1024);
      auto issue_scores = [&](const Coord& c, int pc, int kc) {  // S and dP of pair pc
        // only the key columns that exist (rounded to 32) are produced for a partial key tile
        const int vc = min(kBT, (p.nt - c.j * kBT + 31) / 32 * 32);
        const uint32_t idesc = make_idesc_bf16(kBT, vc, kMajorK, kMajorK);
        const uint32_t kv = sKV_a + (kc & 1) * 2 * kBBox;
        const uint32_t qd_ = sQdO_a + (pc & 1) * 2 * kBBox;
        const uint64_t qd = make_smem_desc(qd_, 0, 1024);
        const uint64_t kd = make_smem_desc(kv, 0, 1024);
        const uint64_t dod = make_smem_desc(qd_ + kBBox, 0, 1024);
        const uint64_t vd = make_smem_desc(kv + kBBox, 0, 1024);
#pragma unroll
        for (int k = 0; k < kBDh / 16; ++k) umma_bf16(tS, qd + 2 * k, kd + 2 * k, idesc, k > 0);
#pragma unroll
        for (int k = 0; k < kBDh / 16; ++k) umma_bf16(tdP, dod + 2 * k, vd + 2 * k, idesc, k > 0);
        umma_commit(s_bar);
      };

      Coord cur{(int)blockIdx.x, 0, 0}, nxt = cur, kvn = cur;
      if (my_pairs > 0) {
        load_kv(cur, 0);
        load_qdo(cur, 0);
        mbar_wait(&kv_bar[0], 0);
        mbar_wait(&qdo_bar[0], 0);
        tcgen05_fence_after();
        issue_scores(cur, 0, 0);
        advance(nxt);
        kvn.j = 1;
        if (kvn.j == ntiles) { kvn.j = 0; kvn.bh += gridDim.x; }
      }
      int kc = 0;                       // running key-step id of `cur`
      for (int pc = 0; pc < my_pairs; ++pc) {
        const int i = cur.i, j = cur.j;
        const int ksteps_q = min(kBT, (p.nt - i * kBT + 15) / 16 * 16) / 16;   // valid query groups
        const int ksteps_k = min(kBT, (p.nt - j * kBT + 31) / 32 * 32) / 16;   // valid key groups
        const bool has_next = pc + 1 < my_pairs;
        // all MMAs of pair pc-1 retired: its Q/dO buffer and (if it closed a key step) the
        // K/V buffer of that step may be overwritten
        if (pc >= 1) mbar_wait(g_bar, (pc - 1) & 1);
        if (i == 0 && (kc + 1) * ntiles < my_pairs) {   // prefetch K/V of the next key step
          load_kv(kvn, kc + 1);
          if (++kvn.j == ntiles) { kvn.j = 0; kvn.bh += gridDim.x; }
        }
        if (has_next) load_qdo(nxt, pc + 1);

        mbar_wait(pds_bar, pc & 1);   // P, dS of pair pc are in smem; S/dP TMEM consumed
        tcgen05_fence_after();
        {
          constexpr uint32_t idesc_t = make_idesc_bf16(kBT, kBDh, kMajorMN, kMajorMN);
          constexpr uint32_t idesc_q = make_idesc_bf16(kBT, kBDh, kMajorK, kMajorMN);
          const uint32_t kv = sKV_a + (kc & 1) * 2 * kBBox;
          const uint32_t qd_ = sQdO_a + (pc & 1) * 2 * kBBox;
          const uint64_t desc_kmn = make_smem_desc(kv, 8192, 1024);          // K_j, MN-major B
          const uint64_t desc_qmn = make_smem_desc(qd_, 8192, 1024);         // Q_i, MN-major B
          const uint64_t desc_domn = make_smem_desc(qd_ + kBBox, 8192, 1024);  // dO_i, MN-major B
          // dQ first (contraction over the valid keys): its epilogue overlaps the dV/dK MMAs.
          // Fully unrolled with compile-time offsets; descriptor address units are 16 bytes.
#pragma unroll
          for (int k = 0; k < kBT / 16; ++k) {
            if (k < ksteps_k)
              umma_bf16(tdQ, desc_dsK + ((k >> 2) * (kBBox >> 4) + (k & 3) * 2),
                        desc_kmn + k * 128, idesc_q, k > 0 ? 1u : 0u);
          }
          umma_commit(dq_bar);
#pragma unroll
          for (int k = 0; k < kBT / 16; ++k) {   // contraction over the valid queries
            if (k < ksteps_q) {
              umma_bf16(tdV, desc_pT + k * 128, desc_domn + k * 128, idesc_t,
                        (i > 0 || k > 0) ? 1u : 0u);
              umma_bf16(tdK, desc_dsT + k * 128, desc_qmn + k * 128, idesc_t,
                        (i > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(g_bar);
        if (has_next) {               // S/dP of the next pair queue right behind
          const int nkc = (nxt.i == 0) ? kc + 1 : kc;
          if (nxt.i == 0) mbar_wait(&kv_bar[nkc & 1], (nkc >> 1) & 1);
          mbar_wait(&qdo_bar[(pc + 1) & 1], ((pc + 1) >> 1) & 1);
          tcgen05_fence_after();
          issue_scores(nxt, pc + 1, nkc);
        }
        advance(cur);
        advance(nxt);
        if (cur.i == 0) ++kc;
      }
      if (my_pairs > 0) mbar_wait(g_bar, (my_pairs - 1) & 1);
    }
    __syncwarp();
  } else {
    // ===================== compute warps =====================
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    int pc = 0;
    for (int bh = blockIdx.x; bh < num_bh; bh += gridDim.x) {
      const int b = bh / p.H, h = bh % p.H;
      for (int j = 0; j < ntiles; ++j) {
        // key-tile tables (all compute threads are past every read of the previous tables:
        // the last read precedes their pds arrive of the previous pair, and bar.sync orders)
        if (threadIdx.x < kBT) {
          const int key = j * kBT + threadIdx.x;
          const bool keep =
              key < p.nt && (p.mask ? (p.mask[(long long)b * p.n + key] != 0) : true);
          bwd_sts_f(sMul + threadIdx.x * 4, keep ? p.scale_log2 : 0.f);
          bwd_sts_f(sAdd + threadIdx.x * 4, keep ? 0.f : -INFINITY);
        }
        if (j == 0) {   // per-(b,h) row statistics: +inf lse -> p = 0 for rows beyond n
          for (int q = threadIdx.x; q < 384; q += kBwdComputeWarps * 32) {
            const long long s_idx = ((long long)b * p.H + h) * p.n + q;
            bwd_sts_f(sLse + q * 4, q < p.nt ? p.lse[s_idx] : INFINITY);
            bwd_sts_f(sDelta + q * 4, q < p.nt ? p.delta[s_idx] : 0.f);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");

        for (int i = 0; i < ntiles; ++i, ++pc) {
          const int q_idx = i * kBT + row;
          const bool q_ok = q_idx < p.nt;
          float lse_i, delta_i;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lse_i) : "r"(sLse + q_idx * 4));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(delta_i) : "r"(sDelta + q_idx * 4));
          mbar_wait(s_bar, pc & 1);
          tcgen05_fence_after();
          // P/dS smem of the previous pair is still read by its dV/dK MMAs until g_bar fires
          if (pc >= 1) mbar_wait(g_bar, (pc - 1) & 1);
          // Partial tiles: the MMAs only touch query groups < vr16 and key columns < vc32 (see the
          // control warp), and every output row depends on its own operand row only, so warps /
          // chunks that are pure padding skip their math and leave their smem slots untouched.
          const int vr16 = min(kBT, (p.nt - i * kBT + 15) / 16 * 16);
          const int vc32 = min(kBT, (p.nt - j * kBT + 31) / 32 * 32);
          const bool warp_alive = quarter * 32 < vr16;
#pragma unroll
          for (int cc0 = 0; cc0 < 2; ++cc0) {
            const int c0 = half * 64 + cc0 * 32;
            if (!warp_alive || c0 >= vc32) continue;
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(tS + lane_off + c0, sv);
            tmem_ld_32x32(tdP + lane_off + c0, dv);
            tmem_ld_wait();
            float pr[32], ds[32];
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 m = bwd_lds_f4(sMul + (c0 + e) * 4);
              const float4 a = bwd_lds_f4(sAdd + (c0 + e) * 4);
              const float mm[4] = {m.x, m.y, m.z, m.w}, aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float t = fmaf(__uint_as_float(sv[e + u]), mm[u], aa[u]);
                const float pv = bwd_ex2(t - lse_i);
                pr[e + u] = pv;
                ds[e + u] = pv * (__uint_as_float(dv[e + u]) - delta_i) * p.scale;
              }
            }
            const uint32_t pblk = smem_u32(sP) + (c0 >> 6) * kBBox;
            const uint32_t dblk = smem_u32(sdS) + (c0 >> 6) * kBBox;
            const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              bwd_sts_v4(pblk + swz128(row, chunk0 + cc),
                         pack_bf16x2(pr[cc * 8 + 0], pr[cc * 8 + 1]),
                         pack_bf16x2(pr[cc * 8 + 2], pr[cc * 8 + 3]),
                         pack_bf16x2(pr[cc * 8 + 4], pr[cc * 8 + 5]),
                         pack_bf16x2(pr[cc * 8 + 6], pr[cc * 8 + 7]));
              bwd_sts_v4(dblk + swz128(row, chunk0 + cc),
                         pack_bf16x2(ds[cc * 8 + 0], ds[cc * 8 + 1]),
                         pack_bf16x2(ds[cc * 8 + 2], ds[cc * 8 + 3]),
                         pack_bf16x2(ds[cc * 8 + 4], ds[cc * 8 + 5]),
                         pack_bf16x2(ds[cc * 8 + 6], ds[cc * 8 + 7]));
            }
          }
          fence_proxy_async_smem();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pds_bar);

          // dQ_i partial for this key tile: this thread owns 32 of the 64 columns of its row.
          // The fp32 partial of the previous key tiles is fetched BEFORE waiting for the MMAs.
          const long long tok = (long long)b * p.n + (q_ok ? q_idx : 0);
          float* ws = p.dq_ws ? p.dq_ws + tok * inner + h * kBDh + half * 32 : nullptr;
          float4 prev[8];
          if (j > 0 && q_ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) prev[e] = *reinterpret_cast<const float4*>(ws + e * 4);
          }
          mbar_wait(dq_bar, pc & 1);
          tcgen05_fence_after();
          {
            uint32_t v[32];
            tmem_ld_32x32(tdQ + lane_off + half * 32, v);
            tmem_ld_wait();
            if (q_ok) {
              float f[32];
#pragma unroll
              for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
              if (j > 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  f[e * 4] += prev[e].x; f[e * 4 + 1] += prev[e].y;
                  f[e * 4 + 2] += prev[e].z; f[e * 4 + 3] += prev[e].w;
                }
              }
              if (j < ntiles - 1) {
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                  *reinterpret_cast<float4*>(ws + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
              } else {
                bf16* dst = p.dqkv + tok * p.ld + h * kBDh + half * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                  uint4 o;
                  o.x = pack_bf16x2(f[e], f[e + 1]);
                  o.y = pack_bf16x2(f[e + 2], f[e + 3]);
                  o.z = pack_bf16x2(f[e + 4], f[e + 5]);
                  o.w = pack_bf16x2(f[e + 6], f[e + 7]);
                  *reinterpret_cast<uint4*>(dst + e) = o;
                }
              }
            }
          }
          tcgen05_fence_before();
        }  // i

        // dK_j (half 0) / dV_j (half 1) are complete once the last pair's dV/dK MMAs retired
        {
          const int key = j * kBT + row;
          mbar_wait(g_bar, (pc - 1) & 1);
          tcgen05_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32((half == 0 ? tdK : tdV) + lane_off + c * 32, v);
            tmem_ld_wait();
            if (key < p.nt) {
              bf16* dst = p.dqkv + ((long long)b * p.n + key) * p.ld + (half + 1) * inner +
                          h * kBDh + c * 32;
#pragma unroll
              for (int e = 0; e < 32; e += 8) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(v[e]), __uint_as_float(v[e + 1]));
                o.y = pack_bf16x2(__uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                o.z = pack_bf16x2(__uint_as_float(v[e + 4]), __uint_as_float(v[e + 5]));
                o.w = pack_bf16x2(__uint_as_float(v[e + 6]), __uint_as_float(v[e + 7]));
                *reinterpret_cast<uint4*>(dst + e) = o;
              }
            }
          }
          tcgen05_fence_before();
        }
      }  // j
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (is_control) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace xclip

namespace xclip {
int attn_bwd_small(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask, const void* d_o,
                   int64_t lddo, const float* lse, const float* delta, void* dqkv, int64_t ld_dqkv,
                   int B, int n, int heads, float scale, int causal, cudaStream_t stream);
}
using namespace xclip;

extern "C" int xclip_attn_bwd(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask,
                              const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                              const float* lse, float* delta, void* dqkv, int64_t ld_dqkv,
                              float* dq_workspace, int B, int n, int heads, float scale,
                              int causal, xclip_stream_t stream) {
  int rc = xclip_init();
  if (rc) return rc;
  XCLIP_REQUIRE(qkv && o && d_o && lse && delta && dqkv, "attn_bwd: null pointer");
  XCLIP_REQUIRE(B > 0 && heads > 0 && n > 0 && n <= 320, "attn_bwd: bad sizes B=%d n=%d heads=%d",
                B, n, heads);
  XCLIP_REQUIRE(n <= 128 || dq_workspace != nullptr,
                "attn_bwd: n=%d > 128 needs the fp32 dq workspace [B*n, heads*64]", n);
  XCLIP_REQUIRE(ld_qkv % 8 == 0 && ld_qkv >= 3 * heads * kBDh && ld_dqkv % 8 == 0 &&
                    ld_dqkv >= 3 * heads * kBDh && ldo % 8 == 0 && lddo % 8 == 0 &&
                    ldo >= heads * kBDh && lddo >= heads * kBDh,
                "attn_bwd: bad leading dimensions");
  XCLIP_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(o) |
                  reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dqkv) |
                  reinterpret_cast<uintptr_t>(dq_workspace)) & 15) == 0,
                "attn_bwd: misaligned pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);

  {
    const long long warps = (long long)B * n * ((heads + 3) / 4);
    long long blocks = (warps + 7) / 8;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    attn_delta_kernel<<<(int)blocks, 256, 0, s>>>((const bf16*)o, ldo, (const bf16*)d_o, lddo,
                                                  delta, B, n, heads);
    XCLIP_LAUNCH_CHECK("attn_delta_kernel");
  }
  if (n <= kBT)
    return attn_bwd_small(qkv, ld_qkv, key_mask, d_o, lddo, lse, delta, dqkv, ld_dqkv, B, n, heads,
                          scale, causal, s);
  XCLIP_REQUIRE(!causal, "attn_bwd: the causal mask is only implemented for n <= 128 (got n=%d)", n);

  AttnBwdParams p;
  p.B = B; p.H = heads; p.n = n;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = key_mask; p.lse = lse; p.delta = delta;
  p.dqkv = reinterpret_cast<bf16*>(dqkv); p.ld = ld_dqkv; p.dq_ws = dq_workspace;
  p.nt = n;

  CUtensorMap tq, tdo;
  rc = encode_3d_bf16(&tq, qkv, (uint64_t)(3 * heads * kBDh), (uint64_t)n, (uint64_t)B,
                      (uint64_t)ld_qkv, (uint64_t)n * ld_qkv, kBDh, kBT);
  if (rc) return rc;
  rc = encode_3d_bf16(&tdo, d_o, (uint64_t)(heads * kBDh), (uint64_t)n, (uint64_t)B,
                      (uint64_t)lddo, (uint64_t)n * lddo, kBDh, kBT);
  if (rc) return rc;

  const int smem = 12 * kBBox + 128 + 2 * 128 * 4 + 2 * 384 * 4;
  rc = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_kernel), smem);
  if (rc) return rc;
  long long grid = num_sms();
  if (grid > (long long)B * heads) grid = (long long)B * heads;
  attn_bwd_kernel<<<(int)grid, kBwdThreads, smem, s>>>(tq, tdo, p);
  XCLIP_LAUNCH_CHECK("attn_bwd_kernel");
  return XCLIP_OK;
}
