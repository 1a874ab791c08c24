// Fused multi-head attention forward for 128 < n <= 320 (dim_head = 64), sm_100a tcgen05.
// (n <= 128 is handled by attention_small.cu.)
//
// Replaces the reference's Attention core, x_clip/x_clip.py:217-244:
//   split heads -> q * dh^-0.5 -> einsum QK^T -> masked_fill(~key_mask, -finfo.max) ->
//   softmax(fp32) -> einsum PV -> merge heads
// which materialises [B,h,n,n] scores in HBM; here scores never leave the SM.
//
// One CTA owns one (batch, head) at a time: K and V of that head (<= 3 TMA boxes of 128 tokens)
// stay in shared memory while the CTA walks the query tiles of 128 rows.  The keys of a tile are
// split in two blocks (w0 + w1 columns) that are processed CONCURRENTLY by two groups of softmax
// warps on two S buffers in TMEM; every block keeps its own max / sum and accumulates
// P_blk V_blk into its own 64-column O buffer taken from a ring of three - nothing is rescaled
// in TMEM, the epilogue combines  O = (2^(m0-m) O0 + 2^(m1-m) O1) / (2^(m0-m) l0 + 2^(m1-m) l1).
// 16 softmax warps (4 per scheduler): two warps share a row of a block (column halves) and
// exchange the row max through a 64-thread named barrier; 16-key chunks without masked keys skip
// the mask tables (p = 2^(s*c - m): one FFMA + EX2).  The issue thread keeps the next block's
// S = Q K_blk^T in flight, prefetches K+Q of the next item as soon as the item's last S retired
// and V after its last PV.  TMEM: S 2 x 160 + O 3 x 64 = 512 columns; smem: Q 16 + K 48 + V 48 +
// 2 P buffers x 48 KiB.
//
// Masking follows the reference exactly: a masked key's score is replaced by -FLT_MAX AFTER
// scaling (so a fully masked row would give uniform attention); keys beyond n do not exist.
// Of the ten forward variants written in round 1 this is the one measured fastest on a B200
// (round 2, B=1024 n=257: 0.60 ms vs 0.91 ms for the 8-warp ping-pong kernel); the others and
// the CUDA-core tail-token path (no gain once this kernel was in, backward tail slower and
// failing parity) were deleted.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace xclip {

constexpr int kTile = 128;
constexpr int kDh = 64;
constexpr int kBoxBytes = kTile * kDh * 2;  // 16 KiB: one [128 x 64] bf16 TMA box

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                       uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct AttnPPParams {
  int B, H, n, nkp, w0, w1;
  int q_tiles;           // 128-query tiles per (batch, head)
  float scale_log2;
  const uint8_t* mask;
  bf16* o;
  long long ldo;
  float* lse;
};

constexpr int kPPSub = 3;                       // 64-key sub-blocks per P buffer (block <= 160 keys)
constexpr int kPPBuf = kPPSub * kBoxBytes;      // 48 KiB

// kVariant bit 0: streaming softmax (S read twice from TMEM) instead of the register-resident block;
// bit 1: the row-max exchange synchronises only the two warps sharing a row (named barrier
// 2 + quarter, 64 threads) instead of all 8 softmax warps.
// ---- softmax passes over 16 / 32 columns of a row ------------------------------------------
template <int N, bool kFastPath>
__device__ __forceinline__ void wg_pass_max(const uint32_t (&w)[N], uint32_t ma, uint32_t aa,
                                            float& m_scaled, float& m_raw) {
  if constexpr (kFastPath) {
    float m = m_raw;
#pragma unroll
    for (int i = 0; i < N; i += 4)
      m = fmaxf(fmaxf(m, fmaxf(__uint_as_float(w[i]), __uint_as_float(w[i + 1]))),
                fmaxf(__uint_as_float(w[i + 2]), __uint_as_float(w[i + 3])));
    m_raw = m;
  } else {
    float m = m_scaled;
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 mm = lds_f4(ma + i * 4), ad = lds_f4(aa + i * 4);
      const float t0 = fmaf(__uint_as_float(w[i]), mm.x, ad.x);
      const float t1 = fmaf(__uint_as_float(w[i + 1]), mm.y, ad.y);
      const float t2 = fmaf(__uint_as_float(w[i + 2]), mm.z, ad.z);
      const float t3 = fmaf(__uint_as_float(w[i + 3]), mm.w, ad.w);
      m = fmaxf(fmaxf(m, fmaxf(t0, t1)), fmaxf(t2, t3));
    }
    m_scaled = m;
  }
}
// exp2 of N (16 or 32) columns starting at block column `col`, bf16 P into the swizzled buffer
template <int N, bool kFastPath>
__device__ __forceinline__ float wg_pass_exp(const uint32_t (&w)[N], uint32_t ma, uint32_t aa, float m,
                                             float c, uint32_t pbuf, int row, int col) {
  float sum = 0.f;
  const uint32_t blk = pbuf + (col >> 6) * kBoxBytes;
  const int chunk0 = (col & 63) >> 3;
  const float neg_m = -m;
#pragma unroll
  for (int i = 0; i < N; i += 8) {
    float e[8];
    if constexpr (kFastPath) {
#pragma unroll
      for (int u = 0; u < 8; ++u) e[u] = ex2_approx(fmaf(__uint_as_float(w[i + u]), c, neg_m));
    } else {
      const float4 m0 = lds_f4(ma + i * 4), a0 = lds_f4(aa + i * 4);
      const float4 m1 = lds_f4(ma + (i + 4) * 4), a1 = lds_f4(aa + (i + 4) * 4);
      e[0] = ex2_approx(fmaf(__uint_as_float(w[i]), m0.x, a0.x) - m);
      e[1] = ex2_approx(fmaf(__uint_as_float(w[i + 1]), m0.y, a0.y) - m);
      e[2] = ex2_approx(fmaf(__uint_as_float(w[i + 2]), m0.z, a0.z) - m);
      e[3] = ex2_approx(fmaf(__uint_as_float(w[i + 3]), m0.w, a0.w) - m);
      e[4] = ex2_approx(fmaf(__uint_as_float(w[i + 4]), m1.x, a1.x) - m);
      e[5] = ex2_approx(fmaf(__uint_as_float(w[i + 5]), m1.y, a1.y) - m);
      e[6] = ex2_approx(fmaf(__uint_as_float(w[i + 6]), m1.z, a1.z) - m);
      e[7] = ex2_approx(fmaf(__uint_as_float(w[i + 7]), m1.w, a1.w) - m);
    }
    sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
    sts_v4(blk + swz128(row, chunk0 + (i >> 3)), pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]),
           pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
  }
  return sum;
}

constexpr int kWgTableCols = 320;   // n <= 320 -> nkp <= 320
constexpr int kWgTailBytes = 128 + 2 * kWgTableCols * 4 + 128 + 8 * 128 * 4 + 16 * 128 * 4 + 8 * 128 * 4;

template <int kHalves, bool kFast>
__global__ void __launch_bounds__((8 * kHalves + 1) * 32, 1)
attn_fwd_wg_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnPPParams p) {
  constexpr int kWarps = 8 * kHalves;            // softmax warps; warp kWarps is the control warp
  constexpr int kOCols = 64 / (2 * kHalves);     // O columns per thread in the epilogue
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = (p.n + kTile - 1) / kTile;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kBoxBytes;
  uint8_t* sV = sK + nkb * kBoxBytes;
  uint8_t* sP = sV + nkb * kBoxBytes;           // 2 buffers (one per key block)
  uint8_t* tail = sP + 2 * kPPBuf;
  uint64_t* k_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_bar = k_bar + 1;
  uint64_t* s_bar = k_bar + 2;     // [2]
  uint64_t* p_bar = k_bar + 4;     // [2], 4*kHalves arrivals each (the warps working on one block)
  uint64_t* o_bar = k_bar + 6;     // [3]
  uint64_t* v_bar = k_bar + 9;
  uint64_t* e_bar = k_bar + 10;    // [2] by tile parity, kWarps arrivals: epilogue of the tile done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k_bar + 12);
  const uint32_t sMul = smem_u32(tail + 128);               // [320] f32
  const uint32_t sAdd = sMul + kWgTableCols * 4;            // [320] f32
  const uint32_t sClean = sAdd + kWgTableCols * 4;          // [32] u32: 16-key chunk has no masked key
  const uint32_t sStatM = sClean + 128;                     // [8 block slots][128] f32 block max
  const uint32_t sStatL = sStatM + 8 * 128 * 4;             // [8 block slots][2 halves][128] f32 sums
  const uint32_t sMax = sStatL + 16 * 128 * 4;              // [2 tile parities][2 blocks][2 halves][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool is_control = warp == kWarps;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(k_bar, 1);
    mbar_init(v_bar, 1);
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_bar[i], 1);
      mbar_init(&p_bar[i], 4 * kHalves);
      mbar_init(&e_bar[i], kWarps);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&o_bar[i], 1);
    fence_barrier_init();
  }
  if (is_control) {
    if (lane == 0) tma_prefetch_desc(&tm_qkv);
    tmem_alloc<512>(tmem_slot);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int T = p.q_tiles;
  const int inner = p.H * kDh;

  if (is_control) {
    if (XCLIP_ONE_LANE(lane)) {
      const uint64_t desc_q = make_smem_desc(smem_u32(sQ), 0, 1024);
      const uint64_t desc_k = make_smem_desc(smem_u32(sK), 0, 1024);
      const uint64_t desc_v = make_smem_desc(smem_u32(sV), 8192, 1024);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kDh, kMajorK, kMajorMN);
      const uint32_t idesc_s0 = make_idesc_bf16(kTile, p.w0, kMajorK, kMajorK);
      const uint32_t idesc_s1 = make_idesc_bf16(kTile, p.w1, kMajorK, kMajorK);
      auto issue_pv = [&](uint32_t gb) {        // O[gb % 3] = P(gb) V_blk(gb)
        const int kb = gb & 1;
        const int k0 = kb ? p.w0 : 0, W = kb ? p.w1 : p.w0;
        mbar_wait(&p_bar[gb & 1], (gb >> 1) & 1);
        if (gb >= 3) {                           // ring slot gb%3 held block gb-3: its tile's
          const uint32_t te = (gb - 3) >> 1;     // epilogue must have read it
          mbar_wait(&e_bar[te & 1], (te >> 1) & 1);
        }
        tcgen05_fence_after();
        const uint64_t pd = make_smem_desc(smem_u32(sP) + (gb & 1) * kPPBuf, 0, 1024);
        const uint64_t vd = desc_v + ((k0 * 128) >> 4);
        const uint32_t td = tmem_base + 320 + (gb % 3) * kDh;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          if (k < W / 16)
            umma_bf16(td, pd + ((k >> 2) * (kBoxBytes >> 4) + (k & 3) * 2), vd + k * 128, idesc_pv,
                      k > 0 ? 1u : 0u);
        }
        umma_commit(&o_bar[gb % 3]);
      };
      auto load_kq = [&](int bh2) {
        const int b2 = bh2 / p.H, h2 = bh2 - b2 * p.H;
        mbar_arrive_expect_tx(k_bar, nkb * kBoxBytes);
        for (int i = 0; i < nkb; ++i)
          tma_load_3d(sK + i * kBoxBytes, &tm_qkv, k_bar, inner + h2 * kDh, i * kTile, b2);
        mbar_arrive_expect_tx(q_bar, kBoxBytes);
        tma_load_3d(sQ, &tm_qkv, q_bar, h2 * kDh, 0, b2);
      };
      auto load_v = [&](int bh2) {
        const int b2 = bh2 / p.H, h2 = bh2 - b2 * p.H;
        mbar_arrive_expect_tx(v_bar, nkb * kBoxBytes);
        for (int i = 0; i < nkb; ++i)
          tma_load_3d(sV + i * kBoxBytes, &tm_qkv, v_bar, 2 * inner + h2 * kDh, i * kTile, b2);
      };
      uint32_t g = 0, tt = 0, kvc = 0;
      const int total = p.B * p.H;
      if ((int)blockIdx.x < total) { load_kq(blockIdx.x); load_v(blockIdx.x); }
      for (int bh = blockIdx.x; bh < total; bh += gridDim.x, ++kvc) {
        const int b = bh / p.H, h = bh - b * p.H;
        const int bh_next = bh + gridDim.x;
        mbar_wait(k_bar, kvc & 1);
        for (int t = 0; t < T; ++t) {
          for (int kb = 0; kb < 2; ++kb, ++g) {
            const int k0 = kb ? p.w0 : 0;
            if (kb == 0) { mbar_wait(q_bar, tt & 1); ++tt; }
            if (g >= 2) mbar_wait(&p_bar[g & 1], ((g - 2) >> 1) & 1);   // S[g&1] consumed
            tcgen05_fence_after();
            {
              const uint64_t kd = desc_k + ((k0 * 128) >> 4);
              const uint32_t ts = tmem_base + (g & 1) * 160;
              const uint32_t idesc = kb ? idesc_s1 : idesc_s0;
#pragma unroll
              for (int k = 0; k < kDh / 16; ++k)
                umma_bf16(ts, desc_q + 2 * k, kd + 2 * k, idesc, k > 0 ? 1u : 0u);
              umma_commit(&s_bar[g & 1]);
            }
            if (kb == 1) {
              if (t + 1 < T) {
                mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
                mbar_arrive_expect_tx(q_bar, kBoxBytes);
                tma_load_3d(sQ, &tm_qkv, q_bar, h * kDh, (t + 1) * kTile, b);
              } else if (bh_next < total) {
                mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
                load_kq(bh_next);
              }
            }
            if (t == 0 && kb == 1) mbar_wait(v_bar, kvc & 1);
            if (!(t == 0 && kb == 0)) issue_pv(g - 1);
          }
        }
        issue_pv(g - 1);
        mbar_wait(&o_bar[(g - 1) % 3], ((g - 1) / 3) & 1);
        if (bh_next < total) load_v(bh_next);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue warps =====================
    const int grp = warp >> 2, quarter = warp & 3;
    const int blk = kHalves == 2 ? (grp >> 1) : grp;          // key block == S / P buffer
    const int half = kHalves == 2 ? (grp & 1) : 0;            // column half inside the block
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int k0 = blk ? p.w0 : 0, W = blk ? p.w1 : p.w0;
    const int nch = W / 16;                                   // 16-key chunks of the block
    const int cb = kHalves == 2 ? (half ? (nch + 1) / 2 : 0) : 0;
    const int ce = kHalves == 2 ? (half ? nch : (nch + 1) / 2) : nch;
    const uint32_t ts = tmem_base + blk * 160 + lane_off;
    const uint32_t pbuf = smem_u32(sP) + blk * kPPBuf;
    const uint32_t ma0 = sMul + k0 * 4, aa0 = sAdd + k0 * 4;
    const uint32_t clean0 = sClean + (k0 >> 4) * 4;
    const int ocol = grp * kOCols;
    const float c_log2 = p.scale_log2;
    uint32_t gt = 0;                                          // global tile counter of this CTA
    for (int bh = blockIdx.x; bh < p.B * p.H; bh += gridDim.x) {
      const int b = bh / p.H, h = bh - b * p.H;
      // mask tables + per-chunk "no masked key" flags.  j advances by whole warps, so the ballot
      // below always sees 32 consecutive key columns with a warp-uniform trip count.
      for (int j = threadIdx.x; j < kWgTableCols; j += kWarps * 32) {
        float mul = 0.f, add = -INFINITY;
        bool clean = false;
        if (j < p.n) {
          const bool keep = p.mask ? (p.mask[(long long)b * p.n + j] != 0) : true;
          mul = keep ? p.scale_log2 : 0.f;
          add = keep ? 0.f : -FLT_MAX;
          clean = keep;
        }
        sts_f(sMul + j * 4, mul);
        sts_f(sAdd + j * 4, add);
        const uint32_t bal = __ballot_sync(0xffffffffu, clean);
        if (lane == 0) {
          sts_u32(sClean + (j >> 4) * 4, (bal & 0xffffu) == 0xffffu ? 1u : 0u);
          sts_u32(sClean + ((j >> 4) + 1) * 4, (bal >> 16) == 0xffffu ? 1u : 0u);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kWarps * 32) : "memory");

      auto epilogue = [&](int t, uint32_t gtile) {
        const uint32_t g0 = 2 * gtile, g1 = g0 + 1;
        mbar_wait(&o_bar[g0 % 3], (g0 / 3) & 1);
        mbar_wait(&o_bar[g1 % 3], (g1 / 3) & 1);
        tcgen05_fence_after();
        const int q_idx = t * kTile + row;
        const float m0 = lds_f(sStatM + ((g0 & 7) * 128 + row) * 4);
        const float m1 = lds_f(sStatM + ((g1 & 7) * 128 + row) * 4);
        float l0 = lds_f(sStatL + (((g0 & 7) * 2) * 128 + row) * 4);
        float l1 = lds_f(sStatL + (((g1 & 7) * 2) * 128 + row) * 4);
        if constexpr (kHalves == 2) {
          l0 += lds_f(sStatL + (((g0 & 7) * 2 + 1) * 128 + row) * 4);
          l1 += lds_f(sStatL + (((g1 & 7) * 2 + 1) * 128 + row) * 4);
        }
        const float m = fmaxf(m0, m1);
        const float a0 = ex2_approx(m0 - m), a1 = ex2_approx(m1 - m);
        const float L = a0 * l0 + a1 * l1;
        const float inv = 1.f / L;
        uint32_t v0[kOCols], v1[kOCols];
        if constexpr (kOCols == 32) {
          tmem_ld_32x32(tmem_base + 320 + (g0 % 3) * kDh + lane_off + ocol, v0);
          tmem_ld_32x32(tmem_base + 320 + (g1 % 3) * kDh + lane_off + ocol, v1);
        } else {
          tmem_ld_32x16(tmem_base + 320 + (g0 % 3) * kDh + lane_off + ocol, v0);
          tmem_ld_32x16(tmem_base + 320 + (g1 % 3) * kDh + lane_off + ocol, v1);
        }
        tmem_ld_wait();
        // both O buffers are in registers: release the ring slots before the global stores
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_bar[gtile & 1]);
        if (q_idx < p.n) {
          if (grp == 0) p.lse[((long long)b * p.H + h) * p.n + q_idx] = m + log2f(L);
          bf16* dst = p.o + ((long long)b * p.n + q_idx) * p.ldo + h * kDh + ocol;
          const float c0 = a0 * inv, c1 = a1 * inv;
#pragma unroll
          for (int i = 0; i < kOCols; i += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              f[e] = __uint_as_float(v0[i + e]) * c0 + __uint_as_float(v1[i + e]) * c1;
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(dst + i) = o;
          }
        }
      };

      for (int t = 0; t < T; ++t, ++gt) {
        const uint32_t gb = 2 * gt + blk;                  // this warp's block of tile t
        const bool warp_alive = t * kTile + quarter * 32 < p.n;
        mbar_wait(&s_bar[blk], gt & 1);
        tcgen05_fence_after();
        float m_scaled = -INFINITY, m_raw = -INFINITY, sum = 0.f;
        if (warp_alive) {
          for (int c = cb; c < ce;) {
            if (kHalves == 1 && c + 2 <= ce) {
              uint32_t w[32];
              tmem_ld_32x32(ts + c * 16, w);
              const bool fast = kFast && lds_u32(clean0 + c * 4) != 0 && lds_u32(clean0 + (c + 1) * 4) != 0;
              tmem_ld_wait();
              if (fast) wg_pass_max<32, true>(w, 0, 0, m_scaled, m_raw);
              else wg_pass_max<32, false>(w, ma0 + c * 64, aa0 + c * 64, m_scaled, m_raw);
              c += 2;
            } else {
              uint32_t w[16];
              tmem_ld_32x16(ts + c * 16, w);
              const bool fast = kFast && lds_u32(clean0 + c * 4) != 0;
              tmem_ld_wait();
              if (fast) wg_pass_max<16, true>(w, 0, 0, m_scaled, m_raw);
              else wg_pass_max<16, false>(w, ma0 + c * 64, aa0 + c * 64, m_scaled, m_raw);
              c += 1;
            }
          }
        }
        // scale > 0 (checked by the launcher for the fast variants): max commutes with the scaling
        float m2 = kFast ? fmaxf(m_scaled, m_raw * c_log2) : m_scaled;
        if constexpr (kHalves == 2) {
          const uint32_t slot = sMax + ((((gt & 1) * 2 + blk) * 2) * 128 + row) * 4;
          sts_f(slot + half * 128 * 4, m2);
          asm volatile("bar.sync %0, 64;" ::"r"(2 + blk * 4 + quarter) : "memory");
          m2 = fmaxf(m2, lds_f(slot + (half ^ 1) * 128 * 4));
        }
        // the block's P buffer was last read by PV(gb-2)
        if (gb >= 2) mbar_wait(&o_bar[(gb - 2) % 3], ((gb - 2) / 3) & 1);
        if (warp_alive) {
          for (int c = cb; c < ce;) {
            if (kHalves == 1 && c + 2 <= ce) {
              uint32_t w[32];
              tmem_ld_32x32(ts + c * 16, w);
              const bool fast = kFast && lds_u32(clean0 + c * 4) != 0 && lds_u32(clean0 + (c + 1) * 4) != 0;
              tmem_ld_wait();
              sum += fast ? wg_pass_exp<32, true>(w, 0, 0, m2, c_log2, pbuf, row, c * 16)
                          : wg_pass_exp<32, false>(w, ma0 + c * 64, aa0 + c * 64, m2, c_log2, pbuf, row, c * 16);
              c += 2;
            } else {
              uint32_t w[16];
              tmem_ld_32x16(ts + c * 16, w);
              const bool fast = kFast && lds_u32(clean0 + c * 4) != 0;
              tmem_ld_wait();
              sum += fast ? wg_pass_exp<16, true>(w, 0, 0, m2, c_log2, pbuf, row, c * 16)
                          : wg_pass_exp<16, false>(w, ma0 + c * 64, aa0 + c * 64, m2, c_log2, pbuf, row, c * 16);
              c += 1;
            }
          }
        }
        if (half == 0) sts_f(sStatM + ((gb & 7) * 128 + row) * 4, m2);
        sts_f(sStatL + (((gb & 7) * 2 + half) * 128 + row) * 4, sum);
        fence_proxy_async_smem();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_bar[blk]);
        if (t > 0) epilogue(t - 1, gt - 1);                // deferred behind this tile's softmax
      }
      epilogue(T - 1, gt - 1);
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
int attn_fwd_small(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask, void* o, int64_t ldo,
                   float* lse, int B, int n, int heads, float scale, int causal,
                   cudaStream_t stream);   // attention_small.cu
}
using namespace xclip;

extern "C" int xclip_attn_fwd(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask, void* o,
                              int64_t ldo, float* lse, int B, int n, int heads, float scale,
                              int causal, xclip_stream_t stream) {
  int rc = xclip_init();
  if (rc) return rc;
  XCLIP_REQUIRE(qkv && o && lse, "attn_fwd: null pointer");
  XCLIP_REQUIRE(B > 0 && heads > 0 && n > 0, "attn_fwd: bad sizes B=%d n=%d heads=%d", B, n, heads);
  XCLIP_REQUIRE(n <= 320, "attn_fwd: sequence length %d > 320 is not supported by this kernel", n);
  XCLIP_REQUIRE(scale > 0.f, "attn_fwd: scale must be positive");
  XCLIP_REQUIRE(ld_qkv % 8 == 0 && ld_qkv >= 3 * heads * kDh, "attn_fwd: bad ld_qkv");
  XCLIP_REQUIRE(ldo % 8 == 0 && ldo >= heads * kDh, "attn_fwd: bad ldo");
  XCLIP_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(o) & 15) == 0,
                "attn_fwd: misaligned pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n <= kTile)
    return attn_fwd_small(qkv, ld_qkv, key_mask, o, ldo, lse, B, n, heads, scale, causal, s);
  XCLIP_REQUIRE(!causal, "attn_fwd: the causal mask is only implemented for n <= 128 (got n=%d)", n);

  AttnPPParams p;
  p.B = B; p.H = heads; p.n = n;
  p.nkp = (n + 15) / 16 * 16;
  p.q_tiles = (n + kTile - 1) / kTile;
  p.w0 = ((p.nkp / 2) + 15) / 16 * 16;      // n > 128 -> nkp >= 144 -> w1 >= 64
  p.w1 = p.nkp - p.w0;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = key_mask;
  p.o = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;

  CUtensorMap tm;
  rc = encode_3d_bf16(&tm, qkv, (uint64_t)(3 * heads * kDh), (uint64_t)n, (uint64_t)B,
                      (uint64_t)ld_qkv, (uint64_t)n * ld_qkv, kDh, kTile);
  if (rc) return rc;
  const int nkb = (n + kTile - 1) / kTile;
  const int smem = (1 + 2 * nkb) * kBoxBytes + 2 * kPPBuf + kWgTailBytes;
  auto kern = attn_fwd_wg_kernel<2, true>;
  rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), (1 + 6) * kBoxBytes + 2 * kPPBuf + kWgTailBytes);
  if (rc) return rc;
  long long grid = num_sms();
  if (grid > (long long)B * heads) grid = (long long)B * heads;
  kern<<<(int)grid, 17 * 32, smem, s>>>(tm, p);
  XCLIP_LAUNCH_CHECK("attn_fwd_wg_kernel");
  return XCLIP_OK;
}
