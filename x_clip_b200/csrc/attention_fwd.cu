// Fused multi-head attention forward for short sequences (n <= 320, dim_head = 64), sm_100a.
//
// Replaces the reference's Attention core, x_clip/x_clip.py:217-244:
//   split heads -> q * dh^-0.5 -> einsum QK^T -> masked_fill(~key_mask, -finfo.max) ->
//   softmax(fp32) -> einsum PV -> merge heads
// which materialises [B,h,n,n] scores in HBM; here scores never leave the SM.
//
// One CTA owns one (batch, head): K and V of that head (<= 3 TMA boxes of 128 tokens each)
// stay in shared memory while the CTA walks the query tiles of 128 rows:
//   S = Q K^T      tcgen05.mma, M=128, N=ceil16(n) (all keys at once), accumulator in TMEM
//   softmax        8 warps; a query row (TMEM lane) is shared by two threads that each own half
//                  of the key columns; exact two-pass softmax (whole key range is resident),
//                  masking is branch-free: t = fma(raw, mul[j], add[j]) with per-key tables
//   P -> smem      bf16, written in the SWIZZLE_128B K-major layout the MMA expects
//   O = P V        tcgen05.mma, M=128, N=64, K=ceil16(n); V consumed MN-major straight from
//                  its TMA box (no transpose)
//   epilogue       O / rowsum -> bf16 -> global ; log-sum-exp (base 2, scaled domain) -> global
// Software pipeline: the control thread prefetches Q of the next tile (double buffer) and
// issues S(next) right behind PV(current), so TMA and tensor-pipe latency hide behind the
// softmax/epilogue of the compute warps.
//
// Masking follows the reference exactly: a masked key's score is replaced by -FLT_MAX AFTER
// scaling (so a fully masked row would give uniform attention); keys beyond n do not exist.
#include <stdlib.h>

#include "common.cuh"
#include "host.h"

namespace xclip {

constexpr int kAttnComputeWarps = 8;
constexpr int kAttnThreads = (kAttnComputeWarps + 1) * 32;  // + control warp (TMA + MMA issue)
constexpr int kTile = 128;
constexpr int kDh = 64;
constexpr int kBoxBytes = kTile * kDh * 2;  // 16 KiB: one [128 x 64] bf16 TMA box

struct AttnFwdParams {
  int B, H, n;
  int nkp;         // keys padded to a multiple of 16
  int tmem_cols;   // 128 / 256 / 512
  float scale_log2;  // dim_head^-0.5 * log2(e)
  const uint8_t* mask;  // [B, n] (1 = attend) or null
  bf16* o;
  long long ldo;
  float* lse;  // [B, H, n], base-2 log-sum-exp of the scaled (and masked) scores
};

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_result, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}

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

// t = raw * mul[j] + add[j] for 32 (or 16) consecutive keys; tables are fp32 in shared memory
template <int CNT>
__device__ __forceinline__ void scaled_scores(const uint32_t (&v)[32], uint32_t mul_addr,
                                              uint32_t add_addr, float (&t)[32]) {
#pragma unroll
  for (int i = 0; i < CNT; i += 4) {
    const float4 m = lds_f4(mul_addr + i * 4);
    const float4 a = lds_f4(add_addr + i * 4);
    t[i] = fmaf(__uint_as_float(v[i]), m.x, a.x);
    t[i + 1] = fmaf(__uint_as_float(v[i + 1]), m.y, a.y);
    t[i + 2] = fmaf(__uint_as_float(v[i + 2]), m.z, a.z);
    t[i + 3] = fmaf(__uint_as_float(v[i + 3]), m.w, a.w);
  }
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = (p.n + kTile - 1) / kTile;   // 128-token boxes of K / V
  const int npb = (p.nkp + 63) / 64;            // 64-key blocks of P
  uint8_t* sQ = smem;                           // 2 buffers
  uint8_t* sK = sQ + 2 * kBoxBytes;
  uint8_t* sV = sK + nkb * kBoxBytes;
  uint8_t* sP = sV + nkb * kBoxBytes;
  uint8_t* tail = sP + npb * kBoxBytes;
  uint64_t* kv_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_bar = kv_bar + 1;   // [2]
  uint64_t* s_bar = kv_bar + 3;
  uint64_t* p_bar = kv_bar + 4;
  uint64_t* o_bar = kv_bar + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_bar + 6);
  const uint32_t sMul = smem_u32(tail + 64);            // [384] f32
  const uint32_t sAdd = sMul + 384 * 4;                 // [384] f32
  const uint32_t sMax = sAdd + 384 * 4;                 // [2][128] f32
  const uint32_t sSum = sMax + 2 * 128 * 4;             // [2][128] f32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool is_control = warp == kAttnComputeWarps;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("xclip attn_fwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(kv_bar, 1);
    mbar_init(&q_bar[0], 1);
    mbar_init(&q_bar[1], 1);
    mbar_init(s_bar, 1);
    mbar_init(p_bar, kAttnComputeWarps);
    mbar_init(o_bar, 1);
    fence_barrier_init();
  }
  if (is_control) {
    if (lane == 0) tma_prefetch_desc(&tm_qkv);
    tmem_alloc_dyn(tmem_slot, p.tmem_cols);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + (p.tmem_cols - kDh);

  const int num_q_tiles = (p.n + kTile - 1) / kTile;
  const int inner = p.H * kDh;  // column offset between q | k | v
  uint32_t kv_phase = 0;
  uint32_t tile_count = 0;      // tiles processed by this CTA (drives barrier parities)

  // column split of a row between its two threads, in 32-wide chunks
  const int nchunks = (p.nkp + 31) / 32;
  const int half = warp >> 2;                   // compute warps only
  const int quarter = warp & 3;
  const int c_begin = half == 0 ? 0 : (nchunks + 1) / 2;
  const int c_end = half == 0 ? (nchunks + 1) / 2 : nchunks;

  for (int bh = blockIdx.x; bh < p.B * p.H; bh += gridDim.x) {
    const int b = bh / p.H, h = bh % p.H;

    if (!is_control) {
      // per-key tables: mul = scale (attend) or 0; add = 0, -FLT_MAX (masked) or -inf (no key)
      for (int j = threadIdx.x; j < 384; j += kAttnComputeWarps * 32) {
        float mul = 0.f, add = -INFINITY;
        if (j < p.n) {
          const bool keep = p.mask ? (p.mask[(long long)b * p.n + j] != 0) : true;
          mul = keep ? p.scale_log2 : 0.f;
          add = keep ? 0.f : -FLT_MAX;
        }
        sts_f(sMul + j * 4, mul);
        sts_f(sAdd + j * 4, add);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    } else if (XCLIP_ONE_LANE(lane)) {
      mbar_arrive_expect_tx(kv_bar, 2 * nkb * kBoxBytes);
      for (int i = 0; i < nkb; ++i) {
        tma_load_3d(sK + i * kBoxBytes, &tm_qkv, kv_bar, inner + h * kDh, i * kTile, b);
        tma_load_3d(sV + i * kBoxBytes, &tm_qkv, kv_bar, 2 * inner + h * kDh, i * kTile, b);
      }
    }

    if (is_control) {
      // ===================== control warp: TMA + MMA issue, software pipelined ============
      if (XCLIP_ONE_LANE(lane)) {
        // operand descriptors are built once; the loops only add compile-time offsets
        // (descriptor address units are 16 bytes)
        const uint64_t desc_q0 = make_smem_desc(smem_u32(sQ), 0, 1024);
        const uint64_t desc_k = make_smem_desc(smem_u32(sK), 0, 1024);
        const uint64_t desc_p = make_smem_desc(smem_u32(sP), 0, 1024);
        const uint64_t desc_v = make_smem_desc(smem_u32(sV), 8192, 1024);
        const int n_lo = min(256, p.nkp), n_hi = p.nkp - n_lo;      // S column chunks (<= 256 each)
        const uint32_t idesc_lo = make_idesc_bf16(kTile, n_lo, kMajorK, kMajorK);
        const uint32_t idesc_hi = make_idesc_bf16(kTile, n_hi > 0 ? n_hi : 16, kMajorK, kMajorK);
        auto issue_s = [&](uint32_t tc) {   // S = Q K^T for the tile with running index tc
          const uint64_t qd = desc_q0 + (tc & 1) * (kBoxBytes >> 4);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma_bf16(tmem_base, qd + 2 * k, desc_k + 2 * k, idesc_lo, k > 0 ? 1u : 0u);
          if (n_hi > 0) {
#pragma unroll
            for (int k = 0; k < kDh / 16; ++k)
              umma_bf16(tmem_base + 256, qd + 2 * k, desc_k + (256 * 128 >> 4) + 2 * k, idesc_hi,
                        k > 0 ? 1u : 0u);
          }
          umma_commit(s_bar);
        };
        // prologue: Q of the first tile, then S(0)
        {
          const uint32_t tc = tile_count;
          mbar_arrive_expect_tx(&q_bar[tc & 1], kBoxBytes);
          tma_load_3d(sQ + (tc & 1) * kBoxBytes, &tm_qkv, &q_bar[tc & 1], h * kDh, 0, b);
          mbar_wait(kv_bar, kv_phase);
          mbar_wait(&q_bar[tc & 1], (tc >> 1) & 1);
          tcgen05_fence_after();
          issue_s(tc);
        }
        for (int qt = 0; qt < num_q_tiles; ++qt) {
          const uint32_t tc = tile_count + qt;
          if (qt + 1 < num_q_tiles) {   // prefetch next Q (its buffer was consumed by S(tc-1))
            const uint32_t tn = tc + 1;
            mbar_arrive_expect_tx(&q_bar[tn & 1], kBoxBytes);
            tma_load_3d(sQ + (tn & 1) * kBoxBytes, &tm_qkv, &q_bar[tn & 1], h * kDh,
                        (qt + 1) * kTile, b);
          }
          // O = P V once the softmax warps have written P (and finished reading S)
          mbar_wait(p_bar, tc & 1);
          tcgen05_fence_after();
          constexpr uint32_t idesc_pv = make_idesc_bf16(kTile, kDh, kMajorK, kMajorMN);
          const int ksteps = p.nkp / 16;          // <= 20 (n <= 320)
#pragma unroll
          for (int k = 0; k < 20; ++k) {
            if (k < ksteps)
              umma_bf16(tmem_o, desc_p + ((k >> 2) * (kBoxBytes >> 4) + (k & 3) * 2),
                        desc_v + k * 128, idesc_pv, k > 0 ? 1u : 0u);
          }
          umma_commit(o_bar);
          if (qt + 1 < num_q_tiles) {   // S(next) queues right behind PV(current)
            const uint32_t tn = tc + 1;
            mbar_wait(&q_bar[tn & 1], (tn >> 1) & 1);
            tcgen05_fence_after();
            issue_s(tn);
          }
        }
        // K/V smem is reused by the next (b,h): wait until the last PV retired
        mbar_wait(o_bar, (tile_count + num_q_tiles - 1) & 1);
      }
      __syncwarp();
    } else {
      // ===================== softmax + epilogue warps =====================
      const int row = quarter * 32 + lane;      // TMEM lane == query row inside the tile
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int qt = 0; qt < num_q_tiles; ++qt) {
        const uint32_t tc = tile_count + qt;
        const int q_idx = qt * kTile + row;
        mbar_wait(s_bar, tc & 1);
        tcgen05_fence_after();

        // a warp whose 32 query rows are all beyond n has nothing to compute: O rows depend only
        // on their own P rows, so its smem/TMEM slots may hold anything (it still takes part in
        // every barrier)
        const bool warp_alive = qt * kTile + quarter * 32 < p.n;
        // pass 1: maximum of this thread's half of the row (base-2, scaled + masked scores)
        float m2 = -INFINITY;
        for (int c = warp_alive ? c_begin : c_end; c < c_end; ++c) {
          const int c0 = c * 32;
          uint32_t v[32];
          float t[32];
          if (p.nkp - c0 >= 32) {
            tmem_ld_32x32(t_row + c0, v);
            tmem_ld_wait();
            scaled_scores<32>(v, sMul + c0 * 4, sAdd + c0 * 4, t);
#pragma unroll
            for (int i = 0; i < 32; ++i) m2 = fmaxf(m2, t[i]);
          } else {
            uint32_t w[16];
            tmem_ld_32x16(t_row + c0, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = w[i];
            scaled_scores<16>(v, sMul + c0 * 4, sAdd + c0 * 4, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) m2 = fmaxf(m2, t[i]);
          }
        }
        sts_f(sMax + (half * 128 + row) * 4, m2);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m2 = fmaxf(m2, lds_f(sMax + ((half ^ 1) * 128 + row) * 4));

        // pass 2: probabilities -> bf16 P in smem (SW128 K-major blocks of 64 keys), row sum
        float sum = 0.f;
        for (int c = warp_alive ? c_begin : c_end; c < c_end; ++c) {
          const int c0 = c * 32;
          uint32_t v[32];
          float t[32];
          const bool full = (p.nkp - c0 >= 32);
          if (full) {
            tmem_ld_32x32(t_row + c0, v);
            tmem_ld_wait();
            scaled_scores<32>(v, sMul + c0 * 4, sAdd + c0 * 4, t);
          } else {
            uint32_t w[16];
            tmem_ld_32x16(t_row + c0, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = w[i];
            scaled_scores<16>(v, sMul + c0 * 4, sAdd + c0 * 4, t);
#pragma unroll
            for (int i = 16; i < 32; ++i) t[i] = -INFINITY;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            t[i] = ex2_approx(t[i] - m2);
            sum += t[i];
          }
          const uint32_t blk = smem_u32(sP) + (c0 >> 6) * kBoxBytes;
          const int chunk0 = (c0 & 63) >> 3;  // first 16-byte chunk inside the 128-byte row
          const int nch = full ? 4 : 2;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc < nch) {
              sts_v4(blk + swz128(row, chunk0 + cc), pack_bf16x2(t[cc * 8 + 0], t[cc * 8 + 1]),
                     pack_bf16x2(t[cc * 8 + 2], t[cc * 8 + 3]),
                     pack_bf16x2(t[cc * 8 + 4], t[cc * 8 + 5]),
                     pack_bf16x2(t[cc * 8 + 6], t[cc * 8 + 7]));
            }
          }
        }
        sts_f(sSum + (half * 128 + row) * 4, sum);
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to tcgen05.mma
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_bar);

        // epilogue: this thread converts 32 of the 64 output columns of its row
        mbar_wait(o_bar, tc & 1);
        tcgen05_fence_after();
        sum += lds_f(sSum + ((half ^ 1) * 128 + row) * 4);
        const float inv = 1.f / sum;
        if (half == 0 && q_idx < p.n)
          p.lse[((long long)b * p.H + h) * p.n + q_idx] = m2 + log2f(sum);
        {
          uint32_t v[32];
          tmem_ld_32x32(tmem_o + (static_cast<uint32_t>(quarter * 32) << 16) + half * 32, v);
          tmem_ld_wait();
          if (q_idx < p.n) {
            bf16* dst = p.o + ((long long)b * p.n + q_idx) * p.ldo + h * kDh + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
              o.y = pack_bf16x2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
              o.z = pack_bf16x2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
              o.w = pack_bf16x2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + i) = o;
            }
          }
        }
        tcgen05_fence_before();
      }
    }
    tile_count += num_q_tiles;
    kv_phase ^= 1;
    // every role finished with this (b,h)'s K/V, tables and TMEM before anyone starts the next
    __syncthreads();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (is_control) {
    tcgen05_fence_after();
    tmem_dealloc_dyn(tmem_base, p.tmem_cols);
  }
}


// ---------------------------------------------------------------------------------------------
// Ping-pong variant for 128 < n <= 320 (default; XCLIP_ATTN_PP=0 selects the single-buffer kernel).
//
// The keys of a query tile are split in two blocks (w0 + w1 columns).  Two S buffers in TMEM
// alternate between consecutive blocks (also across tiles and (batch, head) items), so the issue
// thread always has the NEXT block's S = Q K_blk^T in flight while the softmax warps work on the
// current one.  Every block keeps its own max / sum and accumulates P_blk V_blk into its own
// 64-column O buffer taken from a ring of three; nothing is rescaled in TMEM - the epilogue
// combines the two partial outputs:  O = (2^(m0-m) O0 + 2^(m1-m) O1) / (2^(m0-m) l0 + 2^(m1-m) l1).
// TMEM: S 2 x 160 + O 3 x 64 = 512 columns.  smem: Q 16 + K 48 + V 48 + 2 P buffers x 48 KiB.
struct AttnPPParams {
  int B, H, n, nkp, w0, w1;
  int q_tiles;           // 128-query tiles handled here (the tail token of n = 128k+1 is not)
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
template <int kVariant>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_pp_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnPPParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = (p.n + kTile - 1) / kTile;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kBoxBytes;
  uint8_t* sV = sK + nkb * kBoxBytes;
  uint8_t* sP = sV + nkb * kBoxBytes;           // 2 buffers
  uint8_t* tail = sP + 2 * kPPBuf;
  uint64_t* k_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_bar = k_bar + 1;
  uint64_t* s_bar = k_bar + 2;     // [2]
  uint64_t* p_bar = k_bar + 4;     // [2]
  uint64_t* o_bar = k_bar + 6;     // [3]
  uint64_t* v_bar = k_bar + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k_bar + 10);
  const uint32_t sMul = smem_u32(tail + 128);           // [384] f32
  const uint32_t sAdd = sMul + 384 * 4;                 // [384] f32
  const uint32_t sMax = sAdd + 384 * 4;                 // [2 block parities][2 halves][128] f32
  const uint32_t sSum = sMax + 2 * 2 * 128 * 4;         // [4 block slots][2 halves][128] f32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool is_control = warp == kAttnComputeWarps;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(k_bar, 1);
    mbar_init(v_bar, 1);
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_bar[i], 1); mbar_init(&p_bar[i], kAttnComputeWarps); }
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
  // S buffers at columns 0 and 160, O ring at 320 / 384 / 448

  const int T = p.q_tiles;                      // query tiles per (b,h); 2 blocks per tile
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
      // K + Q(tile 0) and V of the NEXT (b,h) are fetched as soon as their smem is dead: K/Q after
      // the last S of this item retired, V after its last PV.
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
      uint32_t g = 0, tt = 0, kvc = 0;          // global block / tile / (b,h) counters of this CTA
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
              if (t + 1 < T) {                  // Q(t) is dead once S(g) retired: fetch Q(t+1)
                mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
                mbar_arrive_expect_tx(q_bar, kBoxBytes);
                tma_load_3d(sQ, &tm_qkv, q_bar, h * kDh, (t + 1) * kTile, b);
              } else if (bh_next < total) {     // K and Q of this item are dead
                mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
                load_kq(bh_next);
              }
            }
            if (t == 0 && kb == 1) mbar_wait(v_bar, kvc & 1);
            if (!(t == 0 && kb == 0)) issue_pv(g - 1);   // PV lags S by one block
          }
        }
        issue_pv(g - 1);
        // V / P smem are reused by the next (b,h): wait until the last PV retired
        mbar_wait(&o_bar[(g - 1) % 3], ((g - 1) / 3) & 1);
        if (bh_next < total) load_v(bh_next);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue warps =====================
    constexpr bool kResident = (kVariant & 1) == 0;
    constexpr bool kPairBar = (kVariant & 2) != 0;
    const int half = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    auto row_barrier = [&]() {
      if constexpr (kPairBar)
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
      else
        asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    uint32_t g = 0;
    for (int bh = blockIdx.x; bh < p.B * p.H; bh += gridDim.x) {
      const int b = bh / p.H, h = bh - b * p.H;
      for (int j = threadIdx.x; j < 384; j += kAttnComputeWarps * 32) {
        float mul = 0.f, add = -INFINITY;
        if (j < p.n) {
          const bool keep = p.mask ? (p.mask[(long long)b * p.n + j] != 0) : true;
          mul = keep ? p.scale_log2 : 0.f;
          add = keep ? 0.f : -FLT_MAX;
        }
        sts_f(sMul + j * 4, mul);
        sts_f(sAdd + j * 4, add);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");

      float m_blk[2] = {0.f, 0.f}, l_own[2] = {0.f, 0.f};
      float pm0 = 0.f, pm1 = 0.f, pl0 = 0.f, pl1 = 0.f;   // previous tile's block statistics

      auto epilogue = [&](int t, uint32_t g0, float m0, float m1, float l0o, float l1o) {
        const uint32_t g1 = g0 + 1;
        mbar_wait(&o_bar[g0 % 3], (g0 / 3) & 1);
        mbar_wait(&o_bar[g1 % 3], (g1 / 3) & 1);
        tcgen05_fence_after();
        const int q_idx = t * kTile + row;
        const float l0 = l0o + lds_f(sSum + (((g0 & 3) * 2 + (half ^ 1)) * 128 + row) * 4);
        const float l1 = l1o + lds_f(sSum + (((g1 & 3) * 2 + (half ^ 1)) * 128 + row) * 4);
        const float m = fmaxf(m0, m1);
        const float a0 = ex2_approx(m0 - m), a1 = ex2_approx(m1 - m);
        const float L = a0 * l0 + a1 * l1;
        const float inv = 1.f / L;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(tmem_base + 320 + (g0 % 3) * kDh + lane_off + half * 32, v0);
        tmem_ld_32x32(tmem_base + 320 + (g1 % 3) * kDh + lane_off + half * 32, v1);
        tmem_ld_wait();
        if (q_idx < p.n) {
          if (half == 0) p.lse[((long long)b * p.H + h) * p.n + q_idx] = m + log2f(L);
          bf16* dst = p.o + ((long long)b * p.n + q_idx) * p.ldo + h * kDh + half * 32;
          const float c0 = a0 * inv, c1 = a1 * inv;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
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
        tcgen05_fence_before();
      };

      for (int t = 0; t < T; ++t) {
        const bool warp_alive = t * kTile + quarter * 32 < p.n;
        for (int kb = 0; kb < 2; ++kb, ++g) {
          const int k0 = kb ? p.w0 : 0, W = kb ? p.w1 : p.w0;
          const int nch = W / 16;
          const int cb = warp_alive ? (half == 0 ? 0 : (nch + 1) / 2) : 0;
          const int ce = warp_alive ? (half == 0 ? (nch + 1) / 2 : nch) : 0;
          const uint32_t ts = tmem_base + (g & 1) * 160 + lane_off;
          mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
          tcgen05_fence_after();
          float m2 = -INFINITY, sum = 0.f;
          const uint32_t pbuf = smem_u32(sP) + (g & 1) * kPPBuf;
          if constexpr (!kResident) {
            // streaming variant: S is read from TMEM twice, 16 columns at a time
            for (int c = cb; c < ce; ++c) {
              uint32_t w[16];
              tmem_ld_32x16(ts + c * 16, w);
              tmem_ld_wait();
              const uint32_t ma = sMul + (k0 + c * 16) * 4, aa = sAdd + (k0 + c * 16) * 4;
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 mm = lds_f4(ma + i * 4), ad = lds_f4(aa + i * 4);
                m2 = fmaxf(m2, fmaf(__uint_as_float(w[i]), mm.x, ad.x));
                m2 = fmaxf(m2, fmaf(__uint_as_float(w[i + 1]), mm.y, ad.y));
                m2 = fmaxf(m2, fmaf(__uint_as_float(w[i + 2]), mm.z, ad.z));
                m2 = fmaxf(m2, fmaf(__uint_as_float(w[i + 3]), mm.w, ad.w));
              }
            }
            sts_f(sMax + (((g & 1) * 2 + half) * 128 + row) * 4, m2);
            row_barrier();
            m2 = fmaxf(m2, lds_f(sMax + (((g & 1) * 2 + (half ^ 1)) * 128 + row) * 4));
            if (g >= 2) mbar_wait(&o_bar[(g - 2) % 3], ((g - 2) / 3) & 1);   // P buffer free
            for (int c = cb; c < ce; ++c) {
              uint32_t w[16];
              tmem_ld_32x16(ts + c * 16, w);
              tmem_ld_wait();
              const uint32_t ma = sMul + (k0 + c * 16) * 4, aa = sAdd + (k0 + c * 16) * 4;
              float e[16];
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 mm = lds_f4(ma + i * 4), ad = lds_f4(aa + i * 4);
                e[i] = ex2_approx(fmaf(__uint_as_float(w[i]), mm.x, ad.x) - m2);
                e[i + 1] = ex2_approx(fmaf(__uint_as_float(w[i + 1]), mm.y, ad.y) - m2);
                e[i + 2] = ex2_approx(fmaf(__uint_as_float(w[i + 2]), mm.z, ad.z) - m2);
                e[i + 3] = ex2_approx(fmaf(__uint_as_float(w[i + 3]), mm.w, ad.w) - m2);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) sum += e[i];
              const int col = c * 16;
              const uint32_t blk = pbuf + (col >> 6) * kBoxBytes;
              const int chunk0 = (col & 63) >> 3;
              sts_v4(blk + swz128(row, chunk0), pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]),
                     pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
              sts_v4(blk + swz128(row, chunk0 + 1), pack_bf16x2(e[8], e[9]),
                     pack_bf16x2(e[10], e[11]), pack_bf16x2(e[12], e[13]), pack_bf16x2(e[14], e[15]));
            }
          } else {
            // The thread's share of the block (<= 5 chunks of 16 keys) stays in registers between
            // the max pass and the exp pass: one TMEM read + one wait per block.
            const int nc = ce - cb;
            uint32_t w[5][16];
#pragma unroll
            for (int q = 0; q < 5; ++q)
              if (q < nc) tmem_ld_32x16(ts + (cb + q) * 16, w[q]);
            tmem_ld_wait();
            m2 = -INFINITY;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              if (q < nc) {
                const uint32_t ma = sMul + (k0 + (cb + q) * 16) * 4, aa = sAdd + (k0 + (cb + q) * 16) * 4;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 mm = lds_f4(ma + i * 4), ad = lds_f4(aa + i * 4);
                  const float t0 = fmaf(__uint_as_float(w[q][i]), mm.x, ad.x);
                  const float t1 = fmaf(__uint_as_float(w[q][i + 1]), mm.y, ad.y);
                  const float t2 = fmaf(__uint_as_float(w[q][i + 2]), mm.z, ad.z);
                  const float t3 = fmaf(__uint_as_float(w[q][i + 3]), mm.w, ad.w);
                  m2 = fmaxf(fmaxf(m2, fmaxf(t0, t1)), fmaxf(t2, t3));
                  w[q][i] = __float_as_uint(t0); w[q][i + 1] = __float_as_uint(t1);
                  w[q][i + 2] = __float_as_uint(t2); w[q][i + 3] = __float_as_uint(t3);
                }
              }
            }
            sts_f(sMax + (((g & 1) * 2 + half) * 128 + row) * 4, m2);
            row_barrier();
            m2 = fmaxf(m2, lds_f(sMax + (((g & 1) * 2 + (half ^ 1)) * 128 + row) * 4));

            // P buffer (g&1) was last read by PV(g-2)
            if (g >= 2) mbar_wait(&o_bar[(g - 2) % 3], ((g - 2) / 3) & 1);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              if (q < nc) {
                const int col = (cb + q) * 16;                // column inside the block
                const uint32_t blk = pbuf + (col >> 6) * kBoxBytes;
                const int chunk0 = (col & 63) >> 3;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  float e[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) e[i] = ex2_approx(__uint_as_float(w[q][hh * 8 + i]) - m2);
                  sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
                  sts_v4(blk + swz128(row, chunk0 + hh), pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]),
                         pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
                }
              }
            }
          }
          sts_f(sSum + (((g & 3) * 2 + half) * 128 + row) * 4, sum);
          fence_proxy_async_smem();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_bar[g & 1]);
          m_blk[kb] = m2;
          l_own[kb] = sum;

          // the previous tile's epilogue runs after this tile's first block was handed over
          if (kb == 0 && t > 0) epilogue(t - 1, g - 2, pm0, pm1, pl0, pl1);
        }
        pm0 = m_blk[0]; pm1 = m_blk[1]; pl0 = l_own[0]; pl1 = l_own[1];
      }
      epilogue(T - 1, g - 2, pm0, pm1, pl0, pl1);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (is_control) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Warpgroup ping-pong variants (XCLIP_ATTN_PP_VARIANT=4..7, experimental - DESIGN.md section 9).
// Same TMEM / smem plan and issue-thread schedule as attn_fwd_pp_kernel, different softmax
// mapping: the two key blocks of a tile are processed CONCURRENTLY by different warpgroups
// (block 0 <-> S buffer 0, block 1 <-> S buffer 1), so there is no CTA-wide barrier in the loop and
// the warpgroups only meet in the epilogue through a ring of per-block (max, sum) statistics.
//   kHalves = 1:  8 softmax warps; a row of a block is owned by ONE thread (no max exchange).
//   kHalves = 2: 16 softmax warps (4 per scheduler); two warps share a row of a block (column
//                halves) and exchange the row max through a 64-thread named barrier.
//   kFast: 16-column chunks without masked / padded keys skip the mask tables:
//          p = 2^(s*c - m) is one FFMA + EX2, the max pass works on raw scores.
// Because a warpgroup no longer implies that the others are past their deferred epilogue, the O
// ring gets an explicit "epilogue done" barrier per tile (e_bar) that the issue thread checks
// before a PV overwrites a ring slot.
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

// ---------------------------------------------------------------------------------------------
// 16-warp ping-pong variant (XCLIP_ATTN_PP_VARIANT=8 / 9 with the fast unmasked chunks;
// experimental - DESIGN.md section 9).  The validated ping-pong protocol and issue thread of
// attn_fwd_pp_kernel (S look-ahead of one block, O ring of three) with 16 softmax warps: four
// warps share a query row (a quarter of the block's key chunks each), so four warps per
// scheduler hide the TMEM / MUFU / shared-memory latencies that dominate the 8-warp kernel
// (13 % issue utilisation, profiles/r1_ncu_attn_pp_stalls.md).  The row max is exchanged between
// the four warps of a lane quarter through a 128-thread named barrier.
constexpr int kPP16Warps = 16;
constexpr int kPP16Threads = (kPP16Warps + 1) * 32;
constexpr int kPP16TailBytes = 128 + 2 * kWgTableCols * 4 + 128 + 2 * 4 * 128 * 4 + 4 * 4 * 128 * 4;

template <bool kFast>
__global__ void __launch_bounds__(kPP16Threads, 1)
attn_fwd_pp16_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnPPParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nkb = (p.n + kTile - 1) / kTile;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kBoxBytes;
  uint8_t* sV = sK + nkb * kBoxBytes;
  uint8_t* sP = sV + nkb * kBoxBytes;           // 2 buffers
  uint8_t* tail = sP + 2 * kPPBuf;
  uint64_t* k_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_bar = k_bar + 1;
  uint64_t* s_bar = k_bar + 2;     // [2]
  uint64_t* p_bar = k_bar + 4;     // [2], 16 arrivals
  uint64_t* o_bar = k_bar + 6;     // [3]
  uint64_t* v_bar = k_bar + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(k_bar + 10);
  const uint32_t sMul = smem_u32(tail + 128);               // [320] f32
  const uint32_t sAdd = sMul + kWgTableCols * 4;            // [320] f32
  const uint32_t sClean = sAdd + kWgTableCols * 4;          // [32] u32
  const uint32_t sMax = sClean + 128;                       // [2 block parities][4 groups][128] f32
  const uint32_t sSum = sMax + 2 * 4 * 128 * 4;             // [4 block slots][4 groups][128] f32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool is_control = warp == kPP16Warps;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(k_bar, 1);
    mbar_init(v_bar, 1);
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_bar[i], 1); mbar_init(&p_bar[i], kPP16Warps); }
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
      // K + Q(tile 0) and V of the NEXT (b,h) are fetched as soon as their smem is dead: K/Q after
      // the last S of this item retired, V after its last PV.
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
      uint32_t g = 0, tt = 0, kvc = 0;          // global block / tile / (b,h) counters of this CTA
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
              if (t + 1 < T) {                  // Q(t) is dead once S(g) retired: fetch Q(t+1)
                mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
                mbar_arrive_expect_tx(q_bar, kBoxBytes);
                tma_load_3d(sQ, &tm_qkv, q_bar, h * kDh, (t + 1) * kTile, b);
              } else if (bh_next < total) {     // K and Q of this item are dead
                mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
                load_kq(bh_next);
              }
            }
            if (t == 0 && kb == 1) mbar_wait(v_bar, kvc & 1);
            if (!(t == 0 && kb == 0)) issue_pv(g - 1);   // PV lags S by one block
          }
        }
        issue_pv(g - 1);
        // V / P smem are reused by the next (b,h): wait until the last PV retired
        mbar_wait(&o_bar[(g - 1) % 3], ((g - 1) / 3) & 1);
        if (bh_next < total) load_v(bh_next);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue warps =====================
    const int grp = warp >> 2, quarter = warp & 3;            // grp: column group 0..3 of the row
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int ocol = grp * 16;
    const float c_log2 = p.scale_log2;
    uint32_t g = 0;
    for (int bh = blockIdx.x; bh < p.B * p.H; bh += gridDim.x) {
      const int b = bh / p.H, h = bh - b * p.H;
      for (int j = threadIdx.x; j < kWgTableCols; j += kPP16Warps * 32) {
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
        const uint32_t bal = __ballot_sync(0xffffffffu, clean);   // whole warps, 32 consecutive keys
        if (lane == 0) {
          sts_u32(sClean + (j >> 4) * 4, (bal & 0xffffu) == 0xffffu ? 1u : 0u);
          sts_u32(sClean + ((j >> 4) + 1) * 4, (bal >> 16) == 0xffffu ? 1u : 0u);
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");

      float m_blk[2] = {0.f, 0.f}, l_own[2] = {0.f, 0.f};
      float pm0 = 0.f, pm1 = 0.f, pl0 = 0.f, pl1 = 0.f;   // previous tile's block statistics

      auto epilogue = [&](int t, uint32_t g0, float m0, float m1, float l0o, float l1o) {
        const uint32_t g1 = g0 + 1;
        mbar_wait(&o_bar[g0 % 3], (g0 / 3) & 1);
        mbar_wait(&o_bar[g1 % 3], (g1 / 3) & 1);
        tcgen05_fence_after();
        const int q_idx = t * kTile + row;
        float l0 = l0o, l1 = l1o;
#pragma unroll
        for (int o = 1; o < 4; ++o) {                      // the other three column groups' sums
          l0 += lds_f(sSum + (((g0 & 3) * 4 + ((grp + o) & 3)) * 128 + row) * 4);
          l1 += lds_f(sSum + (((g1 & 3) * 4 + ((grp + o) & 3)) * 128 + row) * 4);
        }
        const float m = fmaxf(m0, m1);
        const float a0 = ex2_approx(m0 - m), a1 = ex2_approx(m1 - m);
        const float L = a0 * l0 + a1 * l1;
        const float inv = 1.f / L;
        uint32_t v0[16], v1[16];
        tmem_ld_32x16(tmem_base + 320 + (g0 % 3) * kDh + lane_off + ocol, v0);
        tmem_ld_32x16(tmem_base + 320 + (g1 % 3) * kDh + lane_off + ocol, v1);
        tmem_ld_wait();
        if (q_idx < p.n) {
          if (grp == 0) p.lse[((long long)b * p.H + h) * p.n + q_idx] = m + log2f(L);
          bf16* dst = p.o + ((long long)b * p.n + q_idx) * p.ldo + h * kDh + ocol;
          const float c0 = a0 * inv, c1 = a1 * inv;
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
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
        tcgen05_fence_before();
      };

      for (int t = 0; t < T; ++t) {
        const bool warp_alive = t * kTile + quarter * 32 < p.n;
        for (int kb = 0; kb < 2; ++kb, ++g) {
          const int k0 = kb ? p.w0 : 0, W = kb ? p.w1 : p.w0;
          const int nch = W / 16;                           // 16-key chunks, split over 4 groups
          const int base = nch >> 2, rem = nch & 3;
          const int cb = warp_alive ? grp * base + min(grp, rem) : 0;
          const int ce = warp_alive ? cb + base + (grp < rem ? 1 : 0) : 0;
          const uint32_t ts = tmem_base + (g & 1) * 160 + lane_off;
          const uint32_t ma0 = sMul + k0 * 4, aa0 = sAdd + k0 * 4;
          const uint32_t clean0 = sClean + (k0 >> 4) * 4;
          mbar_wait(&s_bar[g & 1], (g >> 1) & 1);
          tcgen05_fence_after();
          float m_scaled = -INFINITY, m_raw = -INFINITY, sum = 0.f;
          for (int c = cb; c < ce; ++c) {
            uint32_t w[16];
            tmem_ld_32x16(ts + c * 16, w);
            const bool fast = kFast && lds_u32(clean0 + c * 4) != 0;
            tmem_ld_wait();
            if (fast) wg_pass_max<16, true>(w, 0, 0, m_scaled, m_raw);
            else wg_pass_max<16, false>(w, ma0 + c * 64, aa0 + c * 64, m_scaled, m_raw);
          }
          float m2 = kFast ? fmaxf(m_scaled, m_raw * c_log2) : m_scaled;
          {
            const uint32_t slot = sMax + (((g & 1) * 4) * 128 + row) * 4;
            sts_f(slot + grp * 128 * 4, m2);
            asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
#pragma unroll
            for (int o = 1; o < 4; ++o) m2 = fmaxf(m2, lds_f(slot + ((grp + o) & 3) * 128 * 4));
          }
          // P buffer (g&1) was last read by PV(g-2)
          if (g >= 2) mbar_wait(&o_bar[(g - 2) % 3], ((g - 2) / 3) & 1);
          const uint32_t pbuf = smem_u32(sP) + (g & 1) * kPPBuf;
          for (int c = cb; c < ce; ++c) {
            uint32_t w[16];
            tmem_ld_32x16(ts + c * 16, w);
            const bool fast = kFast && lds_u32(clean0 + c * 4) != 0;
            tmem_ld_wait();
            sum += fast ? wg_pass_exp<16, true>(w, 0, 0, m2, c_log2, pbuf, row, c * 16)
                        : wg_pass_exp<16, false>(w, ma0 + c * 64, aa0 + c * 64, m2, c_log2, pbuf, row, c * 16);
          }
          sts_f(sSum + (((g & 3) * 4 + grp) * 128 + row) * 4, sum);
          fence_proxy_async_smem();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_bar[g & 1]);
          m_blk[kb] = m2;
          l_own[kb] = sum;
          // the previous tile's epilogue runs after this tile's first block was handed over
          if (kb == 0 && t > 0) epilogue(t - 1, g - 2, pm0, pm1, pl0, pl1);
        }
        pm0 = m_blk[0]; pm1 = m_blk[1]; pl0 = l_own[0]; pl1 = l_own[1];
      }
      epilogue(T - 1, g - 2, pm0, pm1, pl0, pl1);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (is_control) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_attn_fwd_tail(const void* qkv, long long ld, const uint8_t* mask, void* o, long long ldo,
                         float* lse, int B, int H, int n, float scale_log2, cudaStream_t stream);

static int launch_attn_fwd_pp(const CUtensorMap& tm, const AttnFwdParams& q, bool tail,
                              cudaStream_t stream) {
  AttnPPParams p;
  p.B = q.B; p.H = q.H; p.n = q.n; p.nkp = q.nkp;
  p.q_tiles = tail ? (q.n - 1) / kTile : (q.n + kTile - 1) / kTile;
  p.w0 = ((q.nkp / 2) + 15) / 16 * 16;
  p.w1 = q.nkp - p.w0;
  p.scale_log2 = q.scale_log2; p.mask = q.mask; p.o = q.o; p.ldo = q.ldo; p.lse = q.lse;
  const int nkb = (q.n + kTile - 1) / kTile;
  static const int variant = [] {
    const char* e = getenv("XCLIP_ATTN_PP_VARIANT");
    return (e && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : 7;
  }();
  XCLIP_REQUIRE(variant < 6 || q.scale_log2 > 0.f, "attn_fwd: the fast-chunk variants need scale > 0");
  // barriers + mask tables + per-row exchange buffers
  const int tail_bytes = variant >= 8 ? kPP16TailBytes
                                      : (variant >= 4 ? kWgTailBytes : 128 + 2 * 384 * 4 + (4 + 8) * 128 * 4);
  const int smem = (1 + 2 * nkb) * kBoxBytes + 2 * kPPBuf + tail_bytes;
  static bool configured = false;
  if (!configured) {
    const int max_smem = (1 + 6) * kBoxBytes + 2 * kPPBuf + kWgTailBytes;
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<0>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<1>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<2>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_pp_kernel<3>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_wg_kernel<1, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_wg_kernel<2, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_wg_kernel<1, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_wg_kernel<2, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_pp16_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_pp16_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  long long grid = num_sms();
  if (grid > (long long)q.B * q.H) grid = (long long)q.B * q.H;
  switch (variant) {
    case 1: attn_fwd_pp_kernel<1><<<(int)grid, kAttnThreads, smem, stream>>>(tm, p); break;
    case 2: attn_fwd_pp_kernel<2><<<(int)grid, kAttnThreads, smem, stream>>>(tm, p); break;
    case 3: attn_fwd_pp_kernel<3><<<(int)grid, kAttnThreads, smem, stream>>>(tm, p); break;
    case 4: attn_fwd_wg_kernel<1, false><<<(int)grid, 9 * 32, smem, stream>>>(tm, p); break;
    case 5: attn_fwd_wg_kernel<2, false><<<(int)grid, 17 * 32, smem, stream>>>(tm, p); break;
    case 6: attn_fwd_wg_kernel<1, true><<<(int)grid, 9 * 32, smem, stream>>>(tm, p); break;
    case 7: attn_fwd_wg_kernel<2, true><<<(int)grid, 17 * 32, smem, stream>>>(tm, p); break;
    case 8: attn_fwd_pp16_kernel<false><<<(int)grid, kPP16Threads, smem, stream>>>(tm, p); break;
    case 9: attn_fwd_pp16_kernel<true><<<(int)grid, kPP16Threads, smem, stream>>>(tm, p); break;
    default: attn_fwd_pp_kernel<0><<<(int)grid, kAttnThreads, smem, stream>>>(tm, p); break;
  }
  XCLIP_LAUNCH_CHECK("attn_fwd_pp_kernel");
  return XCLIP_OK;
}

}  // namespace xclip

using namespace xclip;

namespace xclip {
int attn_fwd_small(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask, void* o, int64_t ldo,
                   float* lse, int B, int n, int heads, float scale, int causal,
                   cudaStream_t stream);   // attention_small.cu
}

extern "C" int xclip_attn_fwd(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask, void* o,
                              int64_t ldo, float* lse, int B, int n, int heads, float scale,
                              int causal, xclip_stream_t stream) {
  int rc = xclip_init();
  if (rc) return rc;
  XCLIP_REQUIRE(qkv && o && lse, "attn_fwd: null pointer");
  XCLIP_REQUIRE(B > 0 && heads > 0 && n > 0, "attn_fwd: bad sizes B=%d n=%d heads=%d", B, n, heads);
  XCLIP_REQUIRE(n <= 320, "attn_fwd: sequence length %d > 320 is not supported by this kernel", n);
  XCLIP_REQUIRE(ld_qkv % 8 == 0 && ld_qkv >= 3 * heads * kDh, "attn_fwd: bad ld_qkv");
  XCLIP_REQUIRE(ldo % 8 == 0 && ldo >= heads * kDh, "attn_fwd: bad ldo");
  XCLIP_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(o) & 15) == 0,
                "attn_fwd: misaligned pointer");
  if (n <= kTile)
    return attn_fwd_small(qkv, ld_qkv, key_mask, o, ldo, lse, B, n, heads, scale, causal,
                          reinterpret_cast<cudaStream_t>(stream));
  XCLIP_REQUIRE(!causal, "attn_fwd: the causal mask is only implemented for n <= 128 (got n=%d)", n);
  AttnFwdParams p;
  p.B = B; p.H = heads; p.n = n;
  p.nkp = (n + 15) / 16 * 16;
  const int need = p.nkp + kDh;
  p.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
  p.scale_log2 = scale * 1.4426950408889634f;
  p.mask = key_mask;
  p.o = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.lse = lse;

  CUtensorMap tm;
  rc = encode_3d_bf16(&tm, qkv, (uint64_t)(3 * heads * kDh), (uint64_t)n, (uint64_t)B,
                      (uint64_t)ld_qkv, (uint64_t)n * ld_qkv, kDh, kTile);
  if (rc) return rc;

  static const bool use_pp = [] { const char* e = getenv("XCLIP_ATTN_PP"); return !(e && e[0] == '0'); }();
  if (use_pp && n > kTile && p.nkp - ((p.nkp / 2) + 15) / 16 * 16 >= 16) {
    // n = 128k + 1 (a CLS token on top of whole tiles): the last query is done on CUDA cores
    const bool tail = attn_tail_enabled() && n % kTile == 1;
    rc = launch_attn_fwd_pp(tm, p, tail, reinterpret_cast<cudaStream_t>(stream));
    if (rc || !tail) return rc;
    return launch_attn_fwd_tail(qkv, ld_qkv, key_mask, o, ldo, lse, B, heads, n, p.scale_log2,
                                reinterpret_cast<cudaStream_t>(stream));
  }

  const int nkb = (n + kTile - 1) / kTile;
  const int npb = (p.nkp + 63) / 64;
  const int smem = (2 + 2 * nkb + npb) * kBoxBytes + 64 + 2 * 384 * 4 + 4 * 128 * 4;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem));
    configured_smem = smem;
  }
  // CTAs per SM are bounded by shared memory, TMEM columns (512 per SM) and registers
  int per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)num_sms() * per_sm;
  if (grid > (long long)B * heads) grid = (long long)B * heads;
  attn_fwd_kernel<<<(int)grid, kAttnThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
  XCLIP_LAUNCH_CHECK("attn_fwd_kernel");
  return XCLIP_OK;
}
