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
    } else if (lane == 0) {
      mbar_arrive_expect_tx(kv_bar, 2 * nkb * kBoxBytes);
      for (int i = 0; i < nkb; ++i) {
        tma_load_3d(sK + i * kBoxBytes, &tm_qkv, kv_bar, inner + h * kDh, i * kTile, b);
        tma_load_3d(sV + i * kBoxBytes, &tm_qkv, kv_bar, 2 * inner + h * kDh, i * kTile, b);
      }
    }

    if (is_control) {
      // ===================== control warp: TMA + MMA issue, software pipelined ============
      if (lane == 0) {
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

}  // namespace xclip

using namespace xclip;

extern "C" int xclip_attn_fwd(const void* qkv, int64_t ld_qkv, const uint8_t* key_mask, void* o,
                              int64_t ldo, float* lse, int B, int n, int heads, float scale,
                              xclip_stream_t stream) {
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
