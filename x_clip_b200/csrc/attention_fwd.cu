// Fused multi-head attention forward for short sequences (n <= 384, dim_head = 64), sm_100a.
//
// Replaces the reference's Attention core, x_clip/x_clip.py:217-244:
//   split heads -> q * dh^-0.5 -> einsum QK^T -> masked_fill(~key_mask, -finfo.max) ->
//   softmax(fp32) -> einsum PV -> merge heads
// which materialises [B,h,n,n] scores in HBM; here scores never leave the SM.
//
// One CTA owns one (batch, head): K and V of that head (<= 3 TMA boxes of 128 tokens each)
// stay in shared memory while the CTA walks the query tiles of 128 rows:
//   S = Q K^T      tcgen05.mma, M=128, N=ceil16(n) (all keys at once), accumulator in TMEM
//   softmax        4 warps, one query row per thread (TMEM lane), exp2 in fp32, no online
//                  rescaling needed because the whole key range is resident
//   P -> smem      bf16, written in the SWIZZLE_128B K-major layout the MMA expects
//   O = P V        tcgen05.mma, M=128, N=64, K=ceil16(n); V consumed MN-major straight from
//                  its TMA box (no transpose)
//   epilogue       O / rowsum -> bf16 -> global ; log-sum-exp (base 2, scaled domain) -> global
//
// Masking follows the reference exactly: a masked key's score is replaced by -FLT_MAX AFTER
// scaling (so a fully masked row would give uniform attention); keys beyond n do not exist.
#include "common.cuh"
#include "host.h"

namespace xclip {

constexpr int kAttnThreads = 160;  // warps 0-3: softmax/epilogue, warp 4: TMA + MMA issue
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

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int nkb = (p.n + kTile - 1) / kTile;   // 128-token boxes of K / V
  const int npb = (p.nkp + 63) / 64;            // 64-key blocks of P
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kBoxBytes;
  uint8_t* sV = sK + nkb * kBoxBytes;
  uint8_t* sP = sV + nkb * kBoxBytes;
  uint8_t* tail = sP + npb * kBoxBytes;
  uint64_t* kv_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_bar = kv_bar + 1;
  uint64_t* s_bar = kv_bar + 2;
  uint64_t* p_bar = kv_bar + 3;
  uint64_t* o_bar = kv_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_bar + 5);
  uint8_t* sMask = tail + 64;  // [nkp] bytes

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(kv_bar, 1);
    mbar_init(q_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(p_bar, 4);
    mbar_init(o_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
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
  uint32_t kv_phase = 0, tile_phase = 0;

  for (int bh = blockIdx.x; bh < p.B * p.H; bh += gridDim.x) {
    const int b = bh / p.H, h = bh % p.H;

    // All roles are done with the previous (b,h): its last o_bar was waited by everybody.
    if (warp < 4) {
      for (int j = threadIdx.x; j < p.nkp; j += 128)
        sMask[j] = (j < p.n) ? (p.mask ? p.mask[(long long)b * p.n + j] : (uint8_t)1) : (uint8_t)0;
      // softmax warps sync among themselves before reading sMask (named barrier 1, 128 threads)
      asm volatile("bar.sync 1, 128;" ::: "memory");
    } else if (lane == 0) {
      mbar_arrive_expect_tx(kv_bar, 2 * nkb * kBoxBytes);
      for (int i = 0; i < nkb; ++i) {
        tma_load_3d(sK + i * kBoxBytes, &tm_qkv, kv_bar, inner + h * kDh, i * kTile, b);
        tma_load_3d(sV + i * kBoxBytes, &tm_qkv, kv_bar, 2 * inner + h * kDh, i * kTile, b);
      }
    }

    for (int qt = 0; qt < num_q_tiles; ++qt) {
      if (warp == 4) {
        // ===================== control warp =====================
        if (lane == 0) {
          mbar_arrive_expect_tx(q_bar, kBoxBytes);
          tma_load_3d(sQ, &tm_qkv, q_bar, h * kDh, qt * kTile, b);
          if (qt == 0) mbar_wait(kv_bar, kv_phase);
          mbar_wait(q_bar, tile_phase);
          tcgen05_fence_after();
          // ---- S = Q K^T : N chunks of <= 256 columns, 4 k-steps of 16 over dim_head
          const uint64_t qd = make_smem_desc(smem_u32(sQ), 0, 1024);
          for (int c0 = 0; c0 < p.nkp; c0 += 256) {
            const int nc = min(256, p.nkp - c0);
            const uint32_t idesc = make_idesc_bf16(kTile, nc, kMajorK, kMajorK);
            const uint64_t kd = make_smem_desc(smem_u32(sK) + c0 * 128, 0, 1024);
#pragma unroll
            for (int k = 0; k < kDh / 16; ++k)
              umma_bf16(tmem_base + c0, desc_advance(qd, k * 32), desc_advance(kd, k * 32), idesc,
                        k > 0 ? 1u : 0u);
          }
          umma_commit(s_bar);
          // ---- O = P V once the softmax warps have written P
          mbar_wait(p_bar, tile_phase);
          tcgen05_fence_after();
          const uint32_t idesc_pv = make_idesc_bf16(kTile, kDh, kMajorK, kMajorMN);
          const int ksteps = p.nkp / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t pd =
                make_smem_desc(smem_u32(sP) + (k >> 2) * kBoxBytes + (k & 3) * 32, 0, 1024);
            const uint64_t vd = make_smem_desc(smem_u32(sV) + k * 2048, 8192, 1024);
            umma_bf16(tmem_o, pd, vd, idesc_pv, k > 0 ? 1u : 0u);
          }
          umma_commit(o_bar);
          mbar_wait(o_bar, tile_phase);  // Q/K/V/P smem and S TMEM are reusable after this
        }
        __syncwarp();
      } else {
        // ===================== softmax + epilogue warps =====================
        const int row = warp * 32 + lane;         // TMEM lane == query row inside the tile
        const int q_idx = qt * kTile + row;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        mbar_wait(s_bar, tile_phase);
        tcgen05_fence_after();

        // pass 1: row maximum of the scaled + masked scores (base-2 domain)
        float m2 = -INFINITY;
        for (int c0 = 0; c0 < p.nkp; c0 += 32) {
          uint32_t v[32];
          if (p.nkp - c0 >= 32) {
            tmem_ld_32x32(t_row + c0, v);
          } else {
            uint32_t w[16];
            tmem_ld_32x16(t_row + c0, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = w[i]; v[16 + i] = 0; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = c0 + i;
            if (j < p.n) {
              const float t = sMask[j] ? __uint_as_float(v[i]) * p.scale_log2 : -FLT_MAX;
              m2 = fmaxf(m2, t);
            }
          }
        }
        // pass 2: probabilities -> bf16 P in smem (SW128 K-major blocks of 64 keys), row sum
        float sum = 0.f;
        for (int c0 = 0; c0 < p.nkp; c0 += 32) {
          uint32_t v[32];
          const bool full = (p.nkp - c0 >= 32);
          if (full) {
            tmem_ld_32x32(t_row + c0, v);
          } else {
            uint32_t w[16];
            tmem_ld_32x16(t_row + c0, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = w[i]; v[16 + i] = 0; }
          }
          tmem_ld_wait();
          float pr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = c0 + i;
            float e = 0.f;
            if (j < p.n) {
              const float t = sMask[j] ? __uint_as_float(v[i]) * p.scale_log2 : -FLT_MAX;
              e = exp2f(t - m2);
            }
            pr[i] = e;
            sum += e;
          }
          uint8_t* blk = sP + (c0 >> 6) * kBoxBytes;
          const int chunk0 = (c0 & 63) >> 3;  // first 16-byte chunk inside the 128-byte row
          const int nchunks = full ? 4 : 2;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (cc < nchunks) {
              uint4 o;
              o.x = pack_bf16x2(pr[cc * 8 + 0], pr[cc * 8 + 1]);
              o.y = pack_bf16x2(pr[cc * 8 + 2], pr[cc * 8 + 3]);
              o.z = pack_bf16x2(pr[cc * 8 + 4], pr[cc * 8 + 5]);
              o.w = pack_bf16x2(pr[cc * 8 + 6], pr[cc * 8 + 7]);
              *reinterpret_cast<uint4*>(blk + swz128(row, chunk0 + cc)) = o;
            }
          }
        }
        fence_proxy_async_smem();   // generic-proxy smem writes -> visible to tcgen05.mma
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_bar);

        // epilogue
        mbar_wait(o_bar, tile_phase);
        tcgen05_fence_after();
        const float inv = 1.f / sum;
        if (q_idx < p.n) {
          p.lse[((long long)b * p.H + h) * p.n + q_idx] = m2 + log2f(sum);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_o + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          if (q_idx < p.n) {
            bf16* dst = p.o + ((long long)b * p.n + q_idx) * p.ldo + h * kDh + c * 32;
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
      tile_phase ^= 1;
    }
    kv_phase ^= 1;
    // every role finished with this (b,h)'s K/V, mask and TMEM before anyone starts the next
    __syncthreads();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
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
  XCLIP_REQUIRE(n <= 384, "attn_fwd: sequence length %d > 384 is not supported by this kernel", n);
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
  const int smem = (1 + 2 * nkb + npb) * kBoxBytes + 64 + 512 + 1024;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    XCLIP_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem));
    configured_smem = smem;
  }
  // CTAs per SM are bounded by shared memory and by TMEM columns (512 per SM)
  int per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)num_sms() * per_sm;
  if (grid > (long long)B * heads) grid = (long long)B * heads;
  attn_fwd_kernel<<<(int)grid, kAttnThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
  XCLIP_LAUNCH_CHECK("attn_fwd_kernel");
  return XCLIP_OK;
}
