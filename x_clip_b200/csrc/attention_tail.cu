// Attention "tail token" kernels (CUDA cores) for sequence lengths n = 128*k + 1.
//
// A CLS token makes the common sequence lengths 129 and 257: one token more than a whole
// number of 128-row tensor-core tiles.  Giving that token its own tile costs a full tile
// latency in the forward kernel and turns 4 (key tile, query tile) pairs into 9 in the
// backward kernel.  Instead the tensor-core kernels cover the queries [0, n-1) (forward) and
// the [0,n-1) x [0,n-1) block (backward), and the kernels below do the O(n * 64) work of the
// last token z = n-1 per (batch, head) in fp32:
//   forward : o_z, lse_z  (query z against all n keys).
//   backward: the column of key z (all queries) and the row of query z (keys < z):
//               p^c_i = 2^(c q_i.k_z - lse_i)        ds^c_i = p^c_i (dO_i.v_z - delta_i) scale
//               p^r_j = 2^(c q_z.k_j - lse_z)        ds^r_j = p^r_j (dO_z.v_j - delta_z) scale
//             dV_z = sum_i p^c_i dO_i    dK_z = sum_i ds^c_i q_i
//             dQ_z = sum_{j<z} ds^r_j k_j + ds^c_z k_z
//             and, for every token t < z, the three scalars (ds^c_t, ds^r_t, p^r_t): the
//             tensor-core kernel adds the rank-1 terms  dQ_t += ds^c_t k_z,  dK_t += ds^r_t q_z,
//             dV_t += p^r_t dO_z  in its epilogues.  The scalars travel in the first 3 floats of
//             token t's slot of the fp32 dQ workspace (only this (b,h) ever touches that slot).
// Same math as x_clip/x_clip.py:217-244 (and its autograd), restricted to one row / column.
#include "common.cuh"
#include "host.h"

namespace xclip {

constexpr int kTailThreads = 128;   // 16 token groups of 8 lanes; a lane owns 8 of the 64 dims

__device__ __forceinline__ void tail_load8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = unpack_bf16x2(w[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}
__device__ __forceinline__ float tail_dot8(const float (&a)[8], const float (&b)[8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
  return s;
}
__device__ __forceinline__ float tail_sum8lanes(float s) {   // over the 8 lanes of a token group
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  return s;
}
__device__ __forceinline__ float tail_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sum a per-lane 8-vector over the 16 token groups of the CTA; result valid in warp 0, lanes 0..7
__device__ __forceinline__ void tail_block_sum8(float (&a)[8], float* s_red /* [4][8][8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] += __shfl_xor_sync(0xffffffffu, a[i], 8);
    a[i] += __shfl_xor_sync(0xffffffffu, a[i], 16);
  }
  __syncthreads();   // s_red may still be read from a previous reduction
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_red[(warp * 8 + lane) * 8 + i] = a[i];
  }
  __syncthreads();
  if (warp == 0 && lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      a[i] = s_red[(0 * 8 + lane) * 8 + i] + s_red[(1 * 8 + lane) * 8 + i] +
             s_red[(2 * 8 + lane) * 8 + i] + s_red[(3 * 8 + lane) * 8 + i];
  }
}
__device__ __forceinline__ void tail_store8(bf16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// ------------------------------------------------------------------------------------------
// forward: one CTA per (b,h)
__global__ void __launch_bounds__(kTailThreads)
attn_fwd_tail_kernel(const bf16* __restrict__ qkv, long long ld, const uint8_t* __restrict__ mask,
                     bf16* __restrict__ o, long long ldo, float* __restrict__ lse, int B, int H,
                     int n, float scale_log2) {
  __shared__ float s_t[320];
  __shared__ float s_red[4 * 8 * 8];
  __shared__ float s_stat[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int inner = H * 64;
  for (int bh = blockIdx.x; bh < B * H; bh += gridDim.x) {
    const int b = bh / H, h = bh - b * H;
    const int z = n - 1;
    const bf16* base = qkv + (long long)b * n * ld + h * 64 + sub * 8;
    float qz[8];
    tail_load8(base + (long long)z * ld, qz);
    float m = -INFINITY;
    // NOTE: every loop that contains a full-mask shuffle runs a warp-uniform number of trips
    // (j0 is uniform; the token index j = j0 + grp is only predicated).
    for (int j0 = 0; j0 < n; j0 += 16) {
      const int j = j0 + grp;
      const bool valid = j < n;
      const int jj = valid ? j : n - 1;
      float kj[8];
      tail_load8(base + (long long)jj * ld + inner, kj);
      const float s = tail_sum8lanes(tail_dot8(qz, kj));
      const bool keep = mask ? (mask[(long long)b * n + jj] != 0) : true;
      const float t = keep ? s * scale_log2 : -FLT_MAX;
      if (valid) {
        if (sub == 0) s_t[j] = t;
        m = fmaxf(m, t);
      }
    }
    m = warp_max(m);
    __syncthreads();                    // s_stat free (previous item fully consumed), s_t visible
    if (lane == 0) s_stat[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float l = 0.f;
    for (int j = grp; j < n; j += 16) {
      float vj[8];
      tail_load8(base + (long long)j * ld + 2 * inner, vj);
      const float pj = tail_ex2(s_t[j] - m);
      l += pj;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, vj[i], acc[i]);
    }
    // every lane of a group holds the same l: sum over the 4 groups of the warp, then the warps
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
    if (lane == 0) s_stat[4 + warp] = l;
    tail_block_sum8(acc, s_red);        // contains __syncthreads (publishes s_stat[4..7] too)
    if (warp == 0 && lane < 8) {
      const float L = s_stat[4] + s_stat[5] + s_stat[6] + s_stat[7];
      const float inv = 1.f / L;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= inv;
      tail_store8(o + ((long long)b * n + z) * ldo + h * 64 + lane * 8, acc);
      if (lane == 0) lse[((long long)b * H + h) * n + z] = m + log2f(L);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// backward: one CTA per (b,h).  Needs delta (rowsum(dO*O)) of every token.
__global__ void __launch_bounds__(kTailThreads)
attn_bwd_tail_kernel(const bf16* __restrict__ qkv, long long ld, const uint8_t* __restrict__ mask,
                     const bf16* __restrict__ d_o, long long lddo, const float* __restrict__ lse,
                     const float* __restrict__ delta, bf16* __restrict__ dqkv, long long ldg,
                     float* __restrict__ ws, int B, int H, int n, float scale, float scale_log2) {
  __shared__ float s_red[4 * 8 * 8];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int inner = H * 64;
  for (int bh = blockIdx.x; bh < B * H; bh += gridDim.x) {
    const int b = bh / H, h = bh - b * H;
    const int z = n - 1;
    const bf16* base = qkv + (long long)b * n * ld + h * 64 + sub * 8;
    const bf16* dbase = d_o + (long long)b * n * lddo + h * 64 + sub * 8;
    const float* lse_bh = lse + ((long long)b * H + h) * n;
    const float* delta_bh = delta + ((long long)b * H + h) * n;
    float qz[8], kz[8], vz[8], doz[8];
    tail_load8(base + (long long)z * ld, qz);
    tail_load8(base + (long long)z * ld + inner, kz);
    tail_load8(base + (long long)z * ld + 2 * inner, vz);
    tail_load8(dbase + (long long)z * lddo, doz);
    const float lse_z = lse_bh[z], delta_z = delta_bh[z];
    const bool keep_z = mask ? (mask[(long long)b * n + z] != 0) : true;
    float dvz[8], dkz[8], dqz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dvz[i] = dkz[i] = dqz[i] = 0.f;

    for (int t0 = 0; t0 < n; t0 += 16) {   // warp-uniform trip count (full-mask shuffles inside)
      const bool valid = t0 + grp < n;
      const int t = valid ? t0 + grp : z;
      float qt[8], kt[8], vt[8], dot[8];
      tail_load8(base + (long long)t * ld, qt);
      tail_load8(base + (long long)t * ld + inner, kt);
      tail_load8(base + (long long)t * ld + 2 * inner, vt);
      tail_load8(dbase + (long long)t * lddo, dot);
      const float lse_t = lse_bh[t], delta_t = delta_bh[t];
      const float a = tail_sum8lanes(tail_dot8(qt, kz));     // s[t, z]
      const float bb = tail_sum8lanes(tail_dot8(dot, vz));   // dP[t, z]
      const float c = tail_sum8lanes(tail_dot8(qz, kt));     // s[z, t]
      const float d = tail_sum8lanes(tail_dot8(doz, vt));    // dP[z, t]
      // column of key z (all queries t, including t == z)
      const float p_c = (valid && keep_z) ? tail_ex2(a * scale_log2 - lse_t) : 0.f;
      const float ds_c = p_c * (bb - delta_t) * scale;
      // row of query z (keys t < z; the corner belongs to the column)
      const bool keep_t = mask ? (mask[(long long)b * n + t] != 0) : true;
      const float p_r = (t < z && keep_t) ? tail_ex2(c * scale_log2 - lse_z) : 0.f;
      const float ds_r = p_r * (d - delta_z) * scale;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dvz[i] = fmaf(p_c, dot[i], dvz[i]);
        dkz[i] = fmaf(ds_c, qt[i], dkz[i]);
        dqz[i] = fmaf(ds_r, kt[i], dqz[i]);
      }
      if (!valid) {
        // padding slot of the last trip: p_c = ds_c = p_r = ds_r = 0, nothing to record
      } else if (t == z) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dqz[i] = fmaf(ds_c, kz[i], dqz[i]);
      } else if (sub == 0) {
        float* w = ws + ((long long)b * n + t) * inner + h * 64;
        w[0] = ds_c;
        w[1] = ds_r;
        w[2] = p_r;
      }
    }
    bf16* gz = dqkv + ((long long)b * n + z) * ldg + h * 64;
    tail_block_sum8(dqz, s_red);
    if (warp == 0 && lane < 8) tail_store8(gz + lane * 8, dqz);
    tail_block_sum8(dkz, s_red);
    if (warp == 0 && lane < 8) tail_store8(gz + inner + lane * 8, dkz);
    tail_block_sum8(dvz, s_red);
    if (warp == 0 && lane < 8) tail_store8(gz + 2 * inner + lane * 8, dvz);
  }
}

int launch_attn_fwd_tail(const void* qkv, long long ld, const uint8_t* mask, void* o, long long ldo,
                         float* lse, int B, int H, int n, float scale_log2, cudaStream_t stream) {
  long long grid = (long long)B * H;
  if (grid > (long long)num_sms() * 16) grid = (long long)num_sms() * 16;
  attn_fwd_tail_kernel<<<(int)grid, kTailThreads, 0, stream>>>(
      reinterpret_cast<const bf16*>(qkv), ld, mask, reinterpret_cast<bf16*>(o), ldo, lse, B, H, n,
      scale_log2);
  XCLIP_LAUNCH_CHECK("attn_fwd_tail_kernel");
  return XCLIP_OK;
}

int launch_attn_bwd_tail(const void* qkv, long long ld, const uint8_t* mask, const void* d_o,
                         long long lddo, const float* lse, const float* delta, void* dqkv,
                         long long ldg, float* ws, int B, int H, int n, float scale,
                         cudaStream_t stream) {
  long long grid = (long long)B * H;
  if (grid > (long long)num_sms() * 16) grid = (long long)num_sms() * 16;
  attn_bwd_tail_kernel<<<(int)grid, kTailThreads, 0, stream>>>(
      reinterpret_cast<const bf16*>(qkv), ld, mask, reinterpret_cast<const bf16*>(d_o), lddo, lse,
      delta, reinterpret_cast<bf16*>(dqkv), ldg, ws, B, H, n, scale,
      scale * 1.4426950408889634f);
  XCLIP_LAUNCH_CHECK("attn_bwd_tail_kernel");
  return XCLIP_OK;
}

}  // namespace xclip
