// Tensor-core (tcgen05 + TMEM) backward of the MLP half of the transformer block (dim 32, MLP width 64):
//   x1 = x + drop(a W_proj^T + b_proj) ; y = x1 + drop(fc2(drop(gelu(fc1(LayerNorm2(x1))))))       (erv_block_fused.cu)
// given dy: recompute the forward, then d a, d x1 and all parameter gradients.
//
// One persistent CTA per SM walks 128-token tiles; 512 threads = 4 per token row, thread (row, part) owns a quarter of
// the row's columns (8 of 32, 16 of 64) in registers, so LayerNorm statistics are a thread-local sum plus one 4-way
// shared-memory exchange.  Every product runs on the tensor pipe:
//   M1  p  = a  W_proj^T      M2  h  = n2 W_1^T       M3  dh = do W_2       M4  dn = dh' W_1       M5  da = dp W_proj
//   W1  dW_proj += dp^T [a|1] W2  dW_1 += dh'^T [n2|1] W3  dW_2 += do^T [hd|1]   (accumulated in TMEM over all tiles of the CTA;
//                                                                            the ones column yields the bias gradients)
// Operands are bf16 hi/lo images in the token-major layout of the attention kernels (erv_tc_common.cuh): the same image
// is the K-major A operand of a row product and the MN-major operand of a token reduction.  A split product takes two
// instructions: x_hi x [w_hi | w_lo] and x_lo x w_hi (N-concatenation), error ~2^-17 per operand.
// Dropout masks are the counter hash of erv_block_common.cuh, identical to the forward kernel's.
#include "erv_block_common.cuh"
#include "erv_tc_common.cuh"

namespace erv {
namespace blk {

constexpr uint32_t CH = kTokCh;  // one 8-column chunk of a 128-token image
// image offsets (chunks): A operands of the token reductions first (an M = 128 MMA reads 16 chunks from its base)
constexpr uint32_t IMG_DP = 0, IMG_DO = 8, IMG_DH = 16, IMG_A = 32, IMG_N = 42, IMG_H = 52, IMG_END = 70;
constexpr uint32_t WPF = IMG_END * CH;          // W_proj forward format  [c-chunk][hi rows | lo rows]     4 KB
constexpr uint32_t W1F = WPF + 4 * 1024;        // W_1 forward format                                       8 KB
constexpr uint32_t W2B = W1F + 4 * 2048;        // W_2 for dh = do W_2: rows i' (hi | lo), K = o            8 KB
constexpr uint32_t W1B = W2B + 16 * 512;        // W_1 for dn = dh W_1: rows c' (hi | lo), K = j            8 KB
constexpr uint32_t WPB = W1B + 8 * 1024;        // W_proj for da = dp W_proj                                4 KB
constexpr uint32_t TC_SMEM = WPB + 8 * 512;     // 172 KB
constexpr uint32_t COL_G = 0, COL_ACCP = 192, COL_ACC1 = 272, COL_ACC2 = 352;  // TMEM columns (G: 192, accumulators 80/80/144)

__device__ __forceinline__ uint16_t bf16_hi_bits(float v) { return (uint16_t)(__float_as_uint(v) >> 16); }
__device__ __forceinline__ uint16_t bf16_rn_bits(float v) { return (uint16_t)((__float_as_uint(v) + 0x8000u) >> 16); }

// W [NOUT][NIN] -> forward-format image: element (j', c) at (c/8)*WCH + (j'/8)*128 + (j'%8)*16 + (c%8)*2, j' = j | NOUT + j
__device__ void stage_w_fwd(uint8_t* dst, const float* __restrict__ W, int NOUT, int NIN) {
  const uint32_t WCH = (uint32_t)(2 * NOUT / 8) * 128;
  for (int i = threadIdx.x; i < NOUT * NIN; i += blockDim.x) {
    const int j = i / NIN, c = i % NIN;
    const float w = __ldg(W + i);
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffff0000u);
    const int jl = NOUT + j;
    *reinterpret_cast<uint16_t*>(dst + (c >> 3) * WCH + (j >> 3) * 128 + (j & 7) * 16 + (c & 7) * 2) = bf16_hi_bits(w);
    *reinterpret_cast<uint16_t*>(dst + (c >> 3) * WCH + (jl >> 3) * 128 + (jl & 7) * 16 + (c & 7) * 2) = bf16_rn_bits(w - hi);
  }
}
// W [NOUT][NIN] -> dX-format image: element (c', j) at (c'/8)*WCH2 + (j/8)*128 + (j%8)*16 + (c'%8)*2, c' = c | NIN + c
__device__ void stage_w_bwd(uint8_t* dst, const float* __restrict__ W, int NOUT, int NIN) {
  const uint32_t WCH2 = (uint32_t)(NOUT / 8) * 128;
  for (int i = threadIdx.x; i < NOUT * NIN; i += blockDim.x) {
    const int j = i / NIN, c = i % NIN;
    const float w = __ldg(W + i);
    const float hi = __uint_as_float(__float_as_uint(w) & 0xffff0000u);
    const int cl = NIN + c;
    *reinterpret_cast<uint16_t*>(dst + (c >> 3) * WCH2 + (j >> 3) * 128 + (j & 7) * 16 + (c & 7) * 2) = bf16_hi_bits(w);
    *reinterpret_cast<uint16_t*>(dst + (cl >> 3) * WCH2 + (j >> 3) * 128 + (j & 7) * 16 + (cl & 7) * 2) = bf16_rn_bits(w - hi);
  }
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = ld4(p), b = ld4(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  st4(p, make_float4(v[0], v[1], v[2], v[3]));
  st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
// 8 fp32 TMEM columns of the caller's lane: D[col0 .. col0+7] + D[col1 .. col1+7] (hi-part + lo-part products)
__device__ __forceinline__ void tmem_sum8(uint32_t taddr0, uint32_t taddr1, float (&v)[8]) {
  uint32_t r0[8], r1[8];
  tmem_ld8_nowait(taddr0, r0);
  tmem_ld8_nowait(taddr1, r1);
  tmem_wait_ld8(r0);
  tmem_wait_ld8(r1);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r0[i]) + __uint_as_float(r1[i]);
}

__global__ void __launch_bounds__(kTcThreads, 1) mlp_bwd_tc_kernel(const MlpArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_g, bar_w;  // row products / token reductions
  __shared__ uint32_t tmem_base_s;
  __shared__ float ex_a[4][128], ex_b[4][128];
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, part = tid >> 7;
  const uint32_t rowoff = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;
  uint8_t* const img = smem;

  // static parts of the images: zero everything, then the ones columns (element 0 of the ones chunk = bf16 1.0)
  for (uint32_t i = tid; i < IMG_END * CH / 16; i += kTcThreads) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_g, 1);
    mbar_init(&bar_w, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (part == 0) {
    *reinterpret_cast<uint16_t*>(img + (IMG_A + 4) * CH + rowoff) = 0x3F80;
    *reinterpret_cast<uint16_t*>(img + (IMG_N + 4) * CH + rowoff) = 0x3F80;
    *reinterpret_cast<uint16_t*>(img + (IMG_H + 8) * CH + rowoff) = 0x3F80;
  }
  stage_w_fwd(smem + WPF, p.w_proj, C, C);
  stage_w_fwd(smem + W1F, p.w1, MLP, C);
  stage_w_bwd(smem + W2B, p.w2, C, MLP);
  stage_w_bwd(smem + W1B, p.w1, MLP, C);
  stage_w_bwd(smem + WPB, p.w_proj, C, C);
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t sb = smem_u32(smem);
  uint32_t ph_g = 0, ph_w = 0;

  // per-thread parameters of its 8 (or 16) columns
  const int c0 = part * 8, h0 = part * 16;
  float bp[8], gam[8], bet[8], b1[16];
  ld8(p.b_proj + c0, bp); ld8(p.ln_w + c0, gam); ld8(p.ln_b + c0, bet);
  {
    float t[8];
    ld8(p.b1 + h0, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) b1[i] = t[i];
    ld8(p.b1 + h0 + 8, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) b1[8 + i] = t[i];
  }
  const bool drop = p.p_drop > 0.f;
  const unsigned long long seed = drop ? (unsigned long long)*p.seed : 0ull;
  const uint32_t thresh = drop ? (uint32_t)fminf(p.p_drop * 4294967296.0f, 4294967295.0f) : 0u;
  const float inv_keep = drop ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const uint32_t s0 = (uint32_t)p.salt * 4u;
  float dgam[8], dbet[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dgam[i] = dbet[i] = 0.f;

  // instruction descriptors: K-major A x K-major B (forward), K-major A x MN-major B (dX), MN-major A x MN-major B (dW)
  const uint32_t id_f64 = make_idesc(FMT_BF16, 128, 64, false, false), id_f32 = make_idesc(FMT_BF16, 128, 32, false, false);
  const uint32_t id_f128 = make_idesc(FMT_BF16, 128, 128, false, false);
  const uint32_t id_x128 = make_idesc(FMT_BF16, 128, 128, false, true), id_x64 = make_idesc(FMT_BF16, 128, 64, false, true);
  const uint32_t id_x32 = make_idesc(FMT_BF16, 128, 32, false, true);
  const uint32_t id_w80 = make_idesc(FMT_BF16, 128, 80, true, true), id_w48 = make_idesc(FMT_BF16, 128, 48, true, true);
  const uint32_t id_w144 = make_idesc(FMT_BF16, 128, 144, true, true);

  // row product: D[cols] = A(image at chunk a_ch, lo at a_ch + a_lo; K = 16*ks) x B(weights), split in two instructions
  auto row_product = [&](uint32_t dcol, uint32_t a_ch, uint32_t a_lo, int ks, uint32_t w_off, uint32_t w_step, uint32_t w_lbo,
                         uint32_t w_sbo, uint32_t id_full, uint32_t id_half) {
    for (int s = 0; s < ks; ++s) {
      const uint64_t bd = make_desc(sb + w_off + (uint32_t)s * w_step, w_lbo, w_sbo);
      mma_f16(tm + dcol, make_desc(sb + (a_ch + 2 * s) * CH, CH, 128), bd, id_full, s > 0);
      mma_f16(tm + dcol, make_desc(sb + (a_ch + a_lo + 2 * s) * CH, CH, 128), bd, id_half, true);
    }
  };
  // token reduction: D[cols] (+)= A^T(image rows at chunk a_ch, lo at + a_lo) x B(image at chunk b_ch, [hi|1|lo])
  auto token_reduction = [&](uint32_t dcol, uint32_t a_ch, uint32_t a_lo, uint32_t b_ch, uint32_t id_full, uint32_t id_half,
                             bool first) {
    for (int s = 0; s < 8; ++s) {
      const uint64_t bd = make_desc(sb + b_ch * CH + (uint32_t)s * 256, 128, CH);
      mma_f16(tm + dcol, make_desc(sb + a_ch * CH + (uint32_t)s * 256, 128, CH), bd, id_full, !(first && s == 0));
      mma_f16(tm + dcol, make_desc(sb + (a_ch + a_lo) * CH + (uint32_t)s * 256, 128, CH), bd, id_half, true);
    }
  };
  auto wait_g = [&]() {
    mbar_wait(&bar_g, ph_g);
    ph_g ^= 1;
    fence_after_sync();
  };
  auto publish = [&]() {  // generic-proxy image writes -> visible to the tensor pipe; all threads
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
  };

  const int ntiles = (p.R + 127) / 128;
  bool first = true;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r = tile * 128 + row;
    const bool valid = r < p.R;
    float a[8], x[8], dy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x[i] = dy[i] = 0.f;
    if (valid) {
      ld8(p.a + (size_t)r * C + c0, a);
      ld8(p.x + (size_t)r * C + c0, x);
      ld8(p.dy + (size_t)r * C + c0, dy);
    }
    if (!first) {  // the previous tile's token reductions still read the images
      mbar_wait(&bar_w, ph_w);
      ph_w ^= 1;
      fence_after_sync();
    }
    // ---- P0: a and do = dy * mask3 images ; M1 (projection) and M3 (dh = do W_2)
    store_split8(img + IMG_A * CH, img + (IMG_A + 6) * CH, (uint32_t)part * CH + rowoff, a);
    {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        d[i] = drop ? dy[i] * drop_scale(seed, s0 + 2, (uint32_t)r * C + c0 + i, thresh, inv_keep) : dy[i];
      store_split8(img + IMG_DO * CH, img + (IMG_DO + 4) * CH, (uint32_t)part * CH + rowoff, d);
    }
    publish();
    if (tid == 0) {
      fence_after_sync();
      row_product(COL_G, IMG_A, 6, 2, WPF, 2 * 1024, 1024, 128, id_f64, id_f32);          // p      -> G[0, 64)
      row_product(COL_G + 64, IMG_DO, 4, 2, W2B, 256, 128, 512, id_x128, id_x64);         // dh raw -> G[64, 192)
      commit(&bar_g);
    }
    wait_g();
    // ---- P1: x1, LayerNorm2 -> n2 image ; keep dh raw ; M2 (fc1)
    float x1[8], m1[8], xh[8], dhr[16];
    {
      float pv[8];
      tmem_sum8(tm + lane_off + COL_G + c0, tm + lane_off + COL_G + 32 + c0, pv);
      float t0[8], t1[8];
      tmem_sum8(tm + lane_off + COL_G + 64 + h0, tm + lane_off + COL_G + 128 + h0, t0);
      tmem_sum8(tm + lane_off + COL_G + 64 + h0 + 8, tm + lane_off + COL_G + 128 + h0 + 8, t1);
#pragma unroll
      for (int i = 0; i < 8; ++i) { dhr[i] = t0[i]; dhr[8 + i] = t1[i]; }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        m1[i] = drop ? drop_scale(seed, s0, (uint32_t)r * C + c0 + i, thresh, inv_keep) : 1.0f;
        x1[i] = x[i] + (pv[i] + bp[i]) * m1[i];
        s += x1[i];
      }
      ex_a[part][row] = s;
    }
    fence_before_sync();
    __syncthreads();
    const float mean = ((ex_a[0][row] + ex_a[1][row]) + (ex_a[2][row] + ex_a[3][row])) * (1.0f / C);
    {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = x1[i] - mean;
        s = fmaf(d, d, s);
      }
      ex_b[part][row] = s;
    }
    __syncthreads();
    const float rstd = rsqrtf(((ex_b[0][row] + ex_b[1][row]) + (ex_b[2][row] + ex_b[3][row])) * (1.0f / C) + p.eps);
    {
      float n2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[i] = (x1[i] - mean) * rstd;
        n2[i] = valid ? xh[i] * gam[i] + bet[i] : 0.f;
      }
      store_split8(img + IMG_N * CH, img + (IMG_N + 6) * CH, (uint32_t)part * CH + rowoff, n2);
    }
    publish();
    if (tid == 0) {
      fence_after_sync();
      row_product(COL_G, IMG_N, 6, 2, W1F, 2 * 2048, 2048, 128, id_f128, id_f64);         // h -> G[0, 128)
      commit(&bar_g);
    }
    wait_g();
    // ---- P2: gelu, masks -> hd and dh' images ; M4 (dn = dh' W_1)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float hv[8], hd[8], dh[8];
      tmem_sum8(tm + lane_off + COL_G + h0 + 8 * half, tm + lane_off + COL_G + 64 + h0 + 8 * half, hv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ms = drop ? drop_scale(seed, s0 + 1, (uint32_t)r * MLP + h0 + 8 * half + i, thresh, inv_keep) : 1.0f;
        float gy, gd;
        gelu_both(hv[i] + b1[8 * half + i], gy, gd);
        hd[i] = valid ? gy * ms : 0.f;
        dh[i] = dhr[8 * half + i] * ms * gd;
      }
      store_split8(img + IMG_H * CH, img + (IMG_H + 10) * CH, (uint32_t)(2 * part + half) * CH + rowoff, hd);
      store_split8(img + IMG_DH * CH, img + (IMG_DH + 8) * CH, (uint32_t)(2 * part + half) * CH + rowoff, dh);
    }
    publish();
    if (tid == 0) {
      fence_after_sync();
      row_product(COL_G, IMG_DH, 8, 4, W1B, 256, 128, 1024, id_x64, id_x32);              // dn -> G[0, 64)
      commit(&bar_g);
    }
    wait_g();
    // ---- P3: LayerNorm2 backward -> dx1 ; dp image ; M5 (da) and the three token reductions
    float dx1[8];
    {
      float dn[8];
      tmem_sum8(tm + lane_off + COL_G + c0, tm + lane_off + COL_G + 32 + c0, dn);
      float sa = 0.f, sb2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dgam[i] = fmaf(dn[i], xh[i], dgam[i]);
        dbet[i] += dn[i];
        dn[i] *= gam[i];  // d xhat
        sa += dn[i];
        sb2 = fmaf(dn[i], xh[i], sb2);
      }
      ex_a[part][row] = sa;
      ex_b[part][row] = sb2;
      fence_before_sync();
      __syncthreads();
      const float a1 = ((ex_a[0][row] + ex_a[1][row]) + (ex_a[2][row] + ex_a[3][row])) * (1.0f / C);
      const float a2 = ((ex_b[0][row] + ex_b[1][row]) + (ex_b[2][row] + ex_b[3][row])) * (1.0f / C);
      float dp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dx1[i] = dy[i] + rstd * (dn[i] - a1 - xh[i] * a2);
        dp[i] = dx1[i] * m1[i];
      }
      store_split8(img + IMG_DP * CH, img + (IMG_DP + 4) * CH, (uint32_t)part * CH + rowoff, dp);
    }
    publish();
    if (tid == 0) {
      fence_after_sync();
      row_product(COL_G + 64, IMG_DP, 4, 2, WPB, 256, 128, 512, id_x64, id_x32);           // da -> G[64, 128)
      commit(&bar_g);
      token_reduction(COL_ACCP, IMG_DP, 4, IMG_A, id_w80, id_w48, first);                  // dW_proj | db_proj
      token_reduction(COL_ACC1, IMG_DH, 8, IMG_N, id_w80, id_w48, first);                  // dW_1 | db_1
      token_reduction(COL_ACC2, IMG_DO, 4, IMG_H, id_w144, id_w80, first);                 // dW_2 | db_2
      commit(&bar_w);
    }
    if (valid) st8(p.dx1 + (size_t)r * C + c0, dx1);
    wait_g();
    // ---- P4: da
    {
      float da[8];
      tmem_sum8(tm + lane_off + COL_G + 64 + c0, tm + lane_off + COL_G + 96 + c0, da);
      if (valid) st8(p.da + (size_t)r * C + c0, da);
    }
    fence_before_sync();
    __syncthreads();  // ex_a / ex_b and the G columns are reused by the next tile
    first = false;
  }
  if (!first) {
    mbar_wait(&bar_w, ph_w);
    ph_w ^= 1;
    fence_after_sync();
  }
  // ---- parameter gradients of this CTA -> partials
  float* part_out = p.part + (size_t)blockIdx.x * P_MLP;
  if (first) {  // this CTA had no tile
    for (int i = tid; i < P_MLP; i += kTcThreads) part_out[i] = 0.f;
  } else {
    if (part == 0 && warp == 0) {  // rows o < 32 of dW_proj: cols [0,32) + [48,80), bias col 32
      float v[16], w[16];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        tmem_ld16(tm + lane_off + COL_ACCP + 16 * q, v);
        tmem_ld16(tm + lane_off + COL_ACCP + 48 + 16 * q, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) part_out[O_PROJ + row * C + 16 * q + i] = v[i] + w[i];
      }
      tmem_ld16(tm + lane_off + COL_ACCP + 32, v);
      part_out[O_BPROJ + row] = v[0];
    } else if (part == 1 && (warp & 3) < 2) {  // rows j < 64 of dW_1
      float v[16], w[16];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        tmem_ld16(tm + lane_off + COL_ACC1 + 16 * q, v);
        tmem_ld16(tm + lane_off + COL_ACC1 + 48 + 16 * q, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) part_out[O_W1 + row * C + 16 * q + i] = v[i] + w[i];
      }
      tmem_ld16(tm + lane_off + COL_ACC1 + 32, v);
      part_out[O_B1 + row] = v[0];
    } else if (part == 2 && (warp & 3) == 0) {  // rows o < 32 of dW_2: cols [0,64) + [80,144), bias col 64
      float v[16], w[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_ld16(tm + lane_off + COL_ACC2 + 16 * q, v);
        tmem_ld16(tm + lane_off + COL_ACC2 + 80 + 16 * q, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) part_out[O_W2 + row * MLP + 16 * q + i] = v[i] + w[i];
      }
      tmem_ld16(tm + lane_off + COL_ACC2 + 64, v);
      part_out[O_B2 + row] = v[0];
    }
    // LayerNorm parameter gradients: column sums over the 128 rows (fixed order)
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem);  // [128][64]: dgam (32) | dbet (32); the images are idle now
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[row * 64 + c0 + i] = dgam[i];
      red[row * 64 + 32 + c0 + i] = dbet[i];
    }
    __syncthreads();
    if (tid < 64) {
      float s = 0.f;
      for (int rr = 0; rr < 128; ++rr) s += red[rr * 64 + tid];
      part_out[(tid < 32 ? O_LNW : O_LNB - 32) + tid] = s;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace blk
}  // namespace erv

using namespace erv;
using namespace erv::blk;

namespace erv {
namespace blk {
bool mlp_bwd_tc_enabled() {
  static const bool off = getenv("ERV_DISABLE_BLOCK_TC") != nullptr;
  return !off;
}
// grid <= workspace slots (the caller sized the workspace for `max_ctas` partial vectors)
int launch_mlp_bwd_tc(const MlpArgs& a, int max_ctas, cudaStream_t st, int* grid_out) {
  const int ntiles = (a.R + 127) / 128;
  int grid = ntiles < kNumSMs ? ntiles : kNumSMs;
  if (grid > max_ctas) grid = max_ctas;
  ERV_CUDA(allow_smem(mlp_bwd_tc_kernel, TC_SMEM));
  mlp_bwd_tc_kernel<<<grid, kTcThreads, TC_SMEM, st>>>(a);
  ERV_LAUNCH_CHECK();
  *grid_out = grid;
  return ERV_OK;
}
}  // namespace blk
}  // namespace erv
