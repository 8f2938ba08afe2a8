// Tensor-core (tcgen05 + TMEM) kernels for the transformer block around the attention core (dim 32, MLP width 64):
//   ln_qkv   qkv = LayerNorm1(x) W_qkv^T (+ b)                              forward and backward
//   mlp      x1 = x + drop(a W_proj^T + b_proj) ; y = x1 + drop(fc2(drop(gelu(fc1(LayerNorm2(x1))))))   forward and backward
// (same functions as erv_block_fused.cu, which keeps the fp32 FFMA2 versions).
//
// One persistent CTA per SM walks 128-token tiles; 512 threads = 4 per token row, thread (row, part) owns a quarter of
// the row's columns (8 of 32, 16 of 64, 24 of 96) in registers, so LayerNorm statistics are a thread-local sum plus one
// 4-way shared-memory exchange.  Every product runs on the tensor pipe from bf16 operand images in the token-major layout
// of the attention kernels (erv_tc_common.cuh): the same image is the K-major A operand of a row product and the
// MN-major operand of a token reduction.
//
// Precision.  Row products (everything that flows on to the next op) use THREE-level bf16 splits of both operands
// (hi + lo + lo2 = 24 significant bits, all rounded to nearest) and keep the six terms down to 2^-24 in three instructions
// by concatenating along N:  a_hi x [w_hi | w_lo | w_lo2],  a_lo x [w_hi | w_lo],  a_lo2 x w_hi.
// Two-level splits (2^-17) were measured to be too coarse here: the attention backward of the ReLU/KERPLE variants is
// ill-conditioned at tokens with a near-zero normaliser and amplified a 2e-5 perturbation of its inputs to 2e-2
// (tests/golden model_performer_relu_most_general).  Token reductions (weight / bias gradients, final outputs, summed over
// all tokens and accumulated in TMEM across the CTA's tiles) use two-level splits: a_hi x [x_hi | 1 | x_lo], a_lo x [x_hi | 1].
// Dropout masks are the counter hash of erv_block_common.cuh, the same in the forward and the backward kernel.
#include "erv_block_common.cuh"
#include "erv_tc_common.cuh"

namespace erv {
namespace blk {

constexpr uint32_t CH = kTokCh;  // one 8-column chunk of a 128-token image (2 KB)

__device__ __forceinline__ uint32_t rn_bf16_word(float v) { return (__float_as_uint(v) + 0x8000u) & 0xffff0000u; }

// Three-level bf16 split of 8 values: hi = rn(v), lo = rn(v - hi), lo2 = rn(v - hi - lo); 16-byte stores into three images.
__device__ __forceinline__ void store_split8_3(uint8_t* hi_img, uint8_t* lo_img, uint8_t* lo2_img, uint32_t off, const float (&v)[8]) {
  uint32_t h[4], l[4], m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // two values per F2FP (cvt.rn.bf16x2.f32): 12 instructions per pair instead of 17 integer ones
    const float a = v[2 * i], b = v[2 * i + 1];
    h[i] = pack_bf16x2_rn(a, b);
    const float ra = a - __uint_as_float(h[i] << 16), rb = b - __uint_as_float(h[i] & 0xffff0000u);
    l[i] = pack_bf16x2_rn(ra, rb);
    m[i] = pack_bf16x2_rn(ra - __uint_as_float(l[i] << 16), rb - __uint_as_float(l[i] & 0xffff0000u));
  }
  *reinterpret_cast<uint4*>(hi_img + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_img + off) = make_uint4(l[0], l[1], l[2], l[3]);
  *reinterpret_cast<uint4*>(lo2_img + off) = make_uint4(m[0], m[1], m[2], m[3]);
}

__device__ __forceinline__ void put3(uint8_t* dst, uint32_t o_hi, uint32_t o_lo, uint32_t o_lo2, float w) {
  const uint32_t hb = rn_bf16_word(w);
  const float r = w - __uint_as_float(hb);
  const uint32_t lb = rn_bf16_word(r);
  const float r2 = r - __uint_as_float(lb);
  *reinterpret_cast<uint16_t*>(dst + o_hi) = (uint16_t)(hb >> 16);
  *reinterpret_cast<uint16_t*>(dst + o_lo) = (uint16_t)(lb >> 16);
  *reinterpret_cast<uint16_t*>(dst + o_lo2) = (uint16_t)((__float_as_uint(r2) + 0x8000u) >> 16);
}
// Rows [j0, j0 + nj) of W [*][NIN] -> forward-format image (B operand of y = x W^T, K-major): element (j', c) at
// (c/8)*WCH + (j'/8)*128 + (j'%8)*16 + (c%8)*2 with j' = jj | nj + jj | 2 nj + jj (hi, lo, lo2), WCH = (3 nj / 8) * 128.
__device__ void stage_w_fwd3(uint8_t* dst, const float* __restrict__ W, int j0, int nj, int NIN) {
  const uint32_t WCH = (uint32_t)(3 * nj / 8) * 128;
  for (int i = threadIdx.x; i < nj * NIN; i += blockDim.x) {
    const int jj = i / NIN, c = i % NIN;
    const uint32_t base = (uint32_t)(c >> 3) * WCH + (c & 7) * 2;
    auto at = [&](int jr) { return base + (uint32_t)(jr >> 3) * 128 + (jr & 7) * 16; };
    put3(dst, at(jj), at(nj + jj), at(2 * nj + jj), __ldg(W + (size_t)(j0 + jj) * NIN + c));
  }
}
// W [NOUT][NIN] -> dX-format image (B operand of dx = dy W, MN-major): element (c', j) at
// (c'/8)*WCH2 + (j/8)*128 + (j%8)*16 + (c'%8)*2 with c' = c | NIN + c | 2 NIN + c, WCH2 = (NOUT / 8) * 128.
__device__ void stage_w_bwd3(uint8_t* dst, const float* __restrict__ W, int NOUT, int NIN) {
  const uint32_t WCH2 = (uint32_t)(NOUT / 8) * 128;
  for (int i = threadIdx.x; i < NOUT * NIN; i += blockDim.x) {
    const int j = i / NIN, c = i % NIN;
    const uint32_t base = (uint32_t)(j >> 3) * 128 + (j & 7) * 16;
    auto at = [&](int cr) { return base + (uint32_t)(cr >> 3) * WCH2 + (cr & 7) * 2; };
    put3(dst, at(c), at(NIN + c), at(2 * NIN + c), __ldg(W + i));
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
// activations the attention core produces / consumes (qkv, dqkv, attention output and its gradient) are bf16 under autocast:
// 8 elements = one 16-byte load / store, converted in registers (no separate cast kernels)
__device__ __forceinline__ void ld8a(const float* p, size_t idx, bool bf16, float (&v)[8]) {
  if (!bf16) { ld8(p + idx, v); return; }
  const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + idx);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void st8a(float* p, size_t idx, bool bf16, const float (&v)[8]) {
  if (!bf16) { st8(p + idx, v); return; }
  uint4 raw;
  uint32_t* w = &raw.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + idx) = raw;
}
// 8 fp32 TMEM columns of the caller's lane, summed over the three column groups of a row product
__device__ __forceinline__ void tmem_sum3(uint32_t t0, uint32_t stride, float (&v)[8]) {
  uint32_t r0[8], r1[8], r2[8];
  tmem_ld8_nowait(t0, r0);
  tmem_ld8_nowait(t0 + stride, r1);
  tmem_ld8_nowait(t0 + 2 * stride, r2);
  tmem_wait_ld8(r0);
  tmem_wait_ld8(r1);
  tmem_wait_ld8(r2);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + __uint_as_float(r2[i]);
}

struct TcCtx { uint32_t tm, sb; };
// Row product D[dcol ..] = A x B with three-level operands.  A: K-major images at chunks a_hi / a_lo / a_lo2 (K = 16 ks);
// B: weight image, k-step stride w_step, descriptor strides (w_lbo, w_sbo); id3 / id2 / id1: N = 3n, 2n, n.
__device__ __forceinline__ void tc_row_product3(const TcCtx& c, uint32_t dcol, uint32_t a_hi, uint32_t a_lo, uint32_t a_lo2, int ks,
                                                uint32_t w_off, uint32_t w_step, uint32_t w_lbo, uint32_t w_sbo, uint32_t id3,
                                                uint32_t id2, uint32_t id1) {
  for (int s = 0; s < ks; ++s) {
    const uint64_t bd = make_desc(c.sb + w_off + (uint32_t)s * w_step, w_lbo, w_sbo);
    mma_f16(c.tm + dcol, make_desc(c.sb + (a_hi + 2 * s) * CH, CH, 128), bd, id3, s > 0);
    mma_f16(c.tm + dcol, make_desc(c.sb + (a_lo + 2 * s) * CH, CH, 128), bd, id2, true);
    mma_f16(c.tm + dcol, make_desc(c.sb + (a_lo2 + 2 * s) * CH, CH, 128), bd, id1, true);
  }
}
// Token reduction D[dcol ..] (+)= A^T x B: A image rows at chunk a_hi (low halves at a_lo), B image at chunk b_ch laid out
// [hi | 1 | 0 | lo]; id_full covers the whole B image, id_half its [hi | 1 | 0] prefix.
__device__ __forceinline__ void tc_token_reduction(const TcCtx& c, uint32_t dcol, uint32_t a_hi, uint32_t a_lo, uint32_t b_ch,
                                                   uint32_t id_full, uint32_t id_half, bool first) {
  for (int s = 0; s < 8; ++s) {
    const uint64_t bd = make_desc(c.sb + b_ch * CH + (uint32_t)s * 256, 128, CH);
    mma_f16(c.tm + dcol, make_desc(c.sb + a_hi * CH + (uint32_t)s * 256, 128, CH), bd, id_full, !(first && s == 0));
    mma_f16(c.tm + dcol, make_desc(c.sb + a_lo * CH + (uint32_t)s * 256, 128, CH), bd, id_half, true);
  }
}
// LayerNorm statistics of a 32-wide row held 8 columns per thread by 4 threads (two barriers)
__device__ __forceinline__ void ln_stats4(const float (&v)[8], float (*ex_a)[128], float (*ex_b)[128], int part, int row,
                                          float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  ex_a[part][row] = s;
  fence_before_sync();
  __syncthreads();
  mean = ((ex_a[0][row] + ex_a[1][row]) + (ex_a[2][row] + ex_a[3][row])) * (1.0f / C);
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = v[i] - mean;
    s = fmaf(d, d, s);
  }
  ex_b[part][row] = s;
  __syncthreads();
  rstd = rsqrtf(((ex_b[0][row] + ex_b[1][row]) + (ex_b[2][row] + ex_b[3][row])) * (1.0f / C) + eps);
}

#define ERV_TC_PROLOGUE(NCOLS)                                                              \
  extern __shared__ __align__(128) uint8_t smem[];                                          \
  __shared__ uint32_t tmem_base_s;                                                          \
  __shared__ float ex_a[4][128], ex_b[4][128];                                              \
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, part = tid >> 7;           \
  const uint32_t rowoff = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;                      \
  uint8_t* const img = smem;                                                                \
  (void)ex_a; (void)ex_b;                                                                   \
  if (warp == 0) tmem_alloc(&tmem_base_s, NCOLS)

// ---- backward of the MLP half ----------------------------------------------------------------------------------------------
// images (chunks).  Gradient images [hi | lo | lo2]; activation images [hi | 1 | 0 | lo] + lo2; hd only feeds a reduction.
// The lo2 parts of a and n2 are only read by the first two products, so they share the space of dp's / dh's lo2 parts.
constexpr uint32_t IMG_DP = 0, IMG_DO = 12, IMG_DH = 24, IMG_A = 48, IMG_N = 58, IMG_H = 68, IMG_END = 86;
constexpr uint32_t A_LO2 = IMG_DP + 8, N_LO2 = IMG_DH + 16;  // aliases (see above)
constexpr uint32_t WPF = IMG_END * CH;          // W_proj forward format (3 levels)   4 * 1536 =  6 KB
constexpr uint32_t W1F = WPF + 4 * 1536;        // W_1 forward format                 4 * 3072 = 12 KB
constexpr uint32_t W2B = W1F + 4 * 3072;        // W_2 dX format: rows i', K = o      24 * 512 = 12 KB
constexpr uint32_t W1B = W2B + 24 * 512;        // W_1 dX format: rows c', K = j      12 * 1024 = 12 KB
constexpr uint32_t WPB = W1B + 12 * 1024;       // W_proj dX format                   12 * 512 =  6 KB
constexpr uint32_t TC_SMEM = WPB + 12 * 512;    // 220 KB
constexpr uint32_t COL_G = 0, COL_ACCP = 192, COL_ACC1 = 272, COL_ACC2 = 352;  // TMEM: products 192 cols, accumulators 80/80/144

__global__ void __launch_bounds__(kTcThreads, 1) mlp_bwd_tc_kernel(const MlpArgs p) {
  ERV_TC_PROLOGUE(512);
  __shared__ __align__(8) uint64_t bar_g, bar_w;  // row products / token reductions
  for (uint32_t i = tid; i < IMG_END * CH / 16; i += kTcThreads) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar_g, 1);
    mbar_init(&bar_w, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (part == 0) {  // ones columns of the [hi | 1 | 0 | lo] images
    *reinterpret_cast<uint16_t*>(img + (IMG_A + 4) * CH + rowoff) = 0x3F80;
    *reinterpret_cast<uint16_t*>(img + (IMG_N + 4) * CH + rowoff) = 0x3F80;
    *reinterpret_cast<uint16_t*>(img + (IMG_H + 8) * CH + rowoff) = 0x3F80;
  }
  stage_w_fwd3(smem + WPF, p.w_proj, 0, C, C);
  stage_w_fwd3(smem + W1F, p.w1, 0, MLP, C);
  stage_w_bwd3(smem + W2B, p.w2, C, MLP);
  stage_w_bwd3(smem + W1B, p.w1, MLP, C);
  stage_w_bwd3(smem + WPB, p.w_proj, C, C);
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const TcCtx cx{tmem_base_s, smem_u32(smem)};
  const uint32_t tm = cx.tm, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_g = 0, ph_w = 0;

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

  // K-major A x K-major B (forward products), K-major A x MN-major B (dX products), MN-major A x MN-major B (reductions)
  const uint32_t f96 = make_idesc(FMT_BF16, 128, 96, false, false), f64 = make_idesc(FMT_BF16, 128, 64, false, false);
  const uint32_t f32 = make_idesc(FMT_BF16, 128, 32, false, false), f192 = make_idesc(FMT_BF16, 128, 192, false, false);
  const uint32_t f128 = make_idesc(FMT_BF16, 128, 128, false, false);
  const uint32_t x192 = make_idesc(FMT_BF16, 128, 192, false, true), x128 = make_idesc(FMT_BF16, 128, 128, false, true);
  const uint32_t x96 = make_idesc(FMT_BF16, 128, 96, false, true), x64 = make_idesc(FMT_BF16, 128, 64, false, true);
  const uint32_t x32 = make_idesc(FMT_BF16, 128, 32, false, true);
  const uint32_t w80 = make_idesc(FMT_BF16, 128, 80, true, true), w48 = make_idesc(FMT_BF16, 128, 48, true, true);
  const uint32_t w144 = make_idesc(FMT_BF16, 128, 144, true, true);

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
  float na[8], nx[8], ndy[8];  // rows of the next tile, in flight while this one is processed
  auto load_rows = [&](int tile) {
    const int rr = tile * 128 + row;
#pragma unroll
    for (int i = 0; i < 8; ++i) na[i] = nx[i] = ndy[i] = 0.f;
    if (tile < ntiles && rr < p.R) {
      ld8a(p.a, (size_t)rr * C + c0, p.act_bf16 != 0, na);
      ld8(p.x + (size_t)rr * C + c0, nx);
      ld8(p.dy + (size_t)rr * C + c0, ndy);
    }
  };
  load_rows(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r = tile * 128 + row;
    const bool valid = r < p.R;
    float a[8], x[8], dy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = na[i]; x[i] = nx[i]; dy[i] = ndy[i]; }
    if (!first) {  // the previous tile's token reductions still read the images
      mbar_wait(&bar_w, ph_w);
      ph_w ^= 1;
      fence_after_sync();
    }
    // ---- P0: a and do = dy * mask3 images ; M3 (dh = do W_2), then M1 (projection) in the same columns
    store_split8_3(img + IMG_A * CH, img + (IMG_A + 6) * CH, img + A_LO2 * CH, (uint32_t)part * CH + rowoff, a);
    {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        d[i] = drop ? dy[i] * drop_scale(seed, s0 + 2, (uint32_t)r * C + c0 + i, thresh, inv_keep) : dy[i];
      store_split8_3(img + IMG_DO * CH, img + (IMG_DO + 4) * CH, img + (IMG_DO + 8) * CH, (uint32_t)part * CH + rowoff, d);
    }
    publish();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, COL_G, IMG_DO, IMG_DO + 4, IMG_DO + 8, 2, W2B, 256, 128, 512, x192, x128, x64);  // dh raw -> [0,192)
      commit(&bar_g);
    }
    load_rows(tile + gridDim.x);
    wait_g();
    float dhr[16];
    {
      float t0[8], t1[8];
      tmem_sum3(tm + lane_off + COL_G + h0, 64, t0);
      tmem_sum3(tm + lane_off + COL_G + h0 + 8, 64, t1);
#pragma unroll
      for (int i = 0; i < 8; ++i) { dhr[i] = t0[i]; dhr[8 + i] = t1[i]; }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, COL_G, IMG_A, IMG_A + 6, A_LO2, 2, WPF, 2 * 1536, 1536, 128, f96, f64, f32);     // p -> [0, 96)
      commit(&bar_g);
    }
    wait_g();
    // ---- P1: x1, LayerNorm2 -> n2 image ; M2 (fc1)
    float x1[8], m1[8], xh[8];
    {
      float pv[8];
      tmem_sum3(tm + lane_off + COL_G + c0, 32, pv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        m1[i] = drop ? drop_scale(seed, s0, (uint32_t)r * C + c0 + i, thresh, inv_keep) : 1.0f;
        x1[i] = x[i] + (pv[i] + bp[i]) * m1[i];
      }
    }
    float mean, rstd;
    ln_stats4(x1, ex_a, ex_b, part, row, p.eps, mean, rstd);
    {
      float n2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[i] = (x1[i] - mean) * rstd;
        n2[i] = valid ? xh[i] * gam[i] + bet[i] : 0.f;
      }
      store_split8_3(img + IMG_N * CH, img + (IMG_N + 6) * CH, img + N_LO2 * CH, (uint32_t)part * CH + rowoff, n2);
    }
    publish();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, COL_G, IMG_N, IMG_N + 6, N_LO2, 2, W1F, 2 * 3072, 3072, 128, f192, f128, f64);   // h -> [0, 192)
      commit(&bar_g);
    }
    wait_g();
    // ---- P2: gelu, masks -> hd and dh' images ; M4 (dn = dh' W_1)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float hv[8], hd[8], dh[8];
      tmem_sum3(tm + lane_off + COL_G + h0 + 8 * half, 64, hv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ms = drop ? drop_scale(seed, s0 + 1, (uint32_t)r * MLP + h0 + 8 * half + i, thresh, inv_keep) : 1.0f;
        float gy, gd;
        gelu_both(hv[i] + b1[8 * half + i], gy, gd);
        hd[i] = valid ? gy * ms : 0.f;
        dh[i] = dhr[8 * half + i] * ms * gd;
      }
      store_split8(img + IMG_H * CH, img + (IMG_H + 10) * CH, (uint32_t)(2 * part + half) * CH + rowoff, hd);
      store_split8_3(img + IMG_DH * CH, img + (IMG_DH + 8) * CH, img + (IMG_DH + 16) * CH,
                     (uint32_t)(2 * part + half) * CH + rowoff, dh);
    }
    publish();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, COL_G, IMG_DH, IMG_DH + 8, IMG_DH + 16, 4, W1B, 256, 128, 1024, x96, x64, x32);  // dn -> [0, 96)
      commit(&bar_g);
    }
    wait_g();
    // ---- P3: LayerNorm2 backward -> dx1 ; dp image ; M5 (da) and the three token reductions
    float dx1[8];
    {
      float dn[8];
      tmem_sum3(tm + lane_off + COL_G + c0, 32, dn);
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
      store_split8_3(img + IMG_DP * CH, img + (IMG_DP + 4) * CH, img + (IMG_DP + 8) * CH, (uint32_t)part * CH + rowoff, dp);
    }
    publish();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, COL_G + 96, IMG_DP, IMG_DP + 4, IMG_DP + 8, 2, WPB, 256, 128, 512, x96, x64, x32);  // da -> [96, 192)
      commit(&bar_g);
      tc_token_reduction(cx, COL_ACCP, IMG_DP, IMG_DP + 4, IMG_A, w80, w48, first);    // dW_proj | db_proj
      tc_token_reduction(cx, COL_ACC1, IMG_DH, IMG_DH + 8, IMG_N, w80, w48, first);    // dW_1 | db_1
      tc_token_reduction(cx, COL_ACC2, IMG_DO, IMG_DO + 4, IMG_H, w144, w80, first);   // dW_2 | db_2
      commit(&bar_w);
    }
    if (valid) st8(p.dx1 + (size_t)r * C + c0, dx1);
    wait_g();
    // ---- P4: da
    {
      float da[8];
      tmem_sum3(tm + lane_off + COL_G + 96 + c0, 32, da);
      if (valid) st8a(p.da, (size_t)r * C + c0, p.act_bf16 != 0, da);
    }
    fence_before_sync();
    __syncthreads();  // exchange buffers and the product columns are reused by the next tile
    first = false;
  }
  if (!first) {
    mbar_wait(&bar_w, ph_w);
    ph_w ^= 1;
    fence_after_sync();
  }
  // ---- parameter gradients of this CTA -> partials
  float* part_out = p.part + (size_t)blockIdx.x * P_MLP;
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
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- forward of the MLP half -------------------------------------------------------------------------------------------------
constexpr uint32_t F_A = 0, F_N = 12, F_H = 24, F_END = 48;  // a, n2 [hi 4 | lo 4 | lo2 4], hd [hi 8 | lo 8 | lo2 8]
constexpr uint32_t F_WP = F_END * CH, F_W1 = F_WP + 4 * 1536, F_W2 = F_W1 + 4 * 3072, F_SMEM = F_W2 + 8 * 1536;  // 126 KB

__global__ void __launch_bounds__(kTcThreads, 1) mlp_fwd_tc_kernel(const MlpArgs p) {
  ERV_TC_PROLOGUE(256);
  __shared__ __align__(8) uint64_t bar_g;
  if (tid == 0) {
    mbar_init(&bar_g, 1);
    mbar_init_fence();
  }
  stage_w_fwd3(smem + F_WP, p.w_proj, 0, C, C);
  stage_w_fwd3(smem + F_W1, p.w1, 0, MLP, C);
  stage_w_fwd3(smem + F_W2, p.w2, 0, C, MLP);
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const TcCtx cx{tmem_base_s, smem_u32(smem)};
  const uint32_t tm = cx.tm, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_g = 0;
  const int c0 = part * 8, h0 = part * 16;
  float bp[8], gam[8], bet[8], b2[8], b1[16];
  ld8(p.b_proj + c0, bp); ld8(p.ln_w + c0, gam); ld8(p.ln_b + c0, bet); ld8(p.b2 + c0, b2);
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
  const uint32_t f96 = make_idesc(FMT_BF16, 128, 96, false, false), f64 = make_idesc(FMT_BF16, 128, 64, false, false);
  const uint32_t f32 = make_idesc(FMT_BF16, 128, 32, false, false), f192 = make_idesc(FMT_BF16, 128, 192, false, false);
  const uint32_t f128 = make_idesc(FMT_BF16, 128, 128, false, false);
  auto publish_and_issue = [&](auto&& issue) {
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      issue();
      commit(&bar_g);
    }
    mbar_wait(&bar_g, ph_g);
    ph_g ^= 1;
    fence_after_sync();
  };
  const int ntiles = (p.R + 127) / 128;
  float a[8], x[8];
  auto load_rows = [&](int tile) {
    const int r = tile * 128 + row;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x[i] = 0.f;
    if (tile < ntiles && r < p.R) {
      ld8a(p.a, (size_t)r * C + c0, p.act_bf16 != 0, a);
      ld8(p.x + (size_t)r * C + c0, x);
    }
  };
  load_rows(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r = tile * 128 + row;
    const bool valid = r < p.R;
    store_split8_3(img + F_A * CH, img + (F_A + 4) * CH, img + (F_A + 8) * CH, (uint32_t)part * CH + rowoff, a);
    float x1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x1[i] = x[i];
    publish_and_issue([&]() { tc_row_product3(cx, 0, F_A, F_A + 4, F_A + 8, 2, F_WP, 2 * 1536, 1536, 128, f96, f64, f32); });
    load_rows(tile + gridDim.x);  // the next tile's rows travel during this tile's phases
    {
      float pv[8];
      tmem_sum3(tm + lane_off + c0, 32, pv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float m = drop ? drop_scale(seed, s0, (uint32_t)r * C + c0 + i, thresh, inv_keep) : 1.0f;
        x1[i] += (pv[i] + bp[i]) * m;
      }
    }
    float mean, rstd;
    ln_stats4(x1, ex_a, ex_b, part, row, p.eps, mean, rstd);
    {
      float n2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) n2[i] = (x1[i] - mean) * rstd * gam[i] + bet[i];
      store_split8_3(img + F_N * CH, img + (F_N + 4) * CH, img + (F_N + 8) * CH, (uint32_t)part * CH + rowoff, n2);
    }
    publish_and_issue([&]() { tc_row_product3(cx, 0, F_N, F_N + 4, F_N + 8, 2, F_W1, 2 * 3072, 3072, 128, f192, f128, f64); });
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float hv[8];
      tmem_sum3(tm + lane_off + h0 + 8 * half, 64, hv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float m = drop ? drop_scale(seed, s0 + 1, (uint32_t)r * MLP + h0 + 8 * half + i, thresh, inv_keep) : 1.0f;
        hv[i] = gelu_f(hv[i] + b1[8 * half + i]) * m;
      }
      store_split8_3(img + F_H * CH, img + (F_H + 8) * CH, img + (F_H + 16) * CH, (uint32_t)(2 * part + half) * CH + rowoff, hv);
    }
    publish_and_issue([&]() { tc_row_product3(cx, 0, F_H, F_H + 8, F_H + 16, 4, F_W2, 2 * 1536, 1536, 128, f96, f64, f32); });
    {
      float o[8];
      tmem_sum3(tm + lane_off + c0, 32, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float m = drop ? drop_scale(seed, s0 + 2, (uint32_t)r * C + c0 + i, thresh, inv_keep) : 1.0f;
        o[i] = x1[i] + (o[i] + b2[i]) * m;
      }
      if (valid) st8(p.y + (size_t)r * C + c0, o);
    }
    fence_before_sync();
    __syncthreads();  // TMEM columns, exchange buffers and images are reused by the next tile
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ---- LayerNorm1 + qkv projection ---------------------------------------------------------------------------------------------
struct LnQkvTcArgs {
  const float* x; const float* ln_w; const float* ln_b; const float* w; const float* b;  // b may be null
  float* qkv;
  const float* dqkv; const float* dres;  // backward inputs (dres may be null)
  float* dx; float* part;
  int R; float eps;
  int act_bf16;  // qkv / dqkv are bf16 (autocast) instead of fp32
};
// forward: the 96 outputs are produced as two 48-wide halves (3 x 96 columns would not fit one instruction's N <= 256)
constexpr uint32_t Q_N = 0, Q_END = 12, Q_W = Q_END * CH, Q_WHALF = 4 * 2304, Q_SMEM = Q_W + 2 * Q_WHALF;  // 42 KB

__global__ void __launch_bounds__(kTcThreads, 1) ln_qkv_fwd_tc_kernel(const LnQkvTcArgs p) {
  ERV_TC_PROLOGUE(512);
  __shared__ __align__(8) uint64_t bar_g;
  if (tid == 0) {
    mbar_init(&bar_g, 1);
    mbar_init_fence();
  }
  stage_w_fwd3(smem + Q_W, p.w, 0, 48, C);
  stage_w_fwd3(smem + Q_W + Q_WHALF, p.w, 48, 48, C);
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const TcCtx cx{tmem_base_s, smem_u32(smem)};
  const uint32_t tm = cx.tm, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_g = 0;
  const int c0 = part * 8, q0 = part * 24;
  float gam[8], bet[8], bq[24];
  ld8(p.ln_w + c0, gam); ld8(p.ln_b + c0, bet);
#pragma unroll
  for (int i = 0; i < 24; ++i) bq[i] = p.b ? __ldg(p.b + q0 + i) : 0.f;
  const uint32_t f144 = make_idesc(FMT_BF16, 128, 144, false, false), f96 = make_idesc(FMT_BF16, 128, 96, false, false);
  const uint32_t f48 = make_idesc(FMT_BF16, 128, 48, false, false);
  const int ntiles = (p.R + 127) / 128;
  float x[8];
  auto load_rows = [&](int tile) {
    const int r = tile * 128 + row;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.f;
    if (tile < ntiles && r < p.R) ld8(p.x + (size_t)r * C + c0, x);
  };
  load_rows(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r = tile * 128 + row;
    float mean, rstd;
    ln_stats4(x, ex_a, ex_b, part, row, p.eps, mean, rstd);
    {
      float n[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) n[i] = (x[i] - mean) * rstd * gam[i] + bet[i];
      store_split8_3(img + Q_N * CH, img + (Q_N + 4) * CH, img + (Q_N + 8) * CH, (uint32_t)part * CH + rowoff, n);
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, 0, Q_N, Q_N + 4, Q_N + 8, 2, Q_W, 2 * 2304, 2304, 128, f144, f96, f48);              // outputs 0..47
      tc_row_product3(cx, 144, Q_N, Q_N + 4, Q_N + 8, 2, Q_W + Q_WHALF, 2 * 2304, 2304, 128, f144, f96, f48);  // outputs 48..95
      commit(&bar_g);
    }
    load_rows(tile + gridDim.x);
    mbar_wait(&bar_g, ph_g);
    ph_g ^= 1;
    fence_after_sync();
    const uint32_t qbase = (uint32_t)(part >> 1) * 144 + (part & 1) * 24;  // this thread's 24 outputs inside their half
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float o[8];
      tmem_sum3(tm + lane_off + qbase + 8 * k, 48, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += bq[8 * k + i];
      if (r < p.R) st8a(p.qkv, (size_t)r * QKV + q0 + 8 * k, p.act_bf16 != 0, o);
    }
    fence_before_sync();
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// backward: dx = dres + LN1_bwd(dqkv W_qkv) ; dW_qkv += dqkv^T [LN1(x) | 1] (TMEM accumulator) ; dln_w ; dln_b
constexpr uint32_t B_DQ = 0, B_N = 36, B_END = 46;  // dqkv [hi 12 | lo 12 | lo2 12], n [hi 4 | 1 | 0 | lo 4]
constexpr uint32_t B_W = B_END * CH;
constexpr uint32_t B_SMEM = B_W + 12 * 1536;         // W_qkv dX format (3 levels): rows c', K = j      110 KB
constexpr uint32_t B_COL_ACC = 96;

__global__ void __launch_bounds__(kTcThreads, 1) ln_qkv_bwd_tc_kernel(const LnQkvTcArgs p) {
  ERV_TC_PROLOGUE(256);
  __shared__ __align__(8) uint64_t bar_g, bar_w;
  for (uint32_t i = tid; i < B_W / 16; i += kTcThreads) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar_g, 1);
    mbar_init(&bar_w, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (part == 0) *reinterpret_cast<uint16_t*>(img + (B_N + 4) * CH + rowoff) = 0x3F80;
  stage_w_bwd3(smem + B_W, p.w, QKV, C);
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const TcCtx cx{tmem_base_s, smem_u32(smem)};
  const uint32_t tm = cx.tm, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_g = 0, ph_w = 0;
  const int c0 = part * 8, q0 = part * 24;
  float gam[8], bet[8], dgam[8], dbet[8];
  ld8(p.ln_w + c0, gam); ld8(p.ln_b + c0, bet);
#pragma unroll
  for (int i = 0; i < 8; ++i) dgam[i] = dbet[i] = 0.f;
  const uint32_t x96 = make_idesc(FMT_BF16, 128, 96, false, true), x64 = make_idesc(FMT_BF16, 128, 64, false, true);
  const uint32_t x32 = make_idesc(FMT_BF16, 128, 32, false, true);
  const uint32_t w80 = make_idesc(FMT_BF16, 128, 80, true, true), w48 = make_idesc(FMT_BF16, 128, 48, true, true);
  const int ntiles = (p.R + 127) / 128;
  bool first = true;
  float nx[8], ndq[24];  // rows of the next tile, in flight while this one is processed
  auto load_rows = [&](int tile) {
    const int rr = tile * 128 + row;
#pragma unroll
    for (int i = 0; i < 8; ++i) nx[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) ndq[i] = 0.f;
    if (tile < ntiles && rr < p.R) {
      ld8(p.x + (size_t)rr * C + c0, nx);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t[8];
        ld8a(p.dqkv, (size_t)rr * QKV + q0 + 8 * k, p.act_bf16 != 0, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) ndq[8 * k + i] = t[i];
      }
    }
  };
  load_rows(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r = tile * 128 + row;
    const bool valid = r < p.R;
    float x[8], dres[8], dq[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = nx[i]; dres[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 24; ++i) dq[i] = ndq[i];
    if (valid && p.dres) ld8(p.dres + (size_t)r * C + c0, dres);
    if (!first) {  // the previous tile's token reduction still reads the images
      mbar_wait(&bar_w, ph_w);
      ph_w ^= 1;
      fence_after_sync();
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = dq[8 * k + i];
      store_split8_3(img + B_DQ * CH, img + (B_DQ + 12) * CH, img + (B_DQ + 24) * CH, (uint32_t)(3 * part + k) * CH + rowoff, t);
    }
    float mean, rstd, xh[8];
    ln_stats4(x, ex_a, ex_b, part, row, p.eps, mean, rstd);
    {
      float n[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[i] = (x[i] - mean) * rstd;
        n[i] = valid ? xh[i] * gam[i] + bet[i] : 0.f;
      }
      store_split8(img + B_N * CH, img + (B_N + 6) * CH, (uint32_t)part * CH + rowoff, n);
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, 0, B_DQ, B_DQ + 12, B_DQ + 24, 6, B_W, 256, 128, 1536, x96, x64, x32);  // dn -> cols [0, 96)
      commit(&bar_g);
      tc_token_reduction(cx, B_COL_ACC, B_DQ, B_DQ + 12, B_N, w80, w48, first);                    // dW_qkv | db_qkv
      commit(&bar_w);
    }
    load_rows(tile + gridDim.x);
    mbar_wait(&bar_g, ph_g);
    ph_g ^= 1;
    fence_after_sync();
    {
      float dn[8];
      tmem_sum3(tm + lane_off + c0, 32, dn);
      float sa = 0.f, sb2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dgam[i] = fmaf(dn[i], xh[i], dgam[i]);
        dbet[i] += dn[i];
        dn[i] *= gam[i];
        sa += dn[i];
        sb2 = fmaf(dn[i], xh[i], sb2);
      }
      ex_a[part][row] = sa;
      ex_b[part][row] = sb2;
      fence_before_sync();
      __syncthreads();
      const float a1 = ((ex_a[0][row] + ex_a[1][row]) + (ex_a[2][row] + ex_a[3][row])) * (1.0f / C);
      const float a2 = ((ex_b[0][row] + ex_b[1][row]) + (ex_b[2][row] + ex_b[3][row])) * (1.0f / C);
      float dx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dx[i] = dres[i] + rstd * (dn[i] - a1 - xh[i] * a2);
      if (valid) st8(p.dx + (size_t)r * C + c0, dx);
    }
    fence_before_sync();
    __syncthreads();
    first = false;
  }
  if (!first) {
    mbar_wait(&bar_w, ph_w);
    ph_w ^= 1;
    fence_after_sync();
  }
  float* part_out = p.part + (size_t)blockIdx.x * P_QKV;
  if (part == 0 && (warp & 3) < 3) {  // rows j < 96 of dW_qkv: cols [0,32) + [48,80), bias col 32
    float v[16], w[16];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      tmem_ld16(tm + lane_off + B_COL_ACC + 16 * q, v);
      tmem_ld16(tm + lane_off + B_COL_ACC + 48 + 16 * q, w);
#pragma unroll
      for (int i = 0; i < 16; ++i) part_out[row * C + 16 * q + i] = v[i] + w[i];
    }
    tmem_ld16(tm + lane_off + B_COL_ACC + 32, v);
    part_out[QKV * C + row] = v[0];
  }
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);  // [128][64]: dgam | dbet
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[row * 64 + c0 + i] = dgam[i];
    red[row * 64 + 32 + c0 + i] = dbet[i];
  }
  __syncthreads();
  if (tid < 64) {
    float s = 0.f;
    for (int rr = 0; rr < 128; ++rr) s += red[rr * 64 + tid];
    part_out[QKV * C + QKV + tid] = s;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}


// ---- patch embedding + CLS + position embedding (base_vit.py:190-223) for 4x4 patches, <= 4 channels ------------------------
// x[b, 0] = cls + pos[0] ; x[b, 1+p] = patch(b, p) W^T + bias + pos[1+p].  A tile is 128 consecutive tokens of [B*N];
// thread (row, part) gathers the 16 pixels of channel `part` of its token's patch (four 16-byte loads), which are exactly
// the k = part*16 .. +15 columns of the patch row (k = c P^2 + i P + j), so the patch image needs no transposition.
struct EmbedTcArgs {
  const float* img; const float* w; const float* b; const float* cls; const float* pos;
  float* out;
  const float* dout; float* part;  // backward: per-CTA partials [32*PD | 32 | 32 | N*32] = dW | db | dcls | dpos
  int B, Cin, S, G, N, PD;
};
constexpr uint32_t E_A = 0, E_END = 24, E_W = E_END * CH, E_SMEM = E_W + 8 * 1536;  // patches [hi 8 | lo 8 | lo2 8] ; W (K padded to 64): 60 KB

__device__ __forceinline__ void gather_patch16(const EmbedTcArgs& p, long t, int part, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.f;
  if (t >= (long)p.B * p.N || part >= p.Cin) return;
  const int b = (int)(t / p.N), n = (int)(t - (long)b * p.N);
  if (n == 0) return;
  const int gy = (n - 1) / p.G, gx = (n - 1) - gy * p.G;
  const float* src = p.img + (((size_t)b * p.Cin + part) * p.S + gy * 4) * p.S + gx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 q = ld4(src + (size_t)i * p.S);
    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) embed_fwd_tc_kernel(const EmbedTcArgs p) {
  ERV_TC_PROLOGUE(128);
  __shared__ __align__(8) uint64_t bar_g;
  for (uint32_t i = tid; i < E_SMEM / 16; i += kTcThreads) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar_g, 1);
    mbar_init_fence();
  }
  __syncthreads();
  stage_w_fwd3(smem + E_W, p.w, 0, C, p.PD);  // columns k >= PD stay zero
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const TcCtx cx{tmem_base_s, smem_u32(smem)};
  const uint32_t tm = cx.tm, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_g = 0;
  const int c0 = part * 8;
  float bias[8], cls[8];
  ld8(p.b + c0, bias); ld8(p.cls + c0, cls);
  const uint32_t f96 = make_idesc(FMT_BF16, 128, 96, false, false), f64 = make_idesc(FMT_BF16, 128, 64, false, false);
  const uint32_t f32 = make_idesc(FMT_BF16, 128, 32, false, false);
  const long total = (long)p.B * p.N;
  const long ntiles = (total + 127) / 128;
  float nv[16];
  gather_patch16(p, (long)blockIdx.x * 128 + row, part, nv);
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long t = tile * 128 + row;
    {
      float h8[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h8[i] = nv[8 * hh + i];
        store_split8_3(img + E_A * CH, img + (E_A + 8) * CH, img + (E_A + 16) * CH, (uint32_t)(2 * part + hh) * CH + rowoff, h8);
      }
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_row_product3(cx, 0, E_A, E_A + 8, E_A + 16, (p.PD + 15) / 16, E_W, 2 * 1536, 1536, 128, f96, f64, f32);
      commit(&bar_g);
    }
    gather_patch16(p, (tile + gridDim.x) * 128 + row, part, nv);
    mbar_wait(&bar_g, ph_g);
    ph_g ^= 1;
    fence_after_sync();
    float o[8];
    tmem_sum3(tm + lane_off + c0, 32, o);
    if (t < total) {
      const int n = (int)(t % p.N);
      float pos[8];
      ld8(p.pos + (size_t)n * C + c0, pos);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (n == 0 ? cls[i] : o[i] + bias[i]) + pos[i];
      st8(p.out + (size_t)t * C + c0, o);
    }
    fence_before_sync();
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

// backward: dW | db (token reduction dout^T [patch | 1] accumulated in TMEM), dpos / dcls (per-CTA shared-memory sums)
constexpr uint32_t EB_DO = 0, EB_P = 8, EB_END = 26, EB_SMEM = EB_END * CH;  // dout [hi 4 | lo 4] ; patches [hi 8 | 1 | 0 | lo 8]: 52 KB

__global__ void __launch_bounds__(kTcThreads, 1) embed_bwd_tc_kernel(const EmbedTcArgs p) {
  ERV_TC_PROLOGUE(256);
  __shared__ __align__(8) uint64_t bar_w;
  __shared__ float dpos_s[128 * C];  // N <= 128 rows
  for (uint32_t i = tid; i < EB_SMEM / 16; i += kTcThreads) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 128 * C; i += kTcThreads) dpos_s[i] = 0.f;
  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init_fence();
  }
  __syncthreads();
  if (part == 0) *reinterpret_cast<uint16_t*>(img + (EB_P + 8) * CH + rowoff) = 0x3F80;  // ones column
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const TcCtx cx{tmem_base_s, smem_u32(smem)};
  const uint32_t tm = cx.tm, lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_w = 0;
  const int c0 = part * 8;
  const uint32_t w144 = make_idesc(FMT_BF16, 128, 144, true, true), w80 = make_idesc(FMT_BF16, 128, 80, true, true);
  const long total = (long)p.B * p.N;
  const long ntiles = (total + 127) / 128;
  bool first = true;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long t = tile * 128 + row;
    const bool valid = t < total;
    const int n = valid ? (int)(t % p.N) : 0;
    float d[8], v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = 0.f;
    if (valid) ld8(p.dout + (size_t)t * C + c0, d);
    gather_patch16(p, t, part, v);
    if (!first) {  // the previous tile's reduction still reads the images
      mbar_wait(&bar_w, ph_w);
      ph_w ^= 1;
      fence_after_sync();
    }
    {
      float dz[8];  // CLS rows take no part in dW / db
#pragma unroll
      for (int i = 0; i < 8; ++i) dz[i] = (valid && n > 0) ? d[i] : 0.f;
      store_split8(img + EB_DO * CH, img + (EB_DO + 4) * CH, (uint32_t)part * CH + rowoff, dz);
      float h8[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int i = 0; i < 8; ++i) h8[i] = v[8 * hh + i];
        store_split8(img + EB_P * CH, img + (EB_P + 10) * CH, (uint32_t)(2 * part + hh) * CH + rowoff, h8);
      }
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      tc_token_reduction(cx, 0, EB_DO, EB_DO + 4, EB_P, w144, w80, first);
      commit(&bar_w);
    }
    // dpos: rows of a tile that share a position are N apart; add them in groups of N consecutive rows (deterministic)
    for (int g0 = 0; g0 < 128; g0 += p.N) {
      if (valid && row >= g0 && row < g0 + p.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dpos_s[n * C + c0 + i] += d[i];
      }
      __syncthreads();
    }
    first = false;
  }
  if (!first) {
    mbar_wait(&bar_w, ph_w);
    ph_w ^= 1;
    fence_after_sync();
  }
  const int P = C * p.PD + 2 * C + p.N * C;
  float* part_out = p.part + (size_t)blockIdx.x * P;
  if (first) {
    for (int i = tid; i < P; i += kTcThreads) part_out[i] = 0.f;
  } else {
    if (part == 0 && warp == 0) {  // rows o < 32: dW[o][k] = D[k] + D[80 + k], db[o] = D[64]
      float a[16], l[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_ld16(tm + lane_off + 16 * q, a);
        tmem_ld16(tm + lane_off + 80 + 16 * q, l);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (16 * q + i < p.PD) part_out[row * p.PD + 16 * q + i] = a[i] + l[i];
      }
      tmem_ld16(tm + lane_off + 64, a);
      part_out[C * p.PD + row] = a[0];
    }
    for (int i = tid; i < C; i += kTcThreads) part_out[C * p.PD + C + i] = dpos_s[i];           // dcls = row 0
    for (int i = tid; i < p.N * C; i += kTcThreads) part_out[C * p.PD + 2 * C + i] = dpos_s[i];  // dpos
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

}  // namespace blk
}  // namespace erv

using namespace erv;
using namespace erv::blk;

namespace erv {
namespace blk {
static std::atomic<int> g_block_tc{-1};  // -1: environment default, 0 / 1: set by erv_block_set_tensor_core()
bool mlp_bwd_tc_enabled() {
  static const bool off = getenv("ERV_DISABLE_BLOCK_TC") != nullptr;
  const int v = g_block_tc.load(std::memory_order_relaxed);
  return v < 0 ? !off : v != 0;
}
void set_block_tc(int v) { g_block_tc.store(v, std::memory_order_relaxed); }
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

int launch_mlp_fwd_tc(const MlpArgs& a, cudaStream_t st) {
  const int ntiles = (a.R + 127) / 128;
  const int grid = ntiles < kNumSMs ? ntiles : kNumSMs;
  ERV_CUDA(allow_smem(mlp_fwd_tc_kernel, F_SMEM));
  mlp_fwd_tc_kernel<<<grid, kTcThreads, F_SMEM, st>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
int launch_ln_qkv_tc(bool bwd, const float* x, const float* ln_w, const float* ln_b, const float* w, const float* b, float* qkv,
                     const float* dqkv, const float* dres, float* dx, float* part, int rows, float eps, int max_ctas,
                     cudaStream_t st, int* grid_out, int act_bf16) {
  LnQkvTcArgs a{};
  a.x = x; a.ln_w = ln_w; a.ln_b = ln_b; a.w = w; a.b = b; a.qkv = qkv; a.dqkv = dqkv; a.dres = dres; a.dx = dx; a.part = part;
  a.R = rows; a.eps = eps; a.act_bf16 = act_bf16;
  const int ntiles = (rows + 127) / 128;
  int grid = ntiles < kNumSMs ? ntiles : kNumSMs;
  if (bwd && grid > max_ctas) grid = max_ctas;
  if (bwd) {
    ERV_CUDA(allow_smem(ln_qkv_bwd_tc_kernel, B_SMEM));
    ln_qkv_bwd_tc_kernel<<<grid, kTcThreads, B_SMEM, st>>>(a);
  } else {
    ERV_CUDA(allow_smem(ln_qkv_fwd_tc_kernel, Q_SMEM));
    ln_qkv_fwd_tc_kernel<<<grid, kTcThreads, Q_SMEM, st>>>(a);
  }
  ERV_LAUNCH_CHECK();
  if (grid_out) *grid_out = grid;
  return ERV_OK;
}

bool embed_tc_eligible(int Cin, int P, int N) { return mlp_bwd_tc_enabled() && P == 4 && Cin >= 1 && Cin <= 4 && N <= 128; }
static int embed_tc_grid(int B, int N) {
  const long tiles = ((long)B * N + 127) / 128;
  return (int)(tiles < kNumSMs ? tiles : kNumSMs);
}
size_t embed_tc_bwd_workspace(int B, int Cin, int N) {
  return align_up((size_t)embed_tc_grid(B, N) * (C * Cin * 16 + 2 * C + N * C) * sizeof(float), 256);
}
int launch_embed_fwd_tc(const float* images, const float* w, const float* b, const float* cls, const float* pos, float* out,
                        int B, int Cin, int S, cudaStream_t st) {
  EmbedTcArgs a{};
  a.img = images; a.w = w; a.b = b; a.cls = cls; a.pos = pos; a.out = out;
  a.B = B; a.Cin = Cin; a.S = S; a.G = S / 4; a.N = a.G * a.G + 1; a.PD = Cin * 16;
  ERV_CUDA(allow_smem(embed_fwd_tc_kernel, E_SMEM));
  embed_fwd_tc_kernel<<<embed_tc_grid(B, a.N), kTcThreads, E_SMEM, st>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}
// gradients are ADDED to dw [32][PD], db [32], dcls [32], dpos [N][32] (the caller zeroes them when it wants plain values)
int launch_embed_bwd_tc(const float* images, const float* dout, float* dw, float* db, float* dcls, float* dpos, int B, int Cin,
                        int S, float* workspace, cudaStream_t st) {
  EmbedTcArgs a{};
  a.img = images; a.dout = dout; a.part = workspace;
  a.B = B; a.Cin = Cin; a.S = S; a.G = S / 4; a.N = a.G * a.G + 1; a.PD = Cin * 16;
  const int grid = embed_tc_grid(B, a.N);
  ERV_CUDA(allow_smem(embed_bwd_tc_kernel, EB_SMEM));
  embed_bwd_tc_kernel<<<grid, kTcThreads, EB_SMEM, st>>>(a);
  ERV_LAUNCH_CHECK();
  float* dst[4] = {dw, db, dcls, dpos};
  const int seg[5] = {0, C * a.PD, C * a.PD + C, C * a.PD + 2 * C, C * a.PD + 2 * C + a.N * C};
  return launch_sum(workspace, nullptr, grid, seg[4], dst, seg, 4, st);
}
}  // namespace blk
}  // namespace erv
