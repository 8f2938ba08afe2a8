// KERPLE Toeplitz-masked linear attention on tcgen05 / TMEM (head_dim 16, num_features <= 64).
//
// Reference: favor_plus.py:197-245 + kerple.py:99-344 (FFT route).  Evaluated as A = (phi_q phi_k^T) (.) exp(b[j-i]),
// num = A v, den = A 1, out = num / (den + 1e-6): algebraically the reference's D1 / D2 read-out.
// Backward: dA = r (dO v^T) + dden 1^T, G = dA (.) c, dphi_q = G phi_k, dphi_k = G^T phi_q,
// dv = A^T (r dO), db[delta] = sum over the diagonal j - i = delta of dA (.) A.
//
// What is new against the CUDA-core tile kernels of erv_tileattn.cu:
//   * q / k are read from the packed qkv buffer, L2-normalised and pushed through the feature map INSIDE the tile kernel:
//     P = x W^T as a 3xTF32 tcgen05 product (N = 64), exp / relu straight out of tensor memory, the result written as the
//     bf16 hi/lo operand image of the next product.  phi never goes to HBM (erv_tileattn.cu staged phi_q, phi_k, dphi_q and
//     dphi_k there through separate kernels).
//   * every N x N product runs on the tensor pipe with bf16 hi/lo splits of both operands (three terms, ~2^-17):
//       S  = phi_q phi_k^T            A, B from shared memory (K-major images), D = 128 TMEM columns
//       O  = A~ [v | 1]               A~ = S (.) c written back over the S columns as bf16 hi/lo and read from TENSOR memory
//       backward: U = dO v^T next to S; G (and A~) written back over U (and S); dphi_q (+)= G phi_k, and in the transposed
//       kernel dphi_k (+)= G^T phi_q, dv (+)= A~^T (r dO) -- the phi images double as MN-major B operands.
//   * tiles are 128 tokens of the flattened (batch, token) axis of one head.  Short sequences (N <= 128) pack
//     floor(128 / N) (batch, head) pairs per tile and mask the cross-pair blocks; long sequences walk 128-token tiles.
//   * 512 threads = 4 per tile row: thread (row, quarter) owns 32 of the 128 columns of a score tile (16 of the 64 features).
#include "erv_feat.cuh"
#include "erv_tile_tc.cuh"

namespace erv {

struct KtArgs {
  const void* qkv;
  void* out;          // fwd: output; bwd: saved output
  const void* dout;
  void* dqkv;
  float* den;         // [B*H][N]
  const float* wt;    // [H][Mp][DH+4] W^T rows (prep_wt_public)
  const float* cexp;  // [H][2N-1] exp(bias)
  float* dbias_part;  // bwd: [H][gridDim.x][2N-1], zeroed by the caller
  int B, N, H, M, kind;
  int ppt;            // pairs per tile (N <= 128); 0: one pair per CTA, 128-row tiles
  int nqt;            // 128-row tiles per pair (ppt == 0)
  float inv_sqrt_m;
  FeatGeom g;
};

constexpr int KMP = 64;              // padded feature count of this path
constexpr uint32_t KI_SBO = (KMP / 8) * 128;  // K-major [128 x 64] bf16 image: 8-row groups 1024 B apart, k-chunks 128 B apart
constexpr uint32_t KI_BYTES = 16 * KI_SBO;    // 16 KB per level

// ---- shared-memory plan (bytes) -----------------------------------------------------------------------------
struct KtPlan {
  uint32_t wh, wl, xh, xl, qi, ki, vi, dmn, doi, vk, ex, n2, cs, rowa, rowb, bins, tiles, total;
};
// tiles: fp32 token / feature tiles of the CUDA-core feature-map backward (backward kernels only)
__host__ __device__ inline KtPlan kt_plan(int ldp, bool bwd, bool dkv) {
  KtPlan P;
  uint32_t o = 0;
  P.wh = o; o += KMP * 16 * 4;
  P.wl = o; o += KMP * 16 * 4;
  P.xh = o; o += 128 * 16 * 4;
  P.xl = o; o += 128 * 16 * 4;
  P.qi = o; o += 2 * KI_BYTES;
  P.ki = o; o += 2 * KI_BYTES;
  P.vi = o; if (!bwd) o += 6 * VI_CH;         // forward: [v_hi | 1 | 0 | v_lo] MN-major
  P.dmn = o; if (dkv) o += 6 * VI_CH;          // dkv: [r dO hi | 0 | 0 | r dO lo] MN-major
  P.doi = o; if (bwd) o += 2 * XD_BYTES;       // dO K-major
  P.vk = o; if (bwd) o += 2 * XD_BYTES;        // v K-major
  P.ex = o; o += 4 * 128 * 4;
  P.n2 = o; o += 2 * 128 * 4;  // double-buffered by call parity (written before the first barrier of a call)
  P.cs = o; o += 256 * 4;
  P.rowa = o; o += 128 * 4;
  P.rowb = o; o += 128 * 4;
  P.bins = o; if (bwd && !dkv) o += 16 * 256 * 4;
  o = (o + 127) & ~127u;
  P.tiles = o;
  if (bwd) o += (uint32_t)(3 * KT * 20 + KT * ldp + 4 + 3 * KT + (KTHREADS / KT) * KT) * 4;  // xr, xs, tmp, phi, inv/m/n2, pm
  P.total = o;
  return P;
}

struct KtCtx {
  uint32_t tm, lane_off;
  int tid, warp, row, quarter;
  uint8_t* sm;
  KtPlan P;
  uint64_t* bar_f;   // feature-map product
  uint32_t ph_f;
};

// W^T rows (global fp32 [Mp][DH+4]) -> TF32 hi/lo K-major images [64 features x 16]
__device__ __forceinline__ void kt_load_w(const KtCtx& c, const float* __restrict__ wt, int Mp) {
  for (int i = c.tid; i < KMP * 16; i += KTHREADS) {
    const int f = i >> 4, d = i & 15;
    const float w = (f < Mp) ? __ldg(wt + (size_t)f * 20 + d) : 0.f;
    const float hi = to_tf32(w), lo = to_tf32(w - hi);
    const uint32_t off = off_kmajor(f, d, 4, 4, 128, 512);
    *reinterpret_cast<float*>(c.sm + c.P.wh + off) = hi;
    *reinterpret_cast<float*>(c.sm + c.P.wl + off) = lo;
  }
}

// Feature map of one 128-token tile on the tensor pipe: rows t0 .. t0+rows-1 of `base` (stride tok_stride elements) are
// L2-normalised (favor_plus.py:200-201, no epsilon), P = x W^T (3xTF32, TMEM columns col_p .. col_p+63),
// phi = exp(P - max_f P - |x|^2/2)/sqrt(M) or relu(P)/sqrt(M), written as K-major bf16 hi/lo images [128 x 64] at dst.
// Contains CTA barriers; every thread calls it.  Rows past `rows` produce finite garbage (callers mask them).
template <typename T>
__device__ __forceinline__ void kt_features(KtCtx& c, const T* __restrict__ base, size_t tok_stride, int t0, int rows,
                                            uint32_t col_p, uint8_t* dst, int M, int kind, float inv_sqrt_m) {
  constexpr int DH = 16;
  float* n2_s = reinterpret_cast<float*>(c.sm + c.P.n2) + (c.ph_f ? 128 : 0);
  float* ex_s = reinterpret_cast<float*>(c.sm + c.P.ex);
  if (c.quarter == 0) {
    float x[DH];
#pragma unroll
    for (int a = 0; a < DH; ++a) x[a] = 0.f;
    float n2 = 0.f;
    if (c.row < rows) {
      load_row<T, DH>(base + (size_t)(t0 + c.row) * tok_stride, x);
      float ss = 0.f;
#pragma unroll
      for (int a = 0; a < DH; ++a) ss = fmaf(x[a], x[a], ss);
      const float inv = 1.0f / sqrtf(ss);
#pragma unroll
      for (int a = 0; a < DH; ++a) {
        x[a] *= inv;
        n2 = fmaf(x[a], x[a], n2);
      }
      n2 *= 0.5f;
    }
    n2_s[c.row] = n2;
    store_x_images<DH>(c.sm + c.P.xh, c.sm + c.P.xl, x, c.row);
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  if (c.warp == 0 && elect_one()) {
    fence_after_sync();
    const uint32_t idesc = make_idesc(FMT_TF32, 128, KMP, false, false);
    bool acc = false;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
      const uint32_t xa = smem_u32(c.sm + (term == 1 ? c.P.xl : c.P.xh));
      const uint32_t wb = smem_u32(c.sm + (term == 2 ? c.P.wl : c.P.wh));
#pragma unroll
      for (int s = 0; s < DH / 8; ++s) {
        mma_tf32(c.tm + col_p, make_desc(xa + s * 256, 128, 512), make_desc(wb + s * 256, 128, 512), idesc, acc);
        acc = true;
      }
    }
    commit(c.bar_f);
  }
  mbar_wait(c.bar_f, c.ph_f);
  c.ph_f ^= 1;
  fence_after_sync();
  float pv[16];
  tmem_ld16(c.tm + c.lane_off + col_p + 16 * c.quarter, pv);
  const int f0 = 16 * c.quarter;
  const bool favor = kind == ERV_FEAT_FAVOR;
  if (favor) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (f0 + i < M) m = fmaxf(m, pv[i]);
    ex_s[c.quarter * 128 + c.row] = m;
  }
  fence_before_sync();
  __syncthreads();  // every thread holds its P values: the columns may be reused
  if (favor) {
    const float mx = fmaxf(fmaxf(ex_s[c.row], ex_s[128 + c.row]), fmaxf(ex_s[256 + c.row], ex_s[384 + c.row]));
    const float kLog2e = 1.4426950408889634f;
    const float shift = fmaf(mx + n2_s[c.row], kLog2e, -log2f(inv_sqrt_m));
#pragma unroll
    for (int i = 0; i < 16; ++i) pv[i] = (f0 + i < M) ? ex2_approx(fmaf(pv[i], kLog2e, -shift)) : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) pv[i] = (f0 + i < M) ? fmaxf(pv[i], 0.f) * inv_sqrt_m : 0.f;
  }
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = pv[8 * cc + e];
    store_split8(dst, dst + KI_BYTES, (uint32_t)(c.row >> 3) * KI_SBO + (uint32_t)(2 * c.quarter + cc) * 128 + (c.row & 7) * 16, v);
  }
}

// Toeplitz coefficients of a tile pair into shared memory: packed tiles index the whole table with (j_n - i_n) + N - 1,
// long sequences a 255-entry slice cs[k] = c[(j0 - i0) + k - 127 + N - 1] with k = (key - query) + 127 in tile coordinates.
__device__ __forceinline__ void kt_load_cs(const KtCtx& c, float* cs, const float* __restrict__ cexp, bool packed, int N,
                                           int j0, int i0) {
  if (packed) {
    for (int k = c.tid; k < 2 * N - 1; k += KTHREADS) cs[k] = __ldg(cexp + k);
  } else if (c.tid < 255) {
    const int d = j0 - i0 + c.tid - 127 + N - 1;
    cs[c.tid] = (d >= 0 && d < 2 * N - 1) ? __ldg(cexp + d) : 0.f;
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(KTHREADS, 1) ktile_fwd_kernel(const KtArgs p) {
  constexpr int DH = 16;
  constexpr uint32_t COL_S = 0, COL_O = 128, COL_P = 192;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_f, bar_s, bar_pv;
  __shared__ uint32_t tmem_base_s;
  KtCtx c;
  c.tid = threadIdx.x; c.warp = c.tid >> 5; c.row = c.tid & 127; c.quarter = c.tid >> 7;
  c.sm = smem; c.P = kt_plan(p.g.ldp, false, false); c.bar_f = &bar_f; c.ph_f = 0;
  const int tid = c.tid, warp = c.warp, row = c.row, quarter = c.quarter;
  const int N = p.N, H = p.H, h = blockIdx.y;
  const bool packed = p.ppt > 0;
  int tq0, rows_q, i0 = 0, nkt, kbase;
  if (packed) {
    const int b0 = blockIdx.x * p.ppt;
    tq0 = b0 * N; rows_q = min(p.ppt, p.B - b0) * N; nkt = 1; kbase = tq0;
  } else {
    const int b = blockIdx.x / p.nqt;
    i0 = (blockIdx.x % p.nqt) * KT;
    tq0 = b * N + i0; rows_q = min(KT, N - i0); nkt = p.nqt; kbase = b * N;
  }
  const size_t tok_stride = (size_t)3 * H * DH;
  const T* qb = static_cast<const T*>(p.qkv) + (size_t)h * DH;
  const T* kb = qb + (size_t)H * DH;
  const T* vb = qb + (size_t)2 * H * DH;
  const float* cexp = p.cexp + (size_t)h * (2 * N - 1);
  float* cs = reinterpret_cast<float*>(smem + c.P.cs);
  uint8_t* qi = smem + c.P.qi; uint8_t* ki = smem + c.P.ki; uint8_t* vi = smem + c.P.vi;

  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    mbar_init(&bar_f, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_pv, 1);
    mbar_init_fence();
  }
  kt_load_w(c, p.wt + (size_t)h * p.g.Mp * (DH + 4), p.g.Mp);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  c.tm = tmem_base_s;
  c.lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tm = c.tm, lane_off = c.lane_off;
  uint32_t ph_s = 0, ph_pv = 0;
  const uint32_t idesc_s = make_idesc(FMT_BF16, 128, 128, false, false);
  const uint32_t idesc_pv_a = make_idesc(FMT_BF16, 128, 48, false, true);  // [v_hi | 1 | 0 | v_lo]
  const uint32_t idesc_pv_b = make_idesc(FMT_BF16, 128, 32, false, true);  // [v_hi | 1 | 0]
  const int ks_feat = (p.M + 15) >> 4;
  const int r_pair = packed ? row / N : 0;
  const int r_n = packed ? row - r_pair * N : i0 + row;
  const bool r_ok = row < rows_q;

  kt_features<T>(c, qb, tok_stride, tq0, rows_q, COL_P, qi, p.M, p.kind, p.inv_sqrt_m);

  for (int kt = 0; kt < nkt; ++kt) {
    const int j0 = kt * KT;
    const int tk0 = packed ? kbase : kbase + j0;
    const int rows_k = packed ? rows_q : min(KT, N - j0);
    if (kt > 0) {  // the previous tile's A~ [v|1] product has read the V image and the S columns
      mbar_wait(&bar_pv, ph_pv);
      ph_pv ^= 1;
      fence_after_sync();
    }
    if (quarter == 1) {  // v rows -> MN-major image
      float v[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) v[a] = 0.f;
      const bool ok = row < rows_k;
      if (ok) load_row<T, DH>(vb + (size_t)(tk0 + row) * tok_stride, v);
      kt_store_row_mnmajor(vi, row, v, ok ? 1.f : 0.f);
    }
    kt_load_cs(c, cs, cexp, packed, N, j0, i0);
    kt_features<T>(c, kb, tok_stride, tk0, rows_k, COL_P, ki, p.M, p.kind, p.inv_sqrt_m);
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {  // S = phi_q phi_k^T
      fence_after_sync();
      bool acc = false;
      for (int term = 0; term < 3; ++term) {
        const uint32_t a = smem_u32(qi + (term == 2 ? KI_BYTES : 0));
        const uint32_t b = smem_u32(ki + (term == 1 ? KI_BYTES : 0));
        for (int s = 0; s < ks_feat; ++s) {
          mma_f16(tm + COL_S, make_desc(a + s * 256, 128, KI_SBO), make_desc(b + s * 256, 128, KI_SBO), idesc_s, acc);
          acc = true;
        }
      }
      commit(&bar_s);
    }
    mbar_wait(&bar_s, ph_s);
    ph_s ^= 1;
    fence_after_sync();
    {  // A~ = S (.) c over this thread's 32 columns, written back over them as bf16 hi (first 16 columns) / lo (last 16)
      const int c0 = 32 * quarter;
      const uint32_t cbase = tm + lane_off + COL_S + c0;
      uint32_t sv[32], hw[16], lw[16];
      tmem_ld32_nowait(cbase, sv);
      tmem_wait_ld();
      int cp = 0, cn = 0;
      if (packed) { cp = c0 / N; cn = c0 - cp * N; }
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float a[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = c0 + i + e;
          bool ok = r_ok && col < rows_k;
          int idx;
          if (packed) {
            ok = ok && cp == r_pair;
            idx = cn - r_n + N - 1;
            if (++cn == N) { cn = 0; ++cp; }
          } else {
            idx = col - row + 127;
          }
          a[e] = ok ? __uint_as_float(sv[i + e]) * cs[ok ? idx : 0] : 0.f;
        }
        split_pack2(a[0], a[1], hw[i >> 1], lw[i >> 1]);
      }
      uint32_t w8[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = hw[8 * q + i];
        tmem_st8(cbase + 8 * q, w8);
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = lw[8 * q + i];
        tmem_st8(cbase + 16 + 8 * q, w8);
      }
      tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {  // O (+)= A~ [v | 1 | 0 | v_lo]
      fence_after_sync();
      const int ksteps = (rows_k + 15) >> 4;
      for (int t = 0; t < ksteps; ++t) {
        const uint32_t ca = tm + COL_S + 32 * (t >> 1) + 8 * (t & 1);  // keys 16 t .. 16 t + 15: quarter t/2, hi words
        const uint64_t bd = make_desc(smem_u32(vi) + (uint32_t)t * 256, 128, VI_CH);
        mma_f16_ts(tm + COL_O, ca, bd, idesc_pv_a, kt > 0 || t > 0);
        mma_f16_ts(tm + COL_O, ca + 16, bd, idesc_pv_b, true);
      }
      commit(&bar_pv);
    }
  }
  mbar_wait(&bar_pv, ph_pv);
  ph_pv ^= 1;
  fence_after_sync();
  if (quarter == 0) {  // out = num / (den + eps); columns [0,16) hi-part, 16 = den, [32,48) lo-part
    float o0[32], o1[16];
    tmem_ld32(tm + lane_off + COL_O, o0);
    tmem_ld16(tm + lane_off + COL_O + 32, o1);
    if (r_ok) {
      const float den = o0[DH];
      const float inv = 1.0f / (den + kEps);
      T* ob = static_cast<T*>(p.out) + ((size_t)(tq0 + row) * H + h) * DH;
#pragma unroll
      for (int cc = 0; cc < DH / 4; ++cc)
        st4(ob + 4 * cc, make_float4((o0[4 * cc] + o1[4 * cc]) * inv, (o0[4 * cc + 1] + o1[4 * cc + 1]) * inv,
                                     (o0[4 * cc + 2] + o1[4 * cc + 2]) * inv, (o0[4 * cc + 3] + o1[4 * cc + 3]) * inv));
      const int t = tq0 + row, b = t / N, n = t - b * N;
      p.den[((size_t)b * H + h) * N + n] = den;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ---- backward --------------------------------------------------------------------------------------------------
// TRANSPOSED = false: CTA owns a QUERY tile, walks key tiles: dphi_q, d bias partials -> dq.
// TRANSPOSED = true : CTA owns a KEY tile, walks query tiles (all score tiles transposed: rows = keys): dphi_k, dv -> dk, dv.
template <typename T, bool TRANSPOSED>
__global__ void __launch_bounds__(KTHREADS, 1) ktile_bwd_kernel(const KtArgs p) {
  constexpr int DH = 16, LDM = DH + 4;
  constexpr uint32_t COL_X = 0, COL_Y = 128, COL_DPHI = 256, COL_DV = 320, COL_P = 448;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_f, bar_s, bar_acc;
  __shared__ uint32_t tmem_base_s;
  KtCtx c;
  c.tid = threadIdx.x; c.warp = c.tid >> 5; c.row = c.tid & 127; c.quarter = c.tid >> 7;
  c.sm = smem; c.P = kt_plan(p.g.ldp, true, TRANSPOSED); c.bar_f = &bar_f; c.ph_f = 0;
  const int tid = c.tid, warp = c.warp, row = c.row, quarter = c.quarter;
  const int N = p.N, H = p.H, h = blockIdx.y;
  const bool packed = p.ppt > 0;
  // own tile (rows of every score tile) and the tiles walked (columns)
  int t_own, rows_own, n_own0 = 0, nwalk, wbase;
  if (packed) {
    const int b0 = blockIdx.x * p.ppt;
    t_own = b0 * N; rows_own = min(p.ppt, p.B - b0) * N; nwalk = 1; wbase = t_own;
  } else {
    const int b = blockIdx.x / p.nqt;
    n_own0 = (blockIdx.x % p.nqt) * KT;
    t_own = b * N + n_own0; rows_own = min(KT, N - n_own0); nwalk = p.nqt; wbase = b * N;
  }
  const size_t tok_stride = (size_t)3 * H * DH, out_stride = (size_t)H * DH;
  const T* qb = static_cast<const T*>(p.qkv) + (size_t)h * DH;
  const T* kb = qb + (size_t)H * DH;
  const T* vb = qb + (size_t)2 * H * DH;
  const T* ob = static_cast<const T*>(p.out) + (size_t)h * DH;
  const T* dob = static_cast<const T*>(p.dout) + (size_t)h * DH;
  const float* cexp = p.cexp + (size_t)h * (2 * N - 1);
  const float* wt = p.wt + (size_t)h * p.g.Mp * LDM;
  float* cs = reinterpret_cast<float*>(smem + c.P.cs);
  float* rowa = reinterpret_cast<float*>(smem + c.P.rowa);
  float* rowb = reinterpret_cast<float*>(smem + c.P.rowb);
  float* bins = reinterpret_cast<float*>(smem + c.P.bins);
  uint8_t* qi = smem + c.P.qi; uint8_t* ki = smem + c.P.ki;
  uint8_t* doi = smem + c.P.doi; uint8_t* vk = smem + c.P.vk; uint8_t* dmn = smem + c.P.dmn;
  uint8_t* own_img = TRANSPOSED ? ki : qi;    // features of the own tile: A operand of the score product
  uint8_t* walk_img = TRANSPOSED ? qi : ki;   // features of the walked tile: B operand, and MN-major B of the dphi product

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_f, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_acc, 1);
    mbar_init_fence();
  }
  kt_load_w(c, wt, p.g.Mp);
  if (!TRANSPOSED)
    for (int i = tid; i < 16 * 256; i += KTHREADS) bins[i] = 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  c.tm = tmem_base_s;
  c.lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tm = c.tm, lane_off = c.lane_off;
  uint32_t ph_s = 0, ph_acc = 0;
  const uint32_t idesc_s = make_idesc(FMT_BF16, 128, 128, false, false);
  const uint32_t idesc_dphi = make_idesc(FMT_BF16, 128, KMP, false, true);
  const uint32_t idesc_dv_a = make_idesc(FMT_BF16, 128, 48, false, true);
  const uint32_t idesc_dv_b = make_idesc(FMT_BF16, 128, 32, false, true);
  const int ks_feat = (p.M + 15) >> 4;
  const int r_pair = packed ? row / N : 0;
  const int r_n = packed ? row - r_pair * N : n_own0 + row;
  const bool r_ok = row < rows_own;

  // per-query statistics of a query tile (rows tq0 ..): r = 1/(den + eps), dden = -(dO . O) r; dO images
  auto query_side = [&](int tq0, int rows_q) {
    if (quarter == 1) {
      float g[DH], o[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) { g[a] = 0.f; o[a] = 0.f; }
      float r = 0.f, dden = 0.f;
      if (row < rows_q) {
        const int t = tq0 + row, b = t / N, n = t - b * N;
        load_row<T, DH>(dob + (size_t)t * out_stride, g);
        load_row<T, DH>(ob + (size_t)t * out_stride, o);
        r = 1.0f / (p.den[((size_t)b * H + h) * N + n] + kEps);
        float dot = 0.f;
#pragma unroll
        for (int a = 0; a < DH; ++a) dot = fmaf(g[a], o[a], dot);
        dden = -dot * r;
      }
      rowa[row] = r;
      rowb[row] = dden;
      kt_store_row_kmajor(doi, row, g);
      if (TRANSPOSED) {
#pragma unroll
        for (int a = 0; a < DH; ++a) g[a] *= r;
        kt_store_row_mnmajor(dmn, row, g, 0.f);
      }
    }
  };
  auto value_side = [&](int tk0, int rows_k) {  // v rows of a key tile -> K-major image
    if (quarter == 2) {
      float v[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) v[a] = 0.f;
      if (row < rows_k) load_row<T, DH>(vb + (size_t)(tk0 + row) * tok_stride, v);
      kt_store_row_kmajor(vk, row, v);
    }
  };

  // ---- own tile
  kt_features<T>(c, TRANSPOSED ? kb : qb, tok_stride, t_own, rows_own, COL_P, own_img, p.M, p.kind, p.inv_sqrt_m);
  if (TRANSPOSED) value_side(t_own, rows_own); else query_side(t_own, rows_own);
  float* dpart = TRANSPOSED ? nullptr : p.dbias_part + ((size_t)h * gridDim.x + blockIdx.x) * (2 * N - 1);

  for (int wi = 0; wi < nwalk; ++wi) {
    const int w0 = wi * KT;                         // position of walked row 0 inside the pair (long sequences)
    const int tw0 = packed ? wbase : wbase + w0;
    const int rows_w = packed ? rows_own : min(KT, N - w0);
    if (wi > 0) {  // the previous accumulation products have read the walked images and the X / Y columns
      mbar_wait(&bar_acc, ph_acc);
      ph_acc ^= 1;
      fence_after_sync();
    }
    if (TRANSPOSED) query_side(tw0, rows_w); else value_side(tw0, rows_w);
    // key - query offset of the tile pair: own = query -> j0 - i0 = w0 - n_own0; own = key -> n_own0 - w0
    kt_load_cs(c, cs, cexp, packed, N, TRANSPOSED ? n_own0 : w0, TRANSPOSED ? w0 : n_own0);
    kt_features<T>(c, TRANSPOSED ? qb : kb, tok_stride, tw0, rows_w, COL_P, walk_img, p.M, p.kind, p.inv_sqrt_m);
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {  // X = phi_own phi_walk^T ; Y = (dO | v)_own (v | dO)_walk^T
      fence_after_sync();
      bool acc = false;
      for (int term = 0; term < 3; ++term) {
        const uint32_t a = smem_u32(own_img + (term == 2 ? KI_BYTES : 0));
        const uint32_t b = smem_u32(walk_img + (term == 1 ? KI_BYTES : 0));
        for (int s = 0; s < ks_feat; ++s) {
          mma_f16(tm + COL_X, make_desc(a + s * 256, 128, KI_SBO), make_desc(b + s * 256, 128, KI_SBO), idesc_s, acc);
          acc = true;
        }
      }
      const uint8_t* ya = TRANSPOSED ? vk : doi;
      const uint8_t* yb = TRANSPOSED ? doi : vk;
      for (int term = 0; term < 3; ++term)
        mma_f16(tm + COL_Y, make_desc(smem_u32(ya + (term == 2 ? XD_BYTES : 0)), 128, XD_SBO),
                make_desc(smem_u32(yb + (term == 1 ? XD_BYTES : 0)), 128, XD_SBO), idesc_s, term > 0);
      commit(&bar_s);
    }
    mbar_wait(&bar_s, ph_s);
    ph_s ^= 1;
    fence_after_sync();
    {  // this thread's 32 columns: a = s c ; dA = u r_i + dden_i ; g = dA c
      const int c0 = 32 * quarter;
      const uint32_t xbase = tm + lane_off + COL_X + c0, ybase = tm + lane_off + COL_Y + c0;
      uint32_t sv[32], uv[32];
      tmem_ld32_nowait(xbase, sv);
      tmem_ld32_nowait(ybase, uv);
      tmem_wait_ld();
      uint32_t ghw[16], glw[16], ahw[16], alw[16];
      int cp = 0, cn = 0;
      if (packed) { cp = c0 / N; cn = c0 - cp * N; }
      const float ra_own = TRANSPOSED ? 0.f : rowa[row], rb_own = TRANSPOSED ? 0.f : rowb[row];
      float* my_bins = bins + warp * 256;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float av[2], gv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = c0 + i + e;
          bool ok = r_ok && col < rows_w;
          int idx;
          if (packed) {
            ok = ok && cp == r_pair;
            idx = TRANSPOSED ? (r_n - cn + N - 1) : (cn - r_n + N - 1);
            if (++cn == N) { cn = 0; ++cp; }
          } else {
            idx = TRANSPOSED ? (row - col + 127) : (col - row + 127);
          }
          const float cc = cs[ok ? idx : 0];
          const float ra = TRANSPOSED ? rowa[col] : ra_own, rb = TRANSPOSED ? rowb[col] : rb_own;
          const float a = ok ? __uint_as_float(sv[i + e]) * cc : 0.f;
          const float dA = ok ? fmaf(__uint_as_float(uv[i + e]), ra, rb) : 0.f;
          av[e] = a;
          gv[e] = dA * cc;
          if (!TRANSPOSED && ok) atomicAdd(my_bins + idx, dA * a);  // lanes of a warp hit distinct diagonals: no conflicts
        }
        split_pack2(gv[0], gv[1], ghw[i >> 1], glw[i >> 1]);
        if (TRANSPOSED) split_pack2(av[0], av[1], ahw[i >> 1], alw[i >> 1]);
      }
      uint32_t w8[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = ghw[8 * q + i];
        tmem_st8(ybase + 8 * q, w8);
#pragma unroll
        for (int i = 0; i < 8; ++i) w8[i] = glw[8 * q + i];
        tmem_st8(ybase + 16 + 8 * q, w8);
        if (TRANSPOSED) {
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = ahw[8 * q + i];
          tmem_st8(xbase + 8 * q, w8);
#pragma unroll
          for (int i = 0; i < 8; ++i) w8[i] = alw[8 * q + i];
          tmem_st8(xbase + 16 + 8 * q, w8);
        }
      }
      tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0 && elect_one()) {  // dphi_own (+)= G phi_walk ; transposed: dv (+)= A~^T (r dO)
      fence_after_sync();
      const int ksteps = (rows_w + 15) >> 4;
      for (int t = 0; t < ksteps; ++t) {
        const uint32_t off = 32 * (t >> 1) + 8 * (t & 1);
        const uint32_t gh = tm + COL_Y + off, gl = gh + 16;
        const uint64_t bh = make_desc(smem_u32(walk_img) + (uint32_t)t * 2 * KI_SBO, KI_SBO, 128);
        const uint64_t bl = make_desc(smem_u32(walk_img + KI_BYTES) + (uint32_t)t * 2 * KI_SBO, KI_SBO, 128);
        const bool acc = wi > 0 || t > 0;
        mma_f16_ts(tm + COL_DPHI, gh, bh, idesc_dphi, acc);
        mma_f16_ts(tm + COL_DPHI, gh, bl, idesc_dphi, true);
        mma_f16_ts(tm + COL_DPHI, gl, bh, idesc_dphi, true);
        if (TRANSPOSED) {
          const uint32_t ah = tm + COL_X + off;
          const uint64_t bd = make_desc(smem_u32(dmn) + (uint32_t)t * 256, 128, VI_CH);
          mma_f16_ts(tm + COL_DV, ah, bd, idesc_dv_a, acc);
          mma_f16_ts(tm + COL_DV, ah + 16, bd, idesc_dv_b, true);
        }
      }
      commit(&bar_acc);
    }
    if (!TRANSPOSED) {  // d bias partial of this tile pair: fixed-order sum over the 16 warps' bins (the barrier before the
      // product issue ordered every warp's adds before these reads)
      const int nb = packed ? 2 * N - 1 : 255;
      if (tid < nb) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          s += bins[w * 256 + tid];
          bins[w * 256 + tid] = 0.f;
        }
        const int d = packed ? tid : (w0 - n_own0) + tid - 127 + N - 1;
        if (d >= 0 && d < 2 * N - 1) dpart[d] += s;  // slice private to this CTA
      }
    }
  }
  mbar_wait(&bar_acc, ph_acc);
  ph_acc ^= 1;
  fence_after_sync();

  // ---- epilogue: dphi_own (TMEM) -> feature-map and L2-norm backward on the fp32 tile helpers (once per CTA)
  float* tl = reinterpret_cast<float*>(smem + c.P.tiles);
  float* xr = tl; float* xs = xr + KT * LDM; float* tmp = xs + KT * LDM;
  float* phi = tmp + KT * LDM;
  float* inv_s = phi + KT * p.g.ldp + 4; float* m_s = inv_s + KT; float* n2_s = m_s + KT; float* pm = n2_s + KT;
  float* red_s = reinterpret_cast<float*>(smem + c.P.qi);  // the operand images are dead: 3 * 128 * 20 floats fit in them
  const RotArgs ra{ERV_ROT_NONE, nullptr, nullptr};
  const T* own_base = TRANSPOSED ? kb : qb;
  if (TRANSPOSED && quarter == 0) {  // dv = hi-part + lo-part columns
    float d0[32], d1[16];
    tmem_ld32(tm + lane_off + COL_DV, d0);
    tmem_ld16(tm + lane_off + COL_DV + 32, d1);
    if (r_ok) {
      T* dvp = static_cast<T*>(p.dqkv) + (size_t)2 * H * DH + (size_t)h * DH + (size_t)(t_own + row) * tok_stride;
#pragma unroll
      for (int cc = 0; cc < DH / 4; ++cc)
        st4(dvp + 4 * cc, make_float4(d0[4 * cc] + d1[4 * cc], d0[4 * cc + 1] + d1[4 * cc + 1], d0[4 * cc + 2] + d1[4 * cc + 2],
                                      d0[4 * cc + 3] + d1[4 * cc + 3]));
    }
  }
  load_tile<T, DH, KT>(xr, own_base, tok_stride, t_own, t_own + rows_own, 0.f);
  __syncthreads();
  prep_tile<DH, KT>(xs, xr, nullptr, inv_s, ra, ERV_PREP_L2NORM, 1.f, t_own, t_own + rows_own);
  __syncthreads();
  feature_tile<DH, KT>(phi, xs, wt, m_s, n2_s, pm, p.g, p.kind, p.inv_sqrt_m);
  {
    float dph[16];
    tmem_ld16(tm + lane_off + COL_DPHI + 16 * quarter, dph);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int f = 16 * quarter + i;
      if (f < p.g.Mp) {
        const float ph_v = phi[row * p.g.ldp + f];
        phi[row * p.g.ldp + f] = (f < p.M && r_ok) ? feature_grad(dph[i], ph_v, p.kind, p.inv_sqrt_m) : 0.f;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  h2_narrow<DH, KT, DH / 4 + 1>(red_s, phi, p.g.ldp, wt, LDM, p.M);
  if (p.kind == ERV_FEAT_FAVOR) {
    for (int i = tid; i < KT * DH; i += KTHREADS) {
      const int t = i / DH, d = i % DH;
      red_s[t * LDM + d] -= xs[t * LDM + d] * red_s[t * LDM + DH];
    }
    __syncthreads();
  }
  T* dst = static_cast<T*>(p.dqkv) + (size_t)(TRANSPOSED ? 1 : 0) * H * DH + (size_t)h * DH;
  prep_tile_bwd<T, DH, KT>(red_s, xs, xr, nullptr, inv_s, tmp, ra, ERV_PREP_L2NORM, 1.f, dst, tok_stride, nullptr, t_own,
                           t_own + rows_own);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// d bias[h][d] = sum over the CTAs of a head of their private partials, fixed order
__global__ void __launch_bounds__(1024) kt_dbias_reduce_kernel(const float* __restrict__ part, float* __restrict__ dbias, int H,
                                                               int nx, int W) {
  __shared__ float red[32][33];
  const size_t i = (size_t)blockIdx.x * 32 + threadIdx.x;
  const bool live = i < (size_t)H * W;
  const int h = live ? (int)(i / W) : 0, d = live ? (int)(i % W) : 0;
  float acc = 0.f;
  if (live)
    for (int u = threadIdx.y; u < nx; u += 32) acc += part[((size_t)h * nx + u) * W + d];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    float sum = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) sum += red[y][threadIdx.x];
    dbias[i] = sum;
  }
}

bool ktile_tc_eligible(int N, int DH, int M) {
  static const bool disabled = getenv("ERV_DISABLE_KTILE_TC") != nullptr;
  return !disabled && DH == 16 && M <= KMP && N >= 2;
}

static void kt_fill(KtArgs& a, int B, int N, int H, int M, int kind) {
  a.B = B; a.N = N; a.H = H; a.M = M; a.kind = kind;
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.g = make_geom(M);
  a.g.TS = KTHREADS / a.g.FT;  // the tile helpers of erv_feat.cuh walk tokens in TS slices: all 512 threads take part
  a.g.nthreads = KTHREADS;
  a.ppt = N <= KT ? KT / N : 0;
  a.nqt = (N + KT - 1) / KT;
}

int ktile_tc_grid_x(int B, int N) { return tile_tc_grid_x(B, N); }

int ktile_tc_forward(const void* qkv, void* out, float* den, const float* wt, const float* cexp, int B, int N, int H, int DH,
                     int M, int kind, int dtype, cudaStream_t st) {
  (void)DH;
  KtArgs a{};
  a.qkv = qkv; a.out = out; a.den = den; a.wt = wt; a.cexp = cexp;
  kt_fill(a, B, N, H, M, kind);
  const size_t smem = kt_plan(a.g.ldp, false, false).total;
  dim3 grid(ktile_tc_grid_x(B, N), H);
  if (dtype == ERV_F32) {
    ERV_CUDA(allow_smem(ktile_fwd_kernel<float>, smem));
    ktile_fwd_kernel<float><<<grid, KTHREADS, smem, st>>>(a);
  } else {
    ERV_CUDA(allow_smem(ktile_fwd_kernel<__nv_bfloat16>, smem));
    ktile_fwd_kernel<__nv_bfloat16><<<grid, KTHREADS, smem, st>>>(a);
  }
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

// dbias_part: at least H * ktile_tc_grid_x(B, N) * (2N-1) floats
int ktile_tc_backward(const void* qkv, const void* out, const float* den, const void* dout, void* dqkv, float* dbias,
                      float* dbias_part, const float* wt, const float* cexp, int B, int N, int H, int DH, int M, int kind,
                      int dtype, cudaStream_t st) {
  (void)DH;
  KtArgs a{};
  a.qkv = qkv; a.out = const_cast<void*>(out); a.den = const_cast<float*>(den); a.dout = dout; a.dqkv = dqkv;
  a.wt = wt; a.cexp = cexp; a.dbias_part = dbias_part;
  kt_fill(a, B, N, H, M, kind);
  const int nx = ktile_tc_grid_x(B, N);
  const int W = 2 * N - 1;
  ERV_CUDA(cudaMemsetAsync(dbias_part, 0, (size_t)H * nx * W * sizeof(float), st));
  dim3 grid(nx, H);
  const size_t smem_q = kt_plan(a.g.ldp, true, false).total, smem_k = kt_plan(a.g.ldp, true, true).total;
#define KT_BWD(TT)                                                                  \
  do {                                                                              \
    ERV_CUDA(allow_smem(ktile_bwd_kernel<TT, false>, smem_q));                      \
    ktile_bwd_kernel<TT, false><<<grid, KTHREADS, smem_q, st>>>(a);                 \
    ERV_LAUNCH_CHECK();                                                             \
    ERV_CUDA(allow_smem(ktile_bwd_kernel<TT, true>, smem_k));                       \
    ktile_bwd_kernel<TT, true><<<grid, KTHREADS, smem_k, st>>>(a);                  \
    ERV_LAUNCH_CHECK();                                                             \
  } while (0)
  if (dtype == ERV_F32) KT_BWD(float); else KT_BWD(__nv_bfloat16);
#undef KT_BWD
  const size_t nb = (size_t)H * W;
  kt_dbias_reduce_kernel<<<(unsigned)((nb + 31) / 32), dim3(32, 32), 0, st>>>(dbias_part, dbias, H, nx, W);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv
