// Softmax attention on tcgen05 / TMEM (head_dim 16): the flash-style baseline of softmax.py:86-115 with the RoPE /
// Circulant-STRING rotation applied to q and k in the tile prologue.
//
// Against the CUDA-core tile kernels of erv_tileattn.cu:
//   * q / k rows are read from the packed qkv buffer, rotated in registers and written straight into the bf16 hi/lo operand
//     images; no rot_pack kernel, no fp32 row workspace in HBM (forward and backward), and the backward rotates dq / dk back
//     in its epilogue (Circulant: d g accumulated into CTA-private slots).
//   * S = q k^T, dP = dO v^T, O = P v, dq = dS k, dk = dS^T q, dv = P^T dO are tcgen05 products with bf16 hi/lo splits of
//     both operands (three terms, ~2^-17 of |q||k|, i.e. ~1e-5 on a logit); P and dS are written back over the S / dP
//     columns as bf16 hi/lo and read from TENSOR memory as the A operand of the next product.
//   * softmax statistics: one key tile (N <= 128) -> max / sum from the registers that hold the scores; several key tiles ->
//     a first sweep computes log-sum-exp (thread-local online max / sum, one exchange at the end), a second sweep
//     accumulates O = exp(s - lse) v without any rescaling of the TMEM accumulator.
//   * tiles as in erv_ktile_tc.cu: 128 tokens of the flattened (batch, token) axis; N <= 128 packs floor(128/N) pairs per
//     tile and masks the cross-pair blocks.  mask / attention dropout / return_attention are applied in the tile.
#include "erv_tile_tc.cuh"

namespace erv {

struct StArgs {
  const void* qkv;
  void* out;              // fwd: output; bwd: saved output
  const void* dout;
  void* dqkv;
  float* lse;             // [B*H][N]
  float* attn_out;        // optional [B*H][N][N]
  const uint8_t* mask;    // optional [B][N][N]
  const float* ta;
  const float* tb;
  float* dg_part;         // circulant bwd: [H][slots][N][DH]
  int rot, slots;
  int B, N, H, ppt, nqt, nx;
  float scale, dropout_p;
  uint64_t seed;
  const unsigned long long* seed_dev;
};

// the same counter-based keep mask as erv_tileattn.cu (forward and backward, either implementation, agree)
__device__ __forceinline__ uint64_t st_seed(const StArgs& p) { return p.seed ^ (p.seed_dev ? __ldg(p.seed_dev) : 0ull); }
__device__ __forceinline__ float st_uniform(uint64_t seed, uint32_t pair, uint32_t i, uint32_t j) {
  uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * ((uint64_t)pair + 1));
  x ^= ((uint64_t)i << 32) | (uint64_t)j;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return (float)(x >> 40) * (1.0f / 16777216.0f);
}

struct StGeo {  // a CTA's tile on the flattened token axis of head h
  int t0, rows, n0, nwalk, wbase, b0;
};
__device__ __forceinline__ StGeo st_geo(const StArgs& p, int tile) {
  StGeo g;
  if (p.ppt > 0) {
    g.b0 = tile * p.ppt;
    g.t0 = g.b0 * p.N; g.rows = min(p.ppt, p.B - g.b0) * p.N; g.n0 = 0; g.nwalk = 1; g.wbase = g.t0;
  } else {
    g.b0 = tile / p.nqt;
    g.n0 = (tile % p.nqt) * KT;
    g.t0 = g.b0 * p.N + g.n0; g.rows = min(KT, p.N - g.n0); g.nwalk = p.nqt; g.wbase = g.b0 * p.N;
  }
  return g;
}

// row `row` of a tile (token t0 + row, position n inside its pair): load, rotate, write the K-major image
template <typename T>
__device__ __forceinline__ void st_rotated_row(uint8_t* img, const T* __restrict__ base, size_t tok_stride, int t0, int rows,
                                               int row, int n, const StArgs& p, int h) {
  float x[16];
#pragma unroll
  for (int a = 0; a < 16; ++a) x[a] = 0.f;
  if (row < rows) {
    load_row<T, 16>(base + (size_t)(t0 + row) * tok_stride, x);
    prologue_row<16>(x, p.rot, p.ta, p.tb, h, n, p.N, 1.0f);
  }
  kt_store_row_kmajor(img, row, x);
}

constexpr float kStLog2e = 1.4426950408889634f;

// ---- forward ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(KTHREADS, 1) stile_fwd_kernel(const StArgs p) {
  constexpr int DH = 16;
  constexpr uint32_t COL_S = 0, COL_O = 256;  // S of up to two key tiles at [0,128) and [128,256), O behind them
  __shared__ __align__(128) uint8_t qi[2 * XD_BYTES];
  __shared__ __align__(128) uint8_t ki[2 * XD_BYTES];
  __shared__ __align__(128) uint8_t vi[6 * VI_CH];
  __shared__ __align__(128) uint8_t vi2[6 * VI_CH];  // value image of the second key tile (single-pass path)
  __shared__ float ml_s[4][128][2];
  __shared__ __align__(8) uint64_t bar_s, bar_pv;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, quarter = tid >> 7;
  const int N = p.N, H = p.H, h = blockIdx.y;
  const bool packed = p.ppt > 0;
  const size_t tok_stride = (size_t)3 * H * DH;
  const T* qb = static_cast<const T*>(p.qkv) + (size_t)h * DH;
  const T* kb = qb + (size_t)H * DH;
  const T* vb = qb + (size_t)2 * H * DH;
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint64_t seed = p.dropout_p > 0.f ? st_seed(p) : 0ull;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_pv, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_s = 0, ph_pv = 0;
  const uint32_t idesc_s = make_idesc(FMT_BF16, 128, 128, false, false);
  const uint32_t idesc_pv_a = make_idesc(FMT_BF16, 128, 48, false, true);
  const uint32_t idesc_pv_b = make_idesc(FMT_BF16, 128, 32, false, true);
  const float sl2 = p.scale * kStLog2e;  // logits in log2 units

  for (int tile = blockIdx.x; tile < p.nx; tile += gridDim.x) {
    const StGeo g = st_geo(p, tile);
    const int r_pair = packed ? row / N : 0;
    const int r_n = packed ? row - r_pair * N : g.n0 + row;
    const int r_b = g.b0 + r_pair;
    const bool r_ok = row < g.rows;
    const uint32_t pair = (uint32_t)(r_b * H + h);
    if (quarter == 0) st_rotated_row<T>(qi, qb, tok_stride, g.t0, g.rows, row, r_n, p, h);

    // scores of this thread's 32 columns in log2 units (masked entries -inf): S (TMEM) -> sc[]
    auto load_scores = [&](float (&sc)[32], int w0, int rows_w, uint32_t scol = 0) {
      uint32_t sv[32];
      tmem_ld32_nowait(tm + lane_off + COL_S + scol + 32 * quarter, sv);
      tmem_wait_ld();
      const int c0 = 32 * quarter;
      int cp = 0, cn = 0;
      if (packed) { cp = c0 / N; cn = c0 - cp * N; }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int col = c0 + i;
        bool ok = r_ok && col < rows_w;
        int j;
        if (packed) {
          ok = ok && cp == r_pair;
          j = cn;
          if (++cn == N) { cn = 0; ++cp; }
        } else {
          j = w0 + col;
        }
        if (ok && p.mask != nullptr && p.mask[((size_t)r_b * N + r_n) * N + j] == 0) ok = false;
        sc[i] = ok ? __uint_as_float(sv[i]) * sl2 : -INFINITY;
      }
    };
    auto issue_scores = [&](uint32_t scol = 0) {
      if (warp == 0 && elect_one()) {
        fence_after_sync();
        for (int term = 0; term < 3; ++term)
          mma_f16(tm + COL_S + scol, make_desc(smem_u32(qi + (term == 2 ? XD_BYTES : 0)), 128, XD_SBO),
                  make_desc(smem_u32(ki + (term == 1 ? XD_BYTES : 0)), 128, XD_SBO), idesc_s, term > 0);
        commit(&bar_s);
      }
      mbar_wait(&bar_s, ph_s);
      ph_s ^= 1;
      fence_after_sync();
    };
    auto key_rows = [&](int tw0, int rows_w, int w0) {
      if (quarter == 1) {
        const int wp = packed ? row / N : 0;
        st_rotated_row<T>(ki, kb, tok_stride, tw0, rows_w, row, packed ? row - wp * N : w0 + row, p, h);
      }
    };
    auto value_rows = [&](int tw0, int rows_w, uint8_t* img = nullptr) {
      if (img == nullptr) img = vi;
      if (quarter == 2) {
        float v[DH];
#pragma unroll
        for (int a = 0; a < DH; ++a) v[a] = 0.f;
        if (row < rows_w) load_row<T, DH>(vb + (size_t)(tw0 + row) * tok_stride, v);
        kt_store_row_mnmajor(img, row, v, 0.f);
      }
    };
    // combine the four threads of a row: log-sum-exp in log2 units (-inf for a fully masked row)
    auto row_lse = [&](float m_t, float l_t) -> float {
      ml_s[quarter][row][0] = m_t;
      ml_s[quarter][row][1] = l_t;
      __syncthreads();
      float M = fmaxf(fmaxf(ml_s[0][row][0], ml_s[1][row][0]), fmaxf(ml_s[2][row][0], ml_s[3][row][0]));
      float L = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float mq = ml_s[q][row][0];
        if (mq > -INFINITY) L += ml_s[q][row][1] * ex2_approx(mq - M);
      }
      return (M > -INFINITY) ? M + log2f(L) : -INFINITY;
    };
    // P = exp(s - lse) (+ dropout, + dump) over this thread's columns, written back as bf16 hi / lo; then O (+)= P v
    auto weights_and_pv = [&](float (&sc)[32], float lse2, int w0, int rows_w, bool first) {
      const int c0 = 32 * quarter;
      uint32_t hw[16], lw[16];
      int cp = 0, cn = 0;
      if (packed) { cp = c0 / N; cn = c0 - cp * N; }
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float a[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = c0 + i + e;
          bool inside = r_ok && col < rows_w;  // a (query, key) position of this row's pair, masked or not
          int j;
          if (packed) {
            inside = inside && cp == r_pair;
            j = cn;
            if (++cn == N) { cn = 0; ++cp; }
          } else {
            j = w0 + col;
          }
          float pv = (sc[i + e] > -INFINITY) ? ex2_approx(sc[i + e] - lse2) : 0.f;
          if (p.dropout_p > 0.f && pv != 0.f)
            pv = (st_uniform(seed, pair, (uint32_t)r_n, (uint32_t)j) >= p.dropout_p) ? pv * keep_scale : 0.f;
          if (p.attn_out != nullptr && inside) p.attn_out[((size_t)pair * N + r_n) * N + j] = pv;
          a[e] = pv;
        }
        split_pack2(a[0], a[1], hw[i >> 1], lw[i >> 1]);
      }
      const uint32_t cbase = tm + lane_off + COL_S + c0;
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
      fence_before_sync();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        fence_after_sync();
        const int ksteps = (rows_w + 15) >> 4;
        for (int t = 0; t < ksteps; ++t) {
          const uint32_t ca = tm + COL_S + 32 * (t >> 1) + 8 * (t & 1);
          const uint64_t bd = make_desc(smem_u32(vi) + (uint32_t)t * 256, 128, VI_CH);
          mma_f16_ts(tm + COL_O, ca, bd, idesc_pv_a, !first || t > 0);
          mma_f16_ts(tm + COL_O, ca + 16, bd, idesc_pv_b, true);
        }
        commit(&bar_pv);
      }
    };

    // unnormalised weights e = exp(s - M) of one key tile (dropout applied), written back over its S columns as bf16 hi / lo;
    // returns the thread's share of the row sum (before dropout)
    auto weights_unnormalised = [&](float (&sc)[32], float M, int w0, uint32_t scol) -> float {
      const int c0 = 32 * quarter;
      uint32_t hw[16], lw[16];
      int cp = 0, cn = 0;
      if (packed) { cp = c0 / N; cn = c0 - cp * N; }
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float a[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          int j;
          if (packed) {
            j = cn;
            if (++cn == N) { cn = 0; ++cp; }
          } else {
            j = w0 + c0 + i + e;
          }
          float pv = (sc[i + e] > -INFINITY) ? ex2_approx(sc[i + e] - M) : 0.f;
          l += pv;
          if (p.dropout_p > 0.f && pv != 0.f)
            pv = (st_uniform(seed, pair, (uint32_t)r_n, (uint32_t)j) >= p.dropout_p) ? pv * keep_scale : 0.f;
          a[e] = pv;
        }
        split_pack2(a[0], a[1], hw[i >> 1], lw[i >> 1]);
      }
      const uint32_t cbase = tm + lane_off + COL_S + scol + c0;
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
      return l;
    };

    float lse2;
    float o_scale = 1.f;
    if (g.nwalk <= 2 && p.attn_out == nullptr) {
      // Single pass for up to two key tiles (N <= 256): both score tiles stay in tensor memory, the row maximum is exchanged
      // first, then ONE exponential per score gives the unnormalised weight e = exp(s - M) and the row sum; O = (e v) / L is
      // scaled in the epilogue.  (The two-sweep path below evaluates every exponential twice and, for two key tiles, every
      // score product twice.)  Not used with return_attention, which needs normalised weights.
      float m_t = -INFINITY;
      for (int wi = 0; wi < g.nwalk; ++wi) {
        const int w0 = wi * KT, rows_w = g.nwalk == 1 ? g.rows : min(KT, N - w0);
        key_rows(g.wbase + w0, rows_w, w0);
        value_rows(g.wbase + w0, rows_w, wi ? vi2 : vi);
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
        issue_scores(128u * wi);  // returns when the product has completed: the key image may be overwritten
        float sc[32];
        load_scores(sc, w0, rows_w, 128u * wi);
#pragma unroll
        for (int i = 0; i < 32; ++i) m_t = fmaxf(m_t, sc[i]);
      }
      ml_s[quarter][row][0] = m_t;
      __syncthreads();
      const float M = fmaxf(fmaxf(ml_s[0][row][0], ml_s[1][row][0]), fmaxf(ml_s[2][row][0], ml_s[3][row][0]));
      float l_t = 0.f;
      for (int wi = 0; wi < g.nwalk; ++wi) {
        const int w0 = wi * KT, rows_w = g.nwalk == 1 ? g.rows : min(KT, N - w0);
        float sc[32];
        load_scores(sc, w0, rows_w, 128u * wi);
        l_t += weights_unnormalised(sc, M, w0, 128u * wi);
      }
      tmem_wait_st();
      ml_s[quarter][row][1] = l_t;
      fence_before_sync();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        fence_after_sync();
        for (int wi = 0; wi < g.nwalk; ++wi) {
          const int rows_w = g.nwalk == 1 ? g.rows : min(KT, N - wi * KT);
          const int ksteps = (rows_w + 15) >> 4;
          for (int t = 0; t < ksteps; ++t) {
            const uint32_t ca = tm + COL_S + 128u * wi + 32 * (t >> 1) + 8 * (t & 1);
            const uint64_t bd = make_desc(smem_u32(wi ? vi2 : vi) + (uint32_t)t * 256, 128, VI_CH);
            mma_f16_ts(tm + COL_O, ca, bd, idesc_pv_a, wi > 0 || t > 0);
            mma_f16_ts(tm + COL_O, ca + 16, bd, idesc_pv_b, true);
          }
        }
        commit(&bar_pv);
      }
      const float L = (ml_s[0][row][1] + ml_s[1][row][1]) + (ml_s[2][row][1] + ml_s[3][row][1]);
      lse2 = (M > -INFINITY && L > 0.f) ? M + log2f(L) : -INFINITY;
      o_scale = L > 0.f ? 1.f / L : 0.f;
    } else if (g.nwalk == 1) {  // one key tile: statistics and weights from the same registers
      key_rows(g.wbase, g.rows, 0);
      value_rows(g.wbase, g.rows);
      fence_smem_to_async();
      fence_before_sync();
      __syncthreads();
      issue_scores();
      float sc[32];
      load_scores(sc, 0, g.rows);
      float m_t = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) m_t = fmaxf(m_t, sc[i]);
      float l_t = 0.f;
      if (m_t > -INFINITY) {
#pragma unroll
        for (int i = 0; i < 32; ++i) l_t += ex2_approx(sc[i] - m_t);
      }
      lse2 = row_lse(m_t, l_t);
      weights_and_pv(sc, lse2, 0, g.rows, true);
    } else {
      float m_t = -INFINITY, l_t = 0.f;
      for (int wi = 0; wi < g.nwalk; ++wi) {  // sweep 1: log-sum-exp
        const int w0 = wi * KT, rows_w = min(KT, N - w0);
        key_rows(g.wbase + w0, rows_w, w0);
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();  // also: every thread has consumed the previous S tile
        issue_scores();
        float sc[32];
        load_scores(sc, w0, rows_w);
        float mx = m_t;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, sc[i]);
        if (mx > -INFINITY) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) s += ex2_approx(sc[i] - mx);
          l_t = l_t * ex2_approx(m_t - mx) + s;
          m_t = mx;
        }
        fence_before_sync();
      }
      lse2 = row_lse(m_t, l_t);
      for (int wi = 0; wi < g.nwalk; ++wi) {  // sweep 2: O = sum exp(s - lse) v
        const int w0 = wi * KT, rows_w = min(KT, N - w0);
        if (wi > 0) {  // the previous P v product has read the V image and the S columns
          mbar_wait(&bar_pv, ph_pv);
          ph_pv ^= 1;
          fence_after_sync();
        }
        key_rows(g.wbase + w0, rows_w, w0);
        value_rows(g.wbase + w0, rows_w);
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
        issue_scores();
        float sc[32];
        load_scores(sc, w0, rows_w);
        weights_and_pv(sc, lse2, w0, rows_w, wi == 0);
      }
    }
    mbar_wait(&bar_pv, ph_pv);
    ph_pv ^= 1;
    fence_after_sync();
    if (quarter == 0) {  // columns [0,16) hi-part, [32,48) lo-part
      float o0[16], o1[16];
      tmem_ld16(tm + lane_off + COL_O, o0);
      tmem_ld16(tm + lane_off + COL_O + 32, o1);
      if (r_ok) {
        T* ob = static_cast<T*>(p.out) + ((size_t)(g.t0 + row) * H + h) * DH;
#pragma unroll
        for (int cc = 0; cc < DH / 4; ++cc)
          st4(ob + 4 * cc, make_float4((o0[4 * cc] + o1[4 * cc]) * o_scale, (o0[4 * cc + 1] + o1[4 * cc + 1]) * o_scale,
                                       (o0[4 * cc + 2] + o1[4 * cc + 2]) * o_scale, (o0[4 * cc + 3] + o1[4 * cc + 3]) * o_scale));
        p.lse[((size_t)r_b * H + h) * N + r_n] = lse2 * 0.6931471805599453f;  // natural log, as the reference's logsumexp
      }
    }
    fence_before_sync();
    __syncthreads();  // images, statistics and TMEM columns are reused by the next tile
  }
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- backward --------------------------------------------------------------------------------------------------
// TRANSPOSED = false: CTA owns a QUERY tile, walks key tiles -> dq.  TRANSPOSED = true: CTA owns a KEY tile, walks query
// tiles (score tiles transposed: rows = keys) -> dk, dv.
template <typename T, bool TRANSPOSED>
__global__ void __launch_bounds__(KTHREADS, 1) stile_bwd_kernel(const StArgs p) {
  constexpr int DH = 16;
  constexpr uint32_t COL_X = 0, COL_Y = 128, COL_DX = 256, COL_DV = 288;
  __shared__ __align__(128) uint8_t qi[2 * XD_BYTES];
  __shared__ __align__(128) uint8_t ki[2 * XD_BYTES];
  __shared__ __align__(128) uint8_t doi[2 * XD_BYTES];
  __shared__ __align__(128) uint8_t vk[2 * XD_BYTES];
  __shared__ __align__(128) uint8_t dmn[TRANSPOSED ? 6 * VI_CH : 128];
  __shared__ float rowa[128], rowb[128];
  __shared__ __align__(8) uint64_t bar_s, bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, quarter = tid >> 7;
  const int N = p.N, H = p.H, h = blockIdx.y;
  const bool packed = p.ppt > 0;
  const size_t tok_stride = (size_t)3 * H * DH, out_stride = (size_t)H * DH;
  const T* qb = static_cast<const T*>(p.qkv) + (size_t)h * DH;
  const T* kb = qb + (size_t)H * DH;
  const T* vb = qb + (size_t)2 * H * DH;
  const T* ob = static_cast<const T*>(p.out) + (size_t)h * DH;
  const T* dob = static_cast<const T*>(p.dout) + (size_t)h * DH;
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  const uint64_t seed = p.dropout_p > 0.f ? st_seed(p) : 0ull;
  uint8_t* own_img = TRANSPOSED ? ki : qi;
  uint8_t* walk_img = TRANSPOSED ? qi : ki;
  float* dg_slot = (p.rot == ERV_ROT_CIRCULANT && p.dg_part) ? p.dg_part + ((size_t)h * p.slots + blockIdx.x) * N * DH : nullptr;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_acc, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_s = 0, ph_acc = 0;
  const uint32_t idesc_s = make_idesc(FMT_BF16, 128, 128, false, false);
  const uint32_t idesc_dx = make_idesc(FMT_BF16, 128, 16, false, true);
  const uint32_t idesc_dv_a = make_idesc(FMT_BF16, 128, 48, false, true);
  const uint32_t idesc_dv_b = make_idesc(FMT_BF16, 128, 32, false, true);
  const float sl2 = p.scale * kStLog2e;

  // statistics and dO images of a query tile: rowa = lse (log2 units), rowb = D = dO . O
  auto query_side = [&](int tq0, int rows_q, int n0) {
    if (quarter == 1) {
      float gq[DH], o[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) { gq[a] = 0.f; o[a] = 0.f; }
      float lse2 = 0.f, dd = 0.f;
      if (row < rows_q) {
        const int t = tq0 + row, b = t / N, n = t - b * N;
        (void)n0;
        load_row<T, DH>(dob + (size_t)t * out_stride, gq);
        load_row<T, DH>(ob + (size_t)t * out_stride, o);
        lse2 = p.lse[((size_t)b * H + h) * N + n] * kStLog2e;
#pragma unroll
        for (int a = 0; a < DH; ++a) dd = fmaf(gq[a], o[a], dd);
      }
      rowa[row] = lse2;
      rowb[row] = dd;
      kt_store_row_kmajor(doi, row, gq);
      if (TRANSPOSED) kt_store_row_mnmajor(dmn, row, gq, 0.f);
    }
  };
  auto value_side = [&](int tk0, int rows_k) {
    if (quarter == 2) {
      float v[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) v[a] = 0.f;
      if (row < rows_k) load_row<T, DH>(vb + (size_t)(tk0 + row) * tok_stride, v);
      kt_store_row_kmajor(vk, row, v);
    }
  };

  for (int tile = blockIdx.x; tile < p.nx; tile += gridDim.x) {
    const StGeo g = st_geo(p, tile);
    const int r_pair = packed ? row / N : 0;
    const int r_n = packed ? row - r_pair * N : g.n0 + row;
    const int r_b = g.b0 + r_pair;
    const bool r_ok = row < g.rows;
    // own rows
    if (quarter == 0) st_rotated_row<T>(own_img, TRANSPOSED ? kb : qb, tok_stride, g.t0, g.rows, row, r_n, p, h);
    if (TRANSPOSED) value_side(g.t0, g.rows); else query_side(g.t0, g.rows, g.n0);

    for (int wi = 0; wi < g.nwalk; ++wi) {
      const int w0 = wi * KT;
      const int tw0 = packed ? g.wbase : g.wbase + w0;
      const int rows_w = packed ? g.rows : min(KT, N - w0);
      if (wi > 0) {  // the previous accumulation products have read the walked images and the X / Y columns
        mbar_wait(&bar_acc, ph_acc);
        ph_acc ^= 1;
        fence_after_sync();
      }
      if (quarter == 3) {
        const int wp = packed ? row / N : 0;
        st_rotated_row<T>(walk_img, TRANSPOSED ? qb : kb, tok_stride, tw0, rows_w, row, packed ? row - wp * N : w0 + row, p, h);
      }
      if (TRANSPOSED) query_side(tw0, rows_w, w0); else value_side(tw0, rows_w);
      fence_smem_to_async();
      fence_before_sync();
      __syncthreads();
      if (warp == 0 && elect_one()) {  // X = own walk^T ; Y = (dO | v)_own (v | dO)_walk^T
        fence_after_sync();
        for (int term = 0; term < 3; ++term)
          mma_f16(tm + COL_X, make_desc(smem_u32(own_img + (term == 2 ? XD_BYTES : 0)), 128, XD_SBO),
                  make_desc(smem_u32(walk_img + (term == 1 ? XD_BYTES : 0)), 128, XD_SBO), idesc_s, term > 0);
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
      {  // P = exp(s - lse_i) ; Pd = P keep ; dS = P (dP keep - D_i) scale
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
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float av[2], gv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c0 + i + e;
            bool ok = r_ok && col < rows_w;
            int cpos;  // position of the column token inside its pair
            if (packed) {
              ok = ok && cp == r_pair;
              cpos = cn;
              if (++cn == N) { cn = 0; ++cp; }
            } else {
              cpos = w0 + col;
            }
            const int qi_n = TRANSPOSED ? cpos : r_n, kj_n = TRANSPOSED ? r_n : cpos;  // query / key positions
            if (ok && p.mask != nullptr && p.mask[((size_t)r_b * N + qi_n) * N + kj_n] == 0) ok = false;
            const float lse2 = TRANSPOSED ? rowa[col] : ra_own, dd = TRANSPOSED ? rowb[col] : rb_own;
            float pr = 0.f, keep = 1.f;
            if (ok) {
              pr = ex2_approx(fmaf(__uint_as_float(sv[i + e]), sl2, -lse2));
              if (p.dropout_p > 0.f)
                keep = (st_uniform(seed, (uint32_t)(r_b * H + h), (uint32_t)qi_n, (uint32_t)kj_n) >= p.dropout_p) ? keep_scale : 0.f;
            }
            av[e] = pr * keep;
            gv[e] = pr * (__uint_as_float(uv[i + e]) * keep - dd) * p.scale;
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
      if (warp == 0 && elect_one()) {  // d(own rows, rotated) (+)= dS walk ; transposed: dv (+)= Pd^T dO
        fence_after_sync();
        const int ksteps = (rows_w + 15) >> 4;
        for (int t = 0; t < ksteps; ++t) {
          const uint32_t off = 32 * (t >> 1) + 8 * (t & 1);
          const uint32_t gh = tm + COL_Y + off, gl = gh + 16;
          const uint64_t bh = make_desc(smem_u32(walk_img) + (uint32_t)t * 2 * XD_SBO, XD_SBO, 128);
          const uint64_t bl = make_desc(smem_u32(walk_img + XD_BYTES) + (uint32_t)t * 2 * XD_SBO, XD_SBO, 128);
          const bool acc = wi > 0 || t > 0;
          mma_f16_ts(tm + COL_DX, gh, bh, idesc_dx, acc);
          mma_f16_ts(tm + COL_DX, gh, bl, idesc_dx, true);
          mma_f16_ts(tm + COL_DX, gl, bh, idesc_dx, true);
          if (TRANSPOSED) {
            const uint32_t ah = tm + COL_X + off;
            const uint64_t bd = make_desc(smem_u32(dmn) + (uint32_t)t * 256, 128, VI_CH);
            mma_f16_ts(tm + COL_DV, ah, bd, idesc_dv_a, acc);
            mma_f16_ts(tm + COL_DV, ah + 16, bd, idesc_dv_b, true);
          }
        }
        commit(&bar_acc);
      }
    }
    mbar_wait(&bar_acc, ph_acc);
    ph_acc ^= 1;
    fence_after_sync();
    // ---- epilogue: rotate the gradient of the rotated rows back, write dq / dk (and dv)
    float dy[DH], dxr[DH];
    if (quarter == 0) {
      tmem_ld16(tm + lane_off + COL_DX, dy);
      if (TRANSPOSED) {
        float d0[16], d1[16];
        tmem_ld16(tm + lane_off + COL_DV, d0);
        tmem_ld16(tm + lane_off + COL_DV + 32, d1);
        if (r_ok) {
          T* dvp = static_cast<T*>(p.dqkv) + (size_t)2 * H * DH + (size_t)h * DH + (size_t)(g.t0 + row) * tok_stride;
#pragma unroll
          for (int cc = 0; cc < DH / 4; ++cc)
            st4(dvp + 4 * cc, make_float4(d0[4 * cc] + d1[4 * cc], d0[4 * cc + 1] + d1[4 * cc + 1], d0[4 * cc + 2] + d1[4 * cc + 2],
                                          d0[4 * cc + 3] + d1[4 * cc + 3]));
        }
      }
    }
    const int which = TRANSPOSED ? 1 : 0;
    const int npass = (dg_slot != nullptr && packed) ? p.ppt : 1;  // Circulant: pairs of a packed tile share table rows
    for (int pp = 0; pp < npass; ++pp) {
      if (quarter == 0 && r_ok && (npass == 1 || r_pair == pp)) {
        const T* xraw = static_cast<const T*>(p.qkv) + (size_t)which * H * DH + (size_t)h * DH + (size_t)(g.t0 + row) * tok_stride;
        prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, r_n, N, dg_slot, xraw);
        T* dst = static_cast<T*>(p.dqkv) + (size_t)which * H * DH + (size_t)h * DH + (size_t)(g.t0 + row) * tok_stride;
#pragma unroll
        for (int cc = 0; cc < DH / 4; ++cc) st4(dst + 4 * cc, make_float4(dxr[4 * cc], dxr[4 * cc + 1], dxr[4 * cc + 2], dxr[4 * cc + 3]));
      }
      if (npass > 1) __syncthreads();
    }
    fence_before_sync();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tm, 512);
}

bool stile_tc_eligible(int N, int DH) {
  static const bool disabled = getenv("ERV_DISABLE_STILE_TC") != nullptr;
  return !disabled && DH == 16 && N >= 2;
}

static void st_fill(StArgs& a, int B, int N, int H, int rot, const float* ta, const float* tb, const uint8_t* mask,
                    float dropout_p, uint64_t seed, const long long* seed_dev) {
  a.B = B; a.N = N; a.H = H; a.rot = rot; a.ta = ta; a.tb = tb; a.mask = mask;
  a.scale = 0.25f;  // Dh^-1/2, Dh = 16
  a.dropout_p = dropout_p; a.seed = seed; a.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  a.ppt = N <= KT ? KT / N : 0;
  a.nqt = (N + KT - 1) / KT;
  a.nx = tile_tc_grid_x(B, N);
}

int stile_tc_forward(const void* qkv, void* out, float* lse, float* attn_out, const uint8_t* mask, int B, int N, int H, int rot,
                     const float* ta, const float* tb, float dropout_p, uint64_t seed, const long long* seed_dev, int dtype,
                     cudaStream_t st) {
  StArgs a{};
  a.qkv = qkv; a.out = out; a.lse = lse; a.attn_out = attn_out;
  st_fill(a, B, N, H, rot, ta, tb, mask, dropout_p, seed, seed_dev);
  dim3 grid(a.nx, H);
  if (dtype == ERV_F32) stile_fwd_kernel<float><<<grid, KTHREADS, 0, st>>>(a);
  else stile_fwd_kernel<__nv_bfloat16><<<grid, KTHREADS, 0, st>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

// dg_part (Circulant): [H][slots][N][16], zeroed by the caller; CTA x uses slot x, so the grid is capped at `slots`
int stile_tc_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, const uint8_t* mask,
                      int B, int N, int H, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                      float dropout_p, uint64_t seed, const long long* seed_dev, int dtype, cudaStream_t st) {
  StArgs a{};
  a.qkv = qkv; a.out = const_cast<void*>(out); a.lse = const_cast<float*>(lse); a.dout = dout; a.dqkv = dqkv;
  a.dg_part = dg_part; a.slots = slots;
  st_fill(a, B, N, H, rot, ta, tb, mask, dropout_p, seed, seed_dev);
  int gx = a.nx;
  if (rot == ERV_ROT_CIRCULANT && dg_part != nullptr && gx > slots) gx = slots;
  dim3 grid(gx, H);
#define ST_BWD(TT)                                                   \
  do {                                                               \
    stile_bwd_kernel<TT, false><<<grid, KTHREADS, 0, st>>>(a);       \
    ERV_LAUNCH_CHECK();                                              \
    stile_bwd_kernel<TT, true><<<grid, KTHREADS, 0, st>>>(a);        \
    ERV_LAUNCH_CHECK();                                              \
  } while (0)
  if (dtype == ERV_F32) ST_BWD(float); else ST_BWD(__nv_bfloat16);
#undef ST_BWD
  return ERV_OK;
}

}  // namespace erv
