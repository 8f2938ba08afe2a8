// Quadratic tile kernels: softmax attention (softmax.py:86-115) and KERPLE's Toeplitz-masked linear
// attention (favor_plus.py:221-245 + kerple.py:99-344, evaluated as A = (phi_q phi_k^T) * exp(b[j-i]);
// SURVEY.md section 0 item 5 shows it is identical to the reference's FFT route).
//
// Both are "scores -> weights -> weighted sum of V" over 64x64 tiles with fp32 row operands
//   softmax: rows = rotated q / k            (KD = Dh),  weights = online softmax (+mask, +dropout)
//   kerple : rows = phi(q) / phi(k)          (KD = M),   weights = scores * c[j-i], plain row sums
// One CTA owns a 64-row query tile (forward, dQ) or a 64-row key tile (dK/dV) of one (batch, head);
// nothing of size N x N is ever written (except the optional return_attention dump).
#include "erv_feat.cuh"

namespace erv {

// implemented in erv_linattn.cu
int launch_feature_map(const void* x, void* dx, size_t sb, size_t sh, size_t sn, float* phi, const float* dphi,
                       int ldphi, const float* wt, int B, int N, int H, int DH, int M, int kind, int prep,
                       float prescale, int dtype, bool bwd, cudaStream_t st);
int la_grid(int B, int H);
int la_slots(int B, int H);
// FFT route of the KERPLE forward for long sequences (erv_kerple_fft.cu)
bool kerple_fft_eligible(int B, int N, int H, int DH, int M);
size_t kerple_fft_ws_bytes(int B, int N, int H, int DH, int M);
int kerple_fft_forward(const void* qkv, void* out, float* den, const float* phi_q, const float* phi_k, int ld,
                       const float* cexp, void* ws, int B, int N, int H, int DH, int M, int dtype, cudaStream_t st);
size_t wt_bytes(int H, int DH, int M);
int prep_wt_public(const float* omega, float* wt, int H, int DH, int M, int kind, cudaStream_t st);
// tcgen05 softmax tiles (erv_stile_tc.cu)
bool stile_tc_eligible(int N, int DH);
int stile_tc_forward(const void* qkv, void* out, float* lse, float* attn_out, const uint8_t* mask, int B, int N, int H, int rot,
                     const float* ta, const float* tb, float dropout_p, uint64_t seed, const long long* seed_dev, int dtype,
                     cudaStream_t st);
int stile_tc_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, const uint8_t* mask,
                      int B, int N, int H, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                      float dropout_p, uint64_t seed, const long long* seed_dev, int dtype, cudaStream_t st);
// tcgen05 KERPLE tiles (erv_ktile_tc.cu)
bool ktile_tc_eligible(int N, int DH, int M);
int ktile_tc_forward(const void* qkv, void* out, float* den, const float* wt, const float* cexp, int B, int N, int H, int DH,
                     int M, int kind, int dtype, cudaStream_t st);
int ktile_tc_backward(const void* qkv, const void* out, const float* den, const void* dout, void* dqkv, float* dbias,
                      float* dbias_part, const float* wt, const float* cexp, int B, int N, int H, int DH, int M, int kind,
                      int dtype, cudaStream_t st);

constexpr int TQ = 64;        // tile rows / cols
constexpr int LDP = TQ + 4;   // stride of the 64x64 weight tiles (16-byte aligned rows)

enum { MODE_SOFTMAX = 0, MODE_KERPLE = 1 };

struct TileArgs {
  const float* qrows;   // [B*H][N][ldr] fp32
  const float* krows;   // [B*H][N][ldr]
  float* dqrows;        // bwd outputs, same layout
  float* dkrows;
  const void* qkv;      // packed activations (V is read from here)
  void* dqkv;           // dV is written here
  void* out;            // fwd: output; bwd: saved output
  const void* dout;
  float* stat;          // [B*H][N]: softmax log-sum-exp / kerple denominator
  float* attn_out;      // optional [B*H][N][N]
  const uint8_t* mask;  // optional [B][N][N]
  const float* cexp;    // kerple: [H][2N-1] exp(bias)
  float* dbias_part;    // kerple bwd: [B*H*nqt][2N-1]
  int B, N, H, KD, ldr; // KD = inner dim padded to 8, ldr = row stride of q/k rows
  float scale, dropout_p;
  uint64_t seed;
  const unsigned long long* seed_dev;  // optional device-resident seed (xor-ed in): new masks on every CUDA-graph replay
};

__device__ __forceinline__ uint64_t eff_seed(const TileArgs& p) { return p.seed ^ (p.seed_dev ? __ldg(p.seed_dev) : 0ull); }
// counter-based keep mask: uniform in [0,1) from (seed, pair, i, j)
__device__ __forceinline__ float rng_uniform(uint64_t seed, uint32_t pair, uint32_t i, uint32_t j) {
  uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * ((uint64_t)pair + 1));
  x ^= ((uint64_t)i << 32) | (uint64_t)j;
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return (float)(x >> 40) * (1.0f / 16777216.0f);
}

// rows n0.. of a [N][ld] fp32 matrix -> tile [64][KD+4], zero padded
__device__ __forceinline__ void load_rows(float* __restrict__ dst, const float* __restrict__ src, int ld, int KD,
                                          int n0, int N) {
  const int LDQ = KD + 4, V = KD / 4;
  for (int i = threadIdx.x; i < TQ * V; i += blockDim.x) {
    int t = i / V, v = i % V, n = n0 + t;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) x = ld4(src + (size_t)n * ld + 4 * v);
    st4(dst + t * LDQ + 4 * v, x);
  }
}

// Number of micro-tile rows (or columns) base, base+16, base+32, base+48 that fall inside the sequence.
__device__ __forceinline__ int valid_count(int base, int N) {
  const int rem = N - base;
  return rem <= 0 ? 0 : min(4, (rem + 15) >> 4);
}

// 4x4 micro-tile of scores: rows ty+16r, cols tx+16c.  nr / nc = valid_count of the thread's rows / columns: sequences of
// 64k+1 tokens end in a tile with one live row (or column), where most threads have nothing to compute and the rest a
// sliver; entries outside the sequence come back as 0.
__device__ __forceinline__ void tile_scores(float (&s)[4][4], const float* __restrict__ Qs, const float* __restrict__ Ks,
                                            int KD, int ty, int tx, int nr, int nc) {
  const int LDQ = KD + 4;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
  if (nr == 0 || nc == 0) return;
  if (nr == 4 && nc == 4) {
    for (int k = 0; k < KD; k += 4) {
      float4 q[4], kk[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) q[r] = ld4(Qs + (ty + 16 * r) * LDQ + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) kk[c] = ld4(Ks + (tx + 16 * c) * LDQ + k);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s[r][c] = fmaf(q[r].x, kk[c].x, s[r][c]);
          s[r][c] = fmaf(q[r].y, kk[c].y, s[r][c]);
          s[r][c] = fmaf(q[r].z, kk[c].z, s[r][c]);
          s[r][c] = fmaf(q[r].w, kk[c].w, s[r][c]);
        }
    }
    return;
  }
  for (int k = 0; k < KD; k += 4) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r >= nr) break;
      const float4 q = ld4(Qs + (ty + 16 * r) * LDQ + k);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c >= nc) break;
        const float4 kk = ld4(Ks + (tx + 16 * c) * LDQ + k);
        s[r][c] = fmaf(q.x, kk.x, s[r][c]);
        s[r][c] = fmaf(q.y, kk.y, s[r][c]);
        s[r][c] = fmaf(q.z, kk.z, s[r][c]);
        s[r][c] = fmaf(q.w, kk.w, s[r][c]);
      }
    }
  }
}

// 4x4 micro-tile of X Y^T for two token tiles [64][DH+4] (dP = dO V^T)
template <int DH>
__device__ __forceinline__ void tile_outer(float (&s)[4][4], const float* __restrict__ X, const float* __restrict__ Y,
                                           int ty, int tx, int nr, int nc) {
  constexpr int LDM = DH + 4;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
  if (nr == 0 || nc == 0) return;
#pragma unroll
  for (int k = 0; k < DH; k += 4) {
    float4 q[4], kk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = ld4(X + (ty + 16 * r) * LDM + k);
#pragma unroll
    for (int c = 0; c < 4; ++c) kk[c] = ld4(Y + (tx + 16 * c) * LDM + k);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        s[r][c] += q[r].x * kk[c].x + q[r].y * kk[c].y + q[r].z * kk[c].z + q[r].w * kk[c].w;
  }
}

__device__ __forceinline__ float half_max(float v) {  // over the 16 lanes that share a row
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct TileSmem {
  int Q, K, V, dO, P, G, c, m, l, alpha, extra, total;
};
__host__ __device__ inline TileSmem tile_layout(int KD, int DH, bool bwd) {
  TileSmem L;
  int o = 0;
  L.Q = o; o += TQ * (KD + 4);
  L.K = o; o += TQ * (KD + 4);
  L.V = o; o += TQ * (DH + 4);
  L.dO = o; if (bwd) o += TQ * (DH + 4);
  L.P = o; o += TQ * LDP;
  L.G = o; if (bwd) o += TQ * LDP;
  L.c = o; o += 128;
  L.m = o; o += TQ;
  L.l = o; o += TQ;
  L.alpha = o; o += TQ;
  L.extra = o; o += TQ;
  L.total = o;
  return L;
}

// ---- forward ----------------------------------------------------------------------------------------
// grid (nqt, B*H); 256 threads.
template <typename T, int DH, int MODE>
__global__ void __launch_bounds__(256) tile_fwd_kernel(const TileArgs p) {
  constexpr int LDM = DH + 4, NU = (DH + 15) / 16;  // PV: thread (row = tid/4, dq = tid%4) owns d = 4*(dq + 4u)
  extern __shared__ __align__(16) float smem[];
  const TileSmem L = tile_layout(p.KD, DH, false);
  float* Qs = smem + L.Q; float* Ks = smem + L.K; float* Vs = smem + L.V; float* Ps = smem + L.P;
  float* cs = smem + L.c; float* ms = smem + L.m; float* ls = smem + L.l; float* as = smem + L.alpha;
  const int pair = blockIdx.y, b = pair / p.H, h = pair % p.H, i0 = blockIdx.x * TQ;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int prow = threadIdx.x >> 2, dq = threadIdx.x & 3;
  const int N = p.N;
  const float* qrows = p.qrows + (size_t)pair * N * p.ldr;
  const float* krows = p.krows + (size_t)pair * N * p.ldr;
  const T* vb = static_cast<const T*>(p.qkv) + qkv_off(b, 0, 2, h, N, p.H, DH);
  const size_t tok_stride = (size_t)3 * p.H * DH;
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;

  load_rows(Qs, qrows, p.ldr, p.KD, i0, N);
  if (threadIdx.x < TQ) { ms[threadIdx.x] = -INFINITY; ls[threadIdx.x] = 0.f; }
  float acc[NU][4];
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[u][e] = 0.f;

  for (int j0 = 0; j0 < N; j0 += TQ) {
    __syncthreads();  // previous tile fully consumed
    load_rows(Ks, krows, p.ldr, p.KD, j0, N);
    load_tile<T, DH, TQ>(Vs, vb, tok_stride, j0, N, 0.f);
    if (MODE == MODE_KERPLE && threadIdx.x < 127) {
      int d = j0 - i0 - 63 + (int)threadIdx.x + N - 1;
      cs[threadIdx.x] = (d >= 0 && d < 2 * N - 1) ? __ldg(p.cexp + (size_t)h * (2 * N - 1) + d) : 0.f;
    }
    __syncthreads();
    float s[4][4];
    tile_scores(s, Qs, Ks, p.KD, ty, tx, valid_count(i0 + ty, N), valid_count(j0 + tx, N));
    if (MODE == MODE_KERPLE) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int il = ty + 16 * r, i = i0 + il;
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jl = tx + 16 * c, j = j0 + jl;
          float a = (i < N && j < N) ? s[r][c] * cs[jl - il + 63] : 0.f;
          Ps[il * LDP + jl] = a;
          rs += a;
        }
        rs = half_sum(rs);
        if (tx == 0) ls[il] += rs;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int il = ty + 16 * r, i = i0 + il;
        float tmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + tx + 16 * c;
          float v = s[r][c] * p.scale;
          if (j >= N) v = -INFINITY;
          else if (p.mask && i < N && p.mask[((size_t)b * N + i) * N + j] == 0) v = -INFINITY;
          s[r][c] = v;
          tmax = fmaxf(tmax, v);
        }
        tmax = half_max(tmax);
        const float m_old = ms[il];
        const float m_new = fmaxf(m_old, tmax);
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jl = tx + 16 * c;
          float pv = (m_new == -INFINITY) ? 0.f : expf(s[r][c] - m_new);
          rs += pv;
          if (p.dropout_p > 0.f)
            pv = (rng_uniform(eff_seed(p), pair, i, j0 + jl) >= p.dropout_p) ? pv * keep_scale : 0.f;
          Ps[il * LDP + jl] = pv;
        }
        rs = half_sum(rs);
        __syncwarp();
        if (tx == 0) {
          const float alpha = (m_new == -INFINITY) ? 1.f : expf(m_old - m_new);
          as[il] = alpha;
          ls[il] = ls[il] * alpha + rs;
          ms[il] = m_new;
        }
      }
    }
    __syncthreads();
    {  // acc[row][d] = acc * alpha + sum_j P[row][j] V[j][d]
      const float alpha = (MODE == MODE_SOFTMAX) ? as[prow] : 1.f;
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[u][e] *= alpha;
      const float* prowp = Ps + prow * LDP;
      const int jend = (i0 + prow < N) ? min(TQ, (N - j0 + 3) & ~3) : 0;  // keys past the sequence carry zero weight
      for (int j = 0; j < jend; j += 4) {
        const float4 pv = ld4(prowp + j);
        const float pj[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int d = 4 * (dq + 4 * u);
            if (d < DH) {
              const float4 v = ld4(Vs + (j + jj) * LDM + d);
              acc[u][0] = fmaf(pj[jj], v.x, acc[u][0]);
              acc[u][1] = fmaf(pj[jj], v.y, acc[u][1]);
              acc[u][2] = fmaf(pj[jj], v.z, acc[u][2]);
              acc[u][3] = fmaf(pj[jj], v.w, acc[u][3]);
            }
          }
      }
    }
  }
  __syncthreads();
  {  // finalize
    const int i = i0 + prow;
    if (i < N) {
      const float l = ls[prow];
      const float den = (MODE == MODE_KERPLE) ? (l + kEps) : l;
      T* ob = static_cast<T*>(p.out) + out_off(b, i, h, N, p.H, DH);
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int d = 4 * (dq + 4 * u);
        if (d < DH) st4(ob + d, make_float4(acc[u][0] / den, acc[u][1] / den, acc[u][2] / den, acc[u][3] / den));
      }
      if (dq == 0) p.stat[(size_t)pair * N + i] = (MODE == MODE_KERPLE) ? l : (ms[prow] + logf(l));
    }
  }
  if (MODE == MODE_SOFTMAX && p.attn_out != nullptr) {  // return_attention=True: second sweep with the final lse
    for (int j0 = 0; j0 < N; j0 += TQ) {
      __syncthreads();
      load_rows(Ks, krows, p.ldr, p.KD, j0, N);
      __syncthreads();
      float s[4][4];
      tile_scores(s, Qs, Ks, p.KD, ty, tx, valid_count(i0 + ty, N), valid_count(j0 + tx, N));
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int il = ty + 16 * r, i = i0 + il;
        if (i >= N) continue;
        const float lse = ms[il] + logf(ls[il]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + tx + 16 * c;
          if (j >= N) continue;
          float v = s[r][c] * p.scale;
          if (p.mask && p.mask[((size_t)b * N + i) * N + j] == 0) v = -INFINITY;
          float pv = expf(v - lse);
          if (p.dropout_p > 0.f) pv = (rng_uniform(eff_seed(p), pair, i, j) >= p.dropout_p) ? pv * keep_scale : 0.f;
          p.attn_out[((size_t)pair * N + i) * N + j] = pv;
        }
      }
    }
  }
}

// ---- backward ---------------------------------------------------------------------------------------
// Shared per-tile math.  Given the score micro-tile s (raw dot products) and dp = dO V^T micro-tile:
//   softmax: P = exp(s*scale - lse_i); Pd = P*keep; dPd = dp; dS = P*(dPd*keep - D_i)*scale
//   kerple : A = s*c; dA = dp*r_i + dden_i; G = dA*c; also W = dA*A (for d bias)
// Writes Pd (or A) to Ps and dS (or G) to Gs.
template <int MODE>
__device__ __forceinline__ void bwd_tile_weights(const TileArgs& p, float (&s)[4][4], float (&dp)[4][4], float* Ps,
                                                 float* Gs, const float* cs, const float* rowA, const float* rowB,
                                                 int pair, int b, int i0, int j0, int ty, int tx, float keep_scale,
                                                 bool q_is_row /*Ps indexed [query][key]*/) {
  const int N = p.N;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int il = ty + 16 * r, i = i0 + il;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int jl = tx + 16 * c, j = j0 + jl;
      float w = 0.f, g = 0.f;
      if (i < N && j < N) {
        if (MODE == MODE_KERPLE) {
          const float cc = cs[jl - il + 63];
          w = s[r][c] * cc;                               // A
          g = (dp[r][c] * rowA[il] + rowB[il]) * cc;      // (dnum.v + dden) * c, rowA = r_i, rowB = dden_i
        } else {
          float v = s[r][c] * p.scale;
          if (p.mask && p.mask[((size_t)b * N + i) * N + j] == 0) v = -INFINITY;
          const float pr = expf(v - rowA[il]);            // rowA = lse_i
          float keep = 1.f;
          if (p.dropout_p > 0.f) keep = (rng_uniform(eff_seed(p), pair, i, j) >= p.dropout_p) ? keep_scale : 0.f;
          w = pr * keep;                                  // Pd
          g = pr * (dp[r][c] * keep - rowB[il]) * p.scale;  // rowB = D_i
        }
      }
      Ps[il * LDP + jl] = w;
      Gs[il * LDP + jl] = g;
    }
  }
  (void)q_is_row;
}

// per-row backward statistics for a query tile, into smem: rowA, rowB (see bwd_tile_weights)
template <typename T, int DH, int MODE>
__device__ __forceinline__ void bwd_row_stats(const TileArgs& p, float* rowA, float* rowB, const float* dOs, int pair,
                                              int b, int h, int i0) {
  constexpr int LDM = DH + 4;
  const int N = p.N;
  if (threadIdx.x < TQ) {
    const int i = i0 + threadIdx.x;
    float a = 0.f, bb = 0.f;
    if (i < N) {
      const T* ob = static_cast<const T*>(p.out) + out_off(b, i, h, N, p.H, DH);
      float dot = 0.f;
      for (int d = 0; d < DH; d += 4) {
        float4 o = ld4(ob + d);
        float4 g = ld4(dOs + threadIdx.x * LDM + d);
        dot += o.x * g.x + o.y * g.y + o.z * g.z + o.w * g.w;
      }
      const float st = p.stat[(size_t)pair * N + i];
      if (MODE == MODE_KERPLE) {
        a = 1.0f / (st + kEps);   // r_i
        bb = -dot * a;            // dden_i
      } else {
        a = st;                   // lse_i
        bb = dot;                 // D_i
      }
    }
    rowA[threadIdx.x] = a;
    rowB[threadIdx.x] = bb;
  }
}

// dQ (or d phi_q): grid (nqt, B*H).  NCH = ceil(KD/64) output column chunks held in registers.
template <typename T, int DH, int MODE, int NCH>
__global__ void __launch_bounds__(256) tile_bwd_dq_kernel(const TileArgs p) {
  extern __shared__ __align__(16) float smem[];
  const TileSmem L = tile_layout(p.KD, DH, true);
  float* Qs = smem + L.Q; float* Ks = smem + L.K; float* Vs = smem + L.V; float* dOs = smem + L.dO;
  float* Ps = smem + L.P; float* Gs = smem + L.G; float* cs = smem + L.c;
  float* rowA = smem + L.m; float* rowB = smem + L.l;
  const int pair = blockIdx.y, b = pair / p.H, h = pair % p.H, i0 = blockIdx.x * TQ;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int N = p.N, KD = p.KD, LDQ = KD + 4;
  const float* qrows = p.qrows + (size_t)pair * N * p.ldr;
  const float* krows = p.krows + (size_t)pair * N * p.ldr;
  const T* vb = static_cast<const T*>(p.qkv) + qkv_off(b, 0, 2, h, N, p.H, DH);
  const T* dob = static_cast<const T*>(p.dout) + out_off(b, 0, h, N, p.H, DH);
  const size_t tok_stride = (size_t)3 * p.H * DH, out_stride = (size_t)p.H * DH;
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;
  float* dpart = (MODE == MODE_KERPLE) ? p.dbias_part + ((size_t)pair * gridDim.x + blockIdx.x) * (2 * N - 1) : nullptr;

  load_rows(Qs, qrows, p.ldr, KD, i0, N);
  load_tile<T, DH, TQ>(dOs, dob, out_stride, i0, N, 0.f);
  __syncthreads();
  bwd_row_stats<T, DH, MODE>(p, rowA, rowB, dOs, pair, b, h, i0);
  float acc[NCH][4][4];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[ch][r][c] = 0.f;

  for (int j0 = 0; j0 < N; j0 += TQ) {
    __syncthreads();
    load_rows(Ks, krows, p.ldr, KD, j0, N);
    load_tile<T, DH, TQ>(Vs, vb, tok_stride, j0, N, 0.f);
    if (MODE == MODE_KERPLE && threadIdx.x < 127) {
      int d = j0 - i0 - 63 + (int)threadIdx.x + N - 1;
      cs[threadIdx.x] = (d >= 0 && d < 2 * N - 1) ? __ldg(p.cexp + (size_t)h * (2 * N - 1) + d) : 0.f;
    }
    __syncthreads();
    float s[4][4], dp[4][4];
    const int nr = valid_count(i0 + ty, N), nc = valid_count(j0 + tx, N);
    tile_scores(s, Qs, Ks, KD, ty, tx, nr, nc);
    tile_outer<DH>(dp, dOs, Vs, ty, tx, nr, nc);
    bwd_tile_weights<MODE>(p, s, dp, Ps, Gs, cs, rowA, rowB, pair, b, i0, j0, ty, tx, keep_scale, true);
    __syncthreads();
    // dQ[i][k] += sum_j G[i][j] K[j][k]; thread: rows ty+16r, cols tx+16c+64ch
    const int jend = (nr > 0) ? min(TQ, (N - j0 + 3) & ~3) : 0;  // G is zero outside the sequence
    for (int j = 0; j < jend; j += 4) {
      float4 g[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) g[r] = ld4(Gs + (ty + 16 * r) * LDP + j);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float* krow = Ks + (j + jj) * LDQ;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = tx + 16 * c + 64 * ch;
            const float kv = (k < KD) ? krow[k] : 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float gv = jj == 0 ? g[r].x : (jj == 1 ? g[r].y : (jj == 2 ? g[r].z : g[r].w));
              acc[ch][r][c] = fmaf(gv, kv, acc[ch][r][c]);
            }
          }
      }
    }
    if (MODE == MODE_KERPLE && threadIdx.x < 127) {
      // d bias[delta] = sum over the diagonal j - i = delta of dA * A = (G * A) / c, with G = dA * c and c = exp(b)
      // constant along a diagonal (db = dc * c, dc = sum_diag dA * P).
      const int k = threadIdx.x;  // diagonal index, delta_local = k - 63
      float sum = 0.f;
      const float cc = cs[k];
      if (cc > 0.f) {
        const int il_lo = max(0, 63 - k), il_hi = min(TQ - 1, 126 - k);
        for (int il = il_lo; il <= il_hi; ++il) {
          const int jl = il + k - 63;
          sum += Gs[il * LDP + jl] * Ps[il * LDP + jl];
        }
        sum /= cc;
        const int d = j0 - i0 - 63 + k + N - 1;
        if (d >= 0 && d < 2 * N - 1) dpart[d] += sum;  // private to this CTA
      }
    }
  }
  // write dQ rows
  float* dq = p.dqrows + (size_t)pair * N * p.ldr;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 16 * r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = tx + 16 * c + 64 * ch;
        if (i < N && k < KD) dq[(size_t)i * p.ldr + k] = acc[ch][r][c];
      }
    }
}

// dK (or d phi_k) and dV: grid (nkt, B*H); loops over query tiles.
template <typename T, int DH, int MODE, int NCH>
__global__ void __launch_bounds__(256) tile_bwd_dkv_kernel(const TileArgs p) {
  constexpr int LDM = DH + 4, NU = (DH + 15) / 16;
  extern __shared__ __align__(16) float smem[];
  const TileSmem L = tile_layout(p.KD, DH, true);
  float* Qs = smem + L.Q; float* Ks = smem + L.K; float* Vs = smem + L.V; float* dOs = smem + L.dO;
  float* Ps = smem + L.P; float* Gs = smem + L.G; float* cs = smem + L.c;
  float* rowA = smem + L.m; float* rowB = smem + L.l;
  const int pair = blockIdx.y, b = pair / p.H, h = pair % p.H, j0 = blockIdx.x * TQ;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int prow = threadIdx.x >> 2, dq4 = threadIdx.x & 3;
  const int N = p.N, KD = p.KD, LDQ = KD + 4;
  const float* qrows = p.qrows + (size_t)pair * N * p.ldr;
  const float* krows = p.krows + (size_t)pair * N * p.ldr;
  const T* vb = static_cast<const T*>(p.qkv) + qkv_off(b, 0, 2, h, N, p.H, DH);
  const T* dob = static_cast<const T*>(p.dout) + out_off(b, 0, h, N, p.H, DH);
  const size_t tok_stride = (size_t)3 * p.H * DH, out_stride = (size_t)p.H * DH;
  const float keep_scale = p.dropout_p > 0.f ? 1.f / (1.f - p.dropout_p) : 1.f;

  load_rows(Ks, krows, p.ldr, KD, j0, N);
  load_tile<T, DH, TQ>(Vs, vb, tok_stride, j0, N, 0.f);
  float acc[NCH][4][4];   // dK: rows = keys ty+16r, cols tx+16c+64ch
  float accv[NU][4];      // dV: row = key prow, d = 4*(dq4 + 4u)
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[ch][r][c] = 0.f;
#pragma unroll
  for (int u = 0; u < NU; ++u)
#pragma unroll
    for (int e = 0; e < 4; ++e) accv[u][e] = 0.f;

  for (int i0 = 0; i0 < N; i0 += TQ) {
    __syncthreads();
    load_rows(Qs, qrows, p.ldr, KD, i0, N);
    load_tile<T, DH, TQ>(dOs, dob, out_stride, i0, N, 0.f);
    if (MODE == MODE_KERPLE && threadIdx.x < 127) {
      int d = j0 - i0 - 63 + (int)threadIdx.x + N - 1;
      cs[threadIdx.x] = (d >= 0 && d < 2 * N - 1) ? __ldg(p.cexp + (size_t)h * (2 * N - 1) + d) : 0.f;
    }
    __syncthreads();
    bwd_row_stats<T, DH, MODE>(p, rowA, rowB, dOs, pair, b, h, i0);
    __syncthreads();
    float s[4][4], dp[4][4];
    const int nr = valid_count(i0 + ty, N), nc = valid_count(j0 + tx, N);
    tile_scores(s, Qs, Ks, KD, ty, tx, nr, nc);  // s[r][c]: query ty+16r, key tx+16c
    tile_outer<DH>(dp, dOs, Vs, ty, tx, nr, nc);
    bwd_tile_weights<MODE>(p, s, dp, Ps, Gs, cs, rowA, rowB, pair, b, i0, j0, ty, tx, keep_scale, true);
    __syncthreads();
    // dK[j][k] += sum_i G[i][j] Q[i][k]; rows j = ty+16r
    const int iend_all = min(TQ, N - i0);  // queries past the sequence contribute nothing
    const int iend = (valid_count(j0 + ty, N) > 0) ? iend_all : 0;
    for (int i = 0; i < iend; ++i) {
      float g[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) g[r] = Gs[i * LDP + ty + 16 * r];
      const float* qrow = Qs + i * LDQ;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = tx + 16 * c + 64 * ch;
          const float qv = (k < KD) ? qrow[k] : 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[ch][r][c] = fmaf(g[r], qv, acc[ch][r][c]);
        }
    }
    // dV[j][d] += sum_i W[i][j] * (dO[i][d] * (kerple ? r_i : 1))
    const int iend_v = (j0 + prow < N) ? iend_all : 0;
    for (int i = 0; i < iend_v; ++i) {
      float w = Ps[i * LDP + prow];
      if (MODE == MODE_KERPLE) w *= rowA[i];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int d = 4 * (dq4 + 4 * u);
        if (d < DH) {
          const float4 g = ld4(dOs + i * LDM + d);
          accv[u][0] = fmaf(w, g.x, accv[u][0]);
          accv[u][1] = fmaf(w, g.y, accv[u][1]);
          accv[u][2] = fmaf(w, g.z, accv[u][2]);
          accv[u][3] = fmaf(w, g.w, accv[u][3]);
        }
      }
    }
  }
  float* dk = p.dkrows + (size_t)pair * N * p.ldr;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + ty + 16 * r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = tx + 16 * c + 64 * ch;
        if (j < N && k < KD) dk[(size_t)j * p.ldr + k] = acc[ch][r][c];
      }
    }
  const int j = j0 + prow;
  if (j < N) {
    T* dvb = static_cast<T*>(p.dqkv) + qkv_off(b, j, 2, h, N, p.H, DH);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int d = 4 * (dq4 + 4 * u);
      if (d < DH) st4(dvb + d, make_float4(accv[u][0], accv[u][1], accv[u][2], accv[u][3]));
    }
  }
}

// ---- small helpers ----------------------------------------------------------------------------------
__global__ void exp_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = expf(x[i]);
}

// dbias[h][d] = sum over (b, qtile) of part[(b*H+h)*nqt + t][d]
// d bias[h][d] = sum over (batch, query tile) of the per-CTA partials.  Block (32, 32): x = diagonal (coalesced), y = slice of
// the B*nqt partials; the 32 slice sums are then added in a fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(1024) dbias_reduce_kernel(const float* __restrict__ part, float* __restrict__ dbias,
                                                            int B, int H, int nqt, int W) {
  __shared__ float red[32][33];
  const size_t i = (size_t)blockIdx.x * 32 + threadIdx.x;
  const bool live = i < (size_t)H * W;
  const int h = live ? (int)(i / W) : 0, d = live ? (int)(i % W) : 0;
  float acc = 0.f;
  if (live)
    for (int u = threadIdx.y; u < B * nqt; u += 32) {
      const int b = u / nqt, t = u % nqt;
      acc += part[(((size_t)b * H + h) * nqt + t) * W + d];
    }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    float sum = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) sum += red[y][threadIdx.x];
    dbias[i] = sum;
  }
}

// q and k of the packed buffer -> rotated fp32 rows [2][B*H][N][DH]   (softmax prologue)
struct RotPackArgs {
  const void* qkv; void* dqkv;
  float* rows;          // fwd out: [2][B*H][N][DH]
  const float* drows;   // bwd in
  const float* ta; const float* tb;
  float* dg_part;
  int B, N, H, rot, slots;
};

template <typename T, int DH, bool BWD>
__global__ void __launch_bounds__(256) rot_pack_kernel(const RotPackArgs p) {
  constexpr int LDM = DH + 4, TT = 32;
  __shared__ __align__(16) float xr[TT * LDM];
  __shared__ __align__(16) float xs[TT * LDM];
  __shared__ __align__(16) float tmp[TT * LDM];
  __shared__ __align__(16) float g2[TT * 2 * DH];
  RotArgs ra{p.rot, p.ta, p.tb};
  const size_t tok_stride = (size_t)3 * p.H * DH;
  const size_t plane = (size_t)p.B * p.H * p.N * DH;
  for (int pair = blockIdx.x; pair < p.B * p.H; pair += gridDim.x) {
    const int b = pair / p.H, h = pair % p.H;
    float* dg_slot = (BWD && p.dg_part) ? p.dg_part + ((size_t)h * p.slots + blockIdx.x / p.H) * p.N * DH : nullptr;
    for (int which = 0; which < 2; ++which) {
      const T* src = static_cast<const T*>(p.qkv) + qkv_off(b, 0, which, h, p.N, p.H, DH);
      for (int n0 = 0; n0 < p.N; n0 += TT) {
        load_tile<T, DH, TT>(xr, src, tok_stride, n0, p.N, 0.f);
        if (p.rot == ERV_ROT_CIRCULANT) load_g_tile<DH, TT>(g2, p.ta, h, n0, p.N);
        if (BWD) load_tile<float, DH, TT>(xs, p.drows + which * plane + (size_t)pair * p.N * DH, DH, n0, p.N, 0.f);
        __syncthreads();
        if (!BWD) {
          prep_tile<DH, TT>(xs, xr, g2, nullptr, ra, ERV_PREP_NONE, 1.f, n0, p.N);
          __syncthreads();
          store_tile<float, DH, TT>(p.rows + which * plane + (size_t)pair * p.N * DH, DH, xs, n0, p.N);
        } else {
          T* dst = static_cast<T*>(p.dqkv) + qkv_off(b, 0, which, h, p.N, p.H, DH);
          prep_tile_bwd<T, DH, TT>(xs, nullptr, xr, g2, nullptr, tmp, ra, ERV_PREP_NONE, 1.f, dst, tok_stride, dg_slot,
                                   n0, p.N);
        }
        __syncthreads();
      }
    }
  }
}

// ---- host launchers ---------------------------------------------------------------------------------
template <typename K>
static int launch_tile(K kernel, dim3 grid, size_t smem, cudaStream_t st, const TileArgs& a) {
  ERV_CUDA(allow_smem(kernel, smem));
  kernel<<<grid, 256, smem, st>>>(a);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

#define TILE_DH_SWITCH(DH_, MACRO)                                                                       \
  switch (DH_) {                                                                                         \
    case 8: MACRO(8); break;                                                                             \
    case 16: MACRO(16); break;                                                                           \
    case 32: MACRO(32); break;                                                                           \
    case 64: MACRO(64); break;                                                                           \
    default: set_error("unsupported head_dim %d (supported: 8, 16, 32, 64)", DH_); return ERV_E_UNSUPPORTED; \
  }

template <typename T, int MODE>
static int tile_forward(const TileArgs& a, int DH, cudaStream_t st) {
  const int nqt = (a.N + TQ - 1) / TQ;
  const size_t smem = (size_t)tile_layout(a.KD, DH, false).total * sizeof(float);
  if (smem > kMaxSmem) { set_error("tile attention: inner dim %d does not fit shared memory", a.KD); return ERV_E_UNSUPPORTED; }
  dim3 grid(nqt, a.B * a.H);
#define FWD_CASE(D) return launch_tile(tile_fwd_kernel<T, D, MODE>, grid, smem, st, a)
  TILE_DH_SWITCH(DH, FWD_CASE)
#undef FWD_CASE
  return ERV_OK;
}

template <typename T, int DH, int MODE>
static int tile_backward_nch(const TileArgs& a, cudaStream_t st) {
  const int nqt = (a.N + TQ - 1) / TQ;
  const size_t smem = (size_t)tile_layout(a.KD, DH, true).total * sizeof(float);
  if (smem > kMaxSmem) { set_error("tile attention backward: inner dim %d does not fit shared memory", a.KD); return ERV_E_UNSUPPORTED; }
  dim3 grid(nqt, a.B * a.H);
  const int nch = (a.KD + 63) / 64;
  int rc;
#define BWD_CASE(NCH_)                                                                     \
  do {                                                                                     \
    rc = launch_tile(tile_bwd_dq_kernel<T, DH, MODE, NCH_>, grid, smem, st, a);           \
    if (rc) return rc;                                                                     \
    return launch_tile(tile_bwd_dkv_kernel<T, DH, MODE, NCH_>, grid, smem, st, a);        \
  } while (0)
  if constexpr (MODE == MODE_SOFTMAX) {
    BWD_CASE(1);
  } else {
    switch (nch) {
      case 1: BWD_CASE(1);
      case 2: BWD_CASE(2);
      case 3: BWD_CASE(3);
      case 4: BWD_CASE(4);
      case 5: BWD_CASE(5);
      default: set_error("kerple backward: num_features %d > 320 unsupported", a.KD); return ERV_E_UNSUPPORTED;
    }
  }
#undef BWD_CASE
}

template <typename T, int MODE>
static int tile_backward(const TileArgs& a, int DH, cudaStream_t st) {
#define BWDD_CASE(D) return tile_backward_nch<T, D, MODE>(a, st)
  TILE_DH_SWITCH(DH, BWDD_CASE)
#undef BWDD_CASE
  return ERV_OK;
}

template <typename T, bool BWD>
static int rot_pack(const RotPackArgs& a, int DH, cudaStream_t st) {
  const int grid = la_grid(a.B, a.H);
#define ROT_CASE(D) rot_pack_kernel<T, D, BWD><<<grid, 256, 0, st>>>(a)
  TILE_DH_SWITCH(DH, ROT_CASE)
#undef ROT_CASE
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

static int check_shape(const char* fn, int B, int N, int H, int DH, int dtype) {
  if (B <= 0 || N <= 0 || H <= 0) { set_error("%s: non-positive shape", fn); return ERV_E_INVALID; }
  if (!(DH == 8 || DH == 16 || DH == 32 || DH == 64)) {
    set_error("%s: unsupported head_dim %d (supported: 8, 16, 32, 64)", fn, DH);
    return ERV_E_UNSUPPORTED;
  }
  if (dtype != ERV_F32 && dtype != ERV_BF16) { set_error("%s: bad dtype %d", fn, dtype); return ERV_E_INVALID; }
  return ERV_OK;
}

}  // namespace erv

using namespace erv;

// ---- softmax ----------------------------------------------------------------------------------------
// workspace: rotated q,k rows fp32 [2][B*H][N][DH]  (+ their gradients for the backward)
extern "C" size_t erv_softmax_attention_workspace(int B, int N, int H, int head_dim, int rot, int backward) {
  (void)rot;
  size_t rows = align_up((size_t)2 * B * H * N * head_dim * sizeof(float), 256);
  return backward ? 2 * rows : rows;
}

extern "C" int erv_softmax_attention_fwd(const void* qkv, void* out, float* lse_out, float* attn_out,
                                         const uint8_t* mask, int B, int N, int H, int head_dim, int rot,
                                         const float* tab_a, const float* tab_b, float dropout_p, uint64_t seed,
                                         const long long* seed_dev, int dtype, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  const char* fn = "erv_softmax_attention_fwd";
  int rc = check_shape(fn, B, N, H, head_dim, dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(qkv && out && lse_out && workspace, "%s: null pointer", fn);
  ERV_CHECK_ARG(rot == ERV_ROT_NONE || tab_a, "%s: rotation table missing", fn);
  ERV_CHECK_ARG(rot != ERV_ROT_ROPE || tab_b, "%s: rope sin table missing", fn);
  ERV_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "%s: dropout_p %f out of range", fn, dropout_p);
  if (workspace_bytes < erv_softmax_attention_workspace(B, N, H, head_dim, rot, 0)) {
    set_error("%s: workspace too small", fn);
    return ERV_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (stile_tc_eligible(N, head_dim))  // tensor-core tiles: rotation in the tile prologue, nothing staged in HBM
    return stile_tc_forward(qkv, out, lse_out, attn_out, mask, B, N, H, rot, tab_a, tab_b, dropout_p, seed, seed_dev, dtype, st);
  RotPackArgs r{qkv, nullptr, (float*)workspace, nullptr, tab_a, tab_b, nullptr, B, N, H, rot, 1};
  rc = dtype == ERV_F32 ? rot_pack<float, false>(r, head_dim, st) : rot_pack<__nv_bfloat16, false>(r, head_dim, st);
  if (rc) return rc;
  TileArgs a{};
  const size_t plane = (size_t)B * H * N * head_dim;
  a.qrows = (const float*)workspace; a.krows = a.qrows + plane;
  a.qkv = qkv; a.out = out; a.stat = lse_out; a.attn_out = attn_out; a.mask = mask;
  a.B = B; a.N = N; a.H = H; a.KD = head_dim; a.ldr = head_dim;
  a.scale = (float)pow((double)head_dim, -0.5); a.dropout_p = dropout_p; a.seed = seed;
  a.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  return dtype == ERV_F32 ? tile_forward<float, MODE_SOFTMAX>(a, head_dim, st)
                          : tile_forward<__nv_bfloat16, MODE_SOFTMAX>(a, head_dim, st);
}

extern "C" int erv_softmax_attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout,
                                         void* dqkv, const uint8_t* mask, int B, int N, int H, int head_dim, int rot,
                                         const float* tab_a, const float* tab_b, float* dg_part, float dropout_p,
                                         uint64_t seed, const long long* seed_dev, int dtype, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  const char* fn = "erv_softmax_attention_bwd";
  int rc = check_shape(fn, B, N, H, head_dim, dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(qkv && out && lse && dout && dqkv && workspace, "%s: null pointer", fn);
  ERV_CHECK_ARG(rot == ERV_ROT_NONE || tab_a, "%s: rotation table missing", fn);
  ERV_CHECK_ARG(rot != ERV_ROT_ROPE || tab_b, "%s: rope sin table missing", fn);
  ERV_CHECK_ARG(rot != ERV_ROT_CIRCULANT || dg_part, "%s: dg_part missing", fn);
  if (workspace_bytes < erv_softmax_attention_workspace(B, N, H, head_dim, rot, 1)) {
    set_error("%s: workspace too small", fn);
    return ERV_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int slots = la_slots(B, H);
  static const bool stile_bwd_off = getenv("ERV_DISABLE_STILE_TC_BWD") != nullptr;
  if (stile_tc_eligible(N, head_dim) && !stile_bwd_off) {
    if (rot == ERV_ROT_CIRCULANT) ERV_CUDA(cudaMemsetAsync(dg_part, 0, (size_t)H * slots * N * head_dim * sizeof(float), st));
    return stile_tc_backward(qkv, out, lse, dout, dqkv, mask, B, N, H, rot, tab_a, tab_b,
                             rot == ERV_ROT_CIRCULANT ? dg_part : nullptr, slots, dropout_p, seed, seed_dev, dtype, st);
  }
  const size_t plane = (size_t)B * H * N * head_dim;
  float* rows = (float*)workspace;
  float* drows = (float*)((char*)workspace + erv_softmax_attention_workspace(B, N, H, head_dim, rot, 0));
  RotPackArgs r{qkv, dqkv, rows, drows, tab_a, tab_b, rot == ERV_ROT_CIRCULANT ? dg_part : nullptr, B, N, H, rot, slots};
  rc = dtype == ERV_F32 ? rot_pack<float, false>(r, head_dim, st) : rot_pack<__nv_bfloat16, false>(r, head_dim, st);
  if (rc) return rc;
  TileArgs a{};
  a.qrows = rows; a.krows = rows + plane; a.dqrows = drows; a.dkrows = drows + plane;
  a.qkv = qkv; a.dqkv = dqkv; a.out = const_cast<void*>(out); a.dout = dout; a.stat = const_cast<float*>(lse);
  a.mask = mask; a.B = B; a.N = N; a.H = H; a.KD = head_dim; a.ldr = head_dim;
  a.scale = (float)pow((double)head_dim, -0.5); a.dropout_p = dropout_p; a.seed = seed;
  a.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  rc = dtype == ERV_F32 ? tile_backward<float, MODE_SOFTMAX>(a, head_dim, st)
                        : tile_backward<__nv_bfloat16, MODE_SOFTMAX>(a, head_dim, st);
  if (rc) return rc;
  if (r.dg_part) ERV_CUDA(cudaMemsetAsync(dg_part, 0, (size_t)H * slots * N * head_dim * sizeof(float), st));
  return dtype == ERV_F32 ? rot_pack<float, true>(r, head_dim, st) : rot_pack<__nv_bfloat16, true>(r, head_dim, st);
}

// ---- KERPLE -----------------------------------------------------------------------------------------
// workspace: W^T | exp(bias) [H][2N-1] | phi_q, phi_k [B*H][N][ldphi] | (bwd) dphi_q, dphi_k | dbias partials
static size_t kerple_ldphi(int M) { return (size_t)(M + 7) / 8 * 8; }
struct KerpleWs { size_t wt, cexp, phi, dphi, dpart, fft, total; };
static KerpleWs kerple_ws(int B, int N, int H, int DH, int M, int backward) {
  KerpleWs w;
  size_t o = 0;
  w.wt = o; o += wt_bytes(H, DH, M);
  w.cexp = o; o += align_up((size_t)H * (2 * N - 1) * sizeof(float), 256);
  size_t phi = align_up((size_t)2 * B * H * N * kerple_ldphi(M) * sizeof(float), 256);
  w.phi = o; o += phi;
  w.dphi = o; if (backward) o += phi;
  w.dpart = o; if (backward) o += align_up((size_t)B * H * ((N + TQ - 1) / TQ) * (2 * N - 1) * sizeof(float), 256);
  w.fft = o; if (!backward && kerple_fft_eligible(B, N, H, DH, M)) o += kerple_fft_ws_bytes(B, N, H, DH, M);
  w.total = o;
  return w;
}

extern "C" size_t erv_kerple_attention_workspace(int B, int N, int H, int head_dim, int M, int backward) {
  return kerple_ws(B, N, H, head_dim, M, backward).total;
}

// W^T rows and exp(bias) into the workspace
static int kerple_tables(char* ws, const KerpleWs& w, const float* omega, const float* bias, int N, int H, int DH, int M,
                         int kind, cudaStream_t st) {
  int rc = prep_wt_public(omega, (float*)(ws + w.wt), H, DH, M, kind, st);
  if (rc) return rc;
  size_t nb = (size_t)H * (2 * N - 1);
  exp_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(bias, (float*)(ws + w.cexp), nb);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

static int kerple_features(const void* qkv, char* ws, const KerpleWs& w, const float* omega, const float* bias, int B,
                           int N, int H, int DH, int M, int kind, int dtype, cudaStream_t st) {
  int rc = kerple_tables(ws, w, omega, bias, N, H, DH, M, kind, st);
  if (rc) return rc;
  const int ld = (int)kerple_ldphi(M);
  const size_t es = dtype == ERV_F32 ? 4 : 2;
  const size_t plane = (size_t)B * H * N * ld;
  for (int which = 0; which < 2; ++which) {
    const char* src = (const char*)qkv + (size_t)which * H * DH * es;
    rc = launch_feature_map(src, nullptr, (size_t)N * 3 * H * DH, (size_t)DH, (size_t)3 * H * DH,
                            (float*)(ws + w.phi) + which * plane, nullptr, ld, (const float*)(ws + w.wt), B, N, H, DH, M,
                            kind, ERV_PREP_L2NORM, 1.f, dtype, false, st);
    if (rc) return rc;
  }
  return ERV_OK;
}

extern "C" int erv_kerple_attention_fwd(const void* qkv, void* out, float* den_out, const float* omega,
                                        const float* rel_pos_bias, int B, int N, int H, int head_dim, int M, int kind,
                                        int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "erv_kerple_attention_fwd";
  int rc = check_shape(fn, B, N, H, head_dim, dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(qkv && out && den_out && omega && rel_pos_bias && workspace, "%s: null pointer", fn);
  ERV_CHECK_ARG(M > 0 && M <= 320, "%s: num_features %d unsupported with KERPLE (max 320)", fn, M);
  KerpleWs w = kerple_ws(B, N, H, head_dim, M, 0);
  if (workspace_bytes < w.total) { set_error("%s: workspace too small", fn); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  if (kerple_fft_eligible(B, N, H, head_dim, M)) {  // long sequences: the reference's FFT route, fused with the read-out
    rc = kerple_features(qkv, ws, w, omega, rel_pos_bias, B, N, H, head_dim, M, kind, dtype, st);
    if (rc) return rc;
    const int ld = (int)kerple_ldphi(M);
    const float* phi = (const float*)(ws + w.phi);
    return kerple_fft_forward(qkv, out, den_out, phi, phi + (size_t)B * H * N * ld, ld, (const float*)(ws + w.cexp),
                              ws + w.fft, B, N, H, head_dim, M, dtype, st);
  }
  if (ktile_tc_eligible(N, head_dim, M)) {  // tensor-core tiles: features computed in the tile, nothing staged in HBM
    rc = kerple_tables(ws, w, omega, rel_pos_bias, N, H, head_dim, M, kind, st);
    if (rc) return rc;
    return ktile_tc_forward(qkv, out, den_out, (const float*)(ws + w.wt), (const float*)(ws + w.cexp), B, N, H, head_dim, M,
                            kind, dtype, st);
  }
  rc = kerple_features(qkv, ws, w, omega, rel_pos_bias, B, N, H, head_dim, M, kind, dtype, st);
  if (rc) return rc;
  const int ld = (int)kerple_ldphi(M);
  const size_t plane = (size_t)B * H * N * ld;
  TileArgs a{};
  a.qrows = (const float*)(ws + w.phi); a.krows = a.qrows + plane;
  a.qkv = qkv; a.out = out; a.stat = den_out; a.cexp = (const float*)(ws + w.cexp);
  a.B = B; a.N = N; a.H = H; a.KD = ld; a.ldr = ld; a.scale = 1.f;
  return dtype == ERV_F32 ? tile_forward<float, MODE_KERPLE>(a, head_dim, st)
                          : tile_forward<__nv_bfloat16, MODE_KERPLE>(a, head_dim, st);
}

extern "C" int erv_kerple_attention_bwd(const void* qkv, const void* out, const float* den, const void* dout,
                                        void* dqkv, float* dbias, const float* omega, const float* rel_pos_bias, int B,
                                        int N, int H, int head_dim, int M, int kind, int dtype, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  const char* fn = "erv_kerple_attention_bwd";
  int rc = check_shape(fn, B, N, H, head_dim, dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(qkv && out && den && dout && dqkv && dbias && omega && rel_pos_bias && workspace, "%s: null pointer", fn);
  ERV_CHECK_ARG(M > 0 && M <= 320, "%s: num_features %d unsupported with KERPLE (max 320)", fn, M);
  KerpleWs w = kerple_ws(B, N, H, head_dim, M, 1);
  if (workspace_bytes < w.total) { set_error("%s: workspace too small", fn); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  static const bool tc_bwd_off = getenv("ERV_DISABLE_KTILE_TC_BWD") != nullptr;
  if (ktile_tc_eligible(N, head_dim, M) && !tc_bwd_off) {
    rc = kerple_tables(ws, w, omega, rel_pos_bias, N, H, head_dim, M, kind, st);
    if (rc) return rc;
    return ktile_tc_backward(qkv, out, den, dout, dqkv, dbias, (float*)(ws + w.dpart), (const float*)(ws + w.wt),
                             (const float*)(ws + w.cexp), B, N, H, head_dim, M, kind, dtype, st);
  }
  rc = kerple_features(qkv, ws, w, omega, rel_pos_bias, B, N, H, head_dim, M, kind, dtype, st);
  if (rc) return rc;
  const int ld = (int)kerple_ldphi(M);
  const size_t plane = (size_t)B * H * N * ld;
  const int nqt = (N + TQ - 1) / TQ;
  ERV_CUDA(cudaMemsetAsync(ws + w.dpart, 0, (size_t)B * H * nqt * (2 * N - 1) * sizeof(float), st));
  TileArgs a{};
  a.qrows = (const float*)(ws + w.phi); a.krows = a.qrows + plane;
  a.dqrows = (float*)(ws + w.dphi); a.dkrows = a.dqrows + plane;
  a.qkv = qkv; a.dqkv = dqkv; a.out = const_cast<void*>(out); a.dout = dout; a.stat = const_cast<float*>(den);
  a.cexp = (const float*)(ws + w.cexp); a.dbias_part = (float*)(ws + w.dpart);
  a.B = B; a.N = N; a.H = H; a.KD = ld; a.ldr = ld; a.scale = 1.f;
  rc = dtype == ERV_F32 ? tile_backward<float, MODE_KERPLE>(a, head_dim, st)
                        : tile_backward<__nv_bfloat16, MODE_KERPLE>(a, head_dim, st);
  if (rc) return rc;
  {
    size_t nb = (size_t)H * (2 * N - 1);
    dbias_reduce_kernel<<<(unsigned)((nb + 31) / 32), dim3(32, 32), 0, st>>>((const float*)(ws + w.dpart), dbias, B, H, nqt,
                                                                      2 * N - 1);
    ERV_LAUNCH_CHECK();
  }
  // d phi -> d q, d k through the feature map and the L2 normalisation
  const size_t es = dtype == ERV_F32 ? 4 : 2;
  for (int which = 0; which < 2; ++which) {
    const char* src = (const char*)qkv + (size_t)which * H * head_dim * es;
    char* dst = (char*)dqkv + (size_t)which * H * head_dim * es;
    rc = launch_feature_map(src, dst, (size_t)N * 3 * H * head_dim, (size_t)head_dim, (size_t)3 * H * head_dim, nullptr,
                            (const float*)(ws + w.dphi) + which * plane, ld, (const float*)(ws + w.wt), B, N, H,
                            head_dim, M, kind, ERV_PREP_L2NORM, 1.f, dtype, true, st);
    if (rc) return rc;
  }
  return ERV_OK;
}
