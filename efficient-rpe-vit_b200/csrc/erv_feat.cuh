// Device building blocks shared by the feature-map kernels and the fused linear-attention kernels.
//
// Tiles (all fp32, shared memory):
//   token tiles   [TT][LDM]      LDM = DH + 4: DH data columns, column DH is an "extra" column
//                                (ones / dden / 0), columns DH+1..DH+3 are zero padding
//   feature tile  [TT][ldp]      ldp = Mp + 1 (odd stride: conflict-free when lanes walk tokens)
//   matrices      [Mp][LDM]      feature-major rows: S (col DH = z), dS (col DH = dz), W^T (col DH = 1 for FAVOR+)
// Three small-GEMM mappings cover every contraction of the path:
//   h1_rows   out[t][f]  = sum_j A[t][j] B[f][j]      thread owns feature f (row of B in registers), walks tokens
//   h1_accum  S[f][j]   += sum_t phi[t][f] V[t][j]    thread owns feature f (row of S in registers), walks tokens
//   h2_narrow R[t][j]    = sum_f A[t][f] B[f][j]      thread owns 4 tokens x 4 columns, features split over k-slices
#pragma once
#include "erv_common.cuh"

namespace erv {

struct FeatGeom {
  int M, Mp;     // features, features padded to a multiple of 32
  int FT, TS;    // h1 mapping: FT feature lanes x TS token slices
  int npass;     // ceil(Mp / FT)
  int ldp;       // feature tile row stride
  int nthreads;  // FT * TS
};

inline FeatGeom make_geom(int M) {
  FeatGeom g;
  g.M = M;
  g.Mp = (M + 31) / 32 * 32;
  g.npass = (g.Mp + 319) / 320;
  g.FT = 32 * (((g.Mp / 32) + g.npass - 1) / g.npass);
  g.TS = 256 / g.FT;
  if (g.TS < 1) g.TS = 1;
  g.ldp = g.Mp + 1;
  g.nthreads = g.FT * g.TS;
  return g;
}

// ---- tile movement -------------------------------------------------------------------------------
// rows n0..n0+TT-1 of a strided [N][DH] view -> token tile; column DH <- extra (0 for rows >= N)
template <typename T, int DH, int TT>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const T* __restrict__ base, size_t tok_stride,
                                          int n0, int N, float extra) {
  constexpr int LDM = DH + 4, V = DH / 4;
  for (int i = threadIdx.x; i < TT * (V + 1); i += blockDim.x) {
    int t = i / (V + 1), v = i % (V + 1), n = n0 + t;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < V) {
      if (n < N) x = ld4(base + (size_t)n * tok_stride + 4 * v);
    } else if (n < N) {
      x.x = extra;
    }
    st4(dst + t * LDM + 4 * v, x);
  }
}

template <typename T, int DH, int TT>
__device__ __forceinline__ void store_tile(T* __restrict__ base, size_t tok_stride, const float* __restrict__ src,
                                           int n0, int N) {
  constexpr int LDM = DH + 4, V = DH / 4;
  for (int i = threadIdx.x; i < TT * V; i += blockDim.x) {
    int t = i / V, v = i % V, n = n0 + t;
    if (n < N) st4(base + (size_t)n * tok_stride + 4 * v, ld4(src + t * LDM + 4 * v));
  }
}

// ---- prologue: rotation (RoPE / Circulant-STRING) then Dh^-1/4 scale or L2 normalisation -------------
struct RotArgs {
  int rot;          // ERV_ROT_*
  const float* ta;  // rope: cos [N][DH/2]; circulant: g [H][N][DH]
  const float* tb;  // rope: sin
};

// doubled circular-convolution table tile: g2[t][i] = g[h][n0+t][i mod DH], i < 2*DH  (identity for n >= N)
template <int DH, int TT>
__device__ __forceinline__ void load_g_tile(float* __restrict__ g2, const float* __restrict__ gtab, int h, int n0,
                                            int N) {
  for (int i = threadIdx.x; i < TT * 2 * DH; i += blockDim.x) {
    int t = i / (2 * DH), j = i % (2 * DH), n = n0 + t;
    int jm = j >= DH ? j - DH : j;
    g2[i] = (n < N) ? __ldg(gtab + ((size_t)h * N + n) * DH + jm) : (jm == 0 ? 1.f : 0.f);
  }
}

// xs = prep(rot(xr)).  Caller syncs before (xr/g2 ready) and after.  inv_s[t] (L2NORM only) = 1/||x_t||.
// Rows n >= N and the padding columns come out as zero.
template <int DH, int TT>
__device__ __forceinline__ void prep_tile(float* __restrict__ xs, const float* __restrict__ xr,
                                          const float* __restrict__ g2, float* __restrict__ inv_s, const RotArgs& ra,
                                          int prep, float prescale, int n0, int N) {
  constexpr int LDM = DH + 4;
  for (int i = threadIdx.x; i < TT * LDM; i += blockDim.x) {
    int t = i / LDM, a = i % LDM, n = n0 + t;
    float y = 0.f;
    if (a < DH && n < N) {
      const float* x = xr + t * LDM;
      if (ra.rot == ERV_ROT_ROPE) {
        int m = a >> 1;
        float c = __ldg(ra.ta + (size_t)n * (DH / 2) + m), s = __ldg(ra.tb + (size_t)n * (DH / 2) + m);
        float xe = x[2 * m], xo = x[2 * m + 1];
        y = (a & 1) ? (xe * s + xo * c) : (xe * c - xo * s);
      } else if (ra.rot == ERV_ROT_CIRCULANT) {
        const float* g = g2 + t * 2 * DH + DH + a;  // g[(a-b) mod DH] = g2[DH + a - b]
#pragma unroll 8
        for (int b = 0; b < DH; ++b) y += g[-b] * x[b];
      } else {
        y = x[a];
      }
      if (prep == ERV_PREP_SCALE) {
        y *= prescale;
      } else if (prep == ERV_PREP_L2NORM) {
        float ss = 0.f;
#pragma unroll 8
        for (int b = 0; b < DH; ++b) ss += x[b] * x[b];
        float inv = 1.0f / sqrtf(ss);  // no epsilon: favor_plus.py:200-201
        y *= inv;
        if (a == 0) inv_s[t] = inv;
      }
    }
    xs[i] = y;
  }
}

// Backward of prep_tile.  dxs = grad wrt xs ([TT][LDM]); the grad wrt the raw tile goes to the strided
// global view `dst`.  tmp is a scratch token tile.  Circulant: the table gradient
// dg[n][m] += sum_a dy[a] xr[(a-m) mod DH] is accumulated into this CTA's private slot dg_slot [N][DH].
// Contains one __syncthreads(); the caller syncs before (inputs ready) and before reusing tmp/xr.
template <typename T, int DH, int TT>
__device__ __forceinline__ void prep_tile_bwd(const float* __restrict__ dxs, const float* __restrict__ xs,
                                              const float* __restrict__ xr, const float* __restrict__ g2,
                                              const float* __restrict__ inv_s, float* __restrict__ tmp,
                                              const RotArgs& ra, int prep, float prescale, T* __restrict__ dst,
                                              size_t tok_stride, float* __restrict__ dg_slot, int n0, int N) {
  constexpr int LDM = DH + 4, V = DH / 4;
  // 1) dy = gradient wrt the rotated, un-scaled vector
  for (int i = threadIdx.x; i < TT * DH; i += blockDim.x) {
    int t = i / DH, a = i % DH;
    float d = dxs[t * LDM + a];
    if (prep == ERV_PREP_SCALE) {
      d *= prescale;
    } else if (prep == ERV_PREP_L2NORM) {
      float dot = 0.f;
#pragma unroll 8
      for (int b = 0; b < DH; ++b) dot += xs[t * LDM + b] * dxs[t * LDM + b];
      d = (d - xs[t * LDM + a] * dot) * inv_s[t];
    }
    tmp[t * LDM + a] = d;
  }
  __syncthreads();
  // 2) transpose of the rotation, straight to global memory
  for (int i = threadIdx.x; i < TT * V; i += blockDim.x) {
    int t = i / V, v = i % V, n = n0 + t;
    if (n >= N) continue;
    const float* dy = tmp + t * LDM;
    float r[4];
    if (ra.rot == ERV_ROT_ROPE) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        int m = 2 * v + p;
        float c = __ldg(ra.ta + (size_t)n * (DH / 2) + m), s = __ldg(ra.tb + (size_t)n * (DH / 2) + m);
        float d_even = dy[2 * m], d_odd = dy[2 * m + 1];
        r[2 * p] = d_even * c + d_odd * s;
        r[2 * p + 1] = d_odd * c - d_even * s;
      }
    } else if (ra.rot == ERV_ROT_CIRCULANT) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        int b = 4 * v + p;
        const float* g = g2 + t * 2 * DH + DH - b;  // g[(a-b) mod DH] = g2[DH + a - b]
        float acc = 0.f;
#pragma unroll 8
        for (int a = 0; a < DH; ++a) acc += g[a] * dy[a];
        r[p] = acc;
      }
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) r[p] = dy[4 * v + p];
    }
    st4(dst + (size_t)n * tok_stride + 4 * v, make_float4(r[0], r[1], r[2], r[3]));
  }
  if (ra.rot == ERV_ROT_CIRCULANT && dg_slot != nullptr) {
    for (int i = threadIdx.x; i < TT * DH; i += blockDim.x) {
      int t = i / DH, m = i % DH, n = n0 + t;
      if (n < 1 || n >= N) continue;  // CLS is not rotated
      const float* dy = tmp + t * LDM;
      const float* x = xr + t * LDM;
      float acc = 0.f;
      for (int a = 0; a < DH; ++a) {
        int j = a - m;
        if (j < 0) j += DH;
        acc += dy[a] * x[j];
      }
      dg_slot[(size_t)n * DH + m] += acc;  // slot is private to this CTA: plain read-modify-write
    }
  }
}

// ---- small GEMMs ---------------------------------------------------------------------------------
// out[t][f] = sum_{j<KD} A[t][j] * B[f][j];  A: token tile (smem), B: rows of length >= KD with stride ldb
// (smem or global).  epi(t, f, acc) consumes the result.  No barriers inside.
template <int KD, int LDA, int TT, class Epi>
__device__ __forceinline__ void h1_rows(const float* __restrict__ A, const float* __restrict__ B, int ldb,
                                        const FeatGeom& g, Epi epi) {
  static_assert(KD % 4 == 0, "KD must be a multiple of 4");
  const int fl = threadIdx.x % g.FT, ts = threadIdx.x / g.FT;
  for (int pass = 0; pass < g.npass; ++pass) {
    const int f = pass * g.FT + fl;
    if (f >= g.Mp) continue;
    float w[KD];
#pragma unroll
    for (int j = 0; j < KD; j += 4) {
      float4 v = ld4(B + (size_t)f * ldb + j);
      w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
    }
#pragma unroll 2
    for (int t = ts; t < TT; t += g.TS) {
      const float4* a = reinterpret_cast<const float4*>(A + t * LDA);
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int j = 0; j < KD / 4; ++j) {
        float4 v = a[j];
        acc0 = fmaf(v.x, w[4 * j], acc0);
        acc1 = fmaf(v.y, w[4 * j + 1], acc1);
        acc0 = fmaf(v.z, w[4 * j + 2], acc0);
        acc1 = fmaf(v.w, w[4 * j + 3], acc1);
      }
      epi(t, f, acc0 + acc1);
    }
  }
}

// S[f][j] += sum_t phi[t][f] * V[t][j] for j <= DH (column DH of V is the "extra" column).
// Token slices commit one after another (deterministic).  Contains barriers; call from all threads.
template <int DH, int TT>
__device__ __forceinline__ void h1_accum(float* __restrict__ S, const float* __restrict__ phi,
                                         const float* __restrict__ Vt, const FeatGeom& g) {
  constexpr int LDM = DH + 4;
  const int fl = threadIdx.x % g.FT, ts = threadIdx.x / g.FT;
  for (int pass = 0; pass < g.npass; ++pass) {
    const int f = pass * g.FT + fl;
    const bool live = f < g.Mp;
    float acc[DH + 1];
#pragma unroll
    for (int j = 0; j <= DH; ++j) acc[j] = 0.f;
    if (live) {
#pragma unroll 2
      for (int t = ts; t < TT; t += g.TS) {
        const float p = phi[t * g.ldp + f];
        const float4* v = reinterpret_cast<const float4*>(Vt + t * LDM);
#pragma unroll
        for (int j = 0; j < DH / 4; ++j) {
          float4 x = v[j];
          acc[4 * j] = fmaf(p, x.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(p, x.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(p, x.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(p, x.w, acc[4 * j + 3]);
        }
        acc[DH] = fmaf(p, Vt[t * LDM + DH], acc[DH]);
      }
    }
    for (int s = 0; s < g.TS; ++s) {
      if (live && ts == s) {
        float* row = S + (size_t)f * LDM;
#pragma unroll
        for (int j = 0; j < DH / 4; ++j) {
          float4 x = ld4(row + 4 * j);
          x.x += acc[4 * j]; x.y += acc[4 * j + 1]; x.z += acc[4 * j + 2]; x.w += acc[4 * j + 3];
          st4(row + 4 * j, x);
        }
        row[DH] += acc[DH];
      }
      if (g.TS > 1) __syncthreads();
    }
  }
}

// R[t][j] = sum_{f<K} A[t][f] * B[f][j], j < 4*NJG.  A: feature tile (smem, stride lda), B: rows with stride ldb
// (smem or global, 16-byte aligned).  Partials go to red[ks][TT][LDM]; on return red[0] holds R.
// Contains barriers.  red must hold min(6, max(1, nthreads/items)) * TT * LDM floats.
template <int DH, int TT, int NJG>
__device__ __forceinline__ int h2_kslices(int nthreads) {
  constexpr int ITEMS = (TT / 4) * NJG;
  int ks = nthreads / ITEMS;
  return ks < 1 ? 1 : (ks > 6 ? 6 : ks);
}

template <int DH, int TT, int NJG>
__device__ __forceinline__ void h2_narrow(float* __restrict__ red, const float* __restrict__ A, int lda,
                                          const float* __restrict__ B, int ldb, int K) {
  constexpr int LDM = DH + 4, ITEMS = (TT / 4) * NJG;
  const int KS = h2_kslices<DH, TT, NJG>(blockDim.x);
  for (int u = threadIdx.x; u < ITEMS * KS; u += blockDim.x) {
    const int ks = u / ITEMS, it = u % ITEMS, tg = it / NJG, jg = it % NJG;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    const float* a0 = A + (size_t)(4 * tg) * lda;
#pragma unroll 2
    for (int f = ks; f < K; f += KS) {
      const float4 b = ld4(B + (size_t)f * ldb + 4 * jg);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = a0[r * lda + f];
        acc[r][0] = fmaf(a, b.x, acc[r][0]);
        acc[r][1] = fmaf(a, b.y, acc[r][1]);
        acc[r][2] = fmaf(a, b.z, acc[r][2]);
        acc[r][3] = fmaf(a, b.w, acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      st4(red + ((size_t)ks * TT + 4 * tg + r) * LDM + 4 * jg, make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
  }
  __syncthreads();
  if (KS > 1) {
    for (int i = threadIdx.x; i < TT * 4 * NJG; i += blockDim.x) {
      int t = i / (4 * NJG), j = i % (4 * NJG);
      float s = red[t * LDM + j];
      for (int ks = 1; ks < KS; ++ks) s += red[((size_t)ks * TT + t) * LDM + j];
      red[t * LDM + j] = s;
    }
    __syncthreads();
  }
}

// dst[t] = reduce over f < K of op(A[t][f], f); thread (t = tid % TT, q = tid / TT) scans a contiguous
// part of the features.  pm: scratch of (nthreads/TT)*TT floats.  Contains barriers.
template <int TT, bool IS_MAX, class Val>
__device__ __forceinline__ void row_reduce(float* __restrict__ dst, float* __restrict__ pm, int K, Val val) {
  const int Q = blockDim.x / TT;
  const int t = threadIdx.x % TT, q = threadIdx.x / TT;
  if (q < Q) {
    const int chunk = (K + Q - 1) / Q;
    const int f0 = q * chunk, f1 = min(K, f0 + chunk);
    float r = IS_MAX ? -INFINITY : 0.f;
    for (int f = f0; f < f1; ++f) {
      float v = val(t, f);
      r = IS_MAX ? fmaxf(r, v) : r + v;
    }
    pm[q * TT + t] = r;
  }
  __syncthreads();
  if (threadIdx.x < TT) {
    float r = pm[threadIdx.x];
    for (int qq = 1; qq < Q; ++qq) r = IS_MAX ? fmaxf(r, pm[qq * TT + threadIdx.x]) : r + pm[qq * TT + threadIdx.x];
    dst[threadIdx.x] = r;
  }
  __syncthreads();
}

// ---- the random-feature map on one tile ------------------------------------------------------------
// phi[t][f] for the prepared tile xs; favor_plus.py:112-140 / relu.py:116-138.  wt: W^T rows [Mp][LDM] (global).
// FAVOR+: phi = exp((P - max_f P) - |x|^2/2) / sqrt(M), max over the M real features (stored in m_s).
// Contains barriers.  Padded features (f >= M) are written as 0.
template <int DH, int TT>
__device__ __forceinline__ void feature_tile(float* __restrict__ phi, const float* __restrict__ xs,
                                             const float* __restrict__ wt, float* __restrict__ m_s,
                                             float* __restrict__ n2_s, float* __restrict__ pm, const FeatGeom& g,
                                             int kind, float inv_sqrt_m) {
  constexpr int LDM = DH + 4;
  const int ldp = g.ldp, M = g.M;
  if (kind == ERV_FEAT_RELU) {
    h1_rows<DH, LDM, TT>(xs, wt, LDM, g, [&](int t, int f, float acc) {
      phi[t * ldp + f] = (f < M) ? fmaxf(acc, 0.f) * inv_sqrt_m : 0.f;
    });
    __syncthreads();
    return;
  }
  h1_rows<DH, LDM, TT>(xs, wt, LDM, g, [&](int t, int f, float acc) { phi[t * ldp + f] = acc; });
  if (threadIdx.x < TT) {
    const float* x = xs + threadIdx.x * LDM;
    float ss = 0.f;
#pragma unroll 8
    for (int b = 0; b < DH; ++b) ss += x[b] * x[b];
    n2_s[threadIdx.x] = ss / 2.0f;
  }
  __syncthreads();
  row_reduce<TT, true>(m_s, pm, M, [&](int t, int f) { return phi[t * ldp + f]; });
  // exp sweep in the h1 mapping (each thread rewrites the elements it produced)
  {
    const int fl = threadIdx.x % g.FT, ts = threadIdx.x / g.FT;
    for (int pass = 0; pass < g.npass; ++pass) {
      const int f = pass * g.FT + fl;
      if (f >= g.Mp) continue;
      for (int t = ts; t < TT; t += g.TS) {
        float p = phi[t * ldp + f];
        phi[t * ldp + f] = (f < M) ? expf((p - m_s[t]) - n2_s[t]) * inv_sqrt_m : 0.f;
      }
    }
  }
  __syncthreads();
}

// G = d phi/d P applied to dphi, in place over the phi tile: FAVOR+ G = dphi * phi; ReLU G = dphi/sqrt(M) where phi > 0
__device__ __forceinline__ float feature_grad(float dphi, float phi, int kind, float inv_sqrt_m) {
  return kind == ERV_FEAT_FAVOR ? dphi * phi : (phi > 0.f ? dphi * inv_sqrt_m : 0.f);
}

}  // namespace erv
