// KERPLE forward as the reference computes it (models/rpe/kerple.py:150-344, models/rpe/fft_utils.py:119-170,
// models/attention/favor_plus.py:221-260): D1 = C (phi_k (x) v), D2 = C phi_k by FFT along the patch axis, fused with the
// read-out num = phi_q . D1, den = phi_q . D2, out = num / (den + 1e-6).  C[i][j] = exp(b[j - i]) is Toeplitz.
//
// Route for long sequences (2048 < N - 1 <= 4096; M > 64 or at least 16 (batch, head) pairs), chosen per shape against the
// tensor-core tile route (erv_ktile_tc.cu) by measurement (profiles/r02_kerple_fft_vs_tile.md).  One 8192-point complex FFT
// lives entirely in the registers of a 512-thread CTA (16 points per thread) and crosses shared memory twice:
//
//   * the CLS token is split off, so the 4096 patches need a circular length of 2*4096 - 1 <= 8192 (a power of two);
//     its row and column are rank-1 terms added by the finalize kernels;
//   * two real columns per complex transform: for the feature pair (m, m+1) and value column d the sequence is
//     z[j] = v[j][d] (phi_k[j][m] + i phi_k[j][m+1]); C is real, so Re / Im of the filtered sequence are the two columns
//     of D1 (d = Dh is the all-ones column: D2);
//   * 8192 = 16 x 16 x 32 as a self-sorting three-phase transform: thread t holds x[t + 512 r], 16-point DFT over r,
//     twiddle, exchange, 16-point DFT, twiddle, exchange, 32-point DFT split over two threads; the result lands as
//     X[t + 512 q] — the same register layout as the input, so the inverse is the same routine applied to conj(X G) and
//     nothing is ever bit-reversed;
//   * the (m, d) column never leaves the CTA: after the inverse the thread holds D[i][m][d], D[i][m+1][d] for its 8 patches
//     i and accumulates phi_q[i][m] D + phi_q[i][m+1] D' in registers over the CTA's feature pairs.  D1 / D2
//     ([B, H, N, M, Dh] in the reference: 49 MB per (batch, head) at N = 4097, M = 44) are never materialised.
//
// A CTA owns (batch*head, value column d, chunk of feature pairs); chunk partials are summed in a fixed order by
// kfft_finalize_kernel, which also adds the CLS column and divides; kfft_cls_kernel computes the CLS row in 32 partial sums.
// phi_q / phi_k come from the feature-map kernels (fp32 rows in the workspace) and are re-laid out feature-pair-major by
// kfft_transpose_kernel, so that a CTA stages the 32 KB it needs per feature pair with coalesced 16-byte cp.async; the
// 64 KB of filter coefficients of the head stay in shared memory.  A numpy model of the algorithm: tests/test_kfft_model.py.
#include <math.h>

#include "erv_common.cuh"
#include "erv_umma.cuh"

namespace erv {
namespace kfft {
using umma::mbar_init;
using umma::mbar_init_fence;
using umma::mbar_wait;
using umma::smem_u32;

constexpr int L = 8192;        // transform length
constexpr int NT = 512;        // threads per CTA: 16 points each
constexpr int NP_MAX = 4096;   // patches per sequence this length serves
constexpr int kClsSeg = 32;    // CTAs per (batch, head) pair that share the CLS row
constexpr int XROW = 33;       // padded row of the second exchange (32 points + 1)
constexpr size_t kSmem = (size_t)256 * XROW * sizeof(float2);  // 67 584 B >= 8192 points
constexpr size_t kSmemFwd = kSmem + 2 * (size_t)NP_MAX * sizeof(float2) + (size_t)L * sizeof(float2);  // + phi_k / phi_q stages + G

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// 4-point forward DFT (w = -i) of (a, b, c, d) in place
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
  a = cadd(s0, s2);
  c = csub(s0, s2);
  b = make_float2(s1.x + s3.y, s1.y - s3.x);  // s1 - i s3
  d = make_float2(s1.x - s3.y, s1.y + s3.x);  // s1 + i s3
}

// 16-point forward DFT in registers, natural order in and out: X[k1 + 4 k2] = sum_n2 w4^(n2 k2) w16^(n2 k1) sum_n1 x[n2 + 4 n1] w4^(n1 k1)
__device__ __forceinline__ void dft16(float2 (&x)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(x[n2], x[n2 + 4], x[n2 + 8], x[n2 + 12]);  // x[n2 + 4 k1]
  // twiddles w16^(n2 k1), w16 = exp(-2 pi i / 16)
  x[5] = cmul(x[5], make_float2(C1, -S1));    // n2 = 1, k1 = 1
  x[9] = cmul(x[9], make_float2(R, -R));      // 1, 2
  x[13] = cmul(x[13], make_float2(S1, -C1));  // 1, 3
  x[6] = cmul(x[6], make_float2(R, -R));      // 2, 1
  x[10] = make_float2(x[10].y, -x[10].x);     // 2, 2: -i
  x[14] = cmul(x[14], make_float2(-R, -R));   // 2, 3
  x[7] = cmul(x[7], make_float2(S1, -C1));    // 3, 1
  x[11] = cmul(x[11], make_float2(-R, -R));   // 3, 2
  x[15] = cmul(x[15], make_float2(-C1, S1));  // 3, 3: w16^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);  // -> X[k1 + 4 k2] at [4 k1 + k2]
  // [4 k1 + k2] -> [k1 + 4 k2]: transpose of the 4 x 4 register tile
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const float2 t = x[4 * a + b];
      x[4 * a + b] = x[4 * b + a];
      x[4 * b + a] = t;
    }
}

// x[k] *= w^k, k = 1..15, from w, w^2, w^4, w^8 (products at most four deep)
__device__ __forceinline__ void twiddle16(float2 (&x)[16], float2 w1, float2 w2, float2 w4, float2 w8) {
  const float2 w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  x[1] = cmul(x[1], w1); x[2] = cmul(x[2], w2); x[3] = cmul(x[3], w3); x[4] = cmul(x[4], w4);
  x[5] = cmul(x[5], w5); x[6] = cmul(x[6], w6); x[7] = cmul(x[7], w7); x[8] = cmul(x[8], w8);
  x[9] = cmul(x[9], cmul(w8, w1)); x[10] = cmul(x[10], cmul(w8, w2)); x[11] = cmul(x[11], cmul(w8, w3));
  x[12] = cmul(x[12], cmul(w8, w4)); x[13] = cmul(x[13], cmul(w8, w5)); x[14] = cmul(x[14], cmul(w8, w6));
  x[15] = cmul(x[15], cmul(w8, w7));
}

// cos / sin of pi u / 16: w_32^u = cos - i sin
__device__ constexpr float kCos32[16] = {1.f, 0.98078528040323044f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                                         0.55557023301960222f, 0.38268343236508977f, 0.19509032201612827f, 0.f, -0.19509032201612827f,
                                         -0.38268343236508977f, -0.55557023301960222f, -0.70710678118654752f, -0.83146961230254524f,
                                         -0.92387953251128674f, -0.98078528040323044f};
__device__ constexpr float kSin32[16] = {0.f, 0.19509032201612827f, 0.38268343236508977f, 0.55557023301960222f, 0.70710678118654752f,
                                         0.83146961230254524f, 0.92387953251128674f, 0.98078528040323044f, 1.f, 0.98078528040323044f,
                                         0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960222f,
                                         0.38268343236508977f, 0.19509032201612827f};
__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct Twiddles {  // per-thread bases: phase A w_L^(t p), phase B w_512^(lane p), p = 1, 2, 4, 8
  float2 a1, a2, a4, a8, b1, b2, b4, b8;
};
__device__ __forceinline__ float2 unit(int num, int den) {  // exp(-2 pi i num / den)
  float s, c;
  sincospif(-2.f * (float)num / (float)den, &s, &c);
  return make_float2(c, s);
}
__device__ __forceinline__ Twiddles make_twiddles() {
  const int t = threadIdx.x, lane = t & 31;
  Twiddles w;
  w.a1 = unit(t, L); w.a2 = unit(2 * t, L); w.a4 = unit(4 * t, L); w.a8 = unit(8 * t, L);
  w.b1 = unit(lane, 512); w.b2 = unit(2 * lane, 512); w.b4 = unit(4 * lane, 512); w.b8 = unit(8 * lane, 512);
  return w;
}

// Forward DFT of 8192 points.  in: x[r] = element t + 512 r of the sequence; out: x[q] = X[t + 512 q].  `sm` holds
// kSmem bytes; the routine synchronises the CTA before each of its shared-memory writes and reads.
__device__ __forceinline__ void fft8192(float2 (&x)[16], float2* sm, const Twiddles& w) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // phase A: n = t + 512 r, k = k1 + 16 k2: y_k1[t] = w_L^(t k1) DFT16_r(x)[k1]
  dft16(x);
  twiddle16(x, w.a1, w.a2, w.a4, w.a8);
  __syncthreads();
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) sm[k1 * 512 + t] = x[k1];
  __syncthreads();
  // phase B: thread (k1 = warp, m2 = lane): 512-point DFT of y_k1 with n2 = m2 + 32 m1, k2 = j1 + 16 j2
#pragma unroll
  for (int m1 = 0; m1 < 16; ++m1) x[m1] = sm[warp * 512 + 32 * m1 + lane];
  dft16(x);
  twiddle16(x, w.b1, w.b2, w.b4, w.b8);
  __syncthreads();
#pragma unroll
  for (int j1 = 0; j1 < 16; ++j1) sm[(j1 * 16 + warp) * XROW + lane] = x[j1];
  __syncthreads();
  // phase C: thread (k1 = t & 15, j1 = (t >> 4) & 15, half = t >> 8): 32-point DFT over m2, outputs j2 = 2 q + half
  {
    const float2* row = sm + (((t >> 4) & 15) * 16 + (t & 15)) * XROW;
    const bool odd = t >= 256;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float2 a = row[u], b = row[u + 16];
      if (odd) {
        x[u] = cmul(csub(a, b), make_float2(kCos32[u], -kSin32[u]));  // (a - b) w_32^u
      } else {
        x[u] = cadd(a, b);
      }
    }
  }
  dft16(x);  // X[k1 + 16 j1 + 256 (2 q + half)] = X[t + 512 q]
}

struct FftArgs {
  const void* qkv;      // [B, N, 3, H, DH]
  const float* phi_q;   // [B*H][N][ld]
  const float* phi_k;
  const float2* coef;   // [H][L]: DFT of the circulant first column / L
  const float2* phiT_q; // [B*H][ld/2][NPp]: patch rows of phi_q, feature-pair-major ((m, m+1) of one patch = one float2)
  const float2* phiT_k;
  float* part;          // [chunks][B*H][DH+1][NP]: chunk partials of num (d < DH) and den (d = DH)
  float* cls_part;      // [B*H][kClsSeg][DH+1]: partial sums of the CLS row
  int B, N, H, DH, M, ld, NP, NPp, chunks, fp_per_chunk;
};

// G[h][k] = DFT(g)[k] / L, g[s] = c[-s] (s = 0..NP-1), g[L - s] = c[s] (s = 1..NP-1): y = g (*) x is y[i] = sum_j c[j - i] x[j]
__global__ void __launch_bounds__(NT, 1) kfft_coef_kernel(const float* __restrict__ cexp, float2* __restrict__ coef, int N) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* sm = reinterpret_cast<float2*>(smem_raw);
  const int h = blockIdx.x, t = threadIdx.x, NP = N - 1;
  const float* c = cexp + (size_t)h * (2 * N - 1) + (N - 1);  // c[delta], delta = j - i
  const Twiddles w = make_twiddles();
  float2 x[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int s = t + 512 * r;
    float v = 0.f;
    if (s < NP) v = __ldg(c - s);
    else if (L - s >= 1 && L - s < NP) v = __ldg(c + (L - s));
    x[r] = make_float2(v, 0.f);
  }
  fft8192(x, sm, w);
#pragma unroll
  for (int q = 0; q < 16; ++q) coef[(size_t)h * L + t + 512 * q] = make_float2(x[q].x * (1.f / L), x[q].y * (1.f / L));
}

// phiT[plane][m / 2][i] = (phi[plane][i + 1][m], phi[plane][i + 1][m + 1]) for the patches i = 0..NP-1 (rows padded to NPp with
// zeros): 32-token x 32-feature tiles through shared memory, coalesced on both sides.  plane = (q | k, batch*head).
__global__ void __launch_bounds__(256) kfft_transpose_kernel(const float* __restrict__ phi, float2* __restrict__ phiT, int N, int ld,
                                                             int NPp) {
  __shared__ float tile[32][33];
  const int plane = blockIdx.z, i0 = blockIdx.x * 32, m0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = phi + (size_t)plane * N * ld;
  float2* dst = phiT + (size_t)plane * (ld / 2) * NPp;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r, m = m0 + tx;
    tile[ty + 8 * r][tx] = (i < N - 1 && m < ld) ? src[(size_t)(i + 1) * ld + m] : 0.f;
  }
  __syncthreads();
  // thread (ty, tx): feature pair m0/2 + ty + 8 r' (16 pairs per tile: two passes), token i0 + tx
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int fpl = ty + 8 * r, m = m0 + 2 * fpl, i = i0 + tx;
    if (m < ld && i < NPp) dst[(size_t)(m / 2) * NPp + i] = make_float2(tile[tx][2 * fpl], tile[tx][2 * fpl + 1]);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT, 1) kfft_fwd_kernel(const FftArgs p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* sm = reinterpret_cast<float2*>(smem_raw);
  const int t = threadIdx.x;
  const int chunk = blockIdx.x, d = blockIdx.y, pair = blockIdx.z;
  const int b = pair / p.H, h = pair % p.H;
  const int N = p.N, NP = p.NP, ld = p.ld;
  const Twiddles w = make_twiddles();
  const float2* G = p.coef + (size_t)h * L;

  float vv[8], acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = t + 512 * r;  // patch index; token i + 1
    acc[r] = 0.f;
    vv[r] = 0.f;
    if (i < NP) vv[r] = d < p.DH ? to_f(static_cast<const T*>(p.qkv)[qkv_off(b, i + 1, 2, h, N, p.H, p.DH) + d]) : 1.f;
  }
  // phi_k[., m..m+1] of the next feature pair, phi_q[., m..m+1] of the current one and the filter coefficients of this head are
  // brought into shared memory by the TMA unit as bulk copies (cp.async.bulk, one instruction per 32 / 64 KB row, completion on
  // an mbarrier) from the feature-pair-major copies (kfft_transpose_kernel) while the transforms run.  (First version: 8-byte
  // strided cp.async from the token-major rows and 16 __ldg of G per transform -- 38 % of the warp-stall samples were
  // long_scoreboard on exactly those two lines; second version: per-thread 16-byte cp.async + two CTA barriers per feature pair.)
  float2* stage_k = sm + 256 * XROW;
  float2* stage_q = stage_k + NP_MAX;
  float2* Gs = stage_q + NP_MAX;
  __shared__ __align__(8) uint64_t bar_k, bar_q, bar_g;
  const int NPp = p.NPp;
  const uint32_t row_bytes = (uint32_t)NPp * sizeof(float2);  // NPp is even: a multiple of 16 bytes
  auto bulk = [&](float2* dst, const float2* src, uint32_t bytes, uint64_t* bar) {  // one thread
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
  };
  const int fp0 = chunk * p.fp_per_chunk, fp1 = min(fp0 + p.fp_per_chunk, (p.M + 1) / 2);
  const float2* tq = p.phiT_q + (size_t)pair * (ld / 2) * NPp;
  const float2* tk = p.phiT_k + (size_t)pair * (ld / 2) * NPp;
  if (t == 0) {
    mbar_init(&bar_k, 1);
    mbar_init(&bar_q, 1);
    mbar_init(&bar_g, 4);
    mbar_init_fence();
  }
  __syncthreads();
  if (t == 0) {
    for (int c = 0; c < 4; ++c)  // 4 x 16 KB: a single 64 KB cp.async.bulk faults (illegal address) on this driver, 32 KB rows are fine
      bulk(Gs + c * (L / 4), G + c * (L / 4), (uint32_t)(L / 4 * sizeof(float2)), &bar_g);
    if (fp0 < fp1) bulk(stage_k, tk + (size_t)fp0 * NPp, row_bytes, &bar_k);
  }
  uint32_t ph_k = 0, ph_q = 0;
  if (fp0 < fp1) mbar_wait(&bar_g, 0);
  for (int fp = fp0; fp < fp1; ++fp) {
    const bool two = 2 * fp + 1 < p.M;
    float2 x[16];
    mbar_wait(&bar_k, ph_k);  // phi_k of this feature pair has landed
    ph_k ^= 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = t + 512 * r;
      float2 f = make_float2(0.f, 0.f);
      if (i < NP) f = stage_k[i];
      if (!two) f.y = 0.f;
      x[r] = make_float2(vv[r] * f.x, vv[r] * f.y);
      x[r + 8] = make_float2(0.f, 0.f);
    }
    fft8192(x, sm, w);
    // every thread is past its stage_k / stage_q reads (the transform synchronises the CTA): refill both
    if (t == 0) {
      if (fp + 1 < fp1) bulk(stage_k, tk + (size_t)(fp + 1) * NPp, row_bytes, &bar_k);
      bulk(stage_q, tq + (size_t)fp * NPp, row_bytes, &bar_q);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float2 y = cmul(x[q], Gs[t + 512 * q]);
      x[q] = make_float2(y.x, -y.y);  // inverse transform as conj(DFT(conj(.)))
    }
    fft8192(x, sm, w);
    mbar_wait(&bar_q, ph_q);  // phi_q of this feature pair
    ph_q ^= 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = t + 512 * r;
      if (i < NP) {
        float2 f = stage_q[i];
        if (!two) f.y = 0.f;
        acc[r] = fmaf(f.x, x[r].x, acc[r]);
        acc[r] = fmaf(-f.y, x[r].y, acc[r]);
      }
    }
  }
  float* dst = p.part + (((size_t)chunk * p.B * p.H + pair) * (p.DH + 1) + d) * NP;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = t + 512 * r;
    if (i < NP) dst[i] = acc[r];
  }
}

// patches i >= 1: sum the chunk partials in a fixed order, add the CLS column c[-i] (phi_q[i] . phi_k[0]) [v_0 | 1], divide;
// i = 0 (CLS row): sum the kClsSeg partials of kfft_cls_kernel in a fixed order, divide
template <typename T, int DH>
__global__ void __launch_bounds__(128) kfft_finalize_kernel(const FftArgs p, const float* __restrict__ cexp, T* __restrict__ out,
                                                            float* __restrict__ den_out) {
  const int pair = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
  const int b = pair / p.H, h = pair % p.H, N = p.N, NP = p.NP;
  if (i >= N) return;
  float nd[DH + 1];
#pragma unroll
  for (int d = 0; d <= DH; ++d) nd[d] = 0.f;
  if (i == 0) {
    for (int sgm = 0; sgm < kClsSeg; ++sgm) {
      const float* src = p.cls_part + ((size_t)pair * kClsSeg + sgm) * (DH + 1);
#pragma unroll
      for (int d = 0; d <= DH; ++d) nd[d] += src[d];
    }
  } else {
    for (int c = 0; c < p.chunks; ++c) {
      const float* src = p.part + (((size_t)c * p.B * p.H + pair) * (DH + 1)) * NP + (i - 1);
#pragma unroll
      for (int d = 0; d <= DH; ++d) nd[d] += src[(size_t)d * NP];
    }
    const float* q = p.phi_q + ((size_t)pair * N + i) * p.ld;
    const float* k0 = p.phi_k + (size_t)pair * N * p.ld;
    float s = 0.f;
    for (int m = 0; m < p.M; ++m) s = fmaf(q[m], __ldg(k0 + m), s);
    s *= __ldg(cexp + (size_t)h * (2 * N - 1) + (N - 1) - i);  // c[0 - i]
    const T* v0 = static_cast<const T*>(p.qkv) + qkv_off(b, 0, 2, h, N, p.H, DH);
#pragma unroll
    for (int d = 0; d < DH; ++d) nd[d] = fmaf(s, to_f(v0[d]), nd[d]);
    nd[DH] += s;
  }
  const float inv = 1.f / (nd[DH] + kEps);
  T* o = out + out_off(b, i, h, N, p.H, DH);
#pragma unroll
  for (int d = 0; d < DH; d += 4) st4(o + d, make_float4(nd[d] * inv, nd[d + 1] * inv, nd[d + 2] * inv, nd[d + 3] * inv));
  den_out[(size_t)pair * N + i] = nd[DH];
}

// CLS row: num[0] = sum_j c[j] (phi_q[0] . phi_k[j]) [v_j | 1].  kClsSeg CTAs per (batch, head), each a strided share of the
// keys; deterministic tree sums, the kClsSeg partials are combined by kfft_finalize_kernel.  (One CTA per pair took 98 us.)
template <typename T, int DH>
__global__ void __launch_bounds__(256) kfft_cls_kernel(const FftArgs p, const float* __restrict__ cexp) {
  __shared__ float q0[320];
  __shared__ float red[8][DH + 1];
  const int sgm = blockIdx.x, pair = blockIdx.y, b = pair / p.H, h = pair % p.H, N = p.N, t = threadIdx.x;
  for (int m = t; m < p.ld; m += 256) q0[m] = m < p.M ? p.phi_q[(size_t)pair * N * p.ld + m] : 0.f;
  __syncthreads();
  float nd[DH + 1];
#pragma unroll
  for (int d = 0; d <= DH; ++d) nd[d] = 0.f;
  for (int j = sgm * 256 + t; j < N; j += kClsSeg * 256) {
    const float* k = p.phi_k + ((size_t)pair * N + j) * p.ld;
    float s0 = 0.f, s1 = 0.f;
    for (int m = 0; m < p.ld; m += 4) {  // ld is a multiple of 8; columns >= M are masked on both sides
      const float4 kv = ld4(k + m);
      s0 = fmaf(q0[m], m < p.M ? kv.x : 0.f, s0); s1 = fmaf(q0[m + 1], m + 1 < p.M ? kv.y : 0.f, s1);
      s0 = fmaf(q0[m + 2], m + 2 < p.M ? kv.z : 0.f, s0); s1 = fmaf(q0[m + 3], m + 3 < p.M ? kv.w : 0.f, s1);
    }
    const float s = (s0 + s1) * __ldg(cexp + (size_t)h * (2 * N - 1) + (N - 1) + j);  // c[j - 0]
    const T* v = static_cast<const T*>(p.qkv) + qkv_off(b, j, 2, h, N, p.H, DH);
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 vv = ld4(v + d);
      nd[d] = fmaf(s, vv.x, nd[d]); nd[d + 1] = fmaf(s, vv.y, nd[d + 1]);
      nd[d + 2] = fmaf(s, vv.z, nd[d + 2]); nd[d + 3] = fmaf(s, vv.w, nd[d + 3]);
    }
    nd[DH] += s;
  }
#pragma unroll
  for (int d = 0; d <= DH; ++d) {
    const float v = warp_sum(nd[d]);
    if ((t & 31) == 0) red[t >> 5][d] = v;
  }
  __syncthreads();
  if (t <= DH) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv][t];
    p.cls_part[((size_t)pair * kClsSeg + sgm) * (DH + 1) + t] = s;
  }
}

}  // namespace kfft

// ---- host side ---------------------------------------------------------------------------------------------------------
static std::atomic<int> g_fft_mode{-1};  // -1: environment default, 0 / 1: set by erv_kerple_set_fft()
bool kerple_fft_eligible(int B, int N, int H, int DH, int M) {
  static const int env_mode = [] {
    const char* e = getenv("ERV_KERPLE_FFT");  // 0: never, 1: whenever the length fits, unset: measured crossover (below)
    return e ? atoi(e) : -1;
  }();
  const int set = g_fft_mode.load(std::memory_order_relaxed), mode = set >= 0 ? set : env_mode;
  if (mode == 0) return false;
  const bool fits = N >= 3 && N - 1 <= kfft::NP_MAX && (DH == 8 || DH == 16 || DH == 32 || DH == 64);
  // measured on B200, forward per (batch, head) pair (profiles/r02_kerple_fft_vs_tile.md): at N = 4097 the FFT route takes
  // 156 us for M = 256 against 510 us on the tile route (CUDA cores for M > 64), and for M = 44 34.2 / 24.8 us at 16 / 64
  // pairs against 40.8 / 35.8 us on the tcgen05 tiles; with 4 pairs (73 vs 53 us) and at N <= 2049 (27 vs 14 us) the tiles win
  return mode == 1 ? fits : (fits && N - 1 > 2048 && (M > 64 || B * H >= 16));
}

static int kfft_chunks(int B, int H, int DH, int M, int* fp_per_chunk) {
  const int nfp = (M + 1) / 2, ctas = B * H * (DH + 1);
  int chunks = (4 * kNumSMs + ctas - 1) / ctas;  // about four waves of CTAs
  chunks = chunks < 1 ? 1 : (chunks > nfp ? nfp : chunks);
  const int per = (nfp + chunks - 1) / chunks;
  *fp_per_chunk = per;
  return (nfp + per - 1) / per;
}

static size_t kfft_ld(int M) { return (size_t)(M + 7) / 8 * 8; }  // kerple_ldphi (erv_tileattn.cu)
static size_t kfft_npp(int N) { return (size_t)(N - 1 + 1) / 2 * 2; }
static size_t kfft_phit_bytes(int B, int N, int H, int M) {  // phi_q and phi_k planes, feature-pair-major
  return align_up((size_t)2 * B * H * (kfft_ld(M) / 2) * kfft_npp(N) * sizeof(float2), 256);
}
static size_t kfft_cls_bytes(int B, int H, int DH) { return align_up((size_t)B * H * kfft::kClsSeg * (DH + 1) * sizeof(float), 256); }

size_t kerple_fft_ws_bytes(int B, int N, int H, int DH, int M) {
  int per;
  const int chunks = kfft_chunks(B, H, DH, M, &per);
  return align_up((size_t)H * kfft::L * sizeof(float2), 256) +
         align_up((size_t)chunks * B * H * (DH + 1) * (N - 1) * sizeof(float), 256) + kfft_phit_bytes(B, N, H, M) +
         kfft_cls_bytes(B, H, DH);
}

template <typename T, int DH>
static int kfft_tail(const kfft::FftArgs& a, const float* cexp, void* out, float* den, cudaStream_t st) {
  kfft::kfft_cls_kernel<T, DH><<<dim3(kfft::kClsSeg, a.B * a.H), 256, 0, st>>>(a, cexp);
  ERV_LAUNCH_CHECK();
  kfft::kfft_finalize_kernel<T, DH><<<dim3((a.N + 127) / 128, a.B * a.H), 128, 0, st>>>(a, cexp, static_cast<T*>(out), den);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

// phi_q / phi_k: [B*H][N][ld] fp32 (kerple_features), cexp: [H][2N-1], ws: kerple_fft_ws_bytes
int kerple_fft_forward(const void* qkv, void* out, float* den, const float* phi_q, const float* phi_k, int ld,
                       const float* cexp, void* ws, int B, int N, int H, int DH, int M, int dtype, cudaStream_t st) {
  kfft::FftArgs a{};
  a.qkv = qkv; a.phi_q = phi_q; a.phi_k = phi_k;
  a.coef = reinterpret_cast<const float2*>(ws);
  a.part = reinterpret_cast<float*>((char*)ws + align_up((size_t)H * kfft::L * sizeof(float2), 256));
  a.B = B; a.N = N; a.H = H; a.DH = DH; a.M = M; a.ld = ld; a.NP = N - 1; a.NPp = (int)kfft_npp(N);
  a.chunks = kfft_chunks(B, H, DH, M, &a.fp_per_chunk);
  ERV_CHECK_ARG((size_t)ld == kfft_ld(M) && phi_k == phi_q + (size_t)B * H * N * ld, "kerple_fft_forward: unexpected phi layout");
  float2* phit = reinterpret_cast<float2*>((char*)a.part + align_up((size_t)a.chunks * B * H * (DH + 1) * (N - 1) * sizeof(float), 256));
  a.phiT_q = phit; a.phiT_k = phit + (size_t)B * H * (ld / 2) * a.NPp;
  a.cls_part = reinterpret_cast<float*>((char*)phit + kfft_phit_bytes(B, N, H, M));
  kfft::kfft_transpose_kernel<<<dim3((a.NPp + 31) / 32, (ld + 31) / 32, 2 * B * H), 256, 0, st>>>(phi_q, phit, N, ld, a.NPp);
  ERV_LAUNCH_CHECK();
  ERV_CUDA(allow_smem(kfft::kfft_coef_kernel, kfft::kSmem));
  kfft::kfft_coef_kernel<<<H, kfft::NT, kfft::kSmem, st>>>(cexp, const_cast<float2*>(a.coef), N);
  ERV_LAUNCH_CHECK();
  const dim3 grid(a.chunks, DH + 1, B * H);
  if (dtype == ERV_F32) {
    ERV_CUDA(allow_smem(kfft::kfft_fwd_kernel<float>, kfft::kSmemFwd));
    kfft::kfft_fwd_kernel<float><<<grid, kfft::NT, kfft::kSmemFwd, st>>>(a);
  } else {
    ERV_CUDA(allow_smem(kfft::kfft_fwd_kernel<__nv_bfloat16>, kfft::kSmemFwd));
    kfft::kfft_fwd_kernel<__nv_bfloat16><<<grid, kfft::NT, kfft::kSmemFwd, st>>>(a);
  }
  ERV_LAUNCH_CHECK();
#define ERV_KFFT_TAIL(DHV)                                                                                      \
  case DHV:                                                                                                     \
    return dtype == ERV_F32 ? kfft_tail<float, DHV>(a, cexp, out, den, st) : kfft_tail<__nv_bfloat16, DHV>(a, cexp, out, den, st);
  switch (DH) {
    ERV_KFFT_TAIL(8)
    ERV_KFFT_TAIL(16)
    ERV_KFFT_TAIL(32)
    ERV_KFFT_TAIL(64)
  }
#undef ERV_KFFT_TAIL
  set_error("kerple_fft_forward: head_dim %d unsupported", DH);
  return ERV_E_INVALID;
}

}  // namespace erv

extern "C" void erv_kerple_set_fft(int mode) { erv::g_fft_mode.store(mode < 0 ? -1 : (mode ? 1 : 0)); }
