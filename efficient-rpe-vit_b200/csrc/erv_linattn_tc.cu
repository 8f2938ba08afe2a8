// Tensor-core (tcgen05 + TMEM) FAVOR+/ReLU linear attention for head_dim 8/16, num_features <= 256.
//
// One persistent CTA (128 threads = 128 TMEM lanes) per SM walks (batch, head) pairs; tokens go through 128-row
// tiles, thread t owns token t of the tile:
//   G1  P[t][f]      = x_t . w_f                     (K = Dh)      3xTF32: x_hi w_hi + x_lo w_hi + x_hi w_lo
//       phi          = exp(P - max_f P - |x|^2/2)/sqrt(M)  or  relu(P)/sqrt(M)   -- row max is thread-local
//   G2  S[f][d]     += sum_t phi_k[t][f] [v|1][t][d]  (K = tokens)  phi in TF32, [v|1] split hi+lo
//   G4  [num|den][t] = sum_f phi_q[t][f] S[f][.]      (K = features)
// Accumulators live in TMEM (P: 256 columns, S: 2x32, num|den: 32); operands are written to shared memory by the
// owning threads in the no-swizzle K-major canonical layout (erv_umma.cuh).  The projection needs fp32-level
// accuracy because it feeds exp(); rounding phi and S to TF32 is averaged over N*M terms and stays far below the
// 1e-4 parity budget, the value rows are split hi+lo (tests/test_parity_gpu.py holds the kernel to 1e-4).
#include "erv_tc_common.cuh"

namespace erv {


// NC = 8-feature chunks per thread: Mp = 32 * NC is compile time so the P registers are statically indexed.
template <typename T, int DH, int NC>
__global__ void __launch_bounds__(kTcThreads, 1) la_tc_fwd_kernel(const LaTcArgs p) {
  using C = TcCfg<DH>;
  // bf16 inputs (autocast, budget 2e-2): the contractions use the hi images only -- one bf16 product instead of the three
  // split terms the fp32 path needs for 1e-4; the projection keeps 3xTF32 (it feeds exp)
  constexpr int kTerms = sizeof(T) == 2 ? 1 : 3;
  constexpr int ND = C::ND;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;  // G1 completions / G2-G4 completions
  __shared__ uint32_t tmem_base_s;
  __shared__ float n2_s[128];
  __shared__ float mx_s[4][128];

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, part = tid >> 7;
  constexpr int Mp = 32 * NC, FQ = 8 * NC;   // padded features; features per thread
  constexpr int nrb = (Mp + 127) / 128;      // 128-feature row blocks of S
  const int M = p.M, N = p.N;
  const uint32_t wbytes = tc_w_bytes(DH, Mp);
  const uint32_t phibytes = (uint32_t)nrb * 16 * kTokCh;  // padded to whole row blocks (G2 reads 128 rows per block)
  const uint32_t s_ch = (uint32_t)(Mp / 8) * 128;         // chunk stride of the [S|z] image (rows = features)
  uint8_t* wh = smem;
  uint8_t* wl = wh + wbytes;
  uint8_t* xh = wl + wbytes;
  uint8_t* xl = xh + C::X_BYTES;
  uint8_t* phi1 = xl + C::X_BYTES;
  uint8_t* phi2 = phi1 + phibytes;
  uint8_t* vreg = phi2 + phibytes;  // K pass: [2 buffers][hi, lo] of (ND/8)*kTokCh; Q pass: [S|z] hi, lo
  const uint32_t vbytes = (uint32_t)(ND / 8) * kTokCh;
  uint8_t* s1 = vreg;
  uint8_t* s2 = vreg + (uint32_t)(ND / 8) * s_ch;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_b, 1);
    mbar_init_fence();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_a = 0, ph_b = 0;
  bool pending_b = false;

  const T* qkv = static_cast<const T*>(p.qkv);
  T* out = static_cast<T*>(p.out);
  const size_t tok_stride = (size_t)3 * p.H * DH;
  const uint32_t idesc_p = make_idesc(FMT_TF32, 128, Mp, false, false);
  const uint32_t idesc_g2 = make_idesc(FMT_BF16, 128, ND, true, true);
  const uint32_t idesc_g4 = make_idesc(FMT_BF16, 128, ND, false, true);
  const float kLog2e = 1.4426950408889634f;
  const float log2_c = log2f(p.inv_sqrt_m);  // 1/sqrt(M) folded into the exponent
  const int fbeg = part * FQ;
  int cur_h = -1;

  for (int pair = blockIdx.x; pair < p.B * p.H; pair += gridDim.x) {
    const int b = pair / p.H, h = pair % p.H;
    if (h != cur_h) {  // stage W^T hi/lo TF32 images for this head: rows f, K = Dh
      cur_h = h;
      const float* om = p.omega + (size_t)h * DH * M;
      for (int i = tid; i < Mp * DH; i += kTcThreads) {
        const int d = i / Mp, f = i % Mp;
        const float w = (f < M) ? __ldg(om + (size_t)d * M + f) : 0.f;
        const float hi = to_tf32(w), lo = to_tf32(w - hi);
        const uint32_t off = off_kmajor(f, d, 4, 4, C::X_LBO, C::X_SBO);
        *reinterpret_cast<float*>(wh + off) = hi;
        *reinterpret_cast<float*>(wl + off) = lo;
      }
    }
    const T* qb = qkv + qkv_off(b, 0, 0, h, N, p.H, DH);
    const T* kb = qkv + qkv_off(b, 0, 1, h, N, p.H, DH);
    const T* vb = qkv + qkv_off(b, 0, 2, h, N, p.H, DH);

    auto emit_out = [&](int n0) {  // out = num / (den + eps) for the tile starting at n0 (reads TMEM: all lanes)
      if (part != 0) return;       // warps 0-3 cover the 128 lanes (warp-uniform branch)
      float r[32];
      tmem_ld32(tm + lane_off + C::COL_O, r);
      const int n = n0 + row;
      if (n < N) {
        const float den = r[DH] + kEps;
        T* ob = out + out_off(b, n, h, N, p.H, DH);
#pragma unroll
        for (int c = 0; c < DH / 4; ++c)
          st4(ob + 4 * c, make_float4(r[4 * c] / den, r[4 * c + 1] / den, r[4 * c + 2] / den, r[4 * c + 3] / den));
      }
    };

    for (int pass = 0; pass < 2; ++pass) {  // pass 0: keys -> S ; pass 1: queries -> out
      int tile = 0;
      for (int n0 = 0; n0 < N; n0 += 128, ++tile) {
        const int n = n0 + row;
        const bool valid = n < N;
        const int nt = min(128, N - n0);
        const int nt16 = (nt + 15) & ~15;                    // rows the contractions actually read
        const bool warp_live = (row & ~31) < nt16;           // warp-uniform: this 32-row group holds read rows
        // ---- step 1: operand images of this tile
        if (part == 0) {
          float x[DH];
          float n2 = INFINITY;  // invalid rows: exponent -inf -> phi = 0
          if (valid) {
            load_row<T, DH>((pass == 0 ? kb : qb) + (size_t)n * tok_stride, x);
            prologue_row<DH, true>(x, p.rot, p.ta, p.tb, h, n, N, p.prescale);
            n2 = 0.f;
#pragma unroll
            for (int a = 0; a < DH; ++a) n2 = fmaf(x[a], x[a], n2);
            n2 *= 0.5f;
          } else {
#pragma unroll
            for (int a = 0; a < DH; ++a) x[a] = 0.f;
          }
          n2_s[row] = n2;
          store_x_images<DH>(xh, xl, x, row);
        } else if (part == 1 && pass == 0 && warp_live) {  // [v | 1] rows, double buffered across tiles
          uint8_t* v1 = vreg + (uint32_t)(tile & 1) * 2 * vbytes;
          uint8_t* v2 = v1 + vbytes;
          float v[DH];
          if (valid) {
            load_row<T, DH>(vb + (size_t)n * tok_stride, v);
          } else {
#pragma unroll
            for (int d = 0; d < DH; ++d) v[d] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < DH / 8; ++c) {
            float ch[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ch[e] = v[8 * c + e];
            store_split8(v1, v2, c * kTokCh + (row >> 3) * 128 + (row & 7) * 16, ch);
          }
          const float ones[8] = {valid ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          store_split8(v1, v2, (DH / 8) * kTokCh + (row >> 3) * 128 + (row & 7) * 16, ones);
        }
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
        // ---- G1: P = x W^T, three TF32 terms
        if (warp == 0 && elect_one()) {
          fence_after_sync();
          bool acc = false;
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint8_t* xa = (term == 1) ? xl : xh;
            const uint8_t* wb = (term == 2) ? wl : wh;
#pragma unroll
            for (int s = 0; s < DH / 8; ++s) {
              mma_tf32(tm, make_desc(smem_u32(xa) + s * 2 * C::X_LBO, C::X_LBO, C::X_SBO),
                       make_desc(smem_u32(wb) + s * 2 * C::X_LBO, C::X_LBO, C::X_SBO), idesc_p, acc);
              acc = true;
            }
          }
          commit(&bar_a);
        }
        mbar_wait(&bar_a, ph_a);
        ph_a ^= 1;
        if (pending_b) {  // tensor-pipe work completes in order: the previous tile's G2/G4 is done as well
          mbar_wait(&bar_b, ph_b);
          ph_b ^= 1;
          pending_b = false;
          fence_after_sync();
          if (pass == 1) emit_out(n0 - 128);
        }
        fence_after_sync();
        // ---- this thread's quarter of the row: P -> registers (one TMEM read), max, phi, hi/lo bf16 images
        uint32_t pr[NC][8];  // fp32 bit patterns of P[row][fbeg + 8c + i]
        if (warp_live) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            tmem_ld8_nowait(tm + lane_off + fbeg + c * 8, pr[c]);
#pragma unroll
          for (int c = 0; c < NC; ++c)
            tmem_wait_ld8(pr[c]);
        }
        float mx = 0.f;
        if (p.kind == ERV_FEAT_FAVOR) {
          float m_part = -INFINITY;
          if (warp_live) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (fbeg + c * 8 + i < M) m_part = fmaxf(m_part, __uint_as_float(pr[c][i]));
              }
          }
          mx_s[part][row] = m_part;
          __syncthreads();
          mx = fmaxf(fmaxf(mx_s[0][row], mx_s[1][row]), fmaxf(mx_s[2][row], mx_s[3][row]));
        }
        if (warp_live) {
          // phi = exp(P - mx - n2)/sqrt(M) = 2^(P*log2e - (mx + n2)*log2e + log2(1/sqrt(M)))
          const float shift = fmaf(mx + n2_s[row], kLog2e, -log2_c);
          const float scale = valid ? p.inv_sqrt_m : 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c)
            {
              float ph_v[8];
              const int f0 = fbeg + c * 8;
              if (p.kind == ERV_FEAT_FAVOR) {
#pragma unroll
                for (int i = 0; i < 8; ++i) ph_v[i] = ex2_approx(fmaf(__uint_as_float(pr[c][i]), kLog2e, -shift));
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) ph_v[i] = fmaxf(__uint_as_float(pr[c][i]), 0.f) * scale;
              }
              if (f0 + 8 > M) {  // padded features contribute nothing
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (f0 + i >= M) ph_v[i] = 0.f;
              }
              store_split8(phi1, phi2, (uint32_t)(f0 >> 3) * kTokCh + (row >> 3) * 128 + (row & 7) * 16, ph_v);
            }
        }
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
          fence_after_sync();
          if (pass == 0) {  // G2: S[rb] (+)= phi^T [v|1], terms hi*hi + hi*lo + lo*hi
            const uint8_t* v1 = vreg + (uint32_t)(tile & 1) * 2 * vbytes;
            const uint8_t* v2 = v1 + vbytes;
            const int ksteps = nt16 / 16;
            for (int rb = 0; rb < nrb; ++rb) {
              bool acc = n0 > 0;
              for (int term = 0; term < kTerms; ++term) {
                const uint8_t* a_img = (term == 2) ? phi2 : phi1;
                const uint8_t* b_img = (term == 1) ? v2 : v1;
                for (int s = 0; s < ksteps; ++s) {
                  mma_f16(tm + C::COL_S + rb * ND,
                          make_desc(smem_u32(a_img) + (uint32_t)rb * 16 * kTokCh + s * 256, 128, kTokCh),
                          make_desc(smem_u32(b_img) + s * 256, 128, kTokCh), idesc_g2, acc);
                  acc = true;
                }
              }
            }
          } else {  // G4: [num|den] = phi [S|z]
            bool acc = false;
            for (int term = 0; term < kTerms; ++term) {
              const uint8_t* a_img = (term == 2) ? phi2 : phi1;
              const uint8_t* b_img = (term == 1) ? s2 : s1;
              for (int s = 0; s < Mp / 16; ++s) {
                mma_f16(tm + C::COL_O, make_desc(smem_u32(a_img) + (uint32_t)s * 2 * kTokCh, kTokCh, 128),
                        make_desc(smem_u32(b_img) + s * 256, 128, s_ch), idesc_g4, acc);
                acc = true;
              }
            }
          }
          commit(&bar_b);
        }
        pending_b = true;
      }
      // ---- end of pass: drain the tensor pipe
      mbar_wait(&bar_b, ph_b);
      ph_b ^= 1;
      pending_b = false;
      fence_after_sync();
      if (pass == 0) {  // S (TMEM, lanes = features) -> [S|z] hi/lo bf16 images: byte(f, d) = (d/8)*s_ch + (f/8)*128 + (f%8)*16
        if (part < nrb) {  // warp-uniform
          float sv[32];
          tmem_ld32(tm + lane_off + C::COL_S + part * ND, sv);
          const int f = part * 128 + row;
          if (f < Mp) {
#pragma unroll
            for (int c = 0; c < ND / 8; ++c) {
              float ch[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) ch[e] = sv[8 * c + e];
              store_split8(s1, s2, c * s_ch + (f >> 3) * 128 + (f & 7) * 16, ch);
            }
            if (p.state != nullptr) {  // [S|z] of this pair, d-major: a warp writes 128-byte lines; the backward skips its K1 sweep
              float* so = p.state + (size_t)pair * (DH + 1) * Mp + f;
#pragma unroll
              for (int d = 0; d <= DH; ++d) so[(size_t)d * Mp] = sv[d];
            }
          }
        }
        fence_before_sync();
      } else {
        emit_out(((N - 1) / 128) * 128);
        fence_before_sync();
      }
    }
    __syncthreads();  // the next pair's [v|1] images overwrite the [S|z] images
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

size_t la_tc_smem_bytes(int DH, int Mp) {
  const int ND = (DH + 1 + 15) / 16 * 16;
  const size_t x_bytes = 16 * (size_t)(DH / 4) * 128;
  size_t v_bytes = 4 * (size_t)(ND / 8) * kTokCh;               // 2 buffers x (hi, lo)
  const size_t s_bytes = 2 * (size_t)(ND / 8) * (Mp / 8) * 128;  // hi, lo
  if (s_bytes > v_bytes) v_bytes = s_bytes;
  const size_t nrb = (Mp + 127) / 128;
  return 2 * (size_t)tc_w_bytes(DH, Mp) + 2 * x_bytes + 2 * nrb * 16 * kTokCh + v_bytes + 128;
}

bool la_tc_eligible(int N, int DH, int M) {
  static const bool disabled = getenv("ERV_DISABLE_TC") != nullptr;
  return !disabled && DH == 16 && M <= 256 && N >= 33;
}

int la_tc_forward(const void* qkv, void* out, const float* omega, int B, int N, int H, int DH, int M, int kind, int rot,
                  const float* ta, const float* tb, int dtype, float* state, cudaStream_t st) {
  LaTcArgs a;
  a.qkv = qkv; a.out = out; a.omega = omega; a.ta = ta; a.tb = tb;
  a.B = B; a.N = N; a.H = H; a.M = M; a.Mp16 = M <= 64 ? 64 : (M <= 128 ? 128 : 256); a.kind = kind; a.rot = rot;
  a.prescale = (float)pow((double)DH, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.state = state;
  const size_t smem = la_tc_smem_bytes(DH, a.Mp16);
  int grid = (kNumSMs / H) * H;  // multiple of H: each CTA stays on one head (W images staged once)
  if (grid < H) grid = H;
  if (grid > B * H) grid = B * H;
#define TC_LAUNCH(TT, NC_)                                                      \
  do {                                                                          \
    ERV_CUDA(allow_smem(la_tc_fwd_kernel<TT, 16, NC_>, smem));                  \
    la_tc_fwd_kernel<TT, 16, NC_><<<grid, kTcThreads, smem, st>>>(a);           \
  } while (0)
  if (dtype == ERV_F32) {
    if (a.Mp16 == 64) TC_LAUNCH(float, 2); else if (a.Mp16 == 128) TC_LAUNCH(float, 4); else TC_LAUNCH(float, 8);
  } else {
    if (a.Mp16 == 64) TC_LAUNCH(__nv_bfloat16, 2); else if (a.Mp16 == 128) TC_LAUNCH(__nv_bfloat16, 4); else TC_LAUNCH(__nv_bfloat16, 8);
  }
#undef TC_LAUNCH
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv
