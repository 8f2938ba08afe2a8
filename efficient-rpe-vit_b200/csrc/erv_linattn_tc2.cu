// Tensor-core (tcgen05 + TMEM) FAVOR+/ReLU linear-attention forward for short sequences (33 <= N <= 65, head_dim 16,
// num_features <= 256): TWO (batch, head) pairs share every 128-row tile.
//
// ViT sequences are patches + CLS (N = 65 at CIFAR patch 4).  With one pair per 128-row tile (erv_linattn_tc.cu) half
// of every tile is padding and the 65th token costs a whole 32-lane quarter.  Here rows 0..63 of a tile belong to
// pair (2g, h) and rows 64..127 to pair (2g+1, h); when N = 65 the last token of each pair (the "lone" token) stays
// out of the tile: one warp per (pair, q|k) computes its feature row in fp32, its rank-1 term is folded into S when S
// leaves TMEM, and its output row is a 256-thread reduction against the finished S.
//
// Per group of two pairs (one persistent CTA per SM, 512 threads = 4 per tile row, each owning a quarter of the row's
// features in registers):
//   G1k  P = k W^T                       3xTF32 tcgen05.mma, N = Mp              -> TMEM cols [0, 256)
//   G1q  P = q W^T                       issued as soon as P(k) sits in registers; runs under the exp phase of the keys
//   G2   S_pair[f][.] = phi_k^T [v|1]    per pair and 128-feature block; hi/lo bf16 split of both operands in TWO
//                                        instructions per k-step: phi_hi x [v_hi|1|v_lo] (N = 48), phi_lo x [v_hi|1] (N = 32)
//   G4   num = phi_q [S_A|S_B]           both pairs in one N = 64 instruction stream: phi_hi x [S_hi(A,B)|S_lo(A,B)],
//                                        phi_lo x S_hi(A,B); rows < 64 read the A columns, rows >= 64 the B columns
//   den  = phi_q . z                     fp32 FMAs from the feature registers (keeps the [S|z] image at N = 64)
// An MMA costs >= 96 cycles however narrow (profiles/r01_tcgen05_mma_cost.md); this layout issues 76 per two pairs
// where the one-pair kernel issues 90 per pair.  Global rows of the next group are fetched into registers while G4 runs.
#include "erv_tc_common.cuh"

namespace erv {

template <typename T, int NC>
__global__ void __launch_bounds__(kTcThreads, 1) la_tc2_fwd_kernel(const LaTcArgs p) {
  constexpr int DH = 16;
  using C = TcCfg<DH>;
  constexpr int Mp = 32 * NC, FQ = 8 * NC, nrb = (Mp + 127) / 128;
  constexpr uint32_t COL_S = 256, COL_O = 256, S_STRIDE = 64;  // O reuses S(A, 0): S is in shared memory by then
  constexpr uint32_t wbytes = (uint32_t)(Mp / 8) * (DH / 4) * 128;
  constexpr uint32_t phibytes = (uint32_t)nrb * 16 * kTokCh;
  constexpr uint32_t s_ch = (uint32_t)(Mp / 8) * 128;  // chunk stride of the S image (rows = features)

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;  // G1 completions / G2, G4 completions
  __shared__ uint32_t tmem_base_s;
  __shared__ float n2_s[2][128];                   // |x|^2/2 of the key / query rows
  __shared__ float ex_s[4][128];                   // row-max exchange between the 4 threads of a row
  __shared__ float den_s[4][128];
  __shared__ __align__(16) float z_s[2][Mp];
  __shared__ __align__(16) float lone_s[2][2][Mp];  // [q|k][pair side][feature]
  __shared__ __align__(16) float lone_v[2][DH];
  __shared__ float red_s[16][DH + 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, part = tid >> 7;
  const int M = p.M, N = p.N, H = p.H, B = p.B;
  const bool lone = N > 64;
  const int Nm = lone ? N - 1 : N;  // tokens per pair that go through the tile
  const int ks = (Nm + 15) >> 4;    // 16-token k-steps per pair
  const int side = row >> 6, n = row & 63;
  const int ngroups = ((B + 1) >> 1) * H;
  const int h = blockIdx.x % H;  // the grid is a multiple of H: a CTA stays on one head
  const bool favor = p.kind == ERV_FEAT_FAVOR;

  uint8_t* wh = smem;
  uint8_t* wl = wh + wbytes;
  uint8_t* xh = wl + wbytes;
  uint8_t* xl = xh + C::X_BYTES;
  uint8_t* phi1 = xl + C::X_BYTES;
  uint8_t* phi2 = phi1 + phibytes;
  uint8_t* vs = phi2 + phibytes;  // keys: [v_hi | 1 | 0 | v_lo] rows (6 chunks of kTokCh); queries: S image (8 chunks of s_ch)

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_b, 1);
    mbar_init_fence();
  }
  {  // W^T hi/lo TF32 images of this head: rows f, K = Dh
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
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_a = 0, ph_b = 0;

  const T* qkv = static_cast<const T*>(p.qkv);
  T* out = static_cast<T*>(p.out);
  const uint32_t idesc_p = make_idesc(FMT_TF32, 128, Mp, false, false);
  const uint32_t idesc_g2a = make_idesc(FMT_BF16, 128, 48, true, true);
  const uint32_t idesc_g2b = make_idesc(FMT_BF16, 128, 32, true, true);
  const uint32_t idesc_g4a = make_idesc(FMT_BF16, 128, 64, false, true);
  const uint32_t idesc_g4b = make_idesc(FMT_BF16, 128, 32, false, true);
  const float kLog2e = 1.4426950408889634f;
  const float log2_c = log2f(p.inv_sqrt_m);  // 1/sqrt(M) folded into the exponent
  const int fbeg = part * FQ;
  const uint32_t rowoff = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;

  // ---- global rows of a group -> registers (consumed at the top of that group's iteration)
  float nx[DH];
  float4 nv4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < DH; ++a) nx[a] = 0.f;
  auto prefetch = [&](int g) {
    if (g >= ngroups) return;
    const int b2 = g / H;
    if (part < 3) {  // 0: key row, 1: value row, 2: query row
      const int b = 2 * b2 + side;
      if (b < B && n < Nm) load_row<T, DH>(qkv + qkv_off(b, n, part == 0 ? 1 : (part == 1 ? 2 : 0), h, N, H, DH), nx);
    } else if (lone) {  // warp (pair side, q|k): the last token's row, same address in every lane
      const int lw = warp & 3, b = 2 * b2 + (lw >> 1);
      if (b < B) {
        load_row<T, DH>(qkv + qkv_off(b, N - 1, lw & 1, h, N, H, DH), nx);
        if ((lw & 1) && lane < 4) nv4 = ld4(qkv + qkv_off(b, N - 1, 2, h, N, H, DH) + 4 * lane);
      }
    }
  };

  auto issue_g1 = [&]() {  // P = x W^T, three TF32 terms (one thread)
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
  };

  // P (TMEM) -> registers -> phi -> hi/lo bf16 images.  q = 0: keys (also launches G1 of the queries once P has been
  // read), q = 1: queries (also the normaliser partial phi . z).
  auto feature_phase = [&](const int q) {
    uint32_t pr[NC][8];  // fp32 bit patterns of P[row][fbeg + 8c + i]
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_ld8_nowait(tm + lane_off + fbeg + c * 8, pr[c]);
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_wait_ld8(pr[c]);
    if (favor) {
      float m_part = -INFINITY;
      if (M < Mp) {  // padded features do not take part in the maximum
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (fbeg + c * 8 + i < M) m_part = fmaxf(m_part, __uint_as_float(pr[c][i]));
      } else {
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) m_part = fmaxf(m_part, __uint_as_float(pr[c][i]));
      }
      ex_s[part][row] = m_part;
    }
    fence_before_sync();
    __syncthreads();  // every thread holds its P values: the P columns may be overwritten
    if (q == 0 && warp == 0 && elect_one()) {
      fence_after_sync();
      issue_g1();  // queries (their x images were written before the barrier)
    }
    float mx = 0.f;
    if (favor) mx = fmaxf(fmaxf(ex_s[0][row], ex_s[1][row]), fmaxf(ex_s[2][row], ex_s[3][row]));
    const float n2 = n2_s[q][row];
    // phi = exp(P - mx - n2)/sqrt(M) = 2^(P*log2e - (mx + n2)*log2e + log2(1/sqrt(M)))
    const float shift = fmaf(mx + n2, kLog2e, -log2_c);
    const float scale = (n2 < INFINITY) ? p.inv_sqrt_m : 0.f;
    float den_part = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float ph_v[8];
      const int f0 = fbeg + c * 8;
      if (favor) {
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
      if (q == 1) {
        const float4 za = ld4(&z_s[side][f0]), zb = ld4(&z_s[side][f0 + 4]);
        den_part = fmaf(ph_v[0], za.x, den_part); den_part = fmaf(ph_v[1], za.y, den_part);
        den_part = fmaf(ph_v[2], za.z, den_part); den_part = fmaf(ph_v[3], za.w, den_part);
        den_part = fmaf(ph_v[4], zb.x, den_part); den_part = fmaf(ph_v[5], zb.y, den_part);
        den_part = fmaf(ph_v[6], zb.z, den_part); den_part = fmaf(ph_v[7], zb.w, den_part);
      }
      store_split8(phi1, phi2, (uint32_t)(f0 >> 3) * kTokCh + rowoff, ph_v);
    }
    if (q == 1) den_s[part][row] = den_part;
  };

  prefetch(blockIdx.x);
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int b2 = g / H;
    const int b = 2 * b2 + side;
    const bool valid = b < B && n < Nm;
    // ---- step 1: consume the prefetched rows
    float xq[DH];  // part 2: prepared query row, written to the x images once G1 of the keys has read them
    float n2q = INFINITY;
    if (part == 0) {
      float x[DH];
      float n2 = INFINITY;  // invalid rows: exponent -inf -> phi = 0
#pragma unroll
      for (int a = 0; a < DH; ++a) x[a] = valid ? nx[a] : 0.f;
      if (valid) {
        prologue_row<DH>(x, p.rot, p.ta, p.tb, h, n, N, p.prescale);
        n2 = 0.f;
#pragma unroll
        for (int a = 0; a < DH; ++a) n2 = fmaf(x[a], x[a], n2);
        n2 *= 0.5f;
      }
      n2_s[0][row] = n2;
      store_x_images<DH>(xh, xl, x, row);
    } else if (part == 1) {  // [v_hi | 1 | 0 | v_lo] row of the G2 operand
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        float ch[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[e] = valid ? nx[8 * c + e] : 0.f;
        store_split8(vs, vs + 4 * kTokCh, c * kTokCh + rowoff, ch);
      }
      *reinterpret_cast<uint4*>(vs + 2 * kTokCh + rowoff) = make_uint4(valid ? 0x00003F80u : 0u, 0u, 0u, 0u);  // bf16 1.0
      *reinterpret_cast<uint4*>(vs + 3 * kTokCh + rowoff) = make_uint4(0u, 0u, 0u, 0u);
    } else if (part == 2) {
#pragma unroll
      for (int a = 0; a < DH; ++a) xq[a] = valid ? nx[a] : 0.f;
      if (valid) {
        prologue_row<DH>(xq, p.rot, p.ta, p.tb, h, n, N, p.prescale);
        n2q = 0.f;
#pragma unroll
        for (int a = 0; a < DH; ++a) n2q = fmaf(xq[a], xq[a], n2q);
        n2q *= 0.5f;
      }
    } else if (lone) {  // feature rows of the lone tokens: warp = (pair side, q|k), lanes over features
      const int lw = warp & 3, sp = lw >> 1, which = lw & 1;
      const bool ok = 2 * b2 + sp < B;
      float x[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) x[a] = ok ? nx[a] : 0.f;
      if (ok) prologue_row<DH>(x, p.rot, p.ta, p.tb, h, N - 1, N, p.prescale);
      float n2 = 0.f;
#pragma unroll
      for (int a = 0; a < DH; ++a) n2 = fmaf(x[a], x[a], n2);
      n2 *= 0.5f;
      float pv[NC];
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int f = lane + 32 * i;
        const uint32_t wo = (uint32_t)(f >> 3) * C::X_SBO + (f & 7) * 16;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 a = ld4(reinterpret_cast<const float*>(wh + wo + c * C::X_LBO));
          const float4 l = ld4(reinterpret_cast<const float*>(wl + wo + c * C::X_LBO));
          acc = fmaf(x[4 * c], a.x + l.x, acc); acc = fmaf(x[4 * c + 1], a.y + l.y, acc);
          acc = fmaf(x[4 * c + 2], a.z + l.z, acc); acc = fmaf(x[4 * c + 3], a.w + l.w, acc);
        }
        pv[i] = acc;
        if (f < M) m = fmaxf(m, acc);
      }
      m = warp_max(m);
      const float shift = fmaf(m + n2, kLog2e, -log2_c);
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int f = lane + 32 * i;
        float v = favor ? ex2_approx(fmaf(pv[i], kLog2e, -shift)) : fmaxf(pv[i], 0.f) * p.inv_sqrt_m;
        if (f >= M || !ok) v = 0.f;
        lone_s[which][sp][f] = v;
      }
      if (which == 1 && lane < 4) st4(&lone_v[sp][4 * lane], ok ? nv4 : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    // ---- G1 (keys)
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      issue_g1();
    }
    mbar_wait(&bar_a, ph_a);
    ph_a ^= 1;
    fence_after_sync();
    if (part == 2) {  // the x images are free again: stage the query rows for the G1 issued inside feature_phase(0)
      n2_s[1][row] = n2q;
      store_x_images<DH>(xh, xl, xq, row);
      fence_smem_to_async();
    }
    feature_phase(0);
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    // ---- G2: S(pair, rb) = phi_k^T [v|1]
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      for (int sp = 0; sp < 2; ++sp)
        for (int rb = 0; rb < nrb; ++rb) {
          const uint32_t d = tm + COL_S + (uint32_t)(sp * 2 + rb) * S_STRIDE;
          for (int s = 0; s < ks; ++s) {
            const uint32_t st = (uint32_t)(sp * 4 + s) * 256;
            const uint64_t bd = make_desc(smem_u32(vs) + st, 128, kTokCh);
            mma_f16(d, make_desc(smem_u32(phi1) + (uint32_t)rb * 16 * kTokCh + st, 128, kTokCh), bd, idesc_g2a, s > 0);
            mma_f16(d, make_desc(smem_u32(phi2) + (uint32_t)rb * 16 * kTokCh + st, 128, kTokCh), bd, idesc_g2b, true);
          }
        }
      commit(&bar_b);
    }
    mbar_wait(&bar_b, ph_b);
    ph_b ^= 1;
    fence_after_sync();
    // ---- S (TMEM, lanes = features) -> S image for G4, z, lone-token terms.  Thread = (pair side, block, feature).
    {
      const int sp = part >> 1, rb = part & 1, f = rb * 128 + row;
      float acc32[32];  // lone query read-out partials: [0, DH) numerator, DH denominator, rest padding
#pragma unroll
      for (int j = 0; j < 32; ++j) acc32[j] = 0.f;
      if (rb < nrb && f < Mp) {  // warp-uniform
        float d0[32], d1[16], sv[DH];
        tmem_ld32(tm + lane_off + COL_S + (uint32_t)(sp * 2 + rb) * S_STRIDE, d0);
        tmem_ld16(tm + lane_off + COL_S + (uint32_t)(sp * 2 + rb) * S_STRIDE + 32, d1);
#pragma unroll
        for (int d = 0; d < DH; ++d) sv[d] = d0[d] + d1[d];
        float z = d0[DH];
        if (lone) {  // rank-1 term of the last key
          const float pk = lone_s[1][sp][f];
#pragma unroll
          for (int d = 0; d < DH; ++d) sv[d] = fmaf(pk, lone_v[sp][d], sv[d]);
          z += pk;
        }
        z_s[sp][f] = z;
#pragma unroll
        for (int c = 0; c < DH / 8; ++c) {
          float ch[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) ch[e] = sv[8 * c + e];
          store_split8(vs + (uint32_t)(2 * sp + c) * s_ch, vs + (uint32_t)(4 + 2 * sp + c) * s_ch,
                       (uint32_t)(f >> 3) * 128 + (f & 7) * 16, ch);
        }
        if (p.state != nullptr && 2 * b2 + sp < B) {  // [S|z] of this pair, d-major so that a warp writes 128 B lines
          float* sp_out = p.state + ((size_t)(2 * b2 + sp) * H + h) * (DH + 1) * Mp + f;
#pragma unroll
          for (int d = 0; d < DH; ++d) sp_out[(size_t)d * Mp] = sv[d];
          sp_out[(size_t)DH * Mp] = z;
        }
        if (lone) {  // the last query's read-out against the finished S
          const float pq = lone_s[0][sp][f];
#pragma unroll
          for (int d = 0; d < DH; ++d) acc32[d] = pq * sv[d];
          acc32[DH] = pq * z;
        }
      }
      if (lone) {
        const float t = warp_sum32(acc32);  // lane j holds the warp total of partial j
        if (lane <= DH) red_s[warp][lane] = t;
      }
    }
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    if (lone && tid < 8) {  // warps 8*sp .. 8*sp+7 hold the partial sums of pair side sp
      const int sp = tid >> 2, c = tid & 3, bb = 2 * b2 + sp;
      if (bb < B) {
        float den = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
        for (int w = 8 * sp; w < 8 * sp + 8; ++w) {
          den += red_s[w][DH];
          o0 += red_s[w][4 * c]; o1 += red_s[w][4 * c + 1]; o2 += red_s[w][4 * c + 2]; o3 += red_s[w][4 * c + 3];
        }
        den += kEps;
        st4(out + out_off(bb, N - 1, h, N, H, DH) + 4 * c, make_float4(o0 / den, o1 / den, o2 / den, o3 / den));
      }
    }
    // ---- queries: P(q) was computed under the keys' exp phase
    mbar_wait(&bar_a, ph_a);
    ph_a ^= 1;
    fence_after_sync();
    feature_phase(1);
    fence_smem_to_async();
    fence_before_sync();
    __syncthreads();
    // ---- G4: [num_A | num_B] = phi_q [S_A | S_B]
    if (warp == 0 && elect_one()) {
      fence_after_sync();
      for (int s = 0; s < Mp / 16; ++s) {
        const uint64_t bd = make_desc(smem_u32(vs) + (uint32_t)s * 256, 128, s_ch);
        mma_f16(tm + COL_O, make_desc(smem_u32(phi1) + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_g4a, s > 0);
        mma_f16(tm + COL_O, make_desc(smem_u32(phi2) + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_g4b, true);
      }
      commit(&bar_b);
    }
    prefetch(g + gridDim.x);  // the next group's rows travel while the tensor pipe works
    mbar_wait(&bar_b, ph_b);
    ph_b ^= 1;
    fence_after_sync();
    if (part == 0) {  // out = num / (den + eps): hi-part + lo-part columns of this row's pair
      float a0[16], a1[16];
      tmem_ld16(tm + lane_off + COL_O + 16 * side, a0);
      tmem_ld16(tm + lane_off + COL_O + 32 + 16 * side, a1);
      if (valid) {
        const float den = (den_s[0][row] + den_s[1][row]) + (den_s[2][row] + den_s[3][row]) + kEps;
        T* ob = out + out_off(b, n, h, N, H, DH);
#pragma unroll
        for (int c = 0; c < DH / 4; ++c)
          st4(ob + 4 * c, make_float4((a0[4 * c] + a1[4 * c]) / den, (a0[4 * c + 1] + a1[4 * c + 1]) / den,
                                      (a0[4 * c + 2] + a1[4 * c + 2]) / den, (a0[4 * c + 3] + a1[4 * c + 3]) / den));
      }
    }
    fence_before_sync();
    __syncthreads();  // TMEM columns and the [v|1] / S images are reused by the next group
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int tc2_mp(int M) { return M <= 64 ? 64 : (M <= 128 ? 128 : 256); }

size_t la_tc2_smem_bytes(int Mp) {
  const size_t nrb = (Mp + 127) / 128;
  const size_t wbytes = (size_t)(Mp / 8) * 4 * 128, xbytes = 16 * 4 * 128;
  const size_t v_bytes = 6 * (size_t)kTokCh, s_bytes = 8 * (size_t)(Mp / 8) * 128;
  return 2 * wbytes + 2 * xbytes + 2 * nrb * 16 * kTokCh + (v_bytes > s_bytes ? v_bytes : s_bytes) + 128;
}

bool la_tc2_eligible(int N, int DH, int M) {
  static const bool disabled = getenv("ERV_DISABLE_TC2") != nullptr;
  return !disabled && DH == 16 && M <= 256 && N >= 33 && N <= 65;
}

int la_tc2_forward(const void* qkv, void* out, const float* omega, int B, int N, int H, int M, int kind, int rot,
                   const float* ta, const float* tb, int dtype, float* state, cudaStream_t st) {
  LaTcArgs a;
  a.qkv = qkv; a.out = out; a.omega = omega; a.ta = ta; a.tb = tb;
  a.B = B; a.N = N; a.H = H; a.M = M; a.Mp16 = tc2_mp(M); a.kind = kind; a.rot = rot;
  a.prescale = (float)pow(16.0, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.state = state;
  const size_t smem = la_tc2_smem_bytes(a.Mp16);
  const int ngroups = ((B + 1) / 2) * H;
  int grid = (kNumSMs / H) * H;  // multiple of H: each CTA stays on one head (W images staged once)
  if (grid < H) grid = H;
  if (grid > ngroups) grid = ngroups;
#define TC2_LAUNCH(TT, NC_)                                                   \
  do {                                                                        \
    ERV_CUDA(allow_smem(la_tc2_fwd_kernel<TT, NC_>, smem));                   \
    la_tc2_fwd_kernel<TT, NC_><<<grid, kTcThreads, smem, st>>>(a);            \
  } while (0)
  if (dtype == ERV_F32) {
    if (a.Mp16 == 64) TC2_LAUNCH(float, 2); else if (a.Mp16 == 128) TC2_LAUNCH(float, 4); else TC2_LAUNCH(float, 8);
  } else {
    if (a.Mp16 == 64) TC2_LAUNCH(__nv_bfloat16, 2); else if (a.Mp16 == 128) TC2_LAUNCH(__nv_bfloat16, 4); else TC2_LAUNCH(__nv_bfloat16, 8);
  }
#undef TC2_LAUNCH
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv
