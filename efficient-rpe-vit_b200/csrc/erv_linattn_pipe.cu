// Warp-specialised, software-pipelined FAVOR+/ReLU linear-attention forward for short sequences (33 <= N <= 65,
// head_dim 16, 128 < num_features <= 256): the successor of la_tc2_fwd_kernel (erv_linattn_tc2.cu), same tile geometry
// (two (batch, head) pairs per 128-row tile, the 65th token of each pair handled outside the tile), same math
// (favor_plus.py:112-140, 247-260), but nothing waits for the tensor pipe any more:
//
//   * warp 16 only issues tcgen05.mma (one elected lane) and is throttled by mbarriers; the 16 compute warps (4 threads
//     per tile row) never issue an MMA, so none of them sits behind the ~60-170 cycles every instruction takes to issue.
//   * every feature image is produced and consumed in 128-feature halves: while the compute warps run the exponentials
//     of one half, the tensor pipe contracts the half written just before (keys: S = phi_k^T [v|1]; queries:
//     num = phi_q [S_A|S_B]) or projects the next rows (P = x W^T).
//   * tensor memory holds two P buffers: [0,256) keys -> later S (4 x 48 columns), [256,512) queries -> later the output
//     accumulator (64 columns).  The next group's key projection is issued as soon as S has left tensor memory, its query
//     projection as soon as the previous output rows have.
//
// Per group of two pairs the compute warps walk four units u = (keys|queries) x (half 0|1); the tensor queue, in order, is
//   G2h0(g)  G1q(g)  G2h1(g)  G4h0(g)  G1k(g+1)  G4h1(g)            (~6.3 k cycles of issue per group)
// against ~7 k cycles of exponentials, splits and epilogues on the compute warps.
//
// The projection uses three-level bf16 splits of x and W (six K = 16 products: 24 significant bits, the same six
// instructions 3xTF32 needs with its K = 8) so that the operand images take 36 KB instead of 48.
#include "erv_pipe_common.cuh"

namespace erv {

enum PipeBar {
  B_FULL_XK = 0, B_FULL_XQ, B_FULL_P0, B_FULL_P1, B_FULL_P2, B_FULL_P3, B_FULL_S0, B_FULL_S1, B_O_FREE,  // 16 arrivals
  B_FULL_RED,                                                                                            // 16 arrivals
  B_FULL_LONE,                                                                                           // 2 arrivals
  B_DONE_G1K, B_DONE_G1Q, B_DONE_G20, B_DONE_G21, B_DONE_G40, B_DONE_G41,                                // tcgen05.commit
  B_COUNT
};

template <typename T, bool FAVOR, bool PADDED>
__global__ void __launch_bounds__(kPipeThreads, 1) la_pipe_fwd_kernel(const LaTcArgs p) {
  constexpr int DH = 16, Mp = 256, HF = 128;
  constexpr uint32_t W_IMG = Mp * DH * 2, X_IMG = 128 * DH * 2;  // one bf16 level of W^T (8 KB) / of the token rows (4 KB)
  constexpr uint32_t XL = 128, XS = 256;                         // K-major [rows x 16]: chunk stride, 8-row group stride
  constexpr uint32_t PHI_IMG = 16 * kTokCh;                      // one feature half, one level: 32 KB
  constexpr uint32_t s_ch = (uint32_t)(Mp / 8) * 128;            // chunk stride of the S image (rows = features)
  // S image chunks (8 columns each): [S_hi A (2) | S_hi B (2) | Z = (zA_hi zB_hi zA_lo zB_lo 0 0 0 0) | S_lo A (2) | S_lo B (2)].
  // G4 reads 10 chunks (the tenth is whatever follows the image: its output columns are never read) with phi_hi and with
  // phi_lo: all four partial products, so that every output column carries the same terms.
  constexpr uint32_t CH_Z = 4, CH_LO = 5;
  constexpr uint32_t COL_PQ = 256, COL_S = 0, S_STRIDE = 48, COL_O = 256;
  constexpr uint32_t O_LO = 40, O_DEN = 32;  // accumulator columns: lo parts, normaliser (hi A, hi B, lo A, lo B)

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t tmem_base_s;
  __shared__ float n2_s[2][128];   // |x|^2/2 of the key / query rows
  __shared__ float ex_s[4][128];   // row-max exchange between the 4 threads of a row
  __shared__ __align__(16) float lone_s[2][2][2][Mp];  // [group parity][q|k][pair side][feature]
  __shared__ __align__(16) float lone_v[2][2][DH];
  __shared__ float lone_n2[2][4];
  __shared__ float lone_mx[2][4];
  __shared__ __align__(16) uint8_t lone_raw[2][2][6][64];  // [lone warp][group parity][qA kA qB kB vA vB]: raw rows (cp.async)
  __shared__ float red_s[16][9];  // per compute warp: lone query read-out partials (8 columns + normaliser)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, part = (tid >> 7) & 3;
  const int M = p.M, N = p.N, H = p.H, B = p.B;
  const bool lone = N > 64;
  const int Nm = lone ? N - 1 : N;
  const int ks = (Nm + 15) >> 4;
  const int side = row >> 6, n = row & 63;
  const int ngroups = ((B + 1) >> 1) * H;
  const int h = blockIdx.x % H;  // the grid is a multiple of H: a CTA stays on one head
  const int n_it = (ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const float kLog2e = 1.4426950408889634f;
  const float log2_c = log2f(p.inv_sqrt_m);  // 1/sqrt(M) folded into the exponent
  // per-token statistics saved behind the [S|z] blocks for the backward: aux[((b*H + h)*3 + j)*N + n], j = 0: normaliser
  // (den + eps), 1 / 2: exponent shift (row maximum + |x|^2/2, in log2 units, 1/sqrt(M) folded in) of the query / key row
  float* aux = p.state ? p.state + (size_t)B * H * (DH + 1) * Mp : nullptr;

  uint8_t* wimg = smem;
  uint8_t* ximg = wimg + 3 * W_IMG;
  uint8_t* phi = ximg + 3 * X_IMG;   // [half][hi|lo]
  uint8_t* simg = phi + 4 * PHI_IMG;  // 9 chunks of s_ch
  uint8_t* vs = simg + 9 * s_ch;      // [v_hi (2 chunks) | 1 | 0 | v_lo (2 chunks)] of kTokCh

  if (warp == 16) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) mbar_init(&bars[i], i < B_FULL_LONE ? 16 : (i == B_FULL_LONE ? 2 : 1));
    mbar_init_fence();
  }
  {  // W^T of this head as three bf16 levels, K-major rows f
    const float* om = p.omega + (size_t)h * DH * M;
    for (int i = tid; i < Mp * DH; i += kPipeThreads) {
      const int d = i / Mp, f = i % Mp;
      const float w = (f < M) ? __ldg(om + (size_t)d * M + f) : 0.f;
      const __nv_bfloat16 w0 = __float2bfloat16_rn(w);
      const float r1 = w - __bfloat162float(w0);
      const __nv_bfloat16 w1 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 w2 = __float2bfloat16_rn(r1 - __bfloat162float(w1));
      const uint32_t off = off_kmajor(f, d, 8, 2, XL, XS);
      *reinterpret_cast<__nv_bfloat16*>(wimg + off) = w0;
      *reinterpret_cast<__nv_bfloat16*>(wimg + W_IMG + off) = w1;
      *reinterpret_cast<__nv_bfloat16*>(wimg + 2 * W_IMG + off) = w2;
    }
    for (int i = tid; i < (int)(kTokCh / 16); i += kPipeThreads)  // the zero chunk of the [v|1] image never changes
      reinterpret_cast<uint4*>(vs + 3 * kTokCh)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < (int)(s_ch / 16); i += kPipeThreads)  // Z chunk: columns 4..7 stay zero
      reinterpret_cast<uint4*>(simg + CH_Z * s_ch)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  // phase trace (tools/trace_pipe.py): CTA 0, one thread each of a compute warp (segment 0), a lone-token warp (1) and the
  // issue warp (2); compiled in only with -DERV_TRACE
#ifdef ERV_TRACE
  int tr_i = 0;
  const int tr_seg = tid == 0 ? 0 : (tid == 544 ? 1 : (tid == 512 ? 2 : -1));
  auto TR = [&](int tag) {
    if (p.trace != nullptr && blockIdx.x == 0 && tr_seg >= 0 && tr_i < 500) {
      p.trace[tr_seg * 1000 + 2 * tr_i] = tag;
      p.trace[tr_seg * 1000 + 2 * tr_i + 1] = clock64();
      ++tr_i;
    }
  };
#else
  auto TR = [](int) {};
#endif

  if (warp == 16) {
    // ================================================== MMA issue warp ==================================================
    const uint32_t idesc_p = make_idesc(FMT_BF16, 128, Mp, false, false);
    const uint32_t idesc_g2a = make_idesc(FMT_BF16, 128, 48, true, true);
    const uint32_t idesc_g2b = make_idesc(FMT_BF16, 128, 32, true, true);
    const uint32_t idesc_g4a = make_idesc(FMT_BF16, 128, 80, false, true);
    const uint32_t idesc_g4b = make_idesc(FMT_BF16, 128, 80, false, true);
    auto issue_g1 = [&](uint32_t col, uint64_t* done) {  // P = x W^T: x0w0 + x0w1 + x1w0 + x0w2 + x1w1 + x2w0
      if (elect_one()) {
        const uint32_t xa = smem_u32(ximg), wa = smem_u32(wimg);
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const int xl = (t == 2 || t == 4) ? 1 : (t == 5 ? 2 : 0);
          const int wl = (t == 1 || t == 4) ? 1 : (t == 3 ? 2 : 0);
          mma_f16(tm + col, make_desc(xa + xl * X_IMG, XL, XS), make_desc(wa + wl * W_IMG, XL, XS), idesc_p, t > 0);
        }
        commit(done);
      }
      __syncwarp();
    };
    auto issue_g2 = [&](int hb) {  // S(pair, half) = phi_k^T [v|1]
      if (elect_one()) {
        const uint32_t ph = smem_u32(phi) + (uint32_t)hb * 2 * PHI_IMG, pl = ph + PHI_IMG, va = smem_u32(vs);
        for (int sp = 0; sp < 2; ++sp) {
          const uint32_t d = tm + COL_S + (uint32_t)(hb * 2 + sp) * S_STRIDE;
          for (int s = 0; s < ks; ++s) {
            const uint32_t st = (uint32_t)(sp * 4 + s) * 256;
            const uint64_t bd = make_desc(va + st, 128, kTokCh);
            mma_f16(d, make_desc(ph + st, 128, kTokCh), bd, idesc_g2a, s > 0);
            mma_f16(d, make_desc(pl + st, 128, kTokCh), bd, idesc_g2b, true);
          }
        }
        commit(&bars[B_DONE_G20 + hb]);
      }
      __syncwarp();
    };
    auto issue_g4 = [&](int hb) {  // [num_A | num_B | den] += phi_q(half) [S_A | S_B | z](half)
      if (elect_one()) {
        const uint32_t ph = smem_u32(phi) + (uint32_t)hb * 2 * PHI_IMG, pl = ph + PHI_IMG, sa = smem_u32(simg);
        for (int s = 0; s < HF / 16; ++s) {
          const uint64_t bd = make_desc(sa + (uint32_t)(hb * (HF / 16) + s) * 256, 128, s_ch);
          mma_f16(tm + COL_O, make_desc(ph + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_g4a, hb > 0 || s > 0);
          mma_f16(tm + COL_O, make_desc(pl + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_g4b, true);
        }
        commit(&bars[B_DONE_G40 + hb]);
      }
      __syncwarp();
    };
    mbar_wait(&bars[B_FULL_XK], 0);
    fence_after_sync();
    issue_g1(0, &bars[B_DONE_G1K]);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t par = it & 1;
      TR(0);
      mbar_wait(&bars[B_FULL_P0], par);
      fence_after_sync();
      TR(1);
      issue_g2(0);
      TR(2);
      mbar_wait(&bars[B_FULL_XQ], par);
      mbar_wait(&bars[B_O_FREE], par);
      fence_after_sync();
      TR(3);
      issue_g1(COL_PQ, &bars[B_DONE_G1Q]);
      TR(4);
      mbar_wait(&bars[B_FULL_P1], par);
      fence_after_sync();
      TR(5);
      issue_g2(1);
      TR(6);
      mbar_wait(&bars[B_FULL_S0], par);
      mbar_wait(&bars[B_FULL_P2], par);
      fence_after_sync();
      TR(7);
      issue_g4(0);
      TR(8);
      mbar_wait(&bars[B_FULL_XK], par ^ 1);
      fence_after_sync();
      TR(9);
      if (it + 1 < n_it) issue_g1(0, &bars[B_DONE_G1K]);
      TR(10);
      mbar_wait(&bars[B_FULL_S1], par);
      mbar_wait(&bars[B_FULL_P3], par);
      fence_after_sync();
      TR(11);
      issue_g4(1);
      TR(12);
    }
  } else if (warp > 16) {
    // ================================================ lone-token warps =================================================
    // Everything of the 65th token of each pair except its rank-1 term in S: feature rows of the lone query / key (fp32 FMAs
    // against W^T rows held in registers: lane = 4 features of this warp's feature half), [v] of the lone key, and the
    // lone query's read-out against the finished S image.  Runs one group ahead of the compute warps.
    if (lone) {
      const int lw = warp - 17;
      const T* qkv = static_cast<const T*>(p.qkv);
      T* out = static_cast<T*>(p.out);
      float wreg[4][DH];
      {
        const float* om = p.omega + (size_t)h * DH * M;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = HF * lw + lane + 32 * i;
#pragma unroll
          for (int d = 0; d < DH; ++d) wreg[i][d] = (f < M) ? __ldg(om + (size_t)d * M + f) : 0.f;
        }
      }
      auto lbar = [&]() { asm volatile("bar.sync 5, 64;\n" ::: "memory"); };
      constexpr int CPR = DH * (int)sizeof(T) / 16;  // 16-byte chunks per row
      auto fetch_rows = [&](int g, int buf) {  // raw rows of the lone tokens of group g -> this warp's staging buffer
        if (g < ngroups && lane < 6 * CPR) {
          const int ri = lane / CPR, ch = lane % CPR;
          const int sp = ri < 4 ? ri >> 1 : ri - 4, which = ri < 4 ? (ri & 1) : 2, b = 2 * (g / H) + sp;
          if (b < B) cp_async16(&lone_raw[lw][buf][ri][ch * 16], qkv + qkv_off(b, N - 1, which, h, N, H, DH) + ch * (16 / (int)sizeof(T)));
        }
        cp_async_commit();
      };
      auto features = [&](int g, int buf) {
        const int b2 = g / H;
        cp_async_wait_all();
        __syncwarp();
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {  // r = 2 * pair side + (0: query, 1: key)
          const int sp = r >> 1, which = r & 1, b = 2 * b2 + sp;
          const bool ok = b < B;
          float x[DH];
#pragma unroll
          for (int a = 0; a < DH; ++a) x[a] = 0.f;
          if (ok) {
            load_row<T, DH>(reinterpret_cast<const T*>(&lone_raw[lw][buf][r][0]), x);
            prologue_row<DH>(x, p.rot, p.ta, p.tb, h, N - 1, N, p.prescale);
          }
          float n2 = 0.f;
#pragma unroll
          for (int a = 0; a < DH; ++a) n2 = fmaf(x[a], x[a], n2);
          if (lane == 0) lone_n2[lw][r] = 0.5f * n2;
          float m = -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = HF * lw + lane + 32 * i;
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc = fmaf(x[d], wreg[i][d], acc);
            lone_s[buf][which][sp][f] = acc;  // raw projection; becomes the feature once the maximum is known
            if (!PADDED || f < M) m = fmaxf(m, acc);
          }
          if (FAVOR) {
            m = warp_max(m);
            if (lane == 0) lone_mx[lw][r] = m;
          }
          if (which == 1 && lw == 0 && lane < 4)
            st4(&lone_v[buf][sp][4 * lane],
                ok ? ld4(reinterpret_cast<const T*>(&lone_raw[lw][buf][4 + sp][0]) + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f));
        }
        lbar();
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
          const int sp = r >> 1, which = r & 1;
          const bool ok = 2 * b2 + sp < B;
          const float shift = fmaf(fmaxf(lone_mx[0][r], lone_mx[1][r]) + lone_n2[lw][r], kLog2e, -log2_c);
          if (aux && ok && lw == 0 && lane == 0) aux[((size_t)(2 * b2 + sp) * H + h) * 3 * N + (1 + which) * N + N - 1] = FAVOR ? shift : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = HF * lw + lane + 32 * i;
            const float pv = lone_s[buf][which][sp][f];
            float v = FAVOR ? ex2_approx(fmaf(pv, kLog2e, -shift)) : fmaxf(pv, 0.f) * p.inv_sqrt_m;
            if ((PADDED && f >= M) || !ok) v = 0.f;
            lone_s[buf][which][sp][f] = v;
          }
        }
        lbar();  // lone_mx is rewritten by the next call
      };
      fetch_rows(blockIdx.x, 0);
      features(blockIdx.x, 0);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FULL_LONE]);
      fetch_rows(blockIdx.x + gridDim.x, 1);
      int it = 0;
      for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        const int gn = g + gridDim.x;
        TR(0);
        if (gn < ngroups) {
          features(gn, par ^ 1);
          fetch_rows(gn + gridDim.x, par);
        }
        TR(1);
        // both warps wait: the buffers the next call of features() overwrites were read by this group's S epilogue
        mbar_wait(&bars[B_FULL_RED], par);
        TR(2);
        if (lw == 0) {  // the lone queries' output rows from the compute warps' partial sums
          const int sp = lane >> 4, d = lane & 15, c = d >> 3, bb = 2 * (g / H) + sp;
          if (bb < B) {
            float den = kEps, o = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              den += red_s[4 * (2 * sp) + w][8];
              o += red_s[4 * (2 * sp + c) + w][d & 7];
            }
            o /= den;
            if (aux && d == 0) aux[((size_t)bb * H + h) * 3 * N + N - 1] = den;
            T* dst = out + out_off(bb, N - 1, h, N, H, DH) + d;
            if (sizeof(T) == 4) *reinterpret_cast<float*>(dst) = o;
            else *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(o);
          }
        }
        TR(3);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_FULL_LONE]);
      }
    }
  } else {
    // ================================================== compute warps ===================================================
    const int q4 = warp & 3;  // tensor-memory lane quarter = row quarter of this warp
    const uint32_t tm_thr = tm + ((uint32_t)(q4 * 32) << 16);
    const T* qkv = static_cast<const T*>(p.qkv);
    T* out = static_cast<T*>(p.out);
    const uint32_t rowoff = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;
    const uint32_t xoff = (uint32_t)(row >> 3) * XS + (row & 7) * 16;
    const uint32_t phi_thr = smem_u32(phi) + (uint32_t)(part * 4) * kTokCh + rowoff;
    const uint32_t simg_a = smem_u32(simg), vs_a = smem_u32(vs), ximg_a = smem_u32(ximg);
    auto qbar = [&]() { asm volatile("bar.sync %0, 128;\n" ::"r"(1 + q4) : "memory"); };  // the 4 warps that share the rows

    float nx[DH];  // global row in flight: parts 0 / 1 hold the next group's k / v rows, part 2 its q row
#pragma unroll
    for (int a = 0; a < DH; ++a) nx[a] = 0.f;

    auto load_kv_rows = [&](int g) {
      const int b = 2 * (g / H) + side;
      if (part < 2 && b < B && n < Nm) load_row<T, DH>(qkv + qkv_off(b, n, part == 0 ? 1 : 2, h, N, H, DH), nx);
    };
    auto load_q_rows = [&](int g) {
      const int b = 2 * (g / H) + side;
      if (part == 2 && b < B && n < Nm) load_row<T, DH>(qkv + qkv_off(b, n, 0, h, N, H, DH), nx);
    };
    // rotation + scale of a q / k row, |x|^2/2, three-level images for the projection
    auto stage_x_row = [&](int g, int which) {
      const bool valid = 2 * (g / H) + side < B && n < Nm;
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
      n2_s[which][row] = n2;
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        float ch[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[e] = x[8 * c + e];
        store_split8_l3(ximg, ximg + X_IMG, ximg + 2 * X_IMG, xoff + c * XL, ch);
      }
    };
    // key rows -> x images, value rows -> [v|1] image
    auto prologue_kv = [&](int g) {
      if (part == 0) {
        stage_x_row(g, 0);
      } else if (part == 1) {
        const bool valid = 2 * (g / H) + side < B && n < Nm;
#pragma unroll
        for (int c = 0; c < DH / 8; ++c) {
          float ch[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) ch[e] = valid ? nx[8 * c + e] : 0.f;
          store_split8_sa(vs_a + c * kTokCh + rowoff, vs_a + (4 + c) * kTokCh + rowoff, ch);
        }
        sts128(vs_a + 2 * kTokCh + rowoff, valid ? 0x00003F80u : 0u, 0u, 0u, 0u);  // bf16 1.0
      }
    };
    // out = num / (den + eps) of the group whose accumulator sits in tensor memory; thread = (row, 4 of the 16 columns)
    auto out_epilogue = [&](int g) {
      const int b = 2 * (g / H) + side;
      float a0[4], a1[4], dn[4];
      tmem_ld4(tm_thr + COL_O + 16 * side + 4 * part, a0);
      tmem_ld4(tm_thr + COL_O + O_LO + 16 * side + 4 * part, a1);
      tmem_ld4(tm_thr + COL_O + O_DEN, dn);
      if (b < B && n < Nm) {
        const float den = (side ? dn[1] + dn[3] : dn[0] + dn[2]) + kEps;
        if (aux && part == 0) aux[((size_t)b * H + h) * 3 * N + n] = den;
        st4(out + out_off(b, n, h, N, H, DH) + 4 * part,
            make_float4((a0[0] + a1[0]) / den, (a0[1] + a1[1]) / den, (a0[2] + a1[2]) / den, (a0[3] + a1[3]) / den));
      }
    };

    // ---- preamble: first group's rows
    load_kv_rows(blockIdx.x);
    prologue_kv(blockIdx.x);
    warp_arrive(&bars[B_FULL_XK]);
    load_q_rows(blockIdx.x);

    float shift = 0.f, scale = 0.f;
    float lacc[9];  // lone query read-out partials of this thread's 8 columns (+ the normaliser)
#pragma unroll
    for (int j = 0; j < 9; ++j) lacc[j] = 0.f;
    int g_prev = -1;
    int it = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int gn = g + gridDim.x;
      const int b2 = g / H;
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        const int isq = u >> 1, hb = u & 1;
        if (u >= 2) {
          // ---- S half (tensor memory, lanes = features) -> S image for G4 (with the z columns), saved state, rank-1 term of
          // the lone key.  thread = (feature, pair side, 8 of the 16 columns)
          if (u == 3 && gn < ngroups) load_kv_rows(gn);
          if (u == 2 && lone) mbar_wait(&bars[B_FULL_LONE], par);
          TR(10 * u + 0);
          mbar_wait(&bars[B_DONE_G20 + hb], par);
          fence_after_sync();
          TR(10 * u + 1);
          const int sp = part >> 1, c = part & 1, f = hb * HF + row;
          const uint32_t sb = tm_thr + COL_S + (uint32_t)(hb * 2 + sp) * S_STRIDE;
          uint32_t r0[8], r1[8], r2[4];
          tmem_ld8_nowait(sb + 8 * c, r0);
          tmem_ld8_nowait(sb + 32 + 8 * c, r1);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                       : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]) : "r"(sb + 16) : "memory");
          tmem_wait_ld8(r0);
          tmem_wait_ld8(r1);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(r2[0]), "+r"(r2[1]), "+r"(r2[2]), "+r"(r2[3])::"memory");
          float sv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sv[i] = __uint_as_float(r0[i]) + __uint_as_float(r1[i]);
          float z = __uint_as_float(r2[0]);
          float pq = 0.f;
          if (lone) {  // rank-1 term of the last key
            const float pk = lone_s[par][1][sp][f];
            pq = lone_s[par][0][sp][f];
            const float4 va = ld4(&lone_v[par][sp][8 * c]), vb = ld4(&lone_v[par][sp][8 * c + 4]);
            sv[0] = fmaf(pk, va.x, sv[0]); sv[1] = fmaf(pk, va.y, sv[1]); sv[2] = fmaf(pk, va.z, sv[2]); sv[3] = fmaf(pk, va.w, sv[3]);
            sv[4] = fmaf(pk, vb.x, sv[4]); sv[5] = fmaf(pk, vb.y, sv[5]); sv[6] = fmaf(pk, vb.z, sv[6]); sv[7] = fmaf(pk, vb.w, sv[7]);
            z += pk;
          }
          const uint32_t fo = (uint32_t)(f >> 3) * 128 + (f & 7) * 16;
          store_split8_sa(simg_a + (uint32_t)(2 * sp + c) * s_ch + fo, simg_a + (CH_LO + 2 * sp + c) * s_ch + fo, sv);
          if (c == 0) {  // z as bf16 hi / lo into the Z chunk: columns sp (hi) and 2 + sp (lo)
            const uint32_t zh = pack_bf16x2_rn(z, 0.f) & 0xffffu;
            const uint32_t zl = pack_bf16x2_rn(z - __uint_as_float(zh << 16), 0.f) & 0xffffu;
            sts16(simg_a + CH_Z * s_ch + fo + 2 * sp, zh);
            sts16(simg_a + CH_Z * s_ch + fo + 4 + 2 * sp, zl);
          }
          if (p.state != nullptr && 2 * b2 + sp < B) {  // [S|z] of this pair, d-major: a warp writes 128-byte lines
            float* so = p.state + (((size_t)(2 * b2 + sp) * H + h) * (DH + 1) + 8 * c) * Mp + f;
#pragma unroll
            for (int i = 0; i < 8; ++i) so[i * Mp] = sv[i];
            if (c == 0) so[DH * Mp] = z;
          }
          warp_arrive(&bars[B_FULL_S0 + hb]);
          if (lone) {  // the last query's read-out against the finished S
#pragma unroll
            for (int i = 0; i < 8; ++i) lacc[i] = fmaf(pq, sv[i], lacc[i]);
            lacc[8] = fmaf(pq, z, lacc[8]);
          }
          TR(10 * u + 2);
          if (u == 3) {
            if (lone) {
              float v16[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v16[j] = j < 9 ? lacc[j] : 0.f;
              const float t = warp_sum16(v16);
              if (!(lane & 1) && (lane >> 1) < 9) red_s[warp][lane >> 1] = t;
#pragma unroll
              for (int j = 0; j < 9; ++j) lacc[j] = 0.f;
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars[B_FULL_RED]);
            }
            // ---- next group's key / value rows
            if (gn < ngroups) prologue_kv(gn);
            warp_arrive(&bars[B_FULL_XK]);
            if (gn < ngroups) load_q_rows(gn);
            TR(10 * u + 3);
          }
        }
        if (hb == 0) {
          // ---- a fresh projection: row maximum (keys: also stage this group's query rows for their projection)
          TR(10 * u + 5);
          mbar_wait(&bars[isq ? B_DONE_G1Q : B_DONE_G1K], par);
          fence_after_sync();
          TR(10 * u + 6);
          if (!isq) {
            if (part == 2) stage_x_row(g, 1);  // the x images are free: G1 of the keys has completed
            warp_arrive(&bars[B_FULL_XQ]);
          }
          if (FAVOR) {
            float m = -INFINITY;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t r[32];
              tmem_ld32_nowait(tm_thr + u * HF + hh * HF + part * 32, r);
              tmem_wait_ld32(r);
              if (PADDED) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (hh * HF + part * 32 + i < M) m = fmaxf(m, __uint_as_float(r[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
              }
            }
            ex_s[part][row] = m;
          }
          TR(10 * u + 7);
          qbar();  // row maxima, n2_s
          TR(10 * u + 8);
          float mx = 0.f;
          if (FAVOR) mx = fmaxf(fmaxf(ex_s[0][row], ex_s[1][row]), fmaxf(ex_s[2][row], ex_s[3][row]));
          const float n2 = n2_s[isq][row];
          // phi = exp(P - mx - n2)/sqrt(M) = 2^(P*log2e - (mx + n2)*log2e + log2(1/sqrt(M)))
          shift = fmaf(mx + n2, kLog2e, -log2_c);
          if (aux && part == 0 && n2 < INFINITY) aux[((size_t)(2 * b2 + side) * H + h) * 3 * N + (isq ? 1 : 2) * N + n] = FAVOR ? shift : 0.f;
          scale = (n2 < INFINITY) ? p.inv_sqrt_m : 0.f;
        }
        if (it > 0) {
          if (u == 0) {  // the feature half about to be overwritten was last read by G4 (half 0) of the previous group
            mbar_wait(&bars[B_DONE_G40], par ^ 1);
          } else if (u == 1) {
            mbar_wait(&bars[B_DONE_G41], par ^ 1);
            fence_after_sync();
            out_epilogue(g_prev);
          }
        }
        if (u == 1) warp_arrive(&bars[B_O_FREE]);
        TR(10 * u + 9);
        // ---- one feature half: P (tensor memory) -> phi -> hi/lo bf16 images
        {
          uint32_t r[32];
          tmem_ld32_nowait(tm_thr + u * HF + part * 32, r);
          tmem_wait_ld32(r);
          const uint32_t pa = phi_thr + (uint32_t)hb * 2 * PHI_IMG;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v[8];
            if (FAVOR) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = ex2_approx(fmaf(__uint_as_float(r[8 * c + i]), kLog2e, -shift));
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = fmaxf(__uint_as_float(r[8 * c + i]), 0.f) * scale;
            }
            if (PADDED) {
              const int f0 = hb * HF + part * 32 + c * 8;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (f0 + i >= M) v[i] = 0.f;
            }
            store_split8_sa(pa + c * kTokCh, pa + PHI_IMG + c * kTokCh, v);
          }
          warp_arrive(&bars[B_FULL_P0 + u]);
          TR(50 + u);
        }
      }
      g_prev = g;
    }
    // ---- drain: output rows of the last group
    if (g_prev >= 0) {
      mbar_wait(&bars[B_DONE_G41], (uint32_t)((it - 1) & 1));
      fence_after_sync();
      out_epilogue(g_prev);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tm, 512);
}

size_t la_pipe_smem_bytes() {
  return 3 * (256 * 16 * 2) + 3 * (128 * 16 * 2) + 4 * 16 * (size_t)kTokCh + 9 * (size_t)(256 / 8) * 128 + 6 * (size_t)kTokCh;
}

bool la_pipe_eligible(int N, int DH, int M) {
  static const bool disabled = getenv("ERV_DISABLE_PIPE") != nullptr;
  return !disabled && DH == 16 && M > 128 && M <= 256 && N >= 33 && N <= 65;
}

int la_pipe_forward(const void* qkv, void* out, const float* omega, int B, int N, int H, int M, int kind, int rot,
                    const float* ta, const float* tb, int dtype, float* state, cudaStream_t st) {
  LaTcArgs a;
  a.qkv = qkv; a.out = out; a.omega = omega; a.ta = ta; a.tb = tb;
  a.B = B; a.N = N; a.H = H; a.M = M; a.Mp16 = 256; a.kind = kind; a.rot = rot;
  a.prescale = (float)pow(16.0, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.state = state;
  a.trace = g_trace;
  const size_t smem = la_pipe_smem_bytes();
  const int ngroups = ((B + 1) / 2) * H;
  int grid = (kNumSMs / H) * H;  // multiple of H: each CTA stays on one head (W images staged once)
  if (grid < H) grid = H;
  if (grid > ngroups) grid = ngroups;
  const bool favor = kind == ERV_FEAT_FAVOR, padded = M < 256;
#define PIPE_LAUNCH(TT, FV, PD)                                              \
  do {                                                                       \
    ERV_CUDA(allow_smem(la_pipe_fwd_kernel<TT, FV, PD>, smem));              \
    la_pipe_fwd_kernel<TT, FV, PD><<<grid, kPipeThreads, smem, st>>>(a);     \
  } while (0)
#define PIPE_LAUNCH_T(TT)                                                    \
  do {                                                                       \
    if (favor) { if (padded) PIPE_LAUNCH(TT, true, true); else PIPE_LAUNCH(TT, true, false); } \
    else { if (padded) PIPE_LAUNCH(TT, false, true); else PIPE_LAUNCH(TT, false, false); }     \
  } while (0)
  if (dtype == ERV_F32) PIPE_LAUNCH_T(float); else PIPE_LAUNCH_T(__nv_bfloat16);
#undef PIPE_LAUNCH_T
#undef PIPE_LAUNCH
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv
