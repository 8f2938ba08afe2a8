// Tensor-core backward of FAVOR+/ReLU linear attention (head_dim 8/16, num_features <= 256).
//
// Same tile structure as the forward (erv_linattn_tc.cu): one persistent CTA per SM, 512 threads = 4 per token row
// of a 128-token tile, thread (row, part) owns a quarter of the row's features in registers.  Three sweeps per
// (batch, head) pair:
//   K1   S[f][.]   = sum_t phi_k[t][f] [v|1][t]                         tcgen05, accumulated in TMEM
//   Q    den = phi_q . z ; a = [dO/(den+eps) | -(dO.O)/(den+eps)]       registers
//        dS[f][.] += sum_t phi_q[t][f] a[t]                             tcgen05 (TMEM, reuses S's columns)
//        dphi_q    = a [S|z]^T (N = M)                                  tcgen05, into the P columns
//        G = dphi (.) dphi/dP ; dx = G [W^T | 1] - x rowsum(G)          registers + one 4-way smem reduction
//   K2   dv = phi_k dS ; dphi_k = [v|1] [dS|dz]^T ; G ; dx              same pieces
// Measured on B200 every tcgen05.mma costs >= 96 cycles however narrow (profiles/r01_tcgen05_mma_cost.md), so only the
// wide products (projection, token-reductions, dphi) go to the tensor core; the 17-column products (den, dv, dx) are
// fp32 FMAs straight from the feature registers.
#include "erv_tc_common.cuh"

namespace erv {

struct LaTcBwdArgs {
  const void* qkv;
  const void* out;
  const void* dout;
  void* dqkv;
  const float* omega;
  const float* ta;
  const float* tb;
  float* dg_part;  // [H][slots][N][DH], circulant only
  const float* state;  // optional [B*H][DH+1][Mp]: [S|z] saved by la_tc_fwd_kernel; the K1 sweep is skipped when present
  int B, N, H, M, Mp, kind, rot, slots;
  float prescale, inv_sqrt_m;
  long long* trace;  // optional: CTA 0 / thread 0 stamps clock64() at phase boundaries (erv_debug_set_trace)
};

long long* g_trace = nullptr;  // erv_debug_set_trace(); also read by erv_linattn_tc2_bwd.cu

__device__ __forceinline__ void unpack8(const uint8_t* hi_img, const uint8_t* lo_img, uint32_t off, float (&v)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi_img + off);
  const uint4 l = *reinterpret_cast<const uint4*>(lo_img + off);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    v[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
  }
}

// NRB feature halves (row blocks of S), CPH 8-feature chunks per thread and half: Mp = NRB * CPH * 32 is compile time
// so the per-thread feature registers are statically indexed.
template <typename T, int DH, int NRB, int CPH>
__global__ void __launch_bounds__(kTcThreads, 1) la_tc_bwd_kernel(const LaTcBwdArgs p) {
  using C = TcCfg<DH>;
  // bf16 inputs (autocast, budget 2e-2): the contractions use the hi images only -- one bf16 product instead of the three
  // split terms the fp32 path needs for 1e-4; the projection keeps 3xTF32 (it feeds exp)
  constexpr int kTerms = sizeof(T) == 2 ? 1 : 3;
  constexpr int ND = C::ND, RW = DH + 4;
  constexpr uint32_t COL_P = 0, COL_S = 256;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_base_s;
  __shared__ float n2_s[128];
  __shared__ float ex_s[4][128];  // 4-way exchanges: row max, den
  __shared__ float z_s[256];

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, part = tid >> 7;
  constexpr int nrb = NRB, HF = CPH * 32, FPH = CPH * 8, Mp = NRB * HF, NC = NRB * CPH;
  const int M = p.M, N = p.N;
  const uint32_t wbytes = tc_w_bytes(DH, Mp);
  const uint32_t s_ch = (uint32_t)(Mp / 8) * 128;
  const uint32_t avbytes = (uint32_t)(ND / 8) * kTokCh;
  uint8_t* wh = smem;
  uint8_t* wl = wh + wbytes;
  uint8_t* xh = wl + wbytes;
  uint8_t* xl = xh + C::X_BYTES;
  uint8_t* phi1 = xl + C::X_BYTES;          // one feature half [128 tokens x 128 features], hi
  uint8_t* phi2 = phi1 + 16 * kTokCh;       // lo
  float* red = reinterpret_cast<float*>(phi1);  // [3][128][RW] fp32, aliases the phi images when they are idle
  uint8_t* av1 = phi2 + 16 * kTokCh;        // [v|1] or [dnum|dden] rows, hi
  uint8_t* av2 = av1 + avbytes;
  uint8_t* s1 = av2 + avbytes;              // [S|z] then [dS|dz]: byte(f, j) = (j/8)*s_ch + (f/8)*128 + (f%8)*16 + (j%8)*2
  uint8_t* s2 = s1 + (uint32_t)(ND / 8) * s_ch;
  float* w32 = reinterpret_cast<float*>(s2 + (uint32_t)(ND / 8) * s_ch);  // [Mp][RW] fp32: W^T rows, column DH = 1 (rowsum)
  float* ds32 = w32 + Mp * RW;                                            // [Mp][RW] fp32: dS rows (written after the Q sweep)

  // chunk c of this thread: half c / CPH, local chunk c % CPH
  auto feat0 = [&](int c) { return (c / CPH) * HF + part * FPH + (c % CPH) * 8; };

  for (int i = tid; i < (int)(2 * avbytes / 16); i += kTcThreads)  // columns DH+1.. of the a/[v|1] images stay zero
    reinterpret_cast<uint4*>(av1)[i] = make_uint4(0, 0, 0, 0);
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
  int tr_i = 0;
  // phase trace (tools/trace_bwd.py): compiled in only with -DERV_TRACE, so the production kernel carries no checks
#ifdef ERV_TRACE
  auto TR = [&](int tag) {
    if (p.trace != nullptr && blockIdx.x == 0 && tid == 0 && tr_i < 1000) {
      p.trace[2 * tr_i] = tag;
      p.trace[2 * tr_i + 1] = clock64();
      ++tr_i;
    }
  };
#else
  auto TR = [](int) {};
  (void)tr_i;
#endif

  const T* qkv = static_cast<const T*>(p.qkv);
  const T* outp = static_cast<const T*>(p.out);
  const T* dout = static_cast<const T*>(p.dout);
  T* dqkv = static_cast<T*>(p.dqkv);
  const size_t tok_stride = (size_t)3 * p.H * DH, out_stride = (size_t)p.H * DH;
  const uint32_t idesc_p = make_idesc(FMT_TF32, 128, Mp, false, false);
  const uint32_t idesc_acc = make_idesc(FMT_BF16, 128, ND, true, true);   // S / dS += phi^T rows
  const uint32_t idesc_dphi = make_idesc(FMT_BF16, 128, Mp, false, false);  // dphi = rows [S|z]^T
  const float kLog2e = 1.4426950408889634f;
  const float log2_c = log2f(p.inv_sqrt_m);
  const bool favor = p.kind == ERV_FEAT_FAVOR;
  int cur_h = -1;

  for (int pair = blockIdx.x; pair < p.B * p.H; pair += gridDim.x) {
    const int b = pair / p.H, h = pair % p.H;
    if (h != cur_h) {
      cur_h = h;
      const float* om = p.omega + (size_t)h * DH * M;
      for (int i = tid; i < Mp * DH; i += kTcThreads) {
        const int d = i / Mp, f = i % Mp;
        const float w = (f < M) ? __ldg(om + (size_t)d * M + f) : 0.f;
        const float hi = to_tf32(w), lo = to_tf32(w - hi);
        const uint32_t off = off_kmajor(f, d, 4, 4, C::X_LBO, C::X_SBO);
        *reinterpret_cast<float*>(wh + off) = hi;
        *reinterpret_cast<float*>(wl + off) = lo;
        w32[f * RW + d] = w;
      }
      for (int i = tid; i < Mp * 4; i += kTcThreads) {
        const int f = i >> 2, j = DH + (i & 3);
        w32[f * RW + j] = (j == DH && f < M) ? 1.f : 0.f;
      }
    }
    float* dg_slot = (p.rot == ERV_ROT_CIRCULANT && p.dg_part)
                         ? p.dg_part + ((size_t)h * p.slots + blockIdx.x / p.H) * N * DH : nullptr;

    if (p.state != nullptr) {  // [S|z] from the forward instead of the K1 sweep: straight into the bf16 images and z_s
      if (part < nrb && row < HF) {
        const int f = part * HF + row;
        const float* so = p.state + (size_t)pair * (DH + 1) * Mp + f;
        float sv[DH + 1];
#pragma unroll
        for (int d = 0; d <= DH; ++d) sv[d] = __ldg(so + (size_t)d * Mp);
#pragma unroll
        for (int c = 0; c < ND / 8; ++c) {
          float ch[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) ch[e] = (8 * c + e <= DH) ? sv[(8 * c + e <= DH) ? 8 * c + e : 0] : 0.f;
          store_split8(s1, s2, c * s_ch + (f >> 3) * 128 + (f & 7) * 16, ch);
        }
        z_s[f] = sv[DH];
      }
      fence_smem_to_async();
      fence_before_sync();
      __syncthreads();
    }
    for (int pass = p.state != nullptr ? 1 : 0; pass < 3; ++pass) {  // 0: K1 (build S), 1: Q (dS, dq), 2: K2 (dv, dk)
      const int which = (pass == 1) ? 0 : 1;
      const T* xb = qkv + qkv_off(b, 0, which, h, N, p.H, DH);
      const T* vb = qkv + qkv_off(b, 0, 2, h, N, p.H, DH);
      T* dxb = dqkv + qkv_off(b, 0, which, h, N, p.H, DH);
      T* dvb = dqkv + qkv_off(b, 0, 2, h, N, p.H, DH);
      for (int n0 = 0; n0 < N; n0 += 128) {
        const int n = n0 + row;
        const bool valid = n < N;
        const int nt16 = (min(128, N - n0) + 15) & ~15;
        const bool warp_live = (row & ~31) < nt16;
        TR(pass * 100 + 0);
        // ---- step 1: operand images; part 0 keeps the prepared row, part 1 the value / gradient row
        float rowv[DH];  // part 0: prepared q or k row; part 1: v (K passes) or dO (Q pass)
        float dot = 0.f;
        if (part == 0) {
          float n2 = INFINITY;
          if (valid) {
            load_row<T, DH>(xb + (size_t)n * tok_stride, rowv);
            prologue_row<DH, true>(rowv, p.rot, p.ta, p.tb, h, n, N, p.prescale);
            n2 = 0.f;
#pragma unroll
            for (int a = 0; a < DH; ++a) n2 = fmaf(rowv[a], rowv[a], n2);
            n2 *= 0.5f;
          } else {
#pragma unroll
            for (int a = 0; a < DH; ++a) rowv[a] = 0.f;
          }
          n2_s[row] = n2;
          store_x_images<DH>(xh, xl, rowv, row);
        } else if (part == 1) {
#pragma unroll
          for (int d = 0; d < DH; ++d) rowv[d] = 0.f;
          if (valid) {
            if (pass == 1) {
              float o[DH];
              load_row<T, DH>(dout + out_off(b, n, h, N, p.H, DH), rowv);
              load_row<T, DH>(outp + out_off(b, n, h, N, p.H, DH), o);
#pragma unroll
              for (int d = 0; d < DH; ++d) dot = fmaf(rowv[d], o[d], dot);
            } else {
              load_row<T, DH>(vb + (size_t)n * tok_stride, rowv);
            }
          }
          if (pass != 1 && warp_live) {  // [v | 1] image
#pragma unroll
            for (int c = 0; c < DH / 8; ++c) {
              float ch[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) ch[e] = rowv[8 * c + e];
              store_split8(av1, av2, c * kTokCh + (row >> 3) * 128 + (row & 7) * 16, ch);
            }
            const float ones[8] = {valid ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_split8(av1, av2, (DH / 8) * kTokCh + (row >> 3) * 128 + (row & 7) * 16, ones);
          }
        }
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
        TR(pass * 100 + 1);
        // ---- G1: P = x W^T (3xTF32)
        if (warp == 0 && elect_one()) {
          fence_after_sync();
          bool acc = false;
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint8_t* xa = (term == 1) ? xl : xh;
            const uint8_t* wb = (term == 2) ? wl : wh;
#pragma unroll
            for (int s = 0; s < DH / 8; ++s) {
              mma_tf32(tm + COL_P, make_desc(smem_u32(xa) + s * 2 * C::X_LBO, C::X_LBO, C::X_SBO),
                       make_desc(smem_u32(wb) + s * 2 * C::X_LBO, C::X_LBO, C::X_SBO), idesc_p, acc);
              acc = true;
            }
          }
          commit(&bar_a);
        }
        mbar_wait(&bar_a, ph_a);
        ph_a ^= 1;
        fence_after_sync();
        TR(pass * 100 + 2);
        // ---- P -> registers, row max, phi (kept in pr as fp32 bit patterns)
        uint32_t pr[NC][8];
        if (warp_live) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            tmem_ld8_nowait(tm + lane_off + COL_P + feat0(c), pr[c]);
#pragma unroll
          for (int c = 0; c < NC; ++c)
            tmem_wait_ld8(pr[c]);
        }
        float mx = 0.f;
        if (favor) {
          float m_part = -INFINITY;
          if (warp_live) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (feat0(c) + i < M) m_part = fmaxf(m_part, __uint_as_float(pr[c][i]));
              }
          }
          ex_s[part][row] = m_part;
          fence_before_sync();
          __syncthreads();
          mx = fmaxf(fmaxf(ex_s[0][row], ex_s[1][row]), fmaxf(ex_s[2][row], ex_s[3][row]));
        } else {
          fence_before_sync();
          __syncthreads();  // all P reads are done: the P columns may be overwritten (dphi)
        }
        if (pass == 2 && warp == 0 && elect_one()) {  // K2: dphi_k = [v|1] [dS|dz]^T can start as soon as P has been consumed
          fence_after_sync();
          bool acc = false;
          for (int term = 0; term < kTerms; ++term) {
            const uint8_t* a_img = (term == 2) ? av2 : av1;
            const uint8_t* b_img = (term == 1) ? s2 : s1;
            for (int s = 0; s < ND / 16; ++s) {
              mma_f16(tm + COL_P, make_desc(smem_u32(a_img) + (uint32_t)s * 2 * kTokCh, kTokCh, 128),
                      make_desc(smem_u32(b_img) + (uint32_t)s * 2 * s_ch, s_ch, 128), idesc_dphi, acc);
              acc = true;
            }
          }
          commit(&bar_b);
        }
        {
          const float shift = fmaf(mx + n2_s[row], kLog2e, -log2_c);
          const float scale = valid ? p.inv_sqrt_m : 0.f;
          if (warp_live) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              {
                const int f0 = feat0(c);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float pv = __uint_as_float(pr[c][i]);
                  float v = favor ? ex2_approx(fmaf(pv, kLog2e, -shift)) : fmaxf(pv, 0.f) * scale;
                  if (f0 + i >= M) v = 0.f;
                  pr[c][i] = __float_as_uint(v);
                }
              }
          } else {
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
              for (int i = 0; i < 8; ++i) pr[c][i] = 0u;
          }
        }
        TR(pass * 100 + 3);
        // stores one feature half of the values held in pr into the phi images
        auto store_half = [&](int hb) {
          if (!warp_live) return;
#pragma unroll
          for (int c = 0; c < NC; ++c)
            if (c / CPH == hb) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(pr[c][i]);
              store_split8(phi1, phi2, (uint32_t)(part * CPH + (c % CPH)) * kTokCh + (row >> 3) * 128 + (row & 7) * 16, v);
            }
        };
        // D[COL_S + hb*ND] (+)= phi_half^T rows  (rows image = av1/av2)
        auto accumulate_half = [&](int hb, bool first) {
          fence_smem_to_async();
          fence_before_sync();
          __syncthreads();
          if (warp == 0 && elect_one()) {
            fence_after_sync();
            bool acc = !first;
            for (int term = 0; term < kTerms; ++term) {
              const uint8_t* a_img = (term == 2) ? phi2 : phi1;
              const uint8_t* b_img = (term == 1) ? av2 : av1;
              for (int s = 0; s < nt16 / 16; ++s) {
                mma_f16(tm + COL_S + hb * ND, make_desc(smem_u32(a_img) + s * 256, 128, kTokCh),
                        make_desc(smem_u32(b_img) + s * 256, 128, kTokCh), idesc_acc, acc);
                acc = true;
              }
            }
          }
        };
        // 4-way reduction over the threads of a row; result valid in part 0
        auto reduce_rows = [&](float (&acc)[RW]) {
          if (part > 0) {
#pragma unroll
            for (int j = 0; j < RW; j += 4)
              st4(red + ((size_t)(part - 1) * 128 + row) * RW + j, make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]));
          }
          __syncthreads();
          if (part == 0) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
              for (int j = 0; j < RW; j += 4) {
                const float4 v = ld4(red + ((size_t)q * 128 + row) * RW + j);
                acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
              }
          }
        };
        // G (in pr) -> gradient wrt the raw q/k row, written to global memory
        auto input_gradient = [&]() {
          float acc[RW];
#pragma unroll
          for (int j = 0; j < RW; ++j) acc[j] = 0.f;
          if (warp_live) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              {
                const int f0 = feat0(c);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float g = __uint_as_float(pr[c][i]);
                  const float* wr = w32 + (f0 + i) * RW;
#pragma unroll
                  for (int cd = 0; cd < DH / 4; ++cd) {
                    const float4 a = ld4(wr + 4 * cd);
                    acc[4 * cd] = fmaf(g, a.x, acc[4 * cd]);
                    acc[4 * cd + 1] = fmaf(g, a.y, acc[4 * cd + 1]);
                    acc[4 * cd + 2] = fmaf(g, a.z, acc[4 * cd + 2]);
                    acc[4 * cd + 3] = fmaf(g, a.w, acc[4 * cd + 3]);
                  }
                  acc[DH] += g;  // ones column of [W^T|1] (padded features carry G = 0): no fifth shared-memory load
                }
              }
          }
          reduce_rows(acc);
          if (part == 0 && valid) {
            float dy[DH];
#pragma unroll
            for (int d = 0; d < DH; ++d)
              dy[d] = (favor ? acc[d] - rowv[d] * acc[DH] : acc[d]) * p.prescale;
            float dxr[DH];
            if (p.rot == ERV_ROT_ROPE) {
#pragma unroll
              for (int m = 0; m < DH / 2; ++m) {
                const float c = __ldg(p.ta + (size_t)n * (DH / 2) + m), s = __ldg(p.tb + (size_t)n * (DH / 2) + m);
                dxr[2 * m] = dy[2 * m] * c + dy[2 * m + 1] * s;
                dxr[2 * m + 1] = dy[2 * m + 1] * c - dy[2 * m] * s;
              }
            } else if (p.rot == ERV_ROT_CIRCULANT) {
              float g[DH];
              load_row<float, DH>(p.ta + ((size_t)h * N + n) * DH, g);
#pragma unroll
              for (int bq = 0; bq < DH; ++bq) {
                float a = 0.f;
#pragma unroll
                for (int aa = 0; aa < DH; ++aa) a = fmaf(g[(aa - bq) & (DH - 1)], dy[aa], a);
                dxr[bq] = a;
              }
              if (dg_slot != nullptr && n >= 1) {
                float xr[DH];
                load_row<T, DH>(xb + (size_t)n * tok_stride, xr);
#pragma unroll
                for (int m = 0; m < DH; ++m) {
                  float a = 0.f;
#pragma unroll
                  for (int aa = 0; aa < DH; ++aa) a = fmaf(dy[aa], xr[(aa - m) & (DH - 1)], a);
                  dg_slot[(size_t)n * DH + m] += a;  // slot private to this CTA, row private to this thread
                }
              }
            } else {
#pragma unroll
              for (int d = 0; d < DH; ++d) dxr[d] = dy[d];
            }
#pragma unroll
            for (int c = 0; c < DH / 4; ++c)
              st4(dxb + (size_t)n * tok_stride + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
          }
        };
        // dphi (TMEM, P columns) -> G = dphi (.) dphi/dP, in place in pr
        auto load_dphi_to_g = [&]() {
          if (!warp_live) return;
#pragma unroll
          for (int c = 0; c < NC; ++c)
            {
              uint32_t r[8];
              tmem_ld8_nowait(tm + lane_off + COL_P + feat0(c), r);
              tmem_wait_ld8(r);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float ph_v = __uint_as_float(pr[c][i]), dph = __uint_as_float(r[i]);
                const float g = favor ? dph * ph_v : (ph_v > 0.f ? dph * p.inv_sqrt_m : 0.f);
                pr[c][i] = __float_as_uint(g);
              }
            }
        };

        if (pass == 0) {
          // ---- K1: S[hb] += phi_k^T [v|1]
          for (int hb = 0; hb < nrb; ++hb) {
            store_half(hb);
            accumulate_half(hb, n0 == 0);
            if (warp == 0 && elect_one()) commit(&bar_b);
            mbar_wait(&bar_b, ph_b);
            ph_b ^= 1;
            fence_after_sync();
            TR(4 + hb);
          }
        } else if (pass == 1) {
          // ---- Q: den, a, dS += phi_q^T a, dphi_q = a [S|z]^T
          float den_part = 0.f;
          if (warp_live) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              {
#pragma unroll
                for (int i = 0; i < 8; ++i) den_part = fmaf(__uint_as_float(pr[c][i]), z_s[feat0(c) + i], den_part);
              }
          }
          __syncthreads();  // row-max exchange fully consumed before ex_s is reused
          ex_s[part][row] = den_part;
          __syncthreads();
          if (part == 1 && warp_live) {
            const float r = 1.0f / ((ex_s[0][row] + ex_s[1][row]) + (ex_s[2][row] + ex_s[3][row]) + kEps);
#pragma unroll
            for (int c = 0; c < DH / 8; ++c) {
              float ch[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) ch[e] = rowv[8 * c + e] * r;
              store_split8(av1, av2, c * kTokCh + (row >> 3) * 128 + (row & 7) * 16, ch);
            }
            const float dd[8] = {-dot * r, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_split8(av1, av2, (DH / 8) * kTokCh + (row >> 3) * 128 + (row & 7) * 16, dd);
          }
          TR(104);
          for (int hb = 0; hb < nrb; ++hb) {
            store_half(hb);
            accumulate_half(hb, n0 == 0);
            if (warp == 0 && elect_one()) {
              if (hb == nrb - 1) {  // dphi_q for the whole row, into the (consumed) P columns
                bool acc = false;
                for (int term = 0; term < kTerms; ++term) {
                  const uint8_t* a_img = (term == 2) ? av2 : av1;
                  const uint8_t* b_img = (term == 1) ? s2 : s1;
                  for (int s = 0; s < ND / 16; ++s) {
                    mma_f16(tm + COL_P, make_desc(smem_u32(a_img) + (uint32_t)s * 2 * kTokCh, kTokCh, 128),
                            make_desc(smem_u32(b_img) + (uint32_t)s * 2 * s_ch, s_ch, 128), idesc_dphi, acc);
                    acc = true;
                  }
                }
              }
              commit(&bar_b);
            }
            mbar_wait(&bar_b, ph_b);
            ph_b ^= 1;
            fence_after_sync();
            TR(105 + hb);
          }
          load_dphi_to_g();
          fence_before_sync();
          TR(107);
          input_gradient();  // red aliases the phi images: the accumulate MMAs above have completed
          __syncthreads();
          TR(108);
        } else {
          // ---- K2: dv = phi_k dS ; dphi_k (already issued) ; dk
          float acc[RW];
#pragma unroll
          for (int j = 0; j < RW; ++j) acc[j] = 0.f;
          if (warp_live) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              {
                const int f0 = feat0(c);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float ph_v = __uint_as_float(pr[c][i]);
                  const float* dr = ds32 + (f0 + i) * RW;
#pragma unroll
                  for (int cd = 0; cd < DH / 4; ++cd) {
                    const float4 a = ld4(dr + 4 * cd);
                    acc[4 * cd] = fmaf(ph_v, a.x, acc[4 * cd]);
                    acc[4 * cd + 1] = fmaf(ph_v, a.y, acc[4 * cd + 1]);
                    acc[4 * cd + 2] = fmaf(ph_v, a.z, acc[4 * cd + 2]);
                    acc[4 * cd + 3] = fmaf(ph_v, a.w, acc[4 * cd + 3]);
                  }
                }
              }
          }
          reduce_rows(acc);
          if (part == 0 && valid) {
#pragma unroll
            for (int c = 0; c < DH / 4; ++c)
              st4(dvb + (size_t)n * tok_stride + 4 * c, make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]));
          }
          TR(204);
          mbar_wait(&bar_b, ph_b);
          ph_b ^= 1;
          fence_after_sync();
          load_dphi_to_g();
          fence_before_sync();
          __syncthreads();  // dv reduction buffer fully consumed
          TR(205);
          input_gradient();
          __syncthreads();
          TR(206);
        }
      }
      // ---- end of sweep: move the TMEM accumulator (S after K1, dS after Q) into the bf16 images
      if (pass < 2) {
        if (part < nrb) {  // warp-uniform
          float sv[32];
          tmem_ld32(tm + lane_off + COL_S + part * ND, sv);
          const int f = part * HF + row;
          if (row < HF) {
#pragma unroll
            for (int c = 0; c < ND / 8; ++c) {
              float ch[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) ch[e] = (8 * c + e <= DH) ? sv[8 * c + e] : 0.f;
              store_split8(s1, s2, c * s_ch + (f >> 3) * 128 + (f & 7) * 16, ch);
            }
            if (pass == 0) {
              z_s[f] = sv[DH];
            } else {
#pragma unroll
              for (int j = 0; j < DH; j += 4) st4(ds32 + f * RW + j, make_float4(sv[j], sv[j + 1], sv[j + 2], sv[j + 3]));
            }
          }
        }
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static int tc_bwd_mp(int M) { return M <= 64 ? 64 : (M <= 128 ? 128 : 256); }

size_t la_tc_bwd_smem_bytes(int DH, int M) {
  const int Mp = tc_bwd_mp(M);
  const int ND = (DH + 1 + 15) / 16 * 16;
  const size_t x_bytes = 16 * (size_t)(DH / 4) * 128;
  return 2 * (size_t)tc_w_bytes(DH, Mp) + 2 * x_bytes + 2 * 16 * (size_t)kTokCh + 2 * (size_t)(ND / 8) * kTokCh +
         2 * (size_t)(ND / 8) * (Mp / 8) * 128 + 2 * (size_t)Mp * (DH + 4) * sizeof(float) + 128;
}

int la_tc_backward(const void* qkv, const void* out, const void* dout, void* dqkv, const float* omega, int B, int N,
                   int H, int DH, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                   int dtype, const float* state, cudaStream_t st) {
  LaTcBwdArgs a;
  a.state = state;
  a.qkv = qkv; a.out = out; a.dout = dout; a.dqkv = dqkv; a.omega = omega; a.ta = ta; a.tb = tb; a.dg_part = dg_part;
  a.B = B; a.N = N; a.H = H; a.M = M; a.Mp = tc_bwd_mp(M); a.kind = kind; a.rot = rot; a.slots = slots;
  a.prescale = (float)pow((double)DH, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.trace = g_trace;
  const size_t smem = la_tc_bwd_smem_bytes(DH, M);
  int grid = (kNumSMs / H) * H;
  if (grid < H) grid = H;
  if (grid > B * H) grid = B * H;
  if (grid / H > slots && dg_part != nullptr) grid = slots * H;  // never more CTAs per head than gradient slots
#define TCB_LAUNCH(TT, NRB_, CPH_)                                                  \
  do {                                                                              \
    ERV_CUDA(allow_smem(la_tc_bwd_kernel<TT, 16, NRB_, CPH_>, smem));               \
    la_tc_bwd_kernel<TT, 16, NRB_, CPH_><<<grid, kTcThreads, smem, st>>>(a);        \
  } while (0)
  if (DH != 16) { set_error("tensor-core backward: head_dim %d not instantiated", DH); return ERV_E_UNSUPPORTED; }
  if (dtype == ERV_F32) {
    if (a.Mp == 64) TCB_LAUNCH(float, 1, 2); else if (a.Mp == 128) TCB_LAUNCH(float, 1, 4); else TCB_LAUNCH(float, 2, 4);
  } else {
    if (a.Mp == 64) TCB_LAUNCH(__nv_bfloat16, 1, 2); else if (a.Mp == 128) TCB_LAUNCH(__nv_bfloat16, 1, 4); else TCB_LAUNCH(__nv_bfloat16, 2, 4);
  }
#undef TCB_LAUNCH
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

void set_tc_trace(long long* p) { g_trace = p; }

}  // namespace erv
