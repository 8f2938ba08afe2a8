// Warp-specialised, pipelined backward of FAVOR+/ReLU linear attention for short sequences (33 <= N <= 65, head_dim 16,
// 128 < num_features <= 256), the companion of la_pipe_fwd_kernel (erv_linattn_pipe.cu) and the successor of
// la_tc2_bwd_kernel.  Same tile geometry (two (batch, head) pairs per 128-row tile, the 65th token of each pair outside the
// tile) and the same math (favor_plus.py:112-140, 247-260 differentiated by hand):
//
//   a = dO / den, a16 = -(dO . O) / den                      (den saved by the forward)
//   Q sweep  P = q W^T ; phi_q = exp(P - shift_q)            (shift saved by the forward: no row maximum, no exchange)
//            dS_pair = phi_q^T [a | a16]                     tcgen05, M = 128 features, K = tokens        "T2"
//            dphi_q  = a S^T (+ a16 z^T in registers)        tcgen05, K = 32 (block-structured a image)  "T3"
//            G = dphi (.) dphi/dP -> bf16 hi/lo in tensor memory ; dq' = G [W^T|1]   (A operand in tensor memory)  "T4"
//   K sweep  P = k W^T ; phi_k ; dv = phi_k [dS_A|dS_B] "T2" ; dphi_k = v dS^T (+ dz) "T3" ; dk' = G [W^T|1] "T4"
//
// Work is cut into four units per group, u = (Q|K sweep) x (128-feature half).  Warp 16 only issues MMAs; the 16 compute warps
// (4 threads per tile row) alternate between the exponentials of one half (C1), the gradient of the other (C2) and the
// row / state epilogues, which sit exactly where a compute warp would otherwise wait for the tensor pipe; warps 17/18 own the
// lone tokens (feature rows, W^T reductions, their gradient rows).  Tensor memory: two 128-column P / dphi / G buffers, the
// dS (4 x 48) / dv (64) accumulators, the dq' / dk' accumulator (48).
#include "erv_pipe_common.cuh"

namespace erv {

struct LaPipeBwdArgs {
  const void* qkv;
  const void* out;
  const void* dout;
  void* dqkv;
  const float* omega;
  const float* ta;
  const float* tb;
  float* dg_part;      // [H][slots][N][DH], circulant only
  const float* state;  // [B*H][DH+1][Mp] finished [S|z], then 3 per-token statistics (erv_linattn_pipe.cu)
  int B, N, H, M, kind, rot, slots;
  float prescale, inv_sqrt_m;
  long long* trace;
};

enum BwdBar {
  F_XQ = 0, F_AVQ, F_S, F_XK, F_DS, F_DQFREE, F_KFREE,  // 16 arrivals (compute warps)
  F_C1,                                                 // + unit
  F_C2 = F_C1 + 4,
  FULL_LONE = F_C2 + 4,                                 // 2 arrivals (lone-token warps)
  D_T1,                                                 // tcgen05.commit, + unit
  D_T2 = D_T1 + 4,
  D_T3 = D_T2 + 4,
  D_T4 = D_T3 + 4,
  BWD_BAR_COUNT = D_T4 + 4
};

template <typename T, bool FAVOR, bool PADDED>
__global__ void __launch_bounds__(kPipeThreads, 1) la_pipe_bwd_kernel(const LaPipeBwdArgs p) {
  constexpr int DH = 16, Mp = 256, HF = 128;
  constexpr uint32_t X_IMG = 128 * DH * 2, XL = 128, XS = 256;  // x images: K-major [128 rows x 16], three bf16 levels
  constexpr uint32_t PHI_IMG = 16 * kTokCh;                     // one feature half, one level: 32 KB
  constexpr uint32_t s_ch = (uint32_t)(Mp / 8) * 128;           // chunk stride of images with 256 feature rows (4 KB)
  // W image chunks (8 columns each, rows = features): [W^T hi (2) | ones column | 0 | W^T lo (2) | W^T lo2 (2)].
  //   K-major operand of the projection (rows f, K = d): LBO = s_ch, SBO = 128, levels at chunks 0 / 4 / 6
  //   MN-major B operand of G [W^T|1] (N = column, K = f): LBO = 128, SBO = s_ch, N = 48 (chunks 0..5) / 32 (chunks 0..3)
  constexpr uint32_t W_LO = 4, W_LO2 = 6;
  constexpr uint32_t AV_SIDE = 6 * kTokCh;  // per pair side: [hi (2) | special | pad | lo (2)] chunks, rows of the other side zero
  constexpr uint32_t COL_ACC = 256, S_STRIDE = 48, COL_DQ = 448;

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[BWD_BAR_COUNT];
  __shared__ uint32_t tmem_base_s;
  __shared__ float a16_s[128];
  __shared__ __align__(16) float z_s[2][Mp];
  __shared__ __align__(16) float dz_s[2][Mp];
  __shared__ __align__(16) float lone_s[2][2][2][Mp];  // [group parity][q|k][pair side][feature]
  __shared__ __align__(16) float lone_g[2][2][Mp];     // [q|k][pair side][feature]: G of the lone rows
  __shared__ __align__(16) float lone_a[2][2][DH + 4];  // [group parity][pair side]: a (16), a16
  __shared__ __align__(16) float lone_v[2][2][DH];
  __shared__ __align__(16) float lone_x[2][2][2][DH];  // [group parity][q|k][pair side]: prepared rows
  __shared__ __align__(16) float lone_xs[2][4][DH];    // [lone warp][row]: prepared rows of the group being projected
  __shared__ __align__(16) float lone_dy[2][DH];
  __shared__ __align__(16) uint8_t lone_raw[2][2][10][64];  // [lone warp][group parity][qA kA qB kB vA vB dOA dOB OA OB]
  __shared__ __align__(16) float lone_aux[2][2][8];  // [lone warp][group parity][shift qA kA qB kB | den A B]: staged statistics
  __shared__ float lone_red[2][2][DH + 4];             // [lone warp][pair side]: partial W^T reductions + row sum
  __shared__ float red_dv[16][DH];                     // per compute warp: dv partials of the lone key

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, part = (tid >> 7) & 3;
  const int M = p.M, N = p.N, H = p.H, B = p.B;
  const bool lone = N > 64;
  const int Nm = lone ? N - 1 : N;
  const int ks = (Nm + 15) >> 4;
  const int side = row >> 6, n = row & 63;
  const int ngroups = ((B + 1) >> 1) * H;
  const int h = blockIdx.x % H;
  const int n_it = (ngroups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const float kLog2e = 1.4426950408889634f;
  const float* aux = p.state + (size_t)B * H * (DH + 1) * Mp;  // [(b*H + h)*3 + j][N]: den, shift_q, shift_k

  uint8_t* wimg = smem;                 // 8 chunks of s_ch
  uint8_t* ximg = wimg + 8 * s_ch;      // 3 levels
  uint8_t* phi = ximg + 3 * X_IMG;      // hi | lo of the running half
  uint8_t* av = phi + 2 * PHI_IMG;      // [a|a16] rows (Q sweep) or [v] rows (K sweep), 2 sides
  uint8_t* simg = av + 2 * AV_SIDE;     // [S_hi A (2) | S_hi B (2) | S_lo A (2) | S_lo B (2)] chunks of s_ch
  uint8_t* dsimg = simg + 8 * s_ch;     // the same for dS

  if (warp == 16) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < BWD_BAR_COUNT; ++i) mbar_init(&bars[i], i < FULL_LONE ? 16 : (i == FULL_LONE ? 2 : 1));
    mbar_init_fence();
  }
  {
    for (int i = tid; i < (int)(8 * s_ch / 16); i += kPipeThreads) reinterpret_cast<uint4*>(wimg)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (int)(2 * AV_SIDE / 16); i += kPipeThreads) reinterpret_cast<uint4*>(av)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const float* om = p.omega + (size_t)h * DH * M;
    for (int i = tid; i < Mp * DH; i += kPipeThreads) {
      const int d = i / Mp, f = i % Mp;
      const float w = (f < M) ? __ldg(om + (size_t)d * M + f) : 0.f;
      const __nv_bfloat16 w0 = __float2bfloat16_rn(w);
      const float r1 = w - __bfloat162float(w0);
      const __nv_bfloat16 w1 = __float2bfloat16_rn(r1);
      const __nv_bfloat16 w2 = __float2bfloat16_rn(r1 - __bfloat162float(w1));
      const uint32_t fo = (uint32_t)(f >> 3) * 128 + (f & 7) * 16 + (d & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(wimg + (uint32_t)(d >> 3) * s_ch + fo) = w0;
      *reinterpret_cast<__nv_bfloat16*>(wimg + (W_LO + (d >> 3)) * s_ch + fo) = w1;
      *reinterpret_cast<__nv_bfloat16*>(wimg + (W_LO2 + (d >> 3)) * s_ch + fo) = w2;
      if (d == 0) *reinterpret_cast<__nv_bfloat16*>(wimg + 2 * s_ch + (uint32_t)(f >> 3) * 128 + (f & 7) * 16) = __float2bfloat16_rn(f < M ? 1.f : 0.f);
    }
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
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
    const uint32_t idesc_p = make_idesc(FMT_BF16, 128, HF, false, false);      // P half = x W_half^T
    const uint32_t idesc_acc48 = make_idesc(FMT_BF16, 128, 48, true, true);    // dS (+)= phi^T rows
    const uint32_t idesc_acc32 = make_idesc(FMT_BF16, 128, 32, true, true);
    const uint32_t idesc_dphi = make_idesc(FMT_BF16, 128, HF, false, false);   // dphi half = rows [S_A ; S_B]_half^T
    const uint32_t idesc_dv64 = make_idesc(FMT_BF16, 128, 64, false, true);    // dv = phi [dS_A|dS_B]
    const uint32_t idesc_dv32 = make_idesc(FMT_BF16, 128, 32, false, true);
    const uint32_t idesc_dq48 = make_idesc(FMT_BF16, 128, 48, false, true);    // dq' = G [W^T|1|0|W^T lo], A from tensor memory
    const uint32_t idesc_dq32 = make_idesc(FMT_BF16, 128, 32, false, true);
    const uint32_t wa = smem_u32(wimg), xa = smem_u32(ximg), pa = smem_u32(phi), ava = smem_u32(av);
    auto issue_t1 = [&](int u) {  // P half: x0w0 + x0w1 + x1w0 + x0w2 + x1w1 + x2w0 (three-level bf16 splits)
      if (elect_one()) {
        const int hb = u & 1;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const int xl = (t == 2 || t == 4) ? 1 : (t == 5 ? 2 : 0);
          const uint32_t wl = (t == 1 || t == 4) ? W_LO : (t == 3 ? W_LO2 : 0);
          mma_f16(tm + hb * HF, make_desc(xa + xl * X_IMG, XL, XS), make_desc(wa + wl * s_ch + (uint32_t)hb * (HF / 8) * 128, s_ch, 128),
                  idesc_p, t > 0);
        }
        commit(&bars[D_T1 + u]);
      }
      __syncwarp();
    };
    auto issue_t3 = [&](int u) {  // dphi half = rows x [B_A ; B_B]_half^T, K = [pair A's 16 | pair B's 16]
      if (elect_one()) {
        const int hb = u & 1;
        const uint32_t ba = smem_u32(u < 2 ? simg : dsimg) + (uint32_t)hb * (HF / 8) * 128;
        bool acc = false;
        for (int sp = 0; sp < 2; ++sp)
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t a_off = (uint32_t)sp * AV_SIDE + (term == 2 ? 4 : 0) * kTokCh;
            const uint32_t b_off = (uint32_t)((term == 1 ? 4 : 0) + 2 * sp) * s_ch;
            mma_f16(tm + hb * HF, make_desc(ava + a_off, kTokCh, 128), make_desc(ba + b_off, s_ch, 128), idesc_dphi, acc);
            acc = true;
          }
        commit(&bars[D_T3 + u]);
      }
      __syncwarp();
    };
    auto issue_t2 = [&](int u) {
      if (elect_one()) {
        const int hb = u & 1;
        if (u < 2) {  // dS(pair, half) = phi_q(half)^T [a | a16]
          for (int sp = 0; sp < 2; ++sp) {
            const uint32_t d = tm + COL_ACC + (uint32_t)(sp * 2 + hb) * S_STRIDE;
            for (int s = 0; s < ks; ++s) {
              const uint32_t st = (uint32_t)(sp * 4 + s) * 256;
              const uint64_t bd = make_desc(ava + (uint32_t)sp * AV_SIDE + st, 128, kTokCh);
              mma_f16(d, make_desc(pa + st, 128, kTokCh), bd, idesc_acc48, s > 0);
              mma_f16(d, make_desc(pa + PHI_IMG + st, 128, kTokCh), bd, idesc_acc32, true);
            }
          }
        } else {  // dv += phi_k(half) [dS_A | dS_B](half)
          const uint32_t da = smem_u32(dsimg);
          for (int s = 0; s < HF / 16; ++s) {
            const uint64_t bd = make_desc(da + (uint32_t)(hb * (HF / 16) + s) * 256, 128, s_ch);
            mma_f16(tm + COL_ACC, make_desc(pa + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_dv64, hb > 0 || s > 0);
            mma_f16(tm + COL_ACC, make_desc(pa + PHI_IMG + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_dv32, true);
          }
        }
        commit(&bars[D_T2 + u]);
      }
      __syncwarp();
    };
    auto issue_t4 = [&](int u) {  // dq' / dk' += G(half) [W^T|1]: A in tensor memory (hi words, then lo words, per 32 features)
      if (elect_one()) {
        const int hb = u & 1;
        for (int pq = 0; pq < 4; ++pq)
          for (int j = 0; j < 2; ++j) {
            const uint32_t fs = (uint32_t)(hb * HF + pq * 32 + 16 * j);
            const uint32_t ca = tm + (uint32_t)(hb * HF + pq * 32 + 8 * j);
            const uint64_t bd = make_desc(wa + fs * 16, 128, s_ch);
            mma_f16_ts(tm + COL_DQ, ca, bd, idesc_dq48, hb > 0 || pq > 0 || j > 0);
            mma_f16_ts(tm + COL_DQ, ca + 16, bd, idesc_dq32, true);
          }
        commit(&bars[D_T4 + u]);
      }
      __syncwarp();
    };
    auto wait = [&](int b, uint32_t par) { mbar_wait(&bars[b], par); };
    wait(F_XQ, 0);
    fence_after_sync();
    issue_t1(0);
    issue_t1(1);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t par = it & 1;
      const bool has_next = it + 1 < n_it;
      TR(0);
      wait(F_C1 + 0, par); wait(F_S, par); wait(F_AVQ, par);
      if (it > 0) wait(F_KFREE, par ^ 1);
      fence_after_sync();
      TR(1);
      issue_t3(0); issue_t2(0);
      wait(F_C2 + 0, par); fence_after_sync();
      TR(2);
      issue_t4(0);
      wait(F_XK, par); fence_after_sync();
      issue_t1(2);
      wait(F_C1 + 1, par); fence_after_sync();
      TR(3);
      issue_t3(1); issue_t2(1);
      wait(F_C2 + 1, par); fence_after_sync();
      TR(4);
      issue_t4(1);
      issue_t1(3);
      wait(F_C1 + 2, par); wait(F_DS, par); fence_after_sync();
      TR(5);
      issue_t3(2); issue_t2(2);
      wait(F_C2 + 2, par); wait(F_DQFREE, par); fence_after_sync();
      TR(6);
      issue_t4(2);
      wait(F_XQ, par ^ 1); fence_after_sync();
      if (has_next) issue_t1(0);
      wait(F_C1 + 3, par); fence_after_sync();
      TR(7);
      issue_t3(3); issue_t2(3);
      wait(F_C2 + 3, par); fence_after_sync();
      TR(8);
      issue_t4(3);
      if (has_next) issue_t1(1);
    }
  } else if (warp > 16) {
    // ================================================ lone-token warps =================================================
    if (lone) {
      const int lw = warp - 17;
      const T* qkv = static_cast<const T*>(p.qkv);
      const T* outp = static_cast<const T*>(p.out);
      const T* dout = static_cast<const T*>(p.dout);
      T* dqkv = static_cast<T*>(p.dqkv);
      float wreg[4][DH];  // W^T rows of features f_i = 128 lw + lane + 32 i
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
        if (g < ngroups) {
          for (int c = lane; c < 10 * CPR; c += 32) {
            const int ri = c / CPR, ch = c % CPR;
            const int sp = ri < 4 ? ri >> 1 : (ri & 1), b = 2 * (g / H) + sp;
            if (b < B) {
              const T* src;
              if (ri < 4) src = qkv + qkv_off(b, N - 1, ri & 1, h, N, H, DH);
              else if (ri < 6) src = qkv + qkv_off(b, N - 1, 2, h, N, H, DH);
              else src = (ri < 8 ? dout : outp) + out_off(b, N - 1, h, N, H, DH);
              cp_async16(&lone_raw[lw][buf][ri][ch * 16], src + ch * (16 / (int)sizeof(T)));
            }
          }
          if (lane < 6) {  // exponent shifts of the lone query / key rows, normalisers of the lone queries
            const int sp = lane < 4 ? lane >> 1 : lane - 4, b = 2 * (g / H) + sp;
            const int j = lane < 4 ? 1 + (lane & 1) : 0;
            if (b < B)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(&lone_aux[lw][buf][lane])),
                           "l"(aux + ((size_t)b * H + h) * 3 * N + j * N + N - 1)
                           : "memory");
          }
        }
        cp_async_commit();
      };
      // feature rows of the lone query / key of both pairs, a / a16 of the lone query, [v] of the lone key
      auto features = [&](int g, int buf) {
        const int b2 = g / H;
        cp_async_wait_all();
        __syncwarp();
        if (lane < 4) {  // lane r prepares row r = 2 * pair side + (0: query, 1: key)
          const int sp = lane >> 1, which = lane & 1;
          float x[DH];
#pragma unroll
          for (int a = 0; a < DH; ++a) x[a] = 0.f;
          if (2 * b2 + sp < B) {
            load_row<T, DH>(reinterpret_cast<const T*>(&lone_raw[lw][buf][lane][0]), x);
            prologue_row<DH>(x, p.rot, p.ta, p.tb, h, N - 1, N, p.prescale);
          }
#pragma unroll
          for (int c = 0; c < DH / 4; ++c) {
            const float4 v = make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
            st4(&lone_xs[lw][lane][4 * c], v);
            if (lw == 0) st4(&lone_x[buf][which][sp][4 * c], v);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // one feature at a time: its W^T row against the four rows
          const int f = HF * lw + lane + 32 * i;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < DH / 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float4 xv = ld4(&lone_xs[lw][r][4 * c]);
              acc[r] = fmaf(xv.x, wreg[i][4 * c], acc[r]); acc[r] = fmaf(xv.y, wreg[i][4 * c + 1], acc[r]);
              acc[r] = fmaf(xv.z, wreg[i][4 * c + 2], acc[r]); acc[r] = fmaf(xv.w, wreg[i][4 * c + 3], acc[r]);
            }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const bool ok = 2 * b2 + (r >> 1) < B;
            float v = FAVOR ? ex2_approx(fmaf(acc[r], kLog2e, -lone_aux[lw][buf][r])) : fmaxf(acc[r], 0.f) * p.inv_sqrt_m;
            if ((PADDED && f >= M) || !ok) v = 0.f;
            lone_s[buf][r & 1][r >> 1][f] = v;
          }
        }
        if (lw == 0) {  // a = dO / den, a16 = -(dO . O) / den of the lone queries; v of the lone keys
          const int sp = lane >> 4, d = lane & 15, b = 2 * b2 + sp;
          const bool ok = b < B;
          const float den = ok ? lone_aux[lw][buf][4 + sp] : 1.f;
          const float dO = ok ? to_f(reinterpret_cast<const T*>(&lone_raw[lw][buf][6 + sp][0])[d]) : 0.f;
          const float O = ok ? to_f(reinterpret_cast<const T*>(&lone_raw[lw][buf][8 + sp][0])[d]) : 0.f;
          float dot = dO * O;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
          lone_a[buf][sp][d] = dO / den;
          if (d == 0) lone_a[buf][sp][DH] = -dot / den;
          lone_v[buf][sp][d] = ok ? to_f(reinterpret_cast<const T*>(&lone_raw[lw][buf][4 + sp][0])[d]) : 0.f;
        }
      };
      // gradient row of a lone query / key: sums of G [W^T|1] over this warp's features, combined by warp 0
      auto w_reduce = [&](int g, int which, int buf) {
        const int b2 = g / H;
#pragma unroll 1
        for (int sp = 0; sp < 2; ++sp) {
          float gv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) gv[i] = lone_g[which][sp][HF * lw + lane + 32 * i];
#pragma unroll
          for (int c = 0; c < 2; ++c) {  // 8 columns at a time keeps the W^T rows in registers
            float v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v8[j] = gv[0] * wreg[0][8 * c + j];
#pragma unroll
              for (int i = 1; i < 4; ++i) v8[j] = fmaf(gv[i], wreg[i][8 * c + j], v8[j]);
            }
            const float t = warp_sum8(v8);
            if (!(lane & 3)) lone_red[lw][sp][8 * c + (lane >> 2)] = t;
          }
          const float rs = warp_sum((gv[0] + gv[1]) + (gv[2] + gv[3]));
          if (lane == 0) lone_red[lw][sp][DH] = rs;
        }
        lbar();
        if (lw == 0) {
          const int sp = lane >> 4, d = lane & 15, bb = 2 * b2 + sp;
          const float acc = lone_red[0][sp][d] + lone_red[1][sp][d], rs = lone_red[0][sp][DH] + lone_red[1][sp][DH];
          lone_dy[sp][d] = (FAVOR ? acc - lone_x[buf][which][sp][d] * rs : acc) * p.prescale;
          __syncwarp();
          if (d == 0 && bb < B) {
            float dy[DH], dxr[DH];
#pragma unroll
            for (int a = 0; a < DH; ++a) dy[a] = lone_dy[sp][a];
            float* slot = (p.rot == ERV_ROT_CIRCULANT && p.dg_part)
                              ? p.dg_part + ((size_t)h * p.slots + (blockIdx.x / H) * 2 + sp) * N * DH : nullptr;
            const T* xraw = qkv + qkv_off(bb, N - 1, which, h, N, H, DH);
            prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, N - 1, N, slot, xraw);
            T* dst = dqkv + qkv_off(bb, N - 1, which, h, N, H, DH);
#pragma unroll
            for (int c = 0; c < DH / 4; ++c) st4(dst + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
          }
        }
        lbar();  // lone_red / lone_dy are rewritten by the next call
      };
      fetch_rows(blockIdx.x, 0);
      features(blockIdx.x, 0);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[FULL_LONE]);
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
        mbar_wait(&bars[F_S], par);  // G of the lone queries (written while the S image was built)
        TR(2);
        w_reduce(g, 0, par);
        TR(5);
        mbar_wait(&bars[F_DS], par);  // G of the lone keys, dv partials (written by the dS epilogue)
        TR(3);
        w_reduce(g, 1, par);
        TR(6);
        if (lw == 0) {
          const int sp = lane >> 4, d = lane & 15, bb = 2 * (g / H) + sp;
          if (bb < B) {
            float dv = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) dv += red_dv[8 * sp + w][d];
            T* dst = dqkv + qkv_off(bb, N - 1, 2, h, N, H, DH) + d;
            if (sizeof(T) == 4) *reinterpret_cast<float*>(dst) = dv;
            else *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(dv);
          }
        }
        TR(4);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[FULL_LONE]);
      }
    }
  } else {
    // ================================================== compute warps ===================================================
    const int q4 = warp & 3;
    const uint32_t tm_thr = tm + ((uint32_t)(q4 * 32) << 16);
    const T* qkv = static_cast<const T*>(p.qkv);
    const T* outp = static_cast<const T*>(p.out);
    const T* dout = static_cast<const T*>(p.dout);
    T* dqkv = static_cast<T*>(p.dqkv);
    const uint32_t rowoff = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;
    const uint32_t xoff = (uint32_t)(row >> 3) * XS + (row & 7) * 16;
    const uint32_t phi_thr = smem_u32(phi) + (uint32_t)(part * 4) * kTokCh + rowoff;
    const uint32_t av_thr = smem_u32(av) + (uint32_t)side * AV_SIDE + rowoff;
    auto wait = [&](int b, uint32_t par) { mbar_wait(&bars[b], par); };
    float* dg_slot = (p.rot == ERV_ROT_CIRCULANT && p.dg_part)
                         ? p.dg_part + ((size_t)h * p.slots + (blockIdx.x / H) * 2 + side) * N * DH : nullptr;

    // rotation + scale of a q / k row, three-level images for the projection (part 0)
    auto stage_x_row = [&](int g, int which) {
      const int b = 2 * (g / H) + side;
      float x[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) x[a] = 0.f;
      if (b < B && n < Nm) {
        load_row<T, DH>(qkv + qkv_off(b, n, which, h, N, H, DH), x);
        prologue_row<DH>(x, p.rot, p.ta, p.tb, h, n, N, p.prescale);
      }
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        float ch[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[e] = x[8 * c + e];
        store_split8_l3(ximg, ximg + X_IMG, ximg + 2 * X_IMG, xoff + c * XL, ch);
      }
    };
    auto load_shift = [&](int g, int which) -> float {  // exponent shift of this thread's row (+inf: row not in the tile)
      const int b = 2 * (g / H) + side;
      return (b < B && n < Nm) ? __ldg(aux + ((size_t)b * H + h) * 3 * N + (1 + which) * N + n) : INFINITY;
    };
    // [a | a16] rows of the Q sweep (part 1): a = dO / den, a16 = -(dO . O) / den
    auto stage_a_row = [&](int g) {
      const int b = 2 * (g / H) + side;
      const bool valid = b < B && n < Nm;
      float dO[DH], a16 = 0.f;
#pragma unroll
      for (int a = 0; a < DH; ++a) dO[a] = 0.f;
      if (valid) {
        float O[DH];
        load_row<T, DH>(dout + out_off(b, n, h, N, H, DH), dO);
        load_row<T, DH>(outp + out_off(b, n, h, N, H, DH), O);
        const float r = 1.0f / __ldg(aux + ((size_t)b * H + h) * 3 * N + n);
        float dot = 0.f;
#pragma unroll
        for (int a = 0; a < DH; ++a) dot = fmaf(dO[a], O[a], dot);
#pragma unroll
        for (int a = 0; a < DH; ++a) dO[a] *= r;
        a16 = -dot * r;
      }
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        float ch[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[e] = dO[8 * c + e];
        store_split8_sa(av_thr + c * kTokCh, av_thr + (4 + c) * kTokCh, ch);
      }
      a16_s[row] = a16;
      const uint32_t hi = __float_as_uint(a16) & 0xffff0000u;
      const float lo = a16 - __uint_as_float(hi);
      sts128(av_thr + 2 * kTokCh, (hi >> 16) | (__float_as_uint(lo) & 0xffff0000u), 0u, 0u, 0u);  // [a16_hi, a16_lo]
    };
    auto stage_v_row = [&](int g) {  // [v] rows of the K sweep (part 1)
      const int b = 2 * (g / H) + side;
      float v[DH];
#pragma unroll
      for (int a = 0; a < DH; ++a) v[a] = 0.f;
      if (b < B && n < Nm) load_row<T, DH>(qkv + qkv_off(b, n, 2, h, N, H, DH), v);
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        float ch[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[e] = v[8 * c + e];
        store_split8_sa(av_thr + c * kTokCh, av_thr + (4 + c) * kTokCh, ch);
      }
    };
    // S image of a group from the state the forward saved; thread = (feature, pair side).  Also G of the lone query.
    auto build_s = [&](int g, int buf) {
      const int sp = part >> 1, f = (part & 1) * HF + row, b = 2 * (g / H) + sp;
      float st[DH + 1];
#pragma unroll
      for (int d = 0; d <= DH; ++d) st[d] = 0.f;
      if (b < B) {
        const float* src = p.state + ((size_t)b * H + h) * (DH + 1) * Mp + f;
#pragma unroll
        for (int d = 0; d <= DH; ++d) st[d] = __ldg(src + (size_t)d * Mp);
      }
      z_s[sp][f] = st[DH];
      const uint32_t fo = (uint32_t)(f >> 3) * 128 + (f & 7) * 16, sa = smem_u32(simg);
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        float ch[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ch[e] = st[8 * c + e];
        store_split8_sa(sa + (uint32_t)(2 * sp + c) * s_ch + fo, sa + (uint32_t)(4 + 2 * sp + c) * s_ch + fo, ch);
      }
      if (lone) {  // dphi of the lone query: a_L . S[f] + a16_L z[f]
        const float pq = lone_s[buf][0][sp][f];
        float dph = lone_a[buf][sp][DH] * st[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) dph = fmaf(lone_a[buf][sp][d], st[d], dph);
        lone_g[0][sp][f] = FAVOR ? dph * pq : (pq > 0.f ? dph * p.inv_sqrt_m : 0.f);
      }
    };

    // Rows and state a later phase of this thread reads from global memory are prefetched one unit ahead (into L1 for the
    // 64-byte token rows, into L2 for the 17 strided state words), so that no register is held and the consuming phase
    // does not sit behind the memory latency.
    auto pf_l1 = [](const void* q) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(q)); };
    auto pf_l2 = [](const void* q) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q)); };
    auto prefetch_unit = [&](int u, int g, int gn, bool has_next) {
      const int b = 2 * (g / H) + side, bn = 2 * (gn / H) + side;
      const bool v = b < B && n < Nm, vn = has_next && bn < B && n < Nm;
      if (u == 0) {
        if (part == 0 && v) pf_l1(qkv + qkv_off(b, n, 1, h, N, H, DH));          // key row: x images after C1(Q0)
      } else if (u == 1) {
        if (part == 1 && v) pf_l1(qkv + qkv_off(b, n, 2, h, N, H, DH));          // value row: end of the Q sweep
      } else if (u == 2) {
        if (part == 0 && vn) pf_l1(qkv + qkv_off(bn, n, 0, h, N, H, DH));        // next group's query row
        if (part == 2 && v) pf_l1(qkv + qkv_off(b, n, 0, h, N, H, DH));          // query row again: dq epilogue
        if (has_next) {                                                          // next group's [S|z] rows
          const int sp = part >> 1, f = (part & 1) * HF + row, bs = 2 * (gn / H) + sp;
          if (bs < B) {
            const float* src = p.state + ((size_t)bs * H + h) * (DH + 1) * Mp + f;
            if ((lane & 7) == 0)  // one prefetch per 32-byte sector of the warp's 128-byte line
#pragma unroll
              for (int d = 0; d <= DH; ++d) pf_l2(src + (size_t)d * Mp);
          }
        }
      } else {
        if (part == 1 && vn) {                                                   // next group's dO / O rows
          pf_l1(dout + out_off(bn, n, h, N, H, DH));
          pf_l1(outp + out_off(bn, n, h, N, H, DH));
        }
        if (part == 2 && v) pf_l1(qkv + qkv_off(b, n, 1, h, N, H, DH));          // key row again: dk epilogue
      }
    };

    // ---- preamble: first group's query rows, [a|a16] rows, S image
    const int g0 = blockIdx.x;
    if (part == 0) stage_x_row(g0, 0);
    warp_arrive(&bars[F_XQ]);
    if (part == 1) stage_a_row(g0);
    warp_arrive(&bars[F_AVQ]);
    if (lone) wait(FULL_LONE, 0);
    build_s(g0, 0);
    warp_arrive(&bars[F_S]);
    float shift_q = load_shift(g0, 0), shift_k = 0.f;
    bar_compute();  // a16_s, z_s

    int it = 0;
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int gn = g + gridDim.x;
      const bool has_next = gn < ngroups;
      const int b = 2 * (g / H) + side;
      const bool valid = b < B && n < Nm;
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        const int isk = u >> 1, hb = u & 1;
        prefetch_unit(u, g, gn, has_next);
        // ---- C1: one feature half: P (tensor memory) -> phi -> hi/lo bf16 images
        TR(10 * u + 0);
        wait(D_T1 + u, par);
        if (u > 0) wait(D_T2 + u - 1, par);  // the previous half's contraction has read the feature images
        fence_after_sync();
        TR(10 * u + 1);
        {
          const float shift = isk ? shift_k : shift_q;
          const float scale = (shift < INFINITY) ? p.inv_sqrt_m : 0.f;
          uint32_t r[32];
          tmem_ld32_nowait(tm_thr + hb * HF + part * 32, r);
          tmem_wait_ld32(r);
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
            store_split8_sa(phi_thr + c * kTokCh, phi_thr + PHI_IMG + c * kTokCh, v);
          }
          warp_arrive(&bars[F_C1 + u]);
        }
        TR(10 * u + 2);
        // ---- work that fills the wait for dphi
        if (u == 0) {  // key rows -> x images (the projections of the queries have both completed)
          wait(D_T1 + 1, par);
          if (part == 0) stage_x_row(g, 1);
          warp_arrive(&bars[F_XK]);
          shift_k = load_shift(g, 1);
        } else if (u == 2) {
          if (part == 2) {  // dq rows: dq' = G [W^T|1] of the Q sweep has completed
            wait(D_T4 + 0, par);
            wait(D_T4 + 1, par);
            fence_after_sync();
            float d0[32], d1[16];
            tmem_ld32(tm_thr + COL_DQ, d0);
            tmem_ld16(tm_thr + COL_DQ + 32, d1);
            if (valid) {
              float dy[DH], dxr[DH];
              const T* xraw = qkv + qkv_off(b, n, 0, h, N, H, DH);
              if (FAVOR) {
                float x[DH];
                load_row<T, DH>(xraw, x);
                prologue_row<DH>(x, p.rot, p.ta, p.tb, h, n, N, p.prescale);
#pragma unroll
                for (int d = 0; d < DH; ++d) dy[d] = ((d0[d] + d1[d]) - x[d] * d0[DH]) * p.prescale;
              } else {
#pragma unroll
                for (int d = 0; d < DH; ++d) dy[d] = (d0[d] + d1[d]) * p.prescale;
              }
              prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, n, N, dg_slot, xraw);
              T* dst = dqkv + qkv_off(b, n, 0, h, N, H, DH);
#pragma unroll
              for (int c = 0; c < DH / 4; ++c) st4(dst + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
            }
          }
          warp_arrive(&bars[F_DQFREE]);
          wait(D_T1 + 3, par);  // the projections of the keys have both completed: next group's query rows
          if (has_next && part == 0) stage_x_row(gn, 0);
          warp_arrive(&bars[F_XQ]);
          if (has_next) shift_q = load_shift(gn, 0);
        } else if (u == 3) {  // next group's S image (the lone-token warps are one group ahead)
          if (has_next) {
            if (lone) wait(FULL_LONE, par ^ 1);
            build_s(gn, par ^ 1);
          }
          warp_arrive(&bars[F_S]);
        }
        // ---- C2: dphi (tensor memory) + rank-1 term -> G = dphi (.) dphi/dP -> bf16 hi/lo words back into the same columns
        TR(10 * u + 3);
        wait(D_T3 + u, par);
        fence_after_sync();
        TR(10 * u + 4);
        {
          uint32_t r[32], hw[16], lw[16];
          tmem_ld32_nowait(tm_thr + hb * HF + part * 32, r);
          const float rscale = isk ? 1.0f : a16_s[row];
          const float* rank1 = (isk ? &dz_s[side][0] : &z_s[side][0]) + hb * HF + part * 32;
          tmem_wait_ld32(r);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 ph = lds128(phi_thr + c * kTokCh), pl = lds128(phi_thr + PHI_IMG + c * kTokCh);
            const float4 za = ld4(rank1 + 8 * c), zb = ld4(rank1 + 8 * c + 4);
            const float phv[8] = {bf_lo(ph.x) + bf_lo(pl.x), bf_hi(ph.x) + bf_hi(pl.x), bf_lo(ph.y) + bf_lo(pl.y), bf_hi(ph.y) + bf_hi(pl.y),
                                  bf_lo(ph.z) + bf_lo(pl.z), bf_hi(ph.z) + bf_hi(pl.z), bf_lo(ph.w) + bf_lo(pl.w), bf_hi(ph.w) + bf_hi(pl.w)};
            const float zz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
            float gq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float dph = fmaf(rscale, zz[i], __uint_as_float(r[8 * c + i]));
              gq[i] = FAVOR ? dph * phv[i] : (phv[i] > 0.f ? dph * p.inv_sqrt_m : 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) split_pack2(gq[2 * i], gq[2 * i + 1], hw[4 * c + i], lw[4 * c + i]);
          }
          tmem_st16(tm_thr + hb * HF + part * 32, hw);
          tmem_st16(tm_thr + hb * HF + part * 32 + 16, lw);
          tmem_wait_st();
          warp_arrive(&bars[F_C2 + u]);
        }
        TR(10 * u + 5);
        if (u == 1) {
          // ---- end of the Q sweep: dS (tensor memory, lanes = features) -> dS image, dz, lone-token terms; [v] rows
          wait(D_T2 + 1, par);
          fence_after_sync();
          TR(16);
          const int sp = part >> 1, hq = part & 1, f = hq * HF + row;
          float d0[32], d1[16], sv[DH];
          tmem_ld32(tm_thr + COL_ACC + (uint32_t)(sp * 2 + hq) * S_STRIDE, d0);
          tmem_ld16(tm_thr + COL_ACC + (uint32_t)(sp * 2 + hq) * S_STRIDE + 32, d1);
#pragma unroll
          for (int d = 0; d < DH; ++d) sv[d] = d0[d] + d1[d];
          float dz = d0[DH] + d0[DH + 1];
          float pk = 0.f;
          if (lone) {  // rank-1 term of the last query
            const float pq = lone_s[par][0][sp][f];
            pk = lone_s[par][1][sp][f];
#pragma unroll
            for (int d = 0; d < DH; ++d) sv[d] = fmaf(pq, lone_a[par][sp][d], sv[d]);
            dz = fmaf(pq, lone_a[par][sp][DH], dz);
          }
          dz_s[sp][f] = dz;
          const uint32_t fo = (uint32_t)(f >> 3) * 128 + (f & 7) * 16, da = smem_u32(dsimg);
#pragma unroll
          for (int c = 0; c < DH / 8; ++c) {
            float ch[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ch[e] = sv[8 * c + e];
            store_split8_sa(da + (uint32_t)(2 * sp + c) * s_ch + fo, da + (uint32_t)(4 + 2 * sp + c) * s_ch + fo, ch);
          }
          if (lone) {  // dv and dphi of the last key
            float dph = dz;
#pragma unroll
            for (int d = 0; d < DH; ++d) dph = fmaf(lone_v[par][sp][d], sv[d], dph);
            lone_g[1][sp][f] = FAVOR ? dph * pk : (pk > 0.f ? dph * p.inv_sqrt_m : 0.f);
#pragma unroll
            for (int d = 0; d < DH; ++d) sv[d] *= pk;
            const float t = warp_sum16(sv);
            if (!(lane & 1)) red_dv[warp][lane >> 1] = t;
          }
          if (part == 1) stage_v_row(g);
          warp_arrive(&bars[F_DS]);
          TR(17);
        } else if (u == 3) {
          // ---- end of the K sweep: dv and dk rows, next group's [a|a16] rows
          wait(D_T2 + 3, par);
          fence_after_sync();
          TR(36);
          if (part == 1) {
            float a0[16], a1[16];
            tmem_ld16(tm_thr + COL_ACC + 16 * side, a0);
            tmem_ld16(tm_thr + COL_ACC + 32 + 16 * side, a1);
            if (valid) {
              T* dvp = dqkv + qkv_off(b, n, 2, h, N, H, DH);
#pragma unroll
              for (int c = 0; c < DH / 4; ++c)
                st4(dvp + 4 * c, make_float4(a0[4 * c] + a1[4 * c], a0[4 * c + 1] + a1[4 * c + 1], a0[4 * c + 2] + a1[4 * c + 2],
                                             a0[4 * c + 3] + a1[4 * c + 3]));
            }
            if (has_next) stage_a_row(gn);
          }
          warp_arrive(&bars[F_AVQ]);
          if (part == 2) {
            wait(D_T4 + 2, par);
            wait(D_T4 + 3, par);
            fence_after_sync();
            float d0[32], d1[16];
            tmem_ld32(tm_thr + COL_DQ, d0);
            tmem_ld16(tm_thr + COL_DQ + 32, d1);
            if (valid) {
              float dy[DH], dxr[DH];
              const T* xraw = qkv + qkv_off(b, n, 1, h, N, H, DH);
              if (FAVOR) {
                float x[DH];
                load_row<T, DH>(xraw, x);
                prologue_row<DH>(x, p.rot, p.ta, p.tb, h, n, N, p.prescale);
#pragma unroll
                for (int d = 0; d < DH; ++d) dy[d] = ((d0[d] + d1[d]) - x[d] * d0[DH]) * p.prescale;
              } else {
#pragma unroll
                for (int d = 0; d < DH; ++d) dy[d] = (d0[d] + d1[d]) * p.prescale;
              }
              prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, n, N, dg_slot, xraw);
              T* dst = dqkv + qkv_off(b, n, 1, h, N, H, DH);
#pragma unroll
              for (int c = 0; c < DH / 4; ++c) st4(dst + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
            }
          }
          warp_arrive(&bars[F_KFREE]);
          TR(37);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tm, 512);
}

size_t la_pipe_bwd_smem_bytes() {
  const size_t s_ch = (256 / 8) * 128;
  return 8 * s_ch + 3 * (128 * 16 * 2) + 2 * 16 * (size_t)kTokCh + 12 * (size_t)kTokCh + 16 * s_ch;
}

int la_pipe_backward(const void* qkv, const void* out, const void* dout, void* dqkv, const float* omega, int B, int N,
                     int H, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                     int dtype, const float* state, cudaStream_t st) {
  LaPipeBwdArgs a;
  a.qkv = qkv; a.out = out; a.dout = dout; a.dqkv = dqkv; a.omega = omega; a.ta = ta; a.tb = tb; a.dg_part = dg_part;
  a.B = B; a.N = N; a.H = H; a.M = M; a.kind = kind; a.rot = rot; a.slots = slots; a.state = state;
  a.trace = g_trace;
  a.prescale = (float)pow(16.0, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  const size_t smem = la_pipe_bwd_smem_bytes();
  const int ngroups = ((B + 1) / 2) * H;
  int grid = (kNumSMs / H) * H;
  if (grid < H) grid = H;
  if (grid > ngroups) grid = ngroups;
  if (dg_part != nullptr && 2 * (grid / H) > slots) grid = (slots / 2) * H;  // two gradient slots per CTA
  if (grid < H) { set_error("pipelined backward: %d gradient slots are too few", slots); return ERV_E_INVALID; }
  const bool favor = kind == ERV_FEAT_FAVOR, padded = M < 256;
#define PIPEB_LAUNCH(TT, FV, PD)                                              \
  do {                                                                        \
    ERV_CUDA(allow_smem(la_pipe_bwd_kernel<TT, FV, PD>, smem));               \
    la_pipe_bwd_kernel<TT, FV, PD><<<grid, kPipeThreads, smem, st>>>(a);      \
  } while (0)
#define PIPEB_LAUNCH_T(TT)                                                    \
  do {                                                                        \
    if (favor) { if (padded) PIPEB_LAUNCH(TT, true, true); else PIPEB_LAUNCH(TT, true, false); } \
    else { if (padded) PIPEB_LAUNCH(TT, false, true); else PIPEB_LAUNCH(TT, false, false); }     \
  } while (0)
  if (dtype == ERV_F32) PIPEB_LAUNCH_T(float); else PIPEB_LAUNCH_T(__nv_bfloat16);
#undef PIPEB_LAUNCH_T
#undef PIPEB_LAUNCH
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv
