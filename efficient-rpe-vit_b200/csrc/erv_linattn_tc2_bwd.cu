// Tensor-core backward of FAVOR+/ReLU linear attention for short sequences (33 <= N <= 65, head_dim 16,
// num_features <= 256): two (batch, head) pairs per 128-row tile, the companion of erv_linattn_tc2.cu.
//
// Rows 0..63 of every tile belong to pair (2g, h), rows 64..127 to pair (2g+1, h); 512 threads = 4 per row, thread
// (row, part) owns a quarter of the row's features in registers.  Three sweeps per group:
//   K1  P = k W^T (3xTF32) ; phi_k ; S_pair[f][.] = phi_k^T [v|1]                          tcgen05 -> TMEM
//   Q   P = q W^T ; phi_q ; den = phi_q . z ; a = [dO/den | -(dO.O)/den]
//       dS_pair = phi_q^T a                                                               tcgen05, S's TMEM columns
//       dphi_q  = a S^T  (both pairs in one K = 32 product: block-structured a image)     tcgen05, P's TMEM columns
//       + a16 z^T (rank 1, registers) ; G = dphi (.) dphi/dP, written back over the consumed dphi columns as bf16 hi/lo
//       dq' = G [W^T|1]   (A operand read from tensor memory, K = features)             tcgen05, 48 spare TMEM columns
//   K2  P = k W^T ; phi_k ; dphi_k = v dS^T (+ dz) ; dv = phi_k [dS_A|dS_B]                tcgen05 ; dk as dq
// bf16 hi/lo splits of both operands keep the contractions at ~2^-17; a split product takes two instructions by
// concatenating along N: phi_hi x [b_hi | b_lo] and phi_lo x b_hi.  The feature images are written one 128-feature half
// at a time (shared memory), each half's contraction running while the next half is stored.
// When N = 65 the last token of each pair (the "lone" token) is handled outside the tiles: its feature rows are
// computed by one warp per (pair, q|k), its rank-1 terms are folded into S / dS when they leave TMEM, and its own
// gradients are three 256-thread reductions against S and dS.
#include "erv_tc_common.cuh"

namespace erv {

struct LaTc2BwdArgs {
  const void* qkv;
  const void* out;
  const void* dout;
  void* dqkv;
  const float* omega;
  const float* ta;
  const float* tb;
  float* dg_part;  // [H][slots][N][DH], circulant only
  const float* state;  // optional [B*H][DH+1][Mp]: [S|z] saved by the forward; the K1 sweep is skipped when present
  int B, N, H, M, kind, rot, slots;
  float prescale, inv_sqrt_m;
  long long* trace;  // optional: CTA 0 / thread 0 stamps clock64() at phase boundaries (erv_debug_set_trace)
};

extern long long* g_trace;

// raw 16-element row held as loaded (conversion to fp32 is deferred so the load stays in flight)
template <typename T> struct RawRow;
template <> struct RawRow<float> { float4 r[4]; };
template <> struct RawRow<__nv_bfloat16> { uint4 r[2]; };
__device__ __forceinline__ void load_raw(const float* p, RawRow<float>& w) {
#pragma unroll
  for (int c = 0; c < 4; ++c) w.r[c] = *reinterpret_cast<const float4*>(p + 4 * c);
}
__device__ __forceinline__ void load_raw(const __nv_bfloat16* p, RawRow<__nv_bfloat16>& w) {
  w.r[0] = *reinterpret_cast<const uint4*>(p);
  w.r[1] = *reinterpret_cast<const uint4*>(p + 8);
}
__device__ __forceinline__ void raw_to_f(const RawRow<float>& w, float (&x)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) { x[4 * c] = w.r[c].x; x[4 * c + 1] = w.r[c].y; x[4 * c + 2] = w.r[c].z; x[4 * c + 3] = w.r[c].w; }
}
__device__ __forceinline__ void raw_to_f(const RawRow<__nv_bfloat16>& w, float (&x)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t u[4] = {w.r[c].x, w.r[c].y, w.r[c].z, w.r[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[8 * c + 2 * i] = __uint_as_float(u[i] << 16);
      x[8 * c + 2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
}

// NRB feature halves, CPH 8-feature chunks per thread and half: Mp = NRB * CPH * 32 is compile time so the per-thread
// feature registers are statically indexed.
template <typename T, int NRB, int CPH>
__global__ void __launch_bounds__(kTcThreads, 1) la_tc2_bwd_kernel(const LaTc2BwdArgs p) {
  constexpr int DH = 16, RW = DH + 4;
  using C = TcCfg<DH>;
  constexpr int nrb = NRB, HF = CPH * 32, FPH = CPH * 8, Mp = NRB * HF, NC = NRB * CPH;
  constexpr uint32_t COL_P = 0, COL_S = 256, S_STRIDE = 48, COL_DQ = 448;  // [448, 496): G [W^T|1] of the running sweep
  constexpr uint32_t wbytes = (uint32_t)(Mp / 8) * (DH / 4) * 128;
  constexpr uint32_t s_ch = (uint32_t)(Mp / 8) * 128;
  constexpr uint32_t halfbytes = 16 * kTokCh;  // one feature half, padded to 128 rows of the M dimension
  constexpr uint32_t avbytes = 12 * kTokCh;    // per pair side: [hi(2) | special(1) | zero(1) | lo(2)] chunks

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_a, bar_b, bar_c, bar_d;  // G1 / token- and feature-contractions / dphi / G [W^T|1]
  __shared__ uint32_t tmem_base_s;
  __shared__ float n2_s[128];
  __shared__ float ex_s[4][128];   // row-max exchange
  __shared__ float den_s[4][128];  // den partials
  __shared__ float a16_s[128];
  // staged rows of the saved forward output (Q sweep); after the dot products the same rows hold dO in fp32
  __shared__ __align__(16) uint8_t o_raw[128 * DH * sizeof(float)];
  __shared__ __align__(16) float xrow_s[128][DH];  // prepared q / k rows (kept out of registers across the sweep)
  __shared__ __align__(16) float z_s[2][Mp];
  __shared__ __align__(16) float dz_s[2][Mp];
  __shared__ __align__(16) float lone_s[2][2][Mp];  // [q|k][pair side][feature]
  __shared__ __align__(16) float lone_x[2][2][DH];  // prepared q / k rows of the lone tokens
  __shared__ __align__(16) float lone_v[2][DH];
  __shared__ __align__(16) float lone_do[2][DH];
  __shared__ __align__(16) float lone_o[2][DH];
  __shared__ float lone_a[2][DH + 1];
  __shared__ float lone_mx[2][2][4];  // [q|k][pair side][feature quarter]: row-max partials
  __shared__ float lone_n2[2][2];
  __shared__ float red1_s[16];
  __shared__ float red2_s[16][DH + 1];
  __shared__ float red3_s[16][2 * DH + 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & 127, part = tid >> 7;
  const int M = p.M, N = p.N, H = p.H, B = p.B;
  const bool lone = N > 64;
  const int Nm = lone ? N - 1 : N;
  const int ks = (Nm + 15) >> 4;
  const int side = row >> 6, n = row & 63;
  const int ngroups = ((B + 1) >> 1) * H;
  const int h = blockIdx.x % H;
  const bool favor = p.kind == ERV_FEAT_FAVOR;
  const bool have_state = p.state != nullptr;

  uint8_t* wh = smem;
  uint8_t* wl = wh + wbytes;
  uint8_t* xh = wl + wbytes;
  uint8_t* xl = xh + C::X_BYTES;
  uint8_t* phi1 = xl + C::X_BYTES;  // one feature half [128 tokens x 128 features], hi
  uint8_t* phi2 = phi1 + halfbytes;  // lo
  uint8_t* av = phi2 + halfbytes;    // [v|1] or [a|a16] rows, block-structured over the two pair sides
  uint8_t* simg = av + avbytes;      // [S_A|S_B] hi, lo then [dS_A|dS_B]: byte(f, j) = (j/8)*s_ch + (f/8)*128 + (f%8)*16 + (j%8)*2
  // [W^T | 1 | 0 | W^T lo] as a bf16 MN-major B operand (K = features): byte(f, j) = (j/8)*s_ch + (f/8)*128 + (f%8)*16 + (j%8)*2,
  // columns j: [0,16) hi, 16 = ones (row sum of G), [32,48) lo
  uint8_t* wimg = simg + 8 * s_ch;

  // chunk c of this thread: half c / CPH, local chunk c % CPH
  auto feat0 = [&](int c) { return (c / CPH) * HF + part * FPH + (c % CPH) * 8; };

  for (int i = tid; i < (int)(avbytes / 16); i += kTcThreads) reinterpret_cast<uint4*>(av)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (int)(2 * halfbytes / 16); i += kTcThreads) reinterpret_cast<uint4*>(phi1)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_a, 1);
    mbar_init(&bar_b, 1);
    mbar_init(&bar_c, 1);
    mbar_init(&bar_d, 1);
    mbar_init_fence();
  }
  {  // W^T of this head: hi/lo TF32 images (rows f, K = Dh) and the bf16 [W^T | 1 | 0 | W^T lo] image
    const float* om = p.omega + (size_t)h * DH * M;
    for (int i = tid; i < (int)(6 * s_ch / 16); i += kTcThreads) reinterpret_cast<uint4*>(wimg)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < Mp * DH; i += kTcThreads) {
      const int d = i / Mp, f = i % Mp;
      const float w = (f < M) ? __ldg(om + (size_t)d * M + f) : 0.f;
      const float hi = to_tf32(w), lo = to_tf32(w - hi);
      const uint32_t off = off_kmajor(f, d, 4, 4, C::X_LBO, C::X_SBO);
      *reinterpret_cast<float*>(wh + off) = hi;
      *reinterpret_cast<float*>(wl + off) = lo;
      const __nv_bfloat16 bh = __float2bfloat16_rn(w), bl = __float2bfloat16_rn(w - __bfloat162float(bh));
      const uint32_t fo = (uint32_t)(f >> 3) * 128 + (f & 7) * 16;
      *reinterpret_cast<__nv_bfloat16*>(wimg + (uint32_t)(d >> 3) * s_ch + fo + (d & 7) * 2) = bh;
      *reinterpret_cast<__nv_bfloat16*>(wimg + (uint32_t)(4 + (d >> 3)) * s_ch + fo + (d & 7) * 2) = bl;
      if (d == 0) *reinterpret_cast<__nv_bfloat16*>(wimg + 2 * s_ch + fo) = __float2bfloat16_rn(f < M ? 1.f : 0.f);
    }
  }
  fence_smem_to_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base_s;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t ph_a = 0, ph_b = 0, ph_c = 0, ph_d = 0;
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
  const uint32_t idesc_p = make_idesc(FMT_TF32, 128, Mp, false, false);
  const uint32_t idesc_acc48 = make_idesc(FMT_BF16, 128, 48, true, true);  // S / dS (+)= phi^T rows
  const uint32_t idesc_acc32 = make_idesc(FMT_BF16, 128, 32, true, true);
  const uint32_t idesc_dphi = make_idesc(FMT_BF16, 128, Mp, false, false);  // dphi = rows [S_A ; S_B]^T
  const uint32_t idesc_dv64 = make_idesc(FMT_BF16, 128, 64, false, true);   // dv = phi [dS_A|dS_B]
  const uint32_t idesc_dv32 = make_idesc(FMT_BF16, 128, 32, false, true);
  const uint32_t idesc_dq48 = make_idesc(FMT_BF16, 128, 48, false, true);   // dq' = G [W^T|1|0|W^T lo], A from tensor memory
  const uint32_t idesc_dq32 = make_idesc(FMT_BF16, 128, 32, false, true);
  const float kLog2e = 1.4426950408889634f;
  const float log2_c = log2f(p.inv_sqrt_m);
  const uint32_t rowoff = (uint32_t)(row >> 3) * 128 + (row & 7) * 16;
  uint8_t* av_side = av + (uint32_t)side * 6 * kTokCh;  // this row's pair side: chunks 0,1 hi | 2 special | 3 zero | 4,5 lo

  // ---- global rows for the next sweep: 16-byte cp.async copies into the (idle) x-image regions, 4 consecutive
  // threads per 64-byte row so that a warp touches 8 rows per instruction; consumed after the next barrier.
  // Staging: xh <- q|k rows, xl <- v|dO rows, o_raw <- saved output rows.  Lone-token rows travel in registers.
  constexpr int RB = DH * (int)sizeof(T), CPR = RB / 16;  // row bytes, 16-byte chunks per row
  RawRow<T> nx;
  float4 nv4 = make_float4(0.f, 0.f, 0.f, 0.f);
  load_raw(qkv, nx);  // defined contents; never used before a real prefetch
  auto prefetch = [&](int g, int pass) {
    if (g >= ngroups) return;
    const int b2 = g / H;
    if (!(pass == 0 && have_state) && tid < 128 * CPR) {
      const int srow = tid / CPR, sch = tid % CPR, sb = 2 * b2 + (srow >> 6), sn = srow & 63;
      if (sb < B && sn < Nm) {
        const int eo = sch * (16 / (int)sizeof(T));
        const uint32_t so = (uint32_t)srow * RB + sch * 16;
        cp_async16(xh + so, qkv + qkv_off(sb, sn, pass == 1 ? 0 : 1, h, N, H, DH) + eo);
        if (pass == 1) {
          cp_async16(xl + so, dout + out_off(sb, sn, h, N, H, DH) + eo);
          cp_async16(o_raw + so, outp + out_off(sb, sn, h, N, H, DH) + eo);
        } else {
          cp_async16(xl + so, qkv + qkv_off(sb, sn, 2, h, N, H, DH) + eo);
        }
      }
    }
    cp_async_commit();
    if ((part == 3 || have_state) && lone && pass == 0) {  // warp (pair side, q|k): rows of the last token
      const int lw = warp & 3, bb = 2 * b2 + (lw >> 1);
      if (bb < B) {
        load_raw(qkv + qkv_off(bb, N - 1, lw & 1, h, N, H, DH), nx);
        if (part != 3) {
        } else if (lw & 1) {
          if (lane < 4) nv4 = ld4(qkv + qkv_off(bb, N - 1, 2, h, N, H, DH) + 4 * lane);
        } else {
          if (lane < 4) nv4 = ld4(dout + out_off(bb, N - 1, h, N, H, DH) + 4 * lane);
          else if (lane < 8) nv4 = ld4(outp + out_off(bb, N - 1, h, N, H, DH) + 4 * (lane - 4));
        }
      }
    }
  };

  prefetch(blockIdx.x, 0);
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int b2 = g / H;
    const int b = 2 * b2 + side;
    const bool valid = b < B && n < Nm;
    float* dg_slot = (p.rot == ERV_ROT_CIRCULANT && p.dg_part)
                         ? p.dg_part + ((size_t)h * p.slots + (blockIdx.x / H) * 2) * N * DH : nullptr;  // + side * N * DH

    for (int pass = 0; pass < 3; ++pass) {  // 0: K1 (build S), 1: Q (dS, dq), 2: K2 (dv, dk)
      const int which = (pass == 1) ? 0 : 1;
      // ---- step 1: consume the prefetched rows (prepared q/k rows -> xrow_s, dO rows of the Q sweep -> o_raw)
      const bool skip = pass == 0 && have_state;  // S comes from the forward: no K1 sweep
      float st[DH + 1];  // skip: this thread's feature row of the saved [S|z], in flight across the first barrier
#pragma unroll
      for (int d = 0; d <= DH; ++d) st[d] = 0.f;
      if (skip) {
        const int sp = part >> 1, hb = part & 1, bb = 2 * b2 + sp;
        if (hb < nrb && row < HF && bb < B) {
          const float* src = p.state + ((size_t)bb * H + h) * (DH + 1) * Mp + hb * HF + row;
#pragma unroll
          for (int d = 0; d <= DH; ++d) st[d] = __ldg(src + (size_t)d * Mp);
        }
      }
      TR(pass * 100 + 0);
      float dot = 0.f;  // part 1, Q sweep: dO . O of this row
      if (!skip) {  // staged rows -> registers, then the staging regions become the x images again
        float rowv[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) rowv[d] = 0.f;
        cp_async_wait_all();
        __syncthreads();
        if (valid && part < 2) {
          RawRow<T> r;
          load_raw(reinterpret_cast<const T*>((part == 0 ? xh : xl) + (uint32_t)row * RB), r);
          raw_to_f(r, rowv);
          if (part == 1 && pass == 1) {
            float o[DH];
            load_raw(reinterpret_cast<const T*>(o_raw + (uint32_t)row * RB), r);
            raw_to_f(r, o);
#pragma unroll
            for (int d = 0; d < DH; ++d) dot = fmaf(rowv[d], o[d], dot);
          }
        }
        if (part == 1 && pass == 1) {  // own row only: no other thread reads it
          float* od = reinterpret_cast<float*>(o_raw) + row * DH;
#pragma unroll
          for (int c = 0; c < DH / 4; ++c) st4(od + 4 * c, make_float4(rowv[4 * c], rowv[4 * c + 1], rowv[4 * c + 2], rowv[4 * c + 3]));
        }
        __syncthreads();
      if (part == 0) {
        float n2 = INFINITY;
        if (valid) {
          prologue_row<DH>(rowv, p.rot, p.ta, p.tb, h, n, N, p.prescale);
          n2 = 0.f;
#pragma unroll
          for (int a = 0; a < DH; ++a) n2 = fmaf(rowv[a], rowv[a], n2);
          n2 *= 0.5f;
        }
        n2_s[row] = n2;
        store_x_images<DH>(xh, xl, rowv, row);
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) st4(&xrow_s[row][4 * c], make_float4(rowv[4 * c], rowv[4 * c + 1], rowv[4 * c + 2], rowv[4 * c + 3]));
      } else if (part == 1) {
        if (pass != 1) {  // [v_hi | 1 | 0 | v_lo]
#pragma unroll
          for (int c = 0; c < DH / 8; ++c) {
            float ch[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ch[e] = rowv[8 * c + e];
            store_split8(av_side, av_side + 4 * kTokCh, c * kTokCh + rowoff, ch);
          }
          *reinterpret_cast<uint4*>(av_side + 2 * kTokCh + rowoff) = make_uint4(valid ? 0x00003F80u : 0u, 0u, 0u, 0u);
        }
      }
      }  // !skip
      if (skip && lone) {
        // feature rows of the lone tokens on all 16 warps: warp = (feature quarter, pair side, q|k).  Raw projections
        // and the quarter's maximum go to shared memory; the exponentials follow after the barrier.
        constexpr int LQ = Mp / 4;
        const int lw = warp & 3, sp = lw >> 1, wq = lw & 1, quarter = part;
        const bool ok = 2 * b2 + sp < B;
        float x[DH];
#pragma unroll
        for (int a = 0; a < DH; ++a) x[a] = 0.f;
        if (ok) {
          raw_to_f(nx, x);
          prologue_row<DH>(x, p.rot, p.ta, p.tb, h, N - 1, N, p.prescale);
        }
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < (LQ + 31) / 32; ++i) {
          const int fl = lane + 32 * i;
          if (fl < LQ) {
            const int f = quarter * LQ + fl;
            const uint32_t wo = (uint32_t)(f >> 3) * C::X_SBO + (f & 7) * 16;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < DH / 4; ++c) {
              const float4 a = ld4(reinterpret_cast<const float*>(wh + wo + c * C::X_LBO));
              const float4 l = ld4(reinterpret_cast<const float*>(wl + wo + c * C::X_LBO));
              acc = fmaf(x[4 * c], a.x + l.x, acc); acc = fmaf(x[4 * c + 1], a.y + l.y, acc);
              acc = fmaf(x[4 * c + 2], a.z + l.z, acc); acc = fmaf(x[4 * c + 3], a.w + l.w, acc);
            }
            lone_s[wq][sp][f] = acc;
            if (f < M) m = fmaxf(m, acc);
          }
        }
        m = warp_max(m);
        if (lane == 0) lone_mx[wq][sp][quarter] = m;
        if (quarter == 3) {
          float n2 = 0.f;
#pragma unroll
          for (int a = 0; a < DH; ++a) n2 = fmaf(x[a], x[a], n2);
          if (lane == 0) {
            lone_n2[wq][sp] = 0.5f * n2;
#pragma unroll
            for (int c = 0; c < DH / 4; ++c) st4(&lone_x[wq][sp][4 * c], make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
          }
          const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (wq == 1) {
            if (lane < 4) st4(&lone_v[sp][4 * lane], ok ? nv4 : zero4);
          } else {
            if (lane < 4) st4(&lone_do[sp][4 * lane], ok ? nv4 : zero4);
            else if (lane < 8) st4(&lone_o[sp][4 * (lane - 4)], ok ? nv4 : zero4);
          }
        }
      } else if (part == 3 && lone && pass == 0) {  // the same on 4 warps (no saved state): warp = (pair side, q|k)
        const int lw = warp & 3, sp = lw >> 1, wq = lw & 1;  // wq: 0 = query, 1 = key
        const bool ok = 2 * b2 + sp < B;
        float x[DH];
#pragma unroll
        for (int a = 0; a < DH; ++a) x[a] = 0.f;
        if (ok) {
          raw_to_f(nx, x);
          prologue_row<DH>(x, p.rot, p.ta, p.tb, h, N - 1, N, p.prescale);
        }
        float n2 = 0.f;
#pragma unroll
        for (int a = 0; a < DH; ++a) n2 = fmaf(x[a], x[a], n2);
        n2 *= 0.5f;
        float pv[NC];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int f = lane + 32 * i;
          const uint32_t wo = (uint32_t)(f >> 3) * C::X_SBO + (f & 7) * 16;  // TF32 images: conflict-free 16-byte reads
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
          lone_s[wq][sp][f] = v;
        }
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < DH / 4; ++c) st4(&lone_x[wq][sp][4 * c], make_float4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]));
        }
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wq == 1) {
          if (lane < 4) st4(&lone_v[sp][4 * lane], ok ? nv4 : zero4);
        } else {
          if (lane < 4) st4(&lone_do[sp][4 * lane], ok ? nv4 : zero4);
          else if (lane < 8) st4(&lone_o[sp][4 * (lane - 4)], ok ? nv4 : zero4);
        }
      }
      fence_smem_to_async();
      fence_before_sync();
      __syncthreads();
      // ---- lone-token gradients whose partial sums were written before this barrier
      if (lone && pass > 0 && (tid == 32 || tid == 288)) {  // not the MMA-issuing thread
        const int sp = tid >> 8, bb = 2 * b2 + sp;
        if (bb < B) {
          float* slot = dg_slot ? dg_slot + (size_t)sp * N * DH : nullptr;
          if (pass == 1) {  // dq of the lone query
            float acc[DH + 1];
#pragma unroll
            for (int j = 0; j <= DH; ++j) acc[j] = 0.f;
            for (int w = 8 * sp; w < 8 * sp + 8; ++w)
#pragma unroll
              for (int j = 0; j <= DH; ++j) acc[j] += red2_s[w][j];
            float dy[DH], dxr[DH];
#pragma unroll
            for (int d = 0; d < DH; ++d) dy[d] = (favor ? acc[d] - lone_x[0][sp][d] * acc[DH] : acc[d]) * p.prescale;
            const T* xraw = qkv + qkv_off(bb, N - 1, 0, h, N, H, DH);
            prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, N - 1, N, slot, xraw);
            T* dst = dqkv + qkv_off(bb, N - 1, 0, h, N, H, DH);
#pragma unroll
            for (int c = 0; c < DH / 4; ++c) st4(dst + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
          } else {  // dv and dk of the lone key
            float acc[2 * DH + 1];
#pragma unroll
            for (int j = 0; j <= 2 * DH; ++j) acc[j] = 0.f;
            for (int w = 8 * sp; w < 8 * sp + 8; ++w)
#pragma unroll
              for (int j = 0; j <= 2 * DH; ++j) acc[j] += red3_s[w][j];
            T* dvp = dqkv + qkv_off(bb, N - 1, 2, h, N, H, DH);
#pragma unroll
            for (int c = 0; c < DH / 4; ++c) st4(dvp + 4 * c, make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]));
            float dy[DH], dxr[DH];
#pragma unroll
            for (int d = 0; d < DH; ++d)
              dy[d] = (favor ? acc[DH + d] - lone_x[1][sp][d] * acc[2 * DH] : acc[DH + d]) * p.prescale;
            const T* xraw = qkv + qkv_off(bb, N - 1, 1, h, N, H, DH);
            prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, N - 1, N, slot, xraw);
            T* dst = dqkv + qkv_off(bb, N - 1, 1, h, N, H, DH);
#pragma unroll
            for (int c = 0; c < DH / 4; ++c) st4(dst + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
          }
        }
      }
      // W^T row f (fp32 to ~21 bits) from the TF32 hi/lo images
      auto load_w_row = [&](int f, float (&wr)[DH]) {
        const uint32_t wo = (uint32_t)(f >> 3) * C::X_SBO + (f & 7) * 16;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 a = ld4(reinterpret_cast<const float*>(wh + wo + c * C::X_LBO));
          const float4 l = ld4(reinterpret_cast<const float*>(wl + wo + c * C::X_LBO));
          wr[4 * c] = a.x + l.x; wr[4 * c + 1] = a.y + l.y; wr[4 * c + 2] = a.z + l.z; wr[4 * c + 3] = a.w + l.w;
        }
      };
      // ---- end of sweep: move the TMEM accumulator (S after K1, dS after Q) into the bf16 images; lone-token terms
      auto sweep_tail = [&](const bool from_state, const float (&st)[DH + 1]) {
        TR(pass * 100 + 20);
        const int sp = part >> 1, hb = part & 1, f = hb * HF + row;
        const bool own = hb < nrb && row < HF;  // warp-uniform
        float sv[DH], zz = 0.f;
#pragma unroll
        for (int d = 0; d < DH; ++d) sv[d] = 0.f;
        float pq = 0.f, pk = 0.f;
        if (own) {
          if (from_state) {  // the forward's finished [S|z] (lone key included)
#pragma unroll
            for (int d = 0; d < DH; ++d) sv[d] = st[d];
            zz = st[DH];
          } else {
            float d0[32], d1[16];
            tmem_ld32(tm + lane_off + COL_S + (uint32_t)(sp * nrb + hb) * S_STRIDE, d0);
            tmem_ld16(tm + lane_off + COL_S + (uint32_t)(sp * nrb + hb) * S_STRIDE + 32, d1);
#pragma unroll
            for (int d = 0; d < DH; ++d) sv[d] = d0[d] + d1[d];
            zz = (pass == 0) ? d0[DH] : d0[DH] + d0[DH + 1];
          }
          if (lone) {
            pq = lone_s[0][sp][f];
            pk = lone_s[1][sp][f];
            if (from_state) {
            } else if (pass == 0) {  // rank-1 term of the last key
#pragma unroll
              for (int d = 0; d < DH; ++d) sv[d] = fmaf(pk, lone_v[sp][d], sv[d]);
              zz += pk;
            } else {  // rank-1 term of the last query
#pragma unroll
              for (int d = 0; d < DH; ++d) sv[d] = fmaf(pq, lone_a[sp][d], sv[d]);
              zz = fmaf(pq, lone_a[sp][DH], zz);
            }
          }
          if (pass == 0) z_s[sp][f] = zz; else dz_s[sp][f] = zz;
#pragma unroll
          for (int c = 0; c < DH / 8; ++c) {
            float ch[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ch[e] = sv[8 * c + e];
            store_split8(simg + (uint32_t)(2 * sp + c) * s_ch, simg + (uint32_t)(4 + 2 * sp + c) * s_ch,
                         (uint32_t)(f >> 3) * 128 + (f & 7) * 16, ch);
          }
        }
        if (lone) {
          if (pass == 0) {
            const float s = warp_sum(pq * zz);  // den of the lone query
            if (lane == 0) red1_s[warp] = s;
            fence_smem_to_async();
            __syncthreads();
            float den = kEps;
            for (int w = 8 * sp; w < 8 * sp + 8; ++w) den += red1_s[w];
            const float r = 1.0f / den;
            float dot = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) dot = fmaf(lone_do[sp][d], lone_o[sp][d], dot);
            const float a16 = -dot * r;
            float dph = a16 * zz;
#pragma unroll
            for (int d = 0; d < DH; ++d) dph = fmaf(lone_do[sp][d] * r, sv[d], dph);
            const float gq = own ? (favor ? dph * pq : (pq > 0.f ? dph * p.inv_sqrt_m : 0.f)) : 0.f;
            float wr[DH];
            load_w_row(own ? f : 0, wr);
            float a32[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) a32[j] = (j < DH) ? gq * wr[j] : (j == DH ? gq : 0.f);
            const float t = warp_sum32(a32);  // lane j: warp total of partial j
            if (lane <= DH) red2_s[warp][lane] = t;
            if ((tid & 255) == 0) {
#pragma unroll
              for (int d = 0; d < DH; ++d) lone_a[sp][d] = lone_do[sp][d] * r;
              lone_a[sp][DH] = a16;
            }
          } else {
            float dph = zz;  // dphi of the lone key: v_L . dS[f] + dz[f]
#pragma unroll
            for (int d = 0; d < DH; ++d) dph = fmaf(lone_v[sp][d], sv[d], dph);
            const float gk = own ? (favor ? dph * pk : (pk > 0.f ? dph * p.inv_sqrt_m : 0.f)) : 0.f;
            float wr[DH];
            load_w_row(own ? f : 0, wr);
            float a32[32];  // [0, DH): dv of the lone key, [DH, 2 DH): G [W^T] ; the row sum goes separately
#pragma unroll
            for (int d = 0; d < DH; ++d) {
              a32[d] = pk * sv[d];
              a32[DH + d] = gk * wr[d];
            }
            const float t = warp_sum32(a32);
            red3_s[warp][lane] = t;
            const float t2 = warp_sum(gk);
            if (lane == 0) red3_s[warp][2 * DH] = t2;
          }
        }
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
      };
      if (skip) {  // the forward saved [S|z]: fetch this thread's feature row and go straight to the end of the sweep
        prefetch(g, 1);
        if (lone) {  // raw projections of the lone rows -> features
          for (int e = tid; e < 4 * Mp; e += kTcThreads) {
            const int f = e % Mp, rw = e / Mp, wq = rw & 1, sp = rw >> 1;
            const float mq = fmaxf(fmaxf(lone_mx[wq][sp][0], lone_mx[wq][sp][1]), fmaxf(lone_mx[wq][sp][2], lone_mx[wq][sp][3]));
            const float pv = lone_s[wq][sp][f];
            float v = favor ? ex2_approx(fmaf(pv, kLog2e, -fmaf(mq + lone_n2[wq][sp], kLog2e, -log2_c)))
                            : fmaxf(pv, 0.f) * p.inv_sqrt_m;
            if (f >= M || 2 * b2 + sp >= B) v = 0.f;
            lone_s[wq][sp][f] = v;
          }
          __syncthreads();
        }
        sweep_tail(true, st);
      } else {
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
#pragma unroll
      for (int c = 0; c < NC; ++c) tmem_ld8_nowait(tm + lane_off + COL_P + feat0(c), pr[c]);
#pragma unroll
      for (int c = 0; c < NC; ++c) tmem_wait_ld8(pr[c]);
      if (favor) {
        float m_part = -INFINITY;
        if (M < Mp) {  // padded features do not take part in the maximum
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (feat0(c) + i < M) m_part = fmaxf(m_part, __uint_as_float(pr[c][i]));
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) m_part = fmaxf(m_part, __uint_as_float(pr[c][i]));
        }
        ex_s[part][row] = m_part;
      }
      fence_before_sync();
      __syncthreads();  // all P reads are done: the P columns may be overwritten (dphi)
      if (pass == 2 && warp == 0 && elect_one()) {  // K2: dphi_k = v [dS_A ; dS_B]^T can start as soon as P has been consumed
        fence_after_sync();
        bool acc = false;
        for (int sp = 0; sp < 2; ++sp)
          for (int term = 0; term < 3; ++term) {
            const uint32_t a_off = (uint32_t)(sp * 6 + (term == 2 ? 4 : 0)) * kTokCh;
            const uint32_t b_off = (uint32_t)((term == 1 ? 4 : 0) + 2 * sp) * s_ch;
            mma_f16(tm + COL_P, make_desc(smem_u32(av) + a_off, kTokCh, 128), make_desc(smem_u32(simg) + b_off, s_ch, 128),
                    idesc_dphi, acc);
            acc = true;
          }
        commit(&bar_c);
      }
      {
        float mx = 0.f;
        if (favor) mx = fmaxf(fmaxf(ex_s[0][row], ex_s[1][row]), fmaxf(ex_s[2][row], ex_s[3][row]));
        const float n2 = n2_s[row];
        const float shift = fmaf(mx + n2, kLog2e, -log2_c);
        const float scale = (n2 < INFINITY) ? p.inv_sqrt_m : 0.f;
        if (favor) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pr[c][i] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(pr[c][i]), kLog2e, -shift)));
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) pr[c][i] = __float_as_uint(fmaxf(__uint_as_float(pr[c][i]), 0.f) * scale);
        }
        if (M < Mp) {  // padded features contribute nothing
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (feat0(c) + i >= M) pr[c][i] = 0u;
        }
      }
      TR(pass * 100 + 3);
      // stores one feature half of the values held in pr into the phi images
      auto store_half = [&](int hb) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c / CPH == hb) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(pr[c][i]);
            store_split8(phi1, phi2, (uint32_t)(part * CPH + (c % CPH)) * kTokCh + rowoff, v);
          }
      };
      // D[COL_S + (sp*nrb + hb)*64] = phi_half^T rows of pair side sp (rows image = av); one thread
      auto issue_accumulate = [&](int hb) {
        for (int sp = 0; sp < 2; ++sp) {
          const uint32_t d = tm + COL_S + (uint32_t)(sp * nrb + hb) * S_STRIDE;
          for (int s = 0; s < ks; ++s) {
            const uint32_t st = (uint32_t)(sp * 4 + s) * 256;
            const uint64_t bd = make_desc(smem_u32(av) + (uint32_t)sp * 6 * kTokCh + st, 128, kTokCh);
            mma_f16(d, make_desc(smem_u32(phi1) + st, 128, kTokCh), bd, idesc_acc48, s > 0);
            mma_f16(d, make_desc(smem_u32(phi2) + st, 128, kTokCh), bd, idesc_acc32, true);
          }
        }
        commit(&bar_b);
      };
      // G (in pr) -> bf16 hi/lo, written over the dphi columns THIS thread has just consumed (its features of half hb sit in the
      // FPH fp32 columns [hb*HF + part*FPH, +FPH): hi words go to the first FPH/2 of them, lo words to the rest), so no other
      // thread's unread dphi is touched.  Then dq' = G [W^T|1] on the tensor pipe with A read from tensor memory: k-step
      // (hb, part, j) covers features hb*HF + part*FPH + 16 j.
      auto store_g_tmem = [&]() {
#pragma unroll
        for (int hb = 0; hb < NRB; ++hb)
#pragma unroll
          for (int cc = 0; cc < CPH; cc += 2) {
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
              for (int i = 0; i < 4; ++i)
                split_pack2(__uint_as_float(pr[hb * CPH + cc + e][2 * i]), __uint_as_float(pr[hb * CPH + cc + e][2 * i + 1]),
                            hw[4 * e + i], lw[4 * e + i]);
            const uint32_t col = tm + lane_off + COL_P + (uint32_t)(hb * HF + part * FPH + cc * 4);
            tmem_st8(col, hw);
            tmem_st8(col + FPH / 2, lw);
          }
        tmem_wait_st();
      };
      auto issue_g_product = [&]() {  // one elected thread
        bool acc = false;
        for (int hb = 0; hb < NRB; ++hb)
          for (int pq = 0; pq < 4; ++pq)
            for (int j = 0; j < FPH / 16; ++j) {
              const uint32_t fs = (uint32_t)(hb * HF + pq * FPH + 16 * j);
              const uint32_t ca = tm + COL_P + (uint32_t)(hb * HF + pq * FPH + 8 * j);
              const uint64_t bd = make_desc(smem_u32(wimg) + fs * 16, 128, s_ch);
              mma_f16_ts(tm + COL_DQ, ca, bd, idesc_dq48, acc);
              mma_f16_ts(tm + COL_DQ, ca + FPH / 2, bd, idesc_dq32, true);
              acc = true;
            }
        commit(&bar_d);
      };
      // dq' (TMEM) -> this row's 17 sums; valid in part 0
      auto load_g_product = [&](float (&acc)[RW]) {
        if (part == 0) {
          float d0[32], d1[16];
          tmem_ld32(tm + lane_off + COL_DQ, d0);
          tmem_ld16(tm + lane_off + COL_DQ + 32, d1);
#pragma unroll
          for (int d = 0; d < DH; ++d) acc[d] = d0[d] + d1[d];
          acc[DH] = d0[DH];
        }
      };
      // reduced sums (part 0) -> gradient wrt the raw q/k row, written to global memory
      auto store_input_gradient = [&](const float (&acc)[RW]) {
        if (part == 0 && valid) {
          float dy[DH], dxr[DH];
#pragma unroll
          for (int d = 0; d < DH; ++d) dy[d] = (favor ? acc[d] - xrow_s[row][d] * acc[DH] : acc[d]) * p.prescale;
          const T* xraw = qkv + qkv_off(b, n, which, h, N, H, DH);
          prologue_row_bwd<T, DH>(dy, dxr, p.rot, p.ta, p.tb, h, n, N, dg_slot ? dg_slot + (size_t)side * N * DH : nullptr, xraw);
          T* dst = dqkv + qkv_off(b, n, which, h, N, H, DH);
#pragma unroll
          for (int c = 0; c < DH / 4; ++c) st4(dst + 4 * c, make_float4(dxr[4 * c], dxr[4 * c + 1], dxr[4 * c + 2], dxr[4 * c + 3]));
        }
      };
      // dphi (TMEM, P columns) + rank-1 term -> G = dphi (.) dphi/dP, in place in pr
      auto load_dphi_to_g = [&](const float* rank1, float rscale) {  // dphi[f] += rscale * rank1[f]
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t r[8];
          tmem_ld8_nowait(tm + lane_off + COL_P + feat0(c), r);
          const float4 za = ld4(rank1 + feat0(c)), zb = ld4(rank1 + feat0(c) + 4);
          const float zz[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
          tmem_wait_ld8(r);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ph_v = __uint_as_float(pr[c][i]);
            const float dph = fmaf(rscale, zz[i], __uint_as_float(r[i]));
            const float gq = favor ? dph * ph_v : (ph_v > 0.f ? dph * p.inv_sqrt_m : 0.f);
            pr[c][i] = __float_as_uint(gq);
          }
        }
      };

      if (pass == 0) {
        // ---- K1: S(pair, hb) = phi_k^T [v|1]
        for (int hb = 0; hb < nrb; ++hb) {
          store_half(hb);
          fence_smem_to_async();
          fence_before_sync();
          __syncthreads();
          if (warp == 0 && elect_one()) {
            fence_after_sync();
            issue_accumulate(hb);
          }
          if (hb == nrb - 1) prefetch(g, 1);
          mbar_wait(&bar_b, ph_b);
          ph_b ^= 1;
          fence_after_sync();
        }
      } else if (pass == 1) {
        // ---- Q: den, a, dS = phi_q^T a, dphi_q = a S^T
        float den_part = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 za = ld4(&z_s[side][feat0(c)]), zb = ld4(&z_s[side][feat0(c) + 4]);
          den_part = fmaf(__uint_as_float(pr[c][0]), za.x, den_part); den_part = fmaf(__uint_as_float(pr[c][1]), za.y, den_part);
          den_part = fmaf(__uint_as_float(pr[c][2]), za.z, den_part); den_part = fmaf(__uint_as_float(pr[c][3]), za.w, den_part);
          den_part = fmaf(__uint_as_float(pr[c][4]), zb.x, den_part); den_part = fmaf(__uint_as_float(pr[c][5]), zb.y, den_part);
          den_part = fmaf(__uint_as_float(pr[c][6]), zb.z, den_part); den_part = fmaf(__uint_as_float(pr[c][7]), zb.w, den_part);
        }
        den_s[part][row] = den_part;
        store_half(0);  // does not depend on den: overlaps the exchange
        __syncthreads();
        TR(104);
        if (part == 1) {
          const float r = 1.0f / ((den_s[0][row] + den_s[1][row]) + (den_s[2][row] + den_s[3][row]) + kEps);
#pragma unroll
          for (int c = 0; c < DH / 8; ++c) {
            float ch[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) ch[e] = reinterpret_cast<const float*>(o_raw)[row * DH + 8 * c + e] * r;
            store_split8(av_side, av_side + 4 * kTokCh, c * kTokCh + rowoff, ch);
          }
          const float a16 = valid ? -dot * r : 0.f;
          a16_s[row] = a16;
          const uint32_t hi = __float_as_uint(a16) & 0xffff0000u;
          const float lo = a16 - __uint_as_float(hi);
          *reinterpret_cast<uint4*>(av_side + 2 * kTokCh + rowoff) =
              make_uint4((hi >> 16) | (__float_as_uint(lo) & 0xffff0000u), 0u, 0u, 0u);  // [a16_hi, a16_lo]
        }
        fence_smem_to_async();
        fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
          fence_after_sync();
          issue_accumulate(0);
          // dphi_q for the whole row, into the (consumed) P columns: K = [pair A's 16 | pair B's 16]
          bool acc = false;
          for (int sp = 0; sp < 2; ++sp)
            for (int term = 0; term < 3; ++term) {
              const uint32_t a_off = (uint32_t)(sp * 6 + (term == 2 ? 4 : 0)) * kTokCh;
              const uint32_t b_off = (uint32_t)((term == 1 ? 4 : 0) + 2 * sp) * s_ch;
              mma_f16(tm + COL_P, make_desc(smem_u32(av) + a_off, kTokCh, 128), make_desc(smem_u32(simg) + b_off, s_ch, 128),
                      idesc_dphi, acc);
              acc = true;
            }
          commit(&bar_c);
        }
        TR(105);
        mbar_wait(&bar_b, ph_b);
        ph_b ^= 1;
        fence_after_sync();
        TR(106);
        if (nrb > 1) {
          store_half(1);
          fence_smem_to_async();
          fence_before_sync();
          __syncthreads();
          if (warp == 0 && elect_one()) {
            fence_after_sync();
            issue_accumulate(1);
          }
        }
        TR(107);
        mbar_wait(&bar_c, ph_c);
        ph_c ^= 1;
        fence_after_sync();
        TR(108);
        load_dphi_to_g(&z_s[side][0], a16_s[row]);
        TR(109);
        store_g_tmem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
          fence_after_sync();
          issue_g_product();
        }
        TR(110);
        if (nrb > 1) {
          mbar_wait(&bar_b, ph_b);  // the second half's contraction has completed
          ph_b ^= 1;
          fence_after_sync();
        }
        prefetch(g, 2);
        TR(111);
      } else {
        // ---- K2: dv = phi_k [dS_A|dS_B] ; dphi_k (already issued) ; dk
        for (int hb = 0; hb < nrb; ++hb) {
          store_half(hb);
          fence_smem_to_async();
          fence_before_sync();
          __syncthreads();
          if (warp == 0 && elect_one()) {
            fence_after_sync();
            for (int s = 0; s < HF / 16; ++s) {
              const uint64_t bd = make_desc(smem_u32(simg) + (uint32_t)(hb * (HF / 16) + s) * 256, 128, s_ch);
              mma_f16(tm + COL_S, make_desc(smem_u32(phi1) + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_dv64, hb > 0 || s > 0);
              mma_f16(tm + COL_S, make_desc(smem_u32(phi2) + (uint32_t)s * 2 * kTokCh, kTokCh, 128), bd, idesc_dv32, true);
            }
            commit(&bar_b);
          }
          if (hb < nrb - 1) {
            mbar_wait(&bar_b, ph_b);
            ph_b ^= 1;
            fence_after_sync();
          }
        }
        TR(207);
        mbar_wait(&bar_c, ph_c);
        ph_c ^= 1;
        fence_after_sync();
        TR(208);
        load_dphi_to_g(&dz_s[side][0], 1.0f);
        TR(209);
        store_g_tmem();
        fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
          fence_after_sync();
          issue_g_product();
        }
        TR(210);
        mbar_wait(&bar_b, ph_b);
        ph_b ^= 1;
        fence_after_sync();
        TR(211);
        prefetch(g + gridDim.x, 0);
        if (part == 0) {  // dv rows: hi-part + lo-part columns of this row's pair
          float a0[16], a1[16];
          tmem_ld16(tm + lane_off + COL_S + 16 * side, a0);
          tmem_ld16(tm + lane_off + COL_S + 32 + 16 * side, a1);
          if (valid) {
            T* dvp = dqkv + qkv_off(b, n, 2, h, N, H, DH);
#pragma unroll
            for (int c = 0; c < DH / 4; ++c)
              st4(dvp + 4 * c, make_float4(a0[4 * c] + a1[4 * c], a0[4 * c + 1] + a1[4 * c + 1], a0[4 * c + 2] + a1[4 * c + 2],
                                           a0[4 * c + 3] + a1[4 * c + 3]));
          }
        }
        {  // dk' = G [W^T|1]
          mbar_wait(&bar_d, ph_d);
          ph_d ^= 1;
          fence_after_sync();
          float acc[RW];
          load_g_product(acc);
          store_input_gradient(acc);
          TR(212);
        }
        fence_before_sync();
        __syncthreads();  // the feature images and the TMEM columns are reused by the next group
        TR(213);
      }
        if (pass < 2) {
          const float none[DH + 1] = {};
          sweep_tail(false, none);
        }
        if (pass == 1) {  // dq' ran on the tensor pipe under the sweep tail
          mbar_wait(&bar_d, ph_d);
          ph_d ^= 1;
          fence_after_sync();
          float acc[RW];
          load_g_product(acc);
          store_input_gradient(acc);
          fence_before_sync();
          TR(112);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static int tc2b_mp(int M) { return M <= 64 ? 64 : (M <= 128 ? 128 : 256); }

size_t la_tc2_bwd_smem_bytes(int M) {
  const int Mp = tc2b_mp(M);
  const size_t wbytes = (size_t)(Mp / 8) * 4 * 128, xbytes = 16 * 4 * 128;
  return 2 * wbytes + 2 * xbytes + 2 * 16 * (size_t)kTokCh + 12 * (size_t)kTokCh + 8 * (size_t)(Mp / 8) * 128 +
         6 * (size_t)(Mp / 8) * 128 + 128;
}

int la_tc2_backward(const void* qkv, const void* out, const void* dout, void* dqkv, const float* omega, int B, int N,
                    int H, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                    int dtype, const float* state, cudaStream_t st) {
  LaTc2BwdArgs a;
  a.qkv = qkv; a.out = out; a.dout = dout; a.dqkv = dqkv; a.omega = omega; a.ta = ta; a.tb = tb; a.dg_part = dg_part;
  a.B = B; a.N = N; a.H = H; a.M = M; a.kind = kind; a.rot = rot; a.slots = slots; a.state = state;
  a.trace = g_trace;
  a.prescale = (float)pow(16.0, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  const int Mp = tc2b_mp(M);
  const size_t smem = la_tc2_bwd_smem_bytes(M);
  const int ngroups = ((B + 1) / 2) * H;
  int grid = (kNumSMs / H) * H;
  if (grid < H) grid = H;
  if (grid > ngroups) grid = ngroups;
  if (dg_part != nullptr && 2 * (grid / H) > slots) grid = (slots / 2) * H;  // two gradient slots per CTA
  if (grid < H) { set_error("tensor-core backward: %d gradient slots are too few", slots); return ERV_E_INVALID; }
#define TC2B_LAUNCH(TT, NRB_, CPH_)                                              \
  do {                                                                           \
    ERV_CUDA(allow_smem(la_tc2_bwd_kernel<TT, NRB_, CPH_>, smem));               \
    la_tc2_bwd_kernel<TT, NRB_, CPH_><<<grid, kTcThreads, smem, st>>>(a);        \
  } while (0)
  if (dtype == ERV_F32) {
    if (Mp == 64) TC2B_LAUNCH(float, 1, 2); else if (Mp == 128) TC2B_LAUNCH(float, 1, 4); else TC2B_LAUNCH(float, 2, 4);
  } else {
    if (Mp == 64) TC2B_LAUNCH(__nv_bfloat16, 1, 2); else if (Mp == 128) TC2B_LAUNCH(__nv_bfloat16, 1, 4); else TC2B_LAUNCH(__nv_bfloat16, 2, 4);
  }
#undef TC2B_LAUNCH
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

}  // namespace erv
