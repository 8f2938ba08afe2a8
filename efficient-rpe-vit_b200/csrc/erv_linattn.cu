// Fused FAVOR+/ReLU linear attention (forward + backward) and the stand-alone random-feature map.
//
// One CTA owns one (batch, head) pair at a time (persistent loop; grid is a multiple of H so a CTA
// always sees the same head).  Tokens stream through shared memory in tiles of TT; the feature tile
// phi[TT][M] never leaves the SM, so HBM traffic is the packed q/k/v read and the out write
// (4*B*N*C*s forward, 12*B*N*C*s forward+backward).
//
//   forward   pass K: S[f][0..DH) += phi(k)^T v, S[f][DH] (= z) += phi(k)^T 1
//             pass Q: [num | den] = phi(q) S ; out = num / (den + 1e-6)
//   backward  pass K: rebuild S;  pass Q: dS, dq;  pass K again: dv, dk      
#include "erv_feat.cuh"

namespace erv {

struct LaArgs {
  const void* qkv;
  void* out;          // fwd: output; bwd: saved forward output (read)
  const void* dout;
  void* dqkv;
  const float* wt;    // [H][Mp][DH+4] prepared W^T (col DH = 1 for FAVOR+)
  const float* ta;
  const float* tb;
  float* dg_part;     // [H][slots][N][DH]
  int B, N, H, kind, rot, slots;
  float prescale, inv_sqrt_m;
  FeatGeom g;
};

struct LaSmem {
  int S, dS, phi, xr, xs, v, a, dO, O, tmp, g2, m, n2, den, inv, pm, red, total;
};

__host__ __device__ inline LaSmem la_layout(int DH, int Mp, int ldp, int nthreads, int TT, bool bwd, bool circ) {
  const int LDM = DH + 4, tile = TT * LDM;
  LaSmem L;
  int o = 0;
  L.S = o; o += Mp * LDM;
  L.dS = o; if (bwd) o += Mp * LDM;
  L.phi = o; o += (TT * ldp + 3) / 4 * 4;
  L.xr = o; o += tile;
  L.xs = o; o += tile;
  L.v = o; o += tile;
  L.a = o; if (bwd) o += tile;
  L.dO = o; if (bwd) o += tile;
  L.O = o; if (bwd) o += tile;
  L.tmp = o; if (bwd) o += tile;
  L.g2 = o; if (circ) o += TT * 2 * DH;
  L.m = o; o += TT;
  L.n2 = o; o += TT;
  L.den = o; o += TT;
  L.inv = o; o += TT;
  L.pm = o; o += (nthreads / TT + 1) * TT;
  int items = (TT / 4) * (DH / 4 + 1);
  int ks = nthreads / items;
  ks = ks < 1 ? 1 : (ks > 6 ? 6 : ks);
  int items2 = (TT / 4) * (DH / 4);
  int ks2 = nthreads / items2;
  ks2 = ks2 < 1 ? 1 : (ks2 > 6 ? 6 : ks2);
  L.red = o; o += (ks > ks2 ? ks : ks2) * tile;
  L.total = o;
  return L;
}

// omega [H][DH][M] -> wt [H][Mp][DH+4]
__global__ void wt_prep_kernel(const float* __restrict__ omega, float* __restrict__ wt, int H, int DH, int M, int Mp,
                               float extra) {
  const int LDM = DH + 4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * Mp * LDM) return;
  int j = (int)(i % LDM);
  int f = (int)((i / LDM) % Mp);
  int h = (int)(i / ((size_t)LDM * Mp));
  float v = 0.f;
  if (f < M) {
    if (j < DH) v = omega[((size_t)h * DH + j) * M + f];
    else if (j == DH) v = extra;
  }
  wt[i] = v;
}

// tile of prepared tokens -> feature tile, shared by every pass
template <typename T, int DH, int TT>
__device__ __forceinline__ void tokens_to_features(const LaArgs& p, const LaSmem& L, float* smem, const T* src,
                                                   size_t tok_stride, int h, int n0, const float* wt) {
  RotArgs ra{p.rot, p.ta, p.tb};
  load_tile<T, DH, TT>(smem + L.xr, src, tok_stride, n0, p.N, 0.f);
  if (p.rot == ERV_ROT_CIRCULANT) load_g_tile<DH, TT>(smem + L.g2, p.ta, h, n0, p.N);
  __syncthreads();
  prep_tile<DH, TT>(smem + L.xs, smem + L.xr, smem + L.g2, smem + L.inv, ra, ERV_PREP_SCALE, p.prescale, n0, p.N);
  __syncthreads();
  feature_tile<DH, TT>(smem + L.phi, smem + L.xs, wt, smem + L.m, smem + L.n2, smem + L.pm, p.g, p.kind, p.inv_sqrt_m);
}

template <typename T, int DH, int TT>
__global__ void __launch_bounds__(320, 1) la_fwd_kernel(const LaArgs p) {
  constexpr int LDM = DH + 4, NJG = DH / 4 + 1;
  extern __shared__ __align__(16) float smem[];
  const LaSmem L = la_layout(DH, p.g.Mp, p.g.ldp, blockDim.x, TT, false, p.rot == ERV_ROT_CIRCULANT);
  float* S_s = smem + L.S;
  float* red_s = smem + L.red;
  const T* qkv = static_cast<const T*>(p.qkv);
  T* out = static_cast<T*>(p.out);
  const size_t tok_stride = (size_t)3 * p.H * DH;
  const size_t out_stride = (size_t)p.H * DH;

  for (int pair = blockIdx.x; pair < p.B * p.H; pair += gridDim.x) {
    const int b = pair / p.H, h = pair % p.H;
    const float* wt = p.wt + (size_t)h * p.g.Mp * LDM;
    const T* qb = qkv + qkv_off(b, 0, 0, h, p.N, p.H, DH);
    const T* kb = qkv + qkv_off(b, 0, 1, h, p.N, p.H, DH);
    const T* vb = qkv + qkv_off(b, 0, 2, h, p.N, p.H, DH);
    for (int i = threadIdx.x; i < p.g.Mp * LDM; i += blockDim.x) S_s[i] = 0.f;
    __syncthreads();
    for (int n0 = 0; n0 < p.N; n0 += TT) {  // keys
      load_tile<T, DH, TT>(smem + L.v, vb, tok_stride, n0, p.N, 1.f);
      tokens_to_features<T, DH, TT>(p, L, smem, kb, tok_stride, h, n0, wt);
      h1_accum<DH, TT>(S_s, smem + L.phi, smem + L.v, p.g);
      __syncthreads();
    }
    for (int n0 = 0; n0 < p.N; n0 += TT) {  // queries
      tokens_to_features<T, DH, TT>(p, L, smem, qb, tok_stride, h, n0, wt);
      h2_narrow<DH, TT, NJG>(red_s, smem + L.phi, p.g.ldp, S_s, LDM, p.g.M);
      T* ob = out + out_off(b, 0, h, p.N, p.H, DH);
      for (int i = threadIdx.x; i < TT * (DH / 4); i += blockDim.x) {
        int t = i / (DH / 4), v = i % (DH / 4), n = n0 + t;
        if (n >= p.N) continue;
        float den = red_s[t * LDM + DH] + kEps;
        float4 x = ld4(red_s + t * LDM + 4 * v);
        st4(ob + (size_t)n * out_stride + 4 * v, make_float4(x.x / den, x.y / den, x.z / den, x.w / den));
      }
      __syncthreads();
    }
  }
}

// shared tail of the two gradient passes: G tile (in phi) -> dxs -> gradient wrt the raw q or k tile
template <typename T, int DH, int TT>
__device__ __forceinline__ void features_bwd_tail(const LaArgs& p, const LaSmem& L, float* smem, const float* wt,
                                                  T* dst, size_t tok_stride, float* dg_slot, int n0) {
  constexpr int LDM = DH + 4, NJG = DH / 4 + 1;
  RotArgs ra{p.rot, p.ta, p.tb};
  float* red_s = smem + L.red;
  h2_narrow<DH, TT, NJG>(red_s, smem + L.phi, p.g.ldp, wt, LDM, p.g.M);
  if (p.kind == ERV_FEAT_FAVOR) {  // dX = G W^T - X * rowsum(G)
    const float* xs = smem + L.xs;
    for (int i = threadIdx.x; i < TT * DH; i += blockDim.x) {
      int t = i / DH, d = i % DH;
      red_s[t * LDM + d] -= xs[t * LDM + d] * red_s[t * LDM + DH];
    }
    __syncthreads();
  }
  prep_tile_bwd<T, DH, TT>(red_s, smem + L.xs, smem + L.xr, smem + L.g2, smem + L.inv, smem + L.tmp, ra,
                           ERV_PREP_SCALE, p.prescale, dst, tok_stride, dg_slot, n0, p.N);
  __syncthreads();
}

template <typename T, int DH, int TT>
__global__ void __launch_bounds__(320, (DH <= 16) ? 2 : 1) la_bwd_kernel(const LaArgs p) {
  constexpr int LDM = DH + 4;
  extern __shared__ __align__(16) float smem[];
  const LaSmem L = la_layout(DH, p.g.Mp, p.g.ldp, blockDim.x, TT, true, p.rot == ERV_ROT_CIRCULANT);
  float* S_s = smem + L.S;
  float* dS_s = smem + L.dS;
  float* phi_s = smem + L.phi;
  float* a_s = smem + L.a;
  float* red_s = smem + L.red;
  const T* qkv = static_cast<const T*>(p.qkv);
  const T* outp = static_cast<const T*>(p.out);
  const T* dout = static_cast<const T*>(p.dout);
  T* dqkv = static_cast<T*>(p.dqkv);
  const size_t tok_stride = (size_t)3 * p.H * DH;
  const size_t out_stride = (size_t)p.H * DH;
  const int ldp = p.g.ldp, kind = p.kind;
  const float c = p.inv_sqrt_m;

  for (int pair = blockIdx.x; pair < p.B * p.H; pair += gridDim.x) {
    const int b = pair / p.H, h = pair % p.H;
    const float* wt = p.wt + (size_t)h * p.g.Mp * LDM;
    float* dg_slot = p.dg_part ? p.dg_part + ((size_t)h * p.slots + blockIdx.x / p.H) * p.N * DH : nullptr;
    const T* qb = qkv + qkv_off(b, 0, 0, h, p.N, p.H, DH);
    const T* kb = qkv + qkv_off(b, 0, 1, h, p.N, p.H, DH);
    const T* vb = qkv + qkv_off(b, 0, 2, h, p.N, p.H, DH);
    T* dqb = dqkv + qkv_off(b, 0, 0, h, p.N, p.H, DH);
    T* dkb = dqkv + qkv_off(b, 0, 1, h, p.N, p.H, DH);
    T* dvb = dqkv + qkv_off(b, 0, 2, h, p.N, p.H, DH);
    for (int i = threadIdx.x; i < p.g.Mp * LDM; i += blockDim.x) { S_s[i] = 0.f; dS_s[i] = 0.f; }
    __syncthreads();
    // ---- pass K: S = phi(k)^T [v | 1]
    for (int n0 = 0; n0 < p.N; n0 += TT) {
      load_tile<T, DH, TT>(smem + L.v, vb, tok_stride, n0, p.N, 1.f);
      tokens_to_features<T, DH, TT>(p, L, smem, kb, tok_stride, h, n0, wt);
      h1_accum<DH, TT>(S_s, phi_s, smem + L.v, p.g);
      __syncthreads();
    }
    // ---- pass Q: dS += phi(q)^T [dnum | dden]; dq
    for (int n0 = 0; n0 < p.N; n0 += TT) {
      load_tile<T, DH, TT>(smem + L.dO, dout + out_off(b, 0, h, p.N, p.H, DH), out_stride, n0, p.N, 0.f);
      load_tile<T, DH, TT>(smem + L.O, outp + out_off(b, 0, h, p.N, p.H, DH), out_stride, n0, p.N, 0.f);
      tokens_to_features<T, DH, TT>(p, L, smem, qb, tok_stride, h, n0, wt);
      row_reduce<TT, false>(smem + L.den, smem + L.pm, p.g.M,
                            [&](int t, int f) { return phi_s[t * ldp + f] * S_s[f * LDM + DH]; });
      {  // a = [dO * r | -(dO . O) * r | 0 0 0],  r = 1 / (den + eps)
        const float* dO = smem + L.dO;
        const float* O = smem + L.O;
        for (int i = threadIdx.x; i < TT * LDM; i += blockDim.x) {
          int t = i / LDM, j = i % LDM;
          float r = 1.0f / (smem[L.den + t] + kEps);
          float v = 0.f;
          if (j < DH) {
            v = dO[t * LDM + j] * r;
          } else if (j == DH) {
            float dot = 0.f;
#pragma unroll 8
            for (int d = 0; d < DH; ++d) dot += dO[t * LDM + d] * O[t * LDM + d];
            v = -dot * r;
          }
          a_s[i] = v;
        }
      }
      __syncthreads();
      h1_accum<DH, TT>(dS_s, phi_s, a_s, p.g);
      __syncthreads();
      // dphi = [dnum | dden] . [S | z]^T, turned into G in place
      h1_rows<LDM, LDM, TT>(a_s, S_s, LDM, p.g, [&](int t, int f, float acc) {
        phi_s[t * ldp + f] = feature_grad(acc, phi_s[t * ldp + f], kind, c);
      });
      __syncthreads();
      features_bwd_tail<T, DH, TT>(p, L, smem, wt, dqb, tok_stride, dg_slot, n0);
    }
    // ---- pass K again: dv = phi(k) dS ; dphi(k) = [v | 1] . [dS | dz]^T ; dk
    for (int n0 = 0; n0 < p.N; n0 += TT) {
      load_tile<T, DH, TT>(smem + L.v, vb, tok_stride, n0, p.N, 1.f);
      tokens_to_features<T, DH, TT>(p, L, smem, kb, tok_stride, h, n0, wt);
      h2_narrow<DH, TT, DH / 4>(red_s, phi_s, ldp, dS_s, LDM, p.g.M);
      store_tile<T, DH, TT>(dvb, tok_stride, red_s, n0, p.N);
      h1_rows<LDM, LDM, TT>(smem + L.v, dS_s, LDM, p.g, [&](int t, int f, float acc) {
        phi_s[t * ldp + f] = feature_grad(acc, phi_s[t * ldp + f], kind, c);
      });
      __syncthreads();
      features_bwd_tail<T, DH, TT>(p, L, smem, wt, dkb, tok_stride, dg_slot, n0);
    }
  }
}

// ---- stand-alone feature map: x (strided [B,H,N,DH] view) -> phi [B*H][N][M] ------------------------
struct FeatArgs {
  const void* x;       // element (b,h,n,0) at x + b*sb + h*sh + n*sn
  void* dx;            // same addressing
  size_t sb, sh, sn;
  float* phi;          // [B*H][N][ldphi] (fwd: out; columns M..ldphi-1 are written as 0)
  const float* dphi;   // [B*H][N][ldphi]
  int ldphi;
  const float* wt;
  int B, N, H, kind, prep;
  float prescale, inv_sqrt_m;
  FeatGeom g;
};

template <typename T, int DH, int TT, bool BWD>
__global__ void __launch_bounds__(320, 1) feat_kernel(const FeatArgs p) {
  constexpr int LDM = DH + 4, NJG = DH / 4 + 1;
  extern __shared__ __align__(16) float smem[];
  const LaSmem L = la_layout(DH, p.g.Mp, p.g.ldp, blockDim.x, TT, BWD, false);
  float* phi_s = smem + L.phi;
  // S region is unused here; layout reuse keeps one sizing function
  const int chunks = (p.N + TT - 1) / TT;
  const int ldp = p.g.ldp, M = p.g.M, kind = p.kind;
  const float c = p.inv_sqrt_m;
  RotArgs ra{ERV_ROT_NONE, nullptr, nullptr};
  for (int unit = blockIdx.x; unit < p.B * p.H * chunks; unit += gridDim.x) {
    const int pair = unit / chunks, n0 = (unit % chunks) * TT;
    const int b = pair / p.H, h = pair % p.H;
    const float* wt = p.wt + (size_t)h * p.g.Mp * LDM;
    const T* src = static_cast<const T*>(p.x) + b * p.sb + h * p.sh;
    load_tile<T, DH, TT>(smem + L.xr, src, p.sn, n0, p.N, 0.f);
    __syncthreads();
    prep_tile<DH, TT>(smem + L.xs, smem + L.xr, nullptr, smem + L.inv, ra, p.prep, p.prescale, n0, p.N);
    __syncthreads();
    feature_tile<DH, TT>(phi_s, smem + L.xs, wt, smem + L.m, smem + L.n2, smem + L.pm, p.g, kind, c);
    if (!BWD) {
      const int ld = p.ldphi;
      float* dst = p.phi + ((size_t)pair * p.N + n0) * ld;
      const int rows = min(TT, p.N - n0);
      for (int i = threadIdx.x; i < rows * ld; i += blockDim.x) {
        int t = i / ld, f = i % ld;
        dst[i] = (f < M) ? phi_s[t * ldp + f] : 0.f;
      }
      __syncthreads();
    } else {
      const float* dphi = p.dphi + ((size_t)pair * p.N + n0) * p.ldphi;
      const int rows = min(TT, p.N - n0);
      for (int i = threadIdx.x; i < TT * M; i += blockDim.x) {
        int t = i / M, f = i % M;
        float d = (t < rows) ? __ldg(dphi + (size_t)t * p.ldphi + f) : 0.f;
        phi_s[t * ldp + f] = feature_grad(d, phi_s[t * ldp + f], kind, c);
      }
      __syncthreads();
      float* red_s = smem + L.red;
      h2_narrow<DH, TT, NJG>(red_s, phi_s, ldp, wt, LDM, M);
      if (kind == ERV_FEAT_FAVOR) {
        const float* xs = smem + L.xs;
        for (int i = threadIdx.x; i < TT * DH; i += blockDim.x) {
          int t = i / DH, d = i % DH;
          red_s[t * LDM + d] -= xs[t * LDM + d] * red_s[t * LDM + DH];
        }
        __syncthreads();
      }
      T* dst = static_cast<T*>(p.dx) + b * p.sb + h * p.sh;
      prep_tile_bwd<T, DH, TT>(red_s, smem + L.xs, smem + L.xr, nullptr, smem + L.inv, smem + L.tmp, ra, p.prep,
                               p.prescale, dst, p.sn, nullptr, n0, p.N);
      __syncthreads();
    }
  }
}

// ---- host-side dispatch -----------------------------------------------------------------------------
static int pick_tt(int DH, const FeatGeom& g, bool bwd, bool circ, size_t* smem_bytes) {
  const int cands[2] = {32, 16};
  for (int i = 0; i < 2; ++i) {
    LaSmem L = la_layout(DH, g.Mp, g.ldp, g.nthreads, cands[i], bwd, circ);
    size_t bytes = (size_t)L.total * sizeof(float);
    if (bytes <= kMaxSmem) {
      *smem_bytes = bytes;
      return cands[i];
    }
  }
  return 0;
}

template <typename K>
static int launch_with_smem(K kernel, int grid, int threads, size_t smem, cudaStream_t st, const void* args_struct,
                            size_t args_size) {
  (void)args_size;
  ERV_CUDA(allow_smem(kernel, smem));
  void* kargs[] = {const_cast<void*>(args_struct)};
  ERV_CUDA(cudaLaunchKernel((const void*)kernel, dim3(grid), dim3(threads), kargs, smem, st));
  count_launch();
  return ERV_OK;
}

#define ERV_DISPATCH_DH_TT(DH_, TT_, FN, ...)                                                  \
  do {                                                                                         \
    if (DH_ == 8 && TT_ == 32) { FN(8, 32, __VA_ARGS__); }                                     \
    else if (DH_ == 8 && TT_ == 16) { FN(8, 16, __VA_ARGS__); }                                \
    else if (DH_ == 16 && TT_ == 32) { FN(16, 32, __VA_ARGS__); }                              \
    else if (DH_ == 16 && TT_ == 16) { FN(16, 16, __VA_ARGS__); }                              \
    else if (DH_ == 32 && TT_ == 32) { FN(32, 32, __VA_ARGS__); }                              \
    else if (DH_ == 32 && TT_ == 16) { FN(32, 16, __VA_ARGS__); }                              \
    else if (DH_ == 64 && TT_ == 32) { FN(64, 32, __VA_ARGS__); }                              \
    else if (DH_ == 64 && TT_ == 16) { FN(64, 16, __VA_ARGS__); }                              \
    else { set_error("unsupported head_dim %d (supported: 8, 16, 32, 64)", DH_); return ERV_E_UNSUPPORTED; } \
  } while (0)

int la_grid(int B, int H) {
  int per_head = (2 * kNumSMs + H - 1) / H;
  if (per_head < 1) per_head = 1;
  if (per_head > B) per_head = B;
  return per_head * H;
}

// gradient slots per head for the Circulant table gradient: one per CTA of the fp32 kernels, two per CTA (one per pair
// side) of the short-sequence tensor-core backward
int la_slots(int B, int H) {
  const int per_head = la_grid(B, H) / H;
  return per_head + (per_head & 1);
}

static int prep_wt(const float* omega, float* wt, int H, int DH, int M, int Mp, int kind, cudaStream_t st) {
  size_t total = (size_t)H * Mp * (DH + 4);
  wt_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(omega, wt, H, DH, M, Mp,
                                                                  kind == ERV_FEAT_FAVOR ? 1.f : 0.f);
  ERV_LAUNCH_CHECK();
  return ERV_OK;
}

size_t wt_bytes(int H, int DH, int M) {
  int Mp = (M + 31) / 32 * 32;
  return align_up((size_t)H * Mp * (DH + 4) * sizeof(float), 256);
}

static int check_common(const char* fn, int B, int N, int H, int DH, int M, int dtype) {
  if (B <= 0 || N <= 0 || H <= 0 || M <= 0) { set_error("%s: non-positive shape", fn); return ERV_E_INVALID; }
  if (!(DH == 8 || DH == 16 || DH == 32 || DH == 64)) {
    set_error("%s: unsupported head_dim %d (supported: 8, 16, 32, 64)", fn, DH);
    return ERV_E_UNSUPPORTED;
  }
  if (M > 1024) { set_error("%s: num_features %d > 1024 unsupported", fn, M); return ERV_E_UNSUPPORTED; }
  if (dtype != ERV_F32 && dtype != ERV_BF16) { set_error("%s: bad dtype %d", fn, dtype); return ERV_E_INVALID; }
  return ERV_OK;
}

// feature-map launcher shared with the KERPLE path (erv_tileattn.cu)
int prep_wt_public(const float* omega, float* wt, int H, int DH, int M, int kind, cudaStream_t st) {
  return prep_wt(omega, wt, H, DH, M, (M + 31) / 32 * 32, kind, st);
}

int launch_feature_map(const void* x, void* dx, size_t sb, size_t sh, size_t sn, float* phi, const float* dphi,
                       int ldphi, const float* wt, int B, int N, int H, int DH, int M, int kind, int prep,
                       float prescale, int dtype, bool bwd, cudaStream_t st) {
  FeatArgs a;
  a.x = x; a.dx = dx; a.sb = sb; a.sh = sh; a.sn = sn; a.phi = phi; a.dphi = dphi; a.ldphi = ldphi; a.wt = wt;
  a.B = B; a.N = N; a.H = H; a.kind = kind; a.prep = prep; a.prescale = prescale;
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.g = make_geom(M);
  size_t smem = 0;
  int TT = pick_tt(DH, a.g, bwd, false, &smem);
  if (!TT) { set_error("feature map: head_dim %d x num_features %d does not fit shared memory", DH, M); return ERV_E_UNSUPPORTED; }
  int chunks = (N + TT - 1) / TT;
  long units = (long)B * H * chunks;
  int grid = (int)(units < 4L * kNumSMs ? units : 4L * kNumSMs);
#define FEAT_LAUNCH(DH_, TT_, BW_)                                                                                   \
  do {                                                                                                               \
    if (dtype == ERV_F32) { int rc = launch_with_smem(feat_kernel<float, DH_, TT_, BW_>, grid, a.g.nthreads, smem, st, &a, sizeof(a)); if (rc) return rc; } \
    else { int rc = launch_with_smem(feat_kernel<__nv_bfloat16, DH_, TT_, BW_>, grid, a.g.nthreads, smem, st, &a, sizeof(a)); if (rc) return rc; }          \
  } while (0)
#define FEAT_FN(DH_, TT_, dummy) do { if (bwd) FEAT_LAUNCH(DH_, TT_, true); else FEAT_LAUNCH(DH_, TT_, false); } while (0)
  ERV_DISPATCH_DH_TT(DH, TT, FEAT_FN, 0);
#undef FEAT_FN
#undef FEAT_LAUNCH
  return ERV_OK;
}

}  // namespace erv

using namespace erv;

extern "C" int erv_circulant_slots(int B, int H) {
  if (B <= 0 || H <= 0) return 1;
  return la_slots(B, H);
}

extern "C" size_t erv_linear_attention_workspace(int B, int N, int H, int head_dim, int M, int rot, int backward) {
  (void)B; (void)N; (void)rot; (void)backward;
  return wt_bytes(H, head_dim, M);
}

namespace erv {  // tensor-core path (erv_linattn_tc.cu)
bool la_tc_eligible(int N, int DH, int M);
bool la_tc2_eligible(int N, int DH, int M);  // two pairs per tile (erv_linattn_tc2.cu)
int la_tc2_forward(const void* qkv, void* out, const float* omega, int B, int N, int H, int M, int kind, int rot,
                   const float* ta, const float* tb, int dtype, float* state, cudaStream_t st);
int tc2_mp(int M);
bool la_pipe_eligible(int N, int DH, int M);  // warp-specialised, pipelined forward (erv_linattn_pipe.cu)
int la_pipe_forward(const void* qkv, void* out, const float* omega, int B, int N, int H, int M, int kind, int rot,
                    const float* ta, const float* tb, int dtype, float* state, cudaStream_t st);
int la_pipe_backward(const void* qkv, const void* out, const void* dout, void* dqkv, const float* omega, int B, int N,
                     int H, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                     int dtype, const float* state, cudaStream_t st);
int la_tc2_backward(const void* qkv, const void* out, const void* dout, void* dqkv, const float* omega, int B, int N,
                    int H, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                    int dtype, const float* state, cudaStream_t st);
int la_tc_forward(const void* qkv, void* out, const float* omega, int B, int N, int H, int DH, int M, int kind, int rot,
                  const float* ta, const float* tb, int dtype, float* state, cudaStream_t st);
int la_tc_backward(const void* qkv, const void* out, const void* dout, void* dqkv, const float* omega, int B, int N,
                   int H, int DH, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part, int slots,
                   int dtype, const float* state, cudaStream_t st);
}  // namespace erv

// Floats of the [S|z] state the forward saves for the backward (0: these shapes recompute S in the backward).
extern "C" size_t erv_linear_attention_state_floats(int B, int N, int H, int head_dim, int M) {
  if (B <= 0 || H <= 0) return 0;
  if (!erv::la_tc2_eligible(N, head_dim, M)) {
    // long sequences (one pair per 128-token tile): [Dh+1][Mp] per pair, Mp = 64 / 128 / 256
    if (!erv::la_tc_eligible(N, head_dim, M)) return 0;
    return (size_t)B * H * (head_dim + 1) * (M <= 64 ? 64 : (M <= 128 ? 128 : 256));
  }
  // the pipelined kernels also keep three per-token statistics (normaliser, exponent shifts of the query / key rows)
  const size_t aux = erv::la_pipe_eligible(N, head_dim, M) ? (size_t)3 * B * N * H : 0;
  return (size_t)B * H * (head_dim + 1) * erv::tc2_mp(M) + aux;
}

static int la_launch(bool bwd, const void* qkv, void* out, const void* dout, void* dqkv, const float* omega, int B,
                     int N, int H, int DH, int M, int kind, int rot, const float* ta, const float* tb, float* dg_part,
                     int dtype, float* state, void* ws, size_t ws_bytes, void* stream) {
  const char* fn = bwd ? "erv_linear_attention_bwd" : "erv_linear_attention_fwd";
  int rc = check_common(fn, B, N, H, DH, M, dtype);
  if (rc) return rc;
  ERV_CHECK_ARG(qkv && out && omega && ws, "%s: null pointer", fn);
  ERV_CHECK_ARG(kind == ERV_FEAT_FAVOR || kind == ERV_FEAT_RELU, "%s: bad kind %d", fn, kind);
  ERV_CHECK_ARG(rot == ERV_ROT_NONE || ta, "%s: rotation table missing", fn);
  ERV_CHECK_ARG(rot != ERV_ROT_ROPE || tb, "%s: rope sin table missing", fn);
  ERV_CHECK_ARG(!bwd || (dout && dqkv), "%s: null gradient pointer", fn);
  ERV_CHECK_ARG(!(bwd && rot == ERV_ROT_CIRCULANT) || dg_part, "%s: dg_part missing", fn);
  if (ws_bytes < wt_bytes(H, DH, M)) { set_error("%s: workspace too small", fn); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  if (!bwd && la_pipe_eligible(N, DH, M))
    return la_pipe_forward(qkv, out, omega, B, N, H, M, kind, rot, ta, tb, dtype, state, st);
  if (!bwd && la_tc2_eligible(N, DH, M))
    return la_tc2_forward(qkv, out, omega, B, N, H, M, kind, rot, ta, tb, dtype, state, st);
  if (!bwd && la_tc_eligible(N, DH, M))
    return la_tc_forward(qkv, out, omega, B, N, H, DH, M, kind, rot, ta, tb, dtype, state, st);
  static const bool tc_bwd_off = getenv("ERV_DISABLE_TC_BWD") != nullptr, tc2_bwd_off = getenv("ERV_DISABLE_TC2_BWD") != nullptr;
  if (bwd && la_tc_eligible(N, DH, M) && !tc_bwd_off) {
    const int slots = la_slots(B, H);
    float* dgp = (rot == ERV_ROT_CIRCULANT) ? dg_part : nullptr;
    if (dgp) ERV_CUDA(cudaMemsetAsync(dgp, 0, (size_t)H * slots * N * DH * sizeof(float), st));
    static const bool pipe_bwd_off = getenv("ERV_DISABLE_PIPE_BWD") != nullptr;
    if (la_pipe_eligible(N, DH, M) && state != nullptr && !pipe_bwd_off)  // needs the statistics the pipelined forward saved
      return la_pipe_backward(qkv, out, dout, dqkv, omega, B, N, H, M, kind, rot, ta, tb, dgp, slots, dtype, state, st);
    if (la_tc2_eligible(N, DH, M) && !tc2_bwd_off)
      return la_tc2_backward(qkv, out, dout, dqkv, omega, B, N, H, M, kind, rot, ta, tb, dgp, slots, dtype, state, st);
    // a state written by the short-sequence forward kernels has their layout: the long-sequence backward then rebuilds S
    return la_tc_backward(qkv, out, dout, dqkv, omega, B, N, H, DH, M, kind, rot, ta, tb, dgp, slots, dtype,
                          la_tc2_eligible(N, DH, M) ? nullptr : state, st);
  }
  LaArgs a;
  a.qkv = qkv; a.out = out; a.dout = dout; a.dqkv = dqkv; a.wt = (const float*)ws; a.ta = ta; a.tb = tb;
  a.dg_part = (rot == ERV_ROT_CIRCULANT) ? dg_part : nullptr;
  a.B = B; a.N = N; a.H = H; a.kind = kind; a.rot = rot;
  a.prescale = (float)pow((double)DH, -0.25);
  a.inv_sqrt_m = (float)(1.0 / sqrt((double)M));
  a.g = make_geom(M);
  rc = prep_wt(omega, (float*)ws, H, DH, M, a.g.Mp, kind, st);
  if (rc) return rc;
  const int grid = la_grid(B, H);
  a.slots = la_slots(B, H);
  if (bwd && a.dg_part) ERV_CUDA(cudaMemsetAsync(dg_part, 0, (size_t)H * a.slots * N * DH * sizeof(float), st));
  size_t smem = 0;
  int TT = pick_tt(DH, a.g, bwd, rot == ERV_ROT_CIRCULANT, &smem);
  if (!TT) { set_error("%s: head_dim %d x num_features %d does not fit shared memory", fn, DH, M); return ERV_E_UNSUPPORTED; }
#define LA_FN(DH_, TT_, dummy)                                                                                       \
  do {                                                                                                               \
    int rc2;                                                                                                         \
    if (dtype == ERV_F32) rc2 = bwd ? launch_with_smem(la_bwd_kernel<float, DH_, TT_>, grid, a.g.nthreads, smem, st, &a, sizeof(a)) \
                                    : launch_with_smem(la_fwd_kernel<float, DH_, TT_>, grid, a.g.nthreads, smem, st, &a, sizeof(a)); \
    else rc2 = bwd ? launch_with_smem(la_bwd_kernel<__nv_bfloat16, DH_, TT_>, grid, a.g.nthreads, smem, st, &a, sizeof(a))          \
                   : launch_with_smem(la_fwd_kernel<__nv_bfloat16, DH_, TT_>, grid, a.g.nthreads, smem, st, &a, sizeof(a));         \
    if (rc2) return rc2;                                                                                             \
  } while (0)
  ERV_DISPATCH_DH_TT(DH, TT, LA_FN, 0);
#undef LA_FN
  return ERV_OK;
}

extern "C" int erv_linear_attention_fwd(const void* qkv, void* out, const float* omega, int B, int N, int H,
                                        int head_dim, int M, int kind, int rot, const float* tab_a, const float* tab_b,
                                        int dtype, float* kv_state, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  return la_launch(false, qkv, out, nullptr, nullptr, omega, B, N, H, head_dim, M, kind, rot, tab_a, tab_b, nullptr,
                   dtype, kv_state, workspace, workspace_bytes, stream);
}

extern "C" int erv_linear_attention_bwd(const void* qkv, const void* out, const void* dout, void* dqkv,
                                        const float* omega, int B, int N, int H, int head_dim, int M, int kind, int rot,
                                        const float* tab_a, const float* tab_b, float* dg_part, int dtype,
                                        const float* kv_state, void* workspace, size_t workspace_bytes, void* stream) {
  return la_launch(true, qkv, const_cast<void*>(out), dout, dqkv, omega, B, N, H, head_dim, M, kind, rot, tab_a, tab_b,
                   dg_part, dtype, const_cast<float*>(kv_state), workspace, workspace_bytes, stream);
}

// x [B,H,N,DH] fp32 contiguous; workspace-free variants allocate nothing: the W^T staging buffer is the
// tail of phi_out/dx?  No: the caller passes it explicitly through erv_feature_map_workspace().
extern "C" size_t erv_feature_map_workspace(int H, int head_dim, int M) { return wt_bytes(H, head_dim, M); }

extern "C" int erv_feature_map_fwd(const float* x, const float* omega, int B, int H, int N, int head_dim, int M,
                                   int kind, float* phi_out, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("erv_feature_map_fwd", B, N, H, head_dim, M, ERV_F32);
  if (rc) return rc;
  ERV_CHECK_ARG(x && omega && phi_out && workspace, "erv_feature_map_fwd: null pointer");
  if (workspace_bytes < wt_bytes(H, head_dim, M)) { set_error("erv_feature_map_fwd: workspace too small"); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  FeatGeom g = make_geom(M);
  rc = prep_wt(omega, (float*)workspace, H, head_dim, M, g.Mp, kind, st);
  if (rc) return rc;
  return launch_feature_map(x, nullptr, (size_t)H * N * head_dim, (size_t)N * head_dim, (size_t)head_dim, phi_out,
                            nullptr, M, (const float*)workspace, B, N, H, head_dim, M, kind, ERV_PREP_NONE, 1.f, ERV_F32,
                            false, st);
}

extern "C" int erv_feature_map_bwd(const float* x, const float* omega, const float* dphi, int B, int H, int N,
                                   int head_dim, int M, int kind, float* dx, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  int rc = check_common("erv_feature_map_bwd", B, N, H, head_dim, M, ERV_F32);
  if (rc) return rc;
  ERV_CHECK_ARG(x && omega && dphi && dx && workspace, "erv_feature_map_bwd: null pointer");
  if (workspace_bytes < wt_bytes(H, head_dim, M)) { set_error("erv_feature_map_bwd: workspace too small"); return ERV_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  FeatGeom g = make_geom(M);
  rc = prep_wt(omega, (float*)workspace, H, head_dim, M, g.Mp, kind, st);
  if (rc) return rc;
  return launch_feature_map(x, dx, (size_t)H * N * head_dim, (size_t)N * head_dim, (size_t)head_dim, nullptr, dphi,
                            M, (const float*)workspace, B, N, H, head_dim, M, kind, ERV_PREP_NONE, 1.f, ERV_F32, true, st);
}
