// K2/K4 in the fp32 parity mode (NERF_MODE_FP32): the coarse/fine NeRF MLPs of src/NeRF.py:248-340 evaluated by
// model_predict (src/UtilsNeuralRadianceField.py:214-234) and the gradients TF autodiff produces for them.
// Plain SIMT fp32 GEMMs (FFMA, fp32 accumulate) so the result matches an fp32 CPU evaluation to ~1e-6; the
// tensor-core BF16 path lives in mlp_tc.cu.  Layer loop and concat handling are on the host side of the C ABI.
#include "common.cuh"

namespace nerf {

// provided by mlp_tc.cu
int mlp_tc_fwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* params, const void* packed, const float* xyz_enc,
               const float* view_enc, int64_t m, float* out4, void* saved, void* workspace, cudaStream_t st, bool half);
int mlp_tc_bwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* params, const void* packed, const float* xyz_enc,
               const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads, float* d_xyz_enc,
               void* workspace, cudaStream_t st, int parts, cudaStream_t side, bool half, const BwdRays* rays = nullptr);
int mlp_tc_fwd_rays(const nerf_net_cfg* cfg, const NetGeom& g, const void* packed, const float* origs4, const float* dirs4,
                    const float* z, int64_t n_rays, int n_samples, float* out4, void* saved, cudaStream_t st, bool half,
                    const StratifiedZ* gen = nullptr, const CameraRays* cam = nullptr);
int64_t mlp_tc_saved_bytes(const NetGeom& g, int64_t m);
int64_t mlp_tc_workspace_bytes(const NetGeom& g, int64_t m, int backward);

constexpr int BM = 128, BN = 128, BK = 8;

enum : int { EPI_BIAS = 1, EPI_LEAKY = 2, EPI_ACCUM = 4, EPI_MASK = 8 };

// C[M,N] = epi(A[M,K] * B),  B = Bsrc[K,N] (TRANSB=false) or Bsrc[N,K]^T (TRANSB=true).
// epi: v = acc (+ C_old if ACCUM) (+ bias[n] if BIAS); LEAKY: v = v>0 ? v : alpha v;
//      MASK: v *= (mask[m,n] > 0 ? 1 : alpha)   (LeakyReLU'(saved activation); derivative at 0 is alpha)
template <bool TRANSB>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bsrc,
                                                   int ldb, float* __restrict__ C, int ldc, int64_t M, int N, int K,
                                                   const float* __restrict__ bias, const float* __restrict__ mask,
                                                   int ldmask, float alpha, int flags) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 1, a_k = (tid & 1) * 4;
  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      int64_t gm = m0 + a_row;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        int gk = k0 + a_k + kk;
        As[a_k + kk][a_row] = (gm < M && gk < K) ? __ldg(A + gm * lda + gk) : 0.f;
      }
    }
    if (TRANSB) {
      int gn = n0 + a_row;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        int gk = k0 + a_k + kk;
        Bs[a_k + kk][a_row] = (gn < N && gk < K) ? __ldg(Bsrc + (int64_t)gn * ldb + gk) : 0.f;
      }
    } else {
      int b_k = tid >> 5, b_n = (tid & 31) * 4;
      int gk = k0 + b_k;
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) {
        int gn = n0 + b_n + nn;
        Bs[b_k][b_n + nn] = (gk < K && gn < N) ? __ldg(Bsrc + (int64_t)gk * ldb + gn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (flags & EPI_ACCUM) v += C[gm * ldc + gn];
      if (flags & EPI_BIAS) v += __ldg(bias + gn);
      if (flags & EPI_LEAKY) v = v > 0.f ? v : alpha * v;
      if (flags & EPI_MASK) v *= (__ldg(mask + gm * ldmask + gn) > 0.f ? 1.f : alpha);
      C[gm * ldc + gn] = v;
    }
  }
}

// dW[K,N] += A[M,K]^T * B[M,N]  (split over M, fp32 atomics);  db[n] += sum_m B[m,n] when db != null.
constexpr int TK = 64, TN = 64, TM = 16;
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                      int ldb, float* __restrict__ dW, int ldw, float* __restrict__ db,
                                                      int64_t M, int N, int K, int64_t rows_per_split) {
  __shared__ float As[TM][TK + 1];
  __shared__ float Bs[TM][TN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * TK, n0 = blockIdx.y * TN;
  const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t m_end = min(M, m_begin + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool do_bias = (db != nullptr) && blockIdx.x == 0 && tid < TN;
  for (int64_t mm = m_begin; mm < m_end; mm += TM) {
#pragma unroll
    for (int r = 0; r < (TM * TK) / 256; ++r) {
      int e = tid + r * 256;
      int row = e / TK, col = e % TK;
      int64_t gm = mm + row;
      As[row][col] = (gm < m_end && k0 + col < K) ? __ldg(A + gm * lda + k0 + col) : 0.f;
      Bs[row][col] = (gm < m_end && n0 + col < N) ? __ldg(B + gm * ldb + n0 + col) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[r][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[r][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (do_bias) {
#pragma unroll
      for (int r = 0; r < TM; ++r) bsum += Bs[r][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gk = k0 + ty + 16 * i;
    if (gk >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx + 16 * j;
      if (gn < N) atomicAdd(dW + (int64_t)gk * ldw + gn, acc[i][j]);
    }
  }
  if (do_bias && n0 + tid < N) atomicAdd(db + n0 + tid, bsum);
}

static int gemm(cudaStream_t st, bool transb, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                int64_t M, int N, int K, const float* bias, const float* mask, int ldmask, float alpha, int flags) {
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN));
  if (transb)
    gemm_kernel<true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask, ldmask, alpha, flags);
  else
    gemm_kernel<false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, mask, ldmask, alpha, flags);
  return 0;
}

static int gemm_tn(cudaStream_t st, const float* A, int lda, const float* B, int ldb, float* dW, int ldw, float* db,
                   int64_t M, int N, int K) {
  int kt = (int)ceil_div(K, TK), nt = (int)ceil_div(N, TN);
  int64_t target_splits = max((int64_t)1, (int64_t)(num_sms() * 4) / (kt * nt));
  int64_t rows = ceil_div(ceil_div(M, target_splits), TM) * TM;
  if (rows < 256) rows = 256;
  int64_t splits = ceil_div(M, rows);
  dim3 grid(kt, nt, (unsigned)splits);
  gemm_tn_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, dW, ldw, db, M, N, K, rows);
  return 0;
}

// ---- fp32 forward / backward layer loops ----------------------------------------------------------------------
// saved layout (floats): H1..H8 each [M,H], then HL [M,HL] (view variant) or H9 [M,H], HL [M,HL] (xyz-only).
static int64_t fp32_saved_floats(const NetGeom& g, int64_t m) {
  return m * ((int64_t)g.hidden * (g.view ? 8 : 9) + g.last_hidden);
}

static int fp32_fwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* P, const float* xyz, const float* view,
                    int64_t M, float* out4, float* saved, float* ws, cudaStream_t st) {
  const int H = g.hidden, HL = g.last_hidden, dx = g.dx, dv = g.dv;
  const float al = cfg->leaky_alpha;
  // activation buffers: saved (training) or 3 rotating workspace buffers + HL (+H9)
  float* Hbuf[10];
  if (saved) {
    for (int i = 0; i < 9; ++i) Hbuf[i] = saved + (int64_t)i * M * H;  // Hbuf[i] = H_{i+1}; Hbuf[8] = HL or H9
    if (!g.view) Hbuf[9] = saved + (int64_t)9 * M * H;                 // HL for xyz-only
  } else {
    float* r[3] = {ws, ws + M * H, ws + 2 * M * H};
    // H1..H4 rotate r0,r1; H4 must survive until layer 5 is done; H8 must survive the heads.
    Hbuf[0] = r[0]; Hbuf[1] = r[1]; Hbuf[2] = r[0]; Hbuf[3] = r[1];
    Hbuf[4] = r[2]; Hbuf[5] = r[0]; Hbuf[6] = r[1]; Hbuf[7] = r[2];
    Hbuf[8] = r[0]; Hbuf[9] = r[1];
  }
  auto W = [&](int l) { return P + g.layers[l].w_off; };
  auto Bv = [&](int l) { return P + g.layers[l].b_off; };
  gemm(st, false, xyz, dx, W(0), H, Hbuf[0], H, M, H, dx, Bv(0), nullptr, 0, al, EPI_BIAS | EPI_LEAKY);
  for (int l = 1; l <= 3; ++l)
    gemm(st, false, Hbuf[l - 1], H, W(l), H, Hbuf[l], H, M, H, H, Bv(l), nullptr, 0, al, EPI_BIAS | EPI_LEAKY);
  // layer 4: concat [xyz ; H4]  (src/NeRF.py:322): kernel rows 0..dx-1 belong to xyz
  gemm(st, false, xyz, dx, W(4), H, Hbuf[4], H, M, H, dx, nullptr, nullptr, 0, al, 0);
  gemm(st, false, Hbuf[3], H, W(4) + (int64_t)dx * H, H, Hbuf[4], H, M, H, H, Bv(4), nullptr, 0, al,
       EPI_ACCUM | EPI_BIAS | EPI_LEAKY);
  for (int l = 5; l <= 7; ++l)
    gemm(st, false, Hbuf[l - 1], H, W(l), H, Hbuf[l], H, M, H, H, Bv(l), nullptr, 0, al, EPI_BIAS | EPI_LEAKY);
  const float* H8 = Hbuf[7];
  if (g.view) {
    float* HLb = Hbuf[8];
    // layer 8: concat [H8 ; view] (src/NeRF.py:329-330): rows 0..H-1 trunk, H..H+dv-1 view
    gemm(st, false, view, dv, W(8) + (int64_t)H * HL, HL, HLb, HL, M, HL, dv, nullptr, nullptr, 0, al, 0);
    gemm(st, false, H8, H, W(8), HL, HLb, HL, M, HL, H, Bv(8), nullptr, 0, al, EPI_ACCUM | EPI_BIAS | EPI_LEAKY);
    gemm(st, false, HLb, HL, W(9), 3, out4, 4, M, 3, HL, Bv(9), nullptr, 0, al, EPI_BIAS);
    gemm(st, false, view, dv, W(10) + (int64_t)H, 1, out4 + 3, 4, M, 1, dv, nullptr, nullptr, 0, al, 0);
    gemm(st, false, H8, H, W(10), 1, out4 + 3, 4, M, 1, H, Bv(10), nullptr, 0, al, EPI_ACCUM | EPI_BIAS);
  } else {
    float* H9 = Hbuf[8];
    float* HLb = Hbuf[9];
    gemm(st, false, H8, H, W(8), H, H9, H, M, H, H, Bv(8), nullptr, 0, al, EPI_BIAS | EPI_LEAKY);
    gemm(st, false, H9, H, W(9), HL, HLb, HL, M, HL, H, Bv(9), nullptr, 0, al, EPI_BIAS | EPI_LEAKY);
    gemm(st, false, HLb, HL, W(10), 3, out4, 4, M, 3, HL, Bv(10), nullptr, 0, al, EPI_BIAS);
    gemm(st, false, H8, H, W(11), 1, out4 + 3, 4, M, 1, H, Bv(11), nullptr, 0, al, EPI_BIAS);
  }
  return 0;
}

static int fp32_bwd(const nerf_net_cfg* cfg, const NetGeom& g, const float* P, const float* xyz, const float* view,
                    const float* saved, const float* d_out4, int64_t M, float* G, float* d_xyz, float* ws,
                    cudaStream_t st) {
  const int H = g.hidden, HL = g.last_hidden, dx = g.dx, dv = g.dv;
  const float al = cfg->leaky_alpha;
  auto Hs = [&](int i) { return saved + (int64_t)(i - 1) * M * H; };  // H_i, i = 1..9
  auto W = [&](int l) { return P + g.layers[l].w_off; };
  auto GW = [&](int l) { return G + g.layers[l].w_off; };
  auto GB = [&](int l) { return G + g.layers[l].b_off; };
  float* dA = ws;              // [M,H]
  float* dB = ws + M * H;      // [M,H]
  float* dL = ws + 2 * M * H;  // [M,HL]
  const float* H8 = Hs(8);
  float* dZ8 = dA;
  if (g.view) {
    const float* HLs = saved + (int64_t)8 * M * H;
    // sigma head (layer 10): input [H8 ; view], upstream d_out4[:,3]
    gemm_tn(st, H8, H, d_out4 + 3, 4, GW(10), 1, GB(10), M, 1, H);
    gemm_tn(st, view, dv, d_out4 + 3, 4, GW(10) + H, 1, nullptr, M, 1, dv);
    // rgb head (layer 9)
    gemm_tn(st, HLs, HL, d_out4, 4, GW(9), 3, GB(9), M, 3, HL);
    gemm(st, true, d_out4, 4, W(9), 3, dL, HL, M, HL, 3, nullptr, HLs, HL, al, EPI_MASK);  // dZ_L
    // last hidden (layer 8): input [H8 ; view]
    gemm_tn(st, H8, H, dL, HL, GW(8), HL, GB(8), M, HL, H);
    gemm_tn(st, view, dv, dL, HL, GW(8) + (int64_t)H * HL, HL, nullptr, M, HL, dv);
    // dH8 = dZ_L W8[0:H]^T + dsigma W10[0:H]^T, then LeakyReLU'(H8)
    gemm(st, true, dL, HL, W(8), HL, dZ8, H, M, H, HL, nullptr, nullptr, 0, al, 0);
    gemm(st, true, d_out4 + 3, 4, W(10), 1, dZ8, H, M, H, 1, nullptr, H8, H, al, EPI_ACCUM | EPI_MASK);
  } else {
    const float* H9 = Hs(9);
    const float* HLs = saved + (int64_t)9 * M * H;
    gemm_tn(st, H8, H, d_out4 + 3, 4, GW(11), 1, GB(11), M, 1, H);                          // sigma head
    gemm_tn(st, HLs, HL, d_out4, 4, GW(10), 3, GB(10), M, 3, HL);                           // rgb head
    gemm(st, true, d_out4, 4, W(10), 3, dL, HL, M, HL, 3, nullptr, HLs, HL, al, EPI_MASK);  // dZ_L
    gemm_tn(st, H9, H, dL, HL, GW(9), HL, GB(9), M, HL, H);
    gemm(st, true, dL, HL, W(9), HL, dB, H, M, H, HL, nullptr, H9, H, al, EPI_MASK);        // dZ9
    gemm_tn(st, H8, H, dB, H, GW(8), H, GB(8), M, H, H);
    gemm(st, true, dB, H, W(8), H, dZ8, H, M, H, H, nullptr, nullptr, 0, al, 0);
    gemm(st, true, d_out4 + 3, 4, W(11), 1, dZ8, H, M, H, 1, nullptr, H8, H, al, EPI_ACCUM | EPI_MASK);
  }
  // trunk layers 7..5:   dZ_{l+1} is in `cur`;  dW_l = H_l^T dZ_{l+1};  dZ_l = (dZ_{l+1} W_l^T) * LeakyReLU'(H_l)
  float* cur = dZ8;
  float* nxt = dB;
  for (int l = 7; l >= 5; --l) {
    gemm_tn(st, Hs(l), H, cur, H, GW(l), H, GB(l), M, H, H);
    gemm(st, true, cur, H, W(l), H, nxt, H, M, H, H, nullptr, Hs(l), H, al, EPI_MASK);
    float* t = cur; cur = nxt; nxt = t;
  }
  // layer 4 (concat [xyz ; H4]); cur = dZ5
  gemm_tn(st, xyz, dx, cur, H, GW(4), H, GB(4), M, H, dx);
  gemm_tn(st, Hs(4), H, cur, H, GW(4) + (int64_t)dx * H, H, nullptr, M, H, H);
  if (d_xyz) gemm(st, true, cur, H, W(4), H, d_xyz, dx, M, dx, H, nullptr, nullptr, 0, al, 0);
  gemm(st, true, cur, H, W(4) + (int64_t)dx * H, H, nxt, H, M, H, H, nullptr, Hs(4), H, al, EPI_MASK);
  { float* t = cur; cur = nxt; nxt = t; }
  for (int l = 3; l >= 1; --l) {
    gemm_tn(st, Hs(l), H, cur, H, GW(l), H, GB(l), M, H, H);
    gemm(st, true, cur, H, W(l), H, nxt, H, M, H, H, nullptr, Hs(l), H, al, EPI_MASK);
    float* t = cur; cur = nxt; nxt = t;
  }
  // layer 0; cur = dZ1
  gemm_tn(st, xyz, dx, cur, H, GW(0), H, GB(0), M, H, dx);
  if (d_xyz) gemm(st, true, cur, H, W(0), H, d_xyz, dx, M, dx, H, nullptr, nullptr, 0, al, EPI_ACCUM);
  return 0;
}

}  // namespace nerf

using namespace nerf;

extern "C" {

int64_t nerf_param_count(const nerf_net_cfg* cfg) {
  NetGeom g;
  if (!make_geom(cfg, &g)) { set_error("nerf_param_count: bad net config"); return NERF_E_ARG; }
  return g.n_params;
}

int32_t nerf_xyz_enc_dim(const nerf_net_cfg* cfg) {
  NetGeom g;
  if (!make_geom(cfg, &g)) { set_error("nerf_xyz_enc_dim: bad net config"); return NERF_E_ARG; }
  return g.dx;
}

int32_t nerf_view_enc_dim(const nerf_net_cfg* cfg) {
  NetGeom g;
  if (!make_geom(cfg, &g)) { set_error("nerf_view_enc_dim: bad net config"); return NERF_E_ARG; }
  return g.dv;
}

int64_t nerf_mlp_saved_bytes(const nerf_net_cfg* cfg, int64_t m, int32_t mode) {
  NetGeom g;
  if (!make_geom(cfg, &g) || m < 0) { set_error("nerf_mlp_saved_bytes: bad argument"); return NERF_E_ARG; }
  if (mode == NERF_MODE_BF16 || mode == NERF_MODE_FP16) return mlp_tc_saved_bytes(g, m);
  return fp32_saved_floats(g, m) * (int64_t)sizeof(float);
}

int64_t nerf_mlp_workspace_bytes(const nerf_net_cfg* cfg, int64_t m, int32_t mode, int32_t backward) {
  NetGeom g;
  if (!make_geom(cfg, &g) || m < 0) { set_error("nerf_mlp_workspace_bytes: bad argument"); return NERF_E_ARG; }
  if (mode == NERF_MODE_BF16 || mode == NERF_MODE_FP16) return mlp_tc_workspace_bytes(g, m, backward);
  int64_t fl = backward ? m * (2 * (int64_t)g.hidden + g.last_hidden) : m * 3 * (int64_t)g.hidden;
  return fl * (int64_t)sizeof(float) + 256;
}

int nerf_mlp_fwd(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null, const float* xyz_enc,
                 const float* view_enc, int64_t m, float* out4, void* saved_or_null, void* workspace, int32_t mode,
                 void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(params && xyz_enc && out4 && (view_enc || !g.view), "null pointer");
  NERF_CHECK_ARG(m >= 0, "negative row count");
  NERF_CHECK_ARG(mode == NERF_MODE_FP32 || mode == NERF_MODE_BF16 || mode == NERF_MODE_FP16, "unknown mode");
  if (m == 0) return NERF_OK;
  if (mode != NERF_MODE_FP32) {
    NERF_CHECK_ARG(packed_or_null, "the tensor-core modes need the packed weights (nerf_pack_weights[_fp16])");
    return mlp_tc_fwd(cfg, g, params, packed_or_null, xyz_enc, view_enc, m, out4, saved_or_null, workspace,
                      (cudaStream_t)stream, mode == NERF_MODE_FP16);
  }
  NERF_CHECK_ARG(saved_or_null || workspace, "inference needs a workspace (nerf_mlp_workspace_bytes)");
  fp32_fwd(cfg, g, params, xyz_enc, view_enc, m, out4, (float*)saved_or_null, (float*)workspace, (cudaStream_t)stream);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_mlp_fwd_rays(const nerf_net_cfg* cfg, const void* packed, const float* origs4, const float* dirs4, const float* z,
                      int64_t n_rays, int32_t n_samples, float* out4, void* saved_or_null, int32_t mode, void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(packed && origs4 && dirs4 && z && out4, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0, "bad shape");
  if (mode != NERF_MODE_BF16 && mode != NERF_MODE_FP16) {
    set_error("nerf_mlp_fwd_rays: only the tensor-core modes fuse the encodings into the MLP kernel "
              "(fp32 mode: nerf_encode_samples + nerf_mlp_fwd)");
    return NERF_E_UNSUPPORTED;
  }
  if (n_rays == 0) return NERF_OK;
  return mlp_tc_fwd_rays(cfg, g, packed, origs4, dirs4, z, n_rays, n_samples, out4, saved_or_null, (cudaStream_t)stream,
                         mode == NERF_MODE_FP16);
}

int nerf_mlp_fwd_rays_stratified(const nerf_net_cfg* cfg, const void* packed, const float* origs4, const float* dirs4,
                                 float z_start, float z_end, uint64_t seed, uint32_t step, uint64_t ray_offset,
                                 int64_t n_rays, int32_t n_samples, float* z_out, float* out4, void* saved_or_null,
                                 int32_t mode, void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(packed && origs4 && dirs4 && z_out && out4, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0, "bad shape");
  if (mode != NERF_MODE_BF16 && mode != NERF_MODE_FP16) {
    set_error("nerf_mlp_fwd_rays_stratified: only the tensor-core modes fuse the sampling into the MLP kernel "
              "(fp32 mode: nerf_stratified_z + nerf_encode_samples + nerf_mlp_fwd)");
    return NERF_E_UNSUPPORTED;
  }
  if (n_rays == 0) return NERF_OK;
  StratifiedZ gen = {z_start, z_end, seed, ray_offset, step, z_out};
  return mlp_tc_fwd_rays(cfg, g, packed, origs4, dirs4, nullptr, n_rays, n_samples, out4, saved_or_null, (cudaStream_t)stream,
                         mode == NERF_MODE_FP16, &gen);
}

int nerf_mlp_fwd_camera(const nerf_net_cfg* cfg, const void* packed, const float* c2w_host, float fov, int32_t h, int32_t w,
                        int64_t ray_begin, int64_t n_rays, int32_t n_samples, const float* z_or_null, float z_start, float z_end,
                        uint64_t seed, uint32_t step, float* z_out_or_null, float* out4, int32_t mode, void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(packed && c2w_host && out4 && (z_or_null || z_out_or_null), "null pointer");
  NERF_CHECK_ARG(h > 0 && w > 0 && ray_begin >= 0 && n_rays >= 0 && ray_begin + n_rays <= (int64_t)h * w && n_samples > 0,
                 "ray range outside the image");
  if (mode != NERF_MODE_BF16 && mode != NERF_MODE_FP16) {
    set_error("nerf_mlp_fwd_camera: only the tensor-core modes generate their rays in the MLP kernel "
              "(fp32 mode: nerf_ray_directions + nerf_encode_samples + nerf_mlp_fwd)");
    return NERF_E_UNSUPPORTED;
  }
  if (n_rays == 0) return NERF_OK;
  CameraRays cam;
  for (int i = 0; i < 16; ++i) cam.c2w[i] = c2w_host[i];
  cam.tan_half_fov = tanf(fov / 2.0f);
  cam.h = h; cam.w = w; cam.ray_begin = ray_begin;
  // depths: given, or drawn in the prologue from the Philox stream of pixel ray_begin + i (nerf_mlp_fwd_rays_stratified)
  StratifiedZ gen = {z_start, z_end, seed, (uint64_t)ray_begin, step, z_out_or_null};
  return mlp_tc_fwd_rays(cfg, g, packed, nullptr, nullptr, z_or_null, n_rays, n_samples, out4, nullptr, (cudaStream_t)stream,
                         mode == NERF_MODE_FP16, z_or_null ? nullptr : &gen, &cam);
}

// parts: bit 0 = input-gradient chain (dZ of every layer, d_xyz_enc), bit 1 = weight gradients from the saved activations
// and the dZ workspace.  The fp32 path computes both in one pass (parts must be 3).
static int mlp_bwd_parts(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null, const float* xyz_enc,
                         const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads,
                         float* d_xyz_enc_or_null, void* workspace, int32_t mode, void* stream, int parts,
                         void* side_stream = nullptr) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(params && saved && d_out4 && grads && workspace, "null pointer");
  // the bf16 path reads the bf16 input panel it saved in the forward pass; only the fp32 path needs the encodings again
  const bool tc = mode == NERF_MODE_BF16 || mode == NERF_MODE_FP16;
  NERF_CHECK_ARG(tc || (xyz_enc && (view_enc || !g.view)), "null pointer");
  NERF_CHECK_ARG(m >= 0, "negative row count");
  NERF_CHECK_ARG(mode == NERF_MODE_FP32 || tc, "unknown mode");
  NERF_CHECK_ARG(tc || parts == 3, "NERF_MODE_FP32 computes both halves of the backward in one pass");
  if (m == 0) return NERF_OK;
  if (tc) {
    NERF_CHECK_ARG(packed_or_null, "the tensor-core modes need the packed weights (nerf_pack_weights[_fp16])");
    return mlp_tc_bwd(cfg, g, params, packed_or_null, xyz_enc, view_enc, saved, d_out4, m, grads, d_xyz_enc_or_null,
                      workspace, (cudaStream_t)stream, parts, (cudaStream_t)side_stream, mode == NERF_MODE_FP16);
  }
  fp32_bwd(cfg, g, params, xyz_enc, view_enc, (const float*)saved, d_out4, m, grads, d_xyz_enc_or_null,
           (float*)workspace, (cudaStream_t)stream);
  NERF_CHECK_LAUNCH();
  return NERF_OK;
}

int nerf_mlp_bwd(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null, const float* xyz_enc,
                 const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads,
                 float* d_xyz_enc_or_null, void* workspace, int32_t mode, void* stream) {
  return mlp_bwd_parts(cfg, params, packed_or_null, xyz_enc, view_enc, saved, d_out4, m, grads, d_xyz_enc_or_null, workspace,
                       mode, stream, 3);
}

int nerf_mlp_bwd_overlapped(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null, const float* xyz_enc,
                            const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads,
                            float* d_xyz_enc_or_null, void* workspace, int32_t mode, void* side_stream, void* stream) {
  return mlp_bwd_parts(cfg, params, packed_or_null, xyz_enc, view_enc, saved, d_out4, m, grads, d_xyz_enc_or_null, workspace,
                       mode, stream, 3, mode != NERF_MODE_FP32 ? side_stream : nullptr);
}

int nerf_mlp_bwd_dx(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null, const float* xyz_enc,
                    const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads,
                    float* d_xyz_enc_or_null, void* workspace, int32_t mode, void* stream) {
  return mlp_bwd_parts(cfg, params, packed_or_null, xyz_enc, view_enc, saved, d_out4, m, grads, d_xyz_enc_or_null, workspace,
                       mode, stream, 1);
}

int nerf_mlp_bwd_rays(const nerf_net_cfg* cfg, const void* packed, const void* saved, const float* d_out4, const float* origs4,
                      const float* dirs4, const float* z, int64_t n_rays, int32_t n_samples, float* grads, float* d_z,
                      int32_t accumulate_d_z, void* workspace, int32_t mode, int32_t parts, void* side_stream_or_null,
                      void* stream) {
  NetGeom g;
  NERF_CHECK_ARG(make_geom(cfg, &g), "bad net config");
  NERF_CHECK_ARG(packed && saved && d_out4 && grads && workspace && origs4 && dirs4 && z && d_z, "null pointer");
  NERF_CHECK_ARG(n_rays >= 0 && n_samples > 0 && parts >= 1 && parts <= 3, "bad shape / parts");
  if (mode != NERF_MODE_BF16 && mode != NERF_MODE_FP16) {
    set_error("nerf_mlp_bwd_rays: tensor-core modes only (NERF_MODE_FP32: nerf_mlp_bwd + nerf_encode_samples_bwd_z)");
    return NERF_E_UNSUPPORTED;
  }
  if (n_rays == 0) return NERF_OK;
  BwdRays rays{(const float4*)origs4, (const float4*)dirs4, z, d_z, n_samples, accumulate_d_z};
  return mlp_tc_bwd(cfg, g, nullptr, packed, nullptr, nullptr, saved, d_out4, n_rays * n_samples, grads, nullptr, workspace,
                    (cudaStream_t)stream, parts, (cudaStream_t)side_stream_or_null, mode == NERF_MODE_FP16, &rays);
}

int nerf_mlp_bwd_dw(const nerf_net_cfg* cfg, const float* params, const void* packed_or_null, const float* xyz_enc,
                    const float* view_enc, const void* saved, const float* d_out4, int64_t m, float* grads,
                    float* d_xyz_enc_or_null, void* workspace, int32_t mode, void* stream) {
  return mlp_bwd_parts(cfg, params, packed_or_null, xyz_enc, view_enc, saved, d_out4, m, grads, d_xyz_enc_or_null, workspace,
                       mode, stream, 2);
}

}  // extern "C"
