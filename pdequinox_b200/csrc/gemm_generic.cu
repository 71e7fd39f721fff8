// Generic CUDA-core (FFMA, true fp32) batched GEMMs.
//
// These carry every shape the tcgen05 fast paths do not (1D/3D grids, odd sizes,
// channel counts that are not multiples of the UMMA tile) and are the fp32
// arithmetic reference the tensor-core paths are tested against on the GPU.
// The truncated DFT stages of SpectralConv are real GEMMs against twiddle tables:
//   r2c (last axis):  C[R, 2m]  = X[R, s] . Tt[s, 2m]        (interleaved re/im columns)
//   c2r (last axis):  Y[R, s]   = Z[R, 2m] . Tc[2m, s]
//   c2c (other axes): C[b][k][q] = sum_n T[k][n] * B[b][n][q]  (complex, cgemm_shared_a)
// and PointwiseLinearConv is W[Co,Ci] . x[b][Ci, N].
#include "common.cuh"

namespace pdeqx {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(SgemmArgs a, int tiles_n) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int64_t tm = tile / tiles_n, tn = tile % tiles_n;
  const int b = blockIdx.y;
  const int64_t m0 = tm * BM, n0 = tn * BN;
  const float* __restrict__ A = a.A + (int64_t)b * a.strideA;
  const float* __restrict__ B = a.B + (int64_t)b * a.strideB;
  const int ty = tid / (BN / TN), tx = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const bool k_contig = (a.lda_k == 1);
  for (int64_t k0 = 0; k0 < a.K; k0 += BK) {
    // A tile
    for (int idx = tid; idx < BM * BK; idx += NT) {
      int m, k;
      if (k_contig) { m = idx / BK; k = idx % BK; } else { m = idx % BM; k = idx / BM; }
      int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < a.M && gk < a.K) v = __ldg(A + gm * a.lda_m + gk * a.lda_k);
      As[k][m] = v;
    }
    for (int idx = tid; idx < BK * BN; idx += NT) {
      int k = idx / BN, n = idx % BN;
      int64_t gk = k0 + k, gn = n0 + n;
      float v = 0.f;
      if (gk < a.K && gn < a.N) v = __ldg(B + gk * a.ldb + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float ar[TM], br[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) ar[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) br[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* __restrict__ C = a.C + (int64_t)b * a.strideC;
  const float* __restrict__ add = a.add ? a.add + (int64_t)b * a.strideC : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t gm = m0 + ty * TM + i;
    if (gm >= a.M) continue;
    float bv = a.bias ? __ldg(a.bias + gm) : 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t gn = n0 + tx * TN + j;
      if (gn >= a.N) continue;
      float v = acc[i][j] + bv;
      if (add) v += __ldg(add + gm * a.ldc + gn);
      C[gm * a.ldc + gn] = apply_act_rt(v, a.act);
    }
  }
}

int sgemm_batched(const SgemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return PDEQX_OK;
  PDEQX_CHECK_ARG(a.batch <= 65535, "sgemm: batch %d exceeds grid.y limit", a.batch);
  if (a.N <= 32) {
    constexpr int BM = 128, BN = 32, BK = 16;
    int tiles_n = (int)ceil_div(a.N, BN);
    int64_t tiles = ceil_div(a.M, BM) * tiles_n;
    PDEQX_CHECK_ARG(tiles < (1ll << 31), "sgemm: too many tiles");
    sgemm_kernel<BM, BN, BK, 4, 4><<<dim3((unsigned)tiles, a.batch), 256, 0, s>>>(a, tiles_n);
  } else {
    constexpr int BM = 64, BN = 64, BK = 16;
    int tiles_n = (int)ceil_div(a.N, BN);
    int64_t tiles = ceil_div(a.M, BM) * tiles_n;
    PDEQX_CHECK_ARG(tiles < (1ll << 31), "sgemm: too many tiles");
    sgemm_kernel<BM, BN, BK, 4, 4><<<dim3((unsigned)tiles, a.batch), 256, 0, s>>>(a, tiles_n);
  }
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

// ---------------------------------------------------------------------------
// complex GEMM with shared A:  C[b][m][n] = sum_k A[m][k] * B[b][k][n]
// ---------------------------------------------------------------------------
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(128)
cgemm_shared_a_kernel(const float2* __restrict__ A, const float2* __restrict__ B, float2* __restrict__ C,
                      int M, int K, int64_t N, int tiles_n) {
  static_assert(BM == 32 && BN == 16, "thread map assumes 32x16 tile, 2x2 per thread");
  __shared__ float2 As[BK][BM + 1];
  __shared__ float2 Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  const int tm = blockIdx.y / tiles_n, tn = blockIdx.y % tiles_n;
  const int m0 = tm * BM;
  const int64_t n0 = (int64_t)tn * BN;
  const float2* Bb = B + b * (int64_t)K * N;
  const int ty = tid / 8, tx = tid % 8;  // rows ty*2.., cols tx*2..
  float2 acc[2][2] = {{{0.f, 0.f}, {0.f, 0.f}}, {{0.f, 0.f}, {0.f, 0.f}}};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int idx = tid; idx < BM * BK; idx += 128) {
      int m = idx / BK, k = idx % BK;
      float2 v = make_float2(0.f, 0.f);
      if (m0 + m < M && k0 + k < K) v = __ldg(A + (int64_t)(m0 + m) * K + k0 + k);
      As[k][m] = v;
    }
    for (int idx = tid; idx < BK * BN; idx += 128) {
      int k = idx / BN, n = idx % BN;
      float2 v = make_float2(0.f, 0.f);
      if (k0 + k < K && n0 + n < N) v = __ldg(Bb + (int64_t)(k0 + k) * N + n0 + n);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float2 a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
      float2 b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
      acc[0][0].x = fmaf(a0.x, b0.x, fmaf(-a0.y, b0.y, acc[0][0].x));
      acc[0][0].y = fmaf(a0.x, b0.y, fmaf(a0.y, b0.x, acc[0][0].y));
      acc[0][1].x = fmaf(a0.x, b1.x, fmaf(-a0.y, b1.y, acc[0][1].x));
      acc[0][1].y = fmaf(a0.x, b1.y, fmaf(a0.y, b1.x, acc[0][1].y));
      acc[1][0].x = fmaf(a1.x, b0.x, fmaf(-a1.y, b0.y, acc[1][0].x));
      acc[1][0].y = fmaf(a1.x, b0.y, fmaf(a1.y, b0.x, acc[1][0].y));
      acc[1][1].x = fmaf(a1.x, b1.x, fmaf(-a1.y, b1.y, acc[1][1].x));
      acc[1][1].y = fmaf(a1.x, b1.y, fmaf(a1.y, b1.x, acc[1][1].y));
    }
    __syncthreads();
  }
  float2* Cb = C + b * (int64_t)M * N;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + ty * 2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int64_t n = n0 + tx * 2 + j;
      if (n < N) Cb[(int64_t)m * N + n] = acc[i][j];
    }
  }
}

int cgemm_shared_a(const float2* A, const float2* B, float2* C, int64_t batch, int M, int K, int64_t N,
                   cudaStream_t s) {
  if (batch <= 0 || M <= 0 || N <= 0) return PDEQX_OK;
  constexpr int BM = 32, BN = 16, BK = 16;
  int tiles_n = (int)ceil_div(N, BN);
  int64_t tiles = ceil_div(M, BM) * tiles_n;
  PDEQX_CHECK_ARG(tiles <= 65535, "cgemm: too many tiles per batch entry (%lld)", (long long)tiles);
  PDEQX_CHECK_ARG(batch < (1ll << 31), "cgemm: batch too large");
  cgemm_shared_a_kernel<BM, BN, BK><<<dim3((unsigned)batch, (unsigned)tiles), 128, 0, s>>>(A, B, C, M, K, N, tiles_n);
  PDEQX_LAUNCHED();
  return PDEQX_OK;
}

}  // namespace pdeqx
