// Token-row Linear layers of the structure-prior fusion: the K / V / first-Q projections of MHCA_ein
// (sgtapose/lib/model/networks/dla.py:868-876, bias-free nn.Linear) and the two Linear layers of
// `cat_layer[i]` applied to cat([out, cur_query], -1) (dla.py:1499-1502, :1006-1018):
//     y[m, n] = act( sum_k x[m, k] * w[n, k] + bias[n] ),   x = [x1 | x2] along k (the concat is never materialised)
// Rows are tokens (B * n: 37,856 at level 0 down to 224 at levels 3-5), K = 16 ... 2048, N = 16 ... 2048: under 2 %
// of the step's FLOPs, fp32 FMA arithmetic like the library SGEMM the reference runs (TF32 off for parity), summed in
// ascending k.  One CTA computes a BM x BN tile through a [BK][BM] / [BK][BN] shared-memory pair (both operands are
// K-contiguous in memory, so tiles are loaded as float4 along k and stored k-major); each thread owns a 4 x 4 block.
// The tile shrinks to 32 x 32 when 64 x 64 tiles would not give every SM a CTA (levels 3-5: 224 rows).
#include "common.cuh"

namespace sgta {

struct TokenLinearP {
  const float *x1, *x2, *w, *bias;
  float* y;
  int M, N, K1, K2, relu;
  int hm_heads, hm_n;      // hm_heads > 0: rows are (b, j) with j < hm_n, columns (h, d): store y as [B, heads, hm_n, N / heads]
};

template <int BM, int BN>
__global__ void __launch_bounds__((BM / 4) * (BN / 4)) token_linear_kernel(const TokenLinearP p) {
  constexpr int BK = 16, NT = (BM / 4) * (BN / 4), PAD = 4;
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % (BN / 4), ty = tid / (BN / 4);
  const int K = p.K1 + p.K2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // global loads of K block i+1 are issued before the FMAs of K block i (register prefetch)
  constexpr int LA = BM * (BK / 4) / NT, LB = BN * (BK / 4) / NT;
  float4 ra[LA], rb[LB];
  auto fetch = [&](int k0) {
    // K1 and K2 are multiples of BK, so a K block lies entirely in x1 or in x2
    const float* xs = k0 < p.K1 ? p.x1 + k0 : p.x2 + (k0 - p.K1);
    const int ldx = k0 < p.K1 ? p.K1 : p.K2;
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int i = tid + u * NT, r = i / (BK / 4), kq = i % (BK / 4);
      ra[u] = m0 + r < p.M ? __ldg(reinterpret_cast<const float4*>(xs + (size_t)(m0 + r) * ldx) + kq)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int i = tid + u * NT, r = i / (BK / 4), kq = i % (BK / 4);
      rb[u] = n0 + r < p.N ? __ldg(reinterpret_cast<const float4*>(p.w + (size_t)(n0 + r) * K + k0) + kq)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int i = tid + u * NT, r = i / (BK / 4), kq = i % (BK / 4);
      As[kq * 4 + 0][r] = ra[u].x; As[kq * 4 + 1][r] = ra[u].y; As[kq * 4 + 2][r] = ra[u].z; As[kq * 4 + 3][r] = ra[u].w;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int i = tid + u * NT, r = i / (BK / 4), kq = i % (BK / 4);
      Bs[kq * 4 + 0][r] = rb[u].x; Bs[kq * 4 + 1][r] = rb[u].y; Bs[kq * 4 + 2][r] = rb[u].z; Bs[kq * 4 + 3][r] = rb[u].w;
    }
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= p.N) return;                                        // N % 4 == 0: a thread's 4 columns are all in or all out
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) break;
    float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
    if (p.relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
    size_t dst = (size_t)m * p.N + n;
    if (p.hm_heads > 0) {
      const int D = p.N / p.hm_heads, b = m / p.hm_n, j = m - b * p.hm_n, h = n / D;
      dst = (((size_t)b * p.hm_heads + h) * p.hm_n + j) * D + (n - h * D);
    }
    *reinterpret_cast<float4*>(p.y + dst) = o;
  }
}

int launch_token_linear(TokenLinearP& p, cudaStream_t st);

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_token_linear(const void* x1, int K1, const void* x2, int K2, const void* w, const void* bias,
                                 void* y, int M, int N, int relu, void* stream) {
  SGTA_REQUIRE(x1 && w && y, "sgta_token_linear: null pointer");
  SGTA_REQUIRE(M > 0 && N > 0 && N % 4 == 0 && K1 > 0 && K1 % 16 == 0 && K2 >= 0 && K2 % 16 == 0 && (K2 == 0 || x2),
               "sgta_token_linear: M > 0, N %% 4 == 0, K1 and K2 multiples of 16 (x2 required when K2 > 0)");
  TokenLinearP p{(const float*)x1, (const float*)x2, (const float*)w, (const float*)bias, (float*)y, M, N, K1, K2, relu, 0, 0};
  return launch_token_linear(p, (cudaStream_t)stream);
}

// The K / V projections with a HEAD-MAJOR result: x [B*n_tokens, K], w [N, K] -> y [B, heads, n_tokens, N / heads]
// (what sgta_attn_forward_kvhm reads: the slab of one (sample, head) is contiguous).
extern "C" int sgta_token_linear_heads(const void* x, int K, const void* w, void* y, int M, int N, int n_tokens, int heads,
                                       void* stream) {
  SGTA_REQUIRE(x && w && y, "sgta_token_linear_heads: null pointer");
  SGTA_REQUIRE(M > 0 && N > 0 && K > 0 && K % 16 == 0 && heads > 0 && N % heads == 0 && (N / heads) % 4 == 0 &&
               n_tokens > 0 && M % n_tokens == 0, "sgta_token_linear_heads: bad shape");
  TokenLinearP p{(const float*)x, nullptr, (const float*)w, nullptr, (float*)y, M, N, K, 0, 0, heads, n_tokens};
  return launch_token_linear(p, (cudaStream_t)stream);
}

namespace sgta {
int launch_token_linear(TokenLinearP& p, cudaStream_t st) {
  const int M = p.M, N = p.N;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  if ((long long)cdiv(M, 64) * cdiv(N, 64) >= sms) {
    launch_k(token_linear_kernel<64, 64>, dim3(cdiv(N, 64), cdiv(M, 64)), 256, 0, st, p);
  } else {
    launch_k(token_linear_kernel<32, 32>, dim3(cdiv(N, 32), cdiv(M, 32)), 64, 0, st, p);
  }
  return check_launch("token_linear_kernel");
}
}  // namespace sgta
