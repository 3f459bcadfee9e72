// Post-attention half of the shared TransformerEncoderLayer, fused per token.
//
// Reference: sgtapose/lib/model/networks/dla.py:734-743 (TransformerEncoderLayer.forward) with
// :728-732 (forward_ffn) and the output projection of MHCA_ein (:886-887):
//     q1  = LayerNorm1( fc(att) + q )                       fc: Linear(hid -> C) + bias
// (fc_w and w2 are passed TRANSPOSED -- [hid][C] and [dffn][C], made once at weight-load time -- so every
// weight row a thread needs is contiguous and the staging is plain 16-byte copies)
//     q2  = LayerNorm3( q1 + W2 relu(W1 q1 + b1) + b2 )     d_ffn = 1024, dropout = identity (eval)
//     qp  = w_q q2                                          next layer's query projection (optional)
// The reference (and the round-0 engine) runs this as ~9 library launches per layer -- two fp32 GEMMs
// whose [tokens, 1024] hidden activation goes through HBM (155 MB per layer at level 0), two
// LayerNorms, bias / ReLU / residual element-wise kernels.  Here one thread owns one token: its C
// channels stay in registers from the fc projection to the last LayerNorm, the hidden units are
// produced and consumed one at a time, the weights are staged in shared memory in chunks of JH hidden
// units ([JH][C] rows of W1 and of W2^T, read back as warp-broadcast LDS.128).
// fp32 FMA throughout (the library path is SIMT fp32 too: TF32 is off for parity).
// Measured and rejected (round 2): two tokens per lane group at C = 16 with twice the lanes per token (every weight row
// read from shared memory feeds two tokens: half the LDS per token, same single wave of CTAs) -- 0.865 vs 0.840 ms
// over the nine launches of a step: the kernel is bound by the dependent FFMA2 chains per hidden unit, not by LDS.
#include "common.cuh"

namespace sgta {

template <int C> struct TmThreads { static constexpr int value = C >= 64 ? 256 : 512; };   // C = 64 needs > 128 registers

template <int C>
__device__ __forceinline__ void layer_norm_inplace(float (&x)[C], const float* __restrict__ g, const float* __restrict__ b,
                                                   float eps) {
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) mean += x[c];
  mean *= (1.f / C);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { const float d = x[c] - mean; var = fmaf(d, d, var); }
  const float inv = rsqrtf(var * (1.f / C) + eps);
#pragma unroll
  for (int c = 0; c < C; ++c) x[c] = fmaf((x[c] - mean) * inv, __ldg(g + c), __ldg(b + c));
}

struct TokenMlpP {
  const float *att, *q, *fc_w, *fc_b, *ln1_w, *ln1_b, *w1, *b1, *w2, *b2, *ln3_w, *ln3_b, *wq;
  float *q_out, *qp_out;
  int T, hid, dffn, jh;
  float eps;
};

// LPT consecutive lanes cooperate on one token (LPT = 1, 2, 4, ... 32): they split the fc inputs, the
// hidden units and the w_q outputs; partial channel vectors are all-reduced with xor-shuffles.  Small
// token counts (level 2: 2016 tokens) would otherwise leave the GPU idle behind 1024-step serial loops.
// Weight rows in shared memory are padded to C + 4 floats so the LPT different rows a warp reads per
// step fall into different banks.
template <int C, int LPT>
__device__ __forceinline__ void allreduce_vec(float (&v)[C]) {
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1)
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
}

// rows [n][C] (global, contiguous) -> [n][C + 4] (shared), 16 bytes per thread and trip
template <int C>
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int n, int tid) {
  constexpr int V = C / 4;
  const int nthreads = blockDim.x;                    // run-time (CTAs may be shrunk): keep four loads in flight by hand
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (int i0 = tid; i0 < n * V; i0 += 4 * nthreads) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n * V) v[u] = __ldg(s4 + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n * V) {
        const int r = i / V, c = i - r * V;
        *reinterpret_cast<float4*>(dst + (size_t)r * (C + 4) + 4 * c) = v[u];
      }
    }
  }
}

// smem: [max(hid, jh) * CS] + [jh * CS] floats, CS = C + 4
template <int C, int LPT>
__global__ void __launch_bounds__(TmThreads<C>::value, 1) token_mlp_kernel(const TokenMlpP p) {
  constexpr int CS = C + 4;
  const int TPB = blockDim.x / LPT;                   // tokens per CTA (blockDim.x <= TmThreads<C>, a multiple of 32)
  extern __shared__ __align__(16) float tm_s[];
  float* sA = tm_s;                                   // fc^T [hid][CS]  /  W1 chunk [jh][CS]  /  wq [hid][CS]
  float* sB = tm_s + (size_t)(p.hid > p.jh ? p.hid : p.jh) * CS;   // W2^T chunk [jh][CS]
  const int tid = threadIdx.x;
  const int sub = tid % LPT;
  const int t = blockIdx.x * TPB + tid / LPT;
  const bool live = t < p.T;
  const int tt = live ? t : p.T - 1;

  // ---- q1 = LN1(fc(att) + q): lane `sub` takes inputs k = sub, sub + LPT, ...
  pdl_trigger();
  stage_rows<C>(sA, p.fc_w, p.hid, tid);                       // fc_w^T [hid][C] -> [hid][CS]  (weights: before the wait)
  pdl_wait();
  __syncthreads();
  float x[C];
#pragma unroll
  for (int c = 0; c < C; ++c) x[c] = sub == 0 ? __ldg(p.fc_b + c) + __ldg(p.q + (size_t)tt * C + c) : 0.f;
  {
    const float* arow = p.att + (size_t)tt * p.hid;
    for (int k = sub; k < p.hid; k += LPT) {
      const float a = __ldg(arow + k);
      const float4* wr = reinterpret_cast<const float4*>(sA + (size_t)k * CS);
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 w = wr[c4];
        x[4 * c4] = fmaf(a, w.x, x[4 * c4]); x[4 * c4 + 1] = fmaf(a, w.y, x[4 * c4 + 1]);
        x[4 * c4 + 2] = fmaf(a, w.z, x[4 * c4 + 2]); x[4 * c4 + 3] = fmaf(a, w.w, x[4 * c4 + 3]);
      }
    }
  }
  allreduce_vec<C, LPT>(x);
  layer_norm_inplace<C>(x, p.ln1_w, p.ln1_b, p.eps);

  // ---- acc = W2 relu(W1 q1 + b1): hidden units in chunks of jh, lane `sub` takes j = sub, sub + LPT, ...
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = sub == 0 ? __ldg(p.b2 + c) : 0.f;
  for (int j0 = 0; j0 < p.dffn; j0 += p.jh) {
    const int nj = min(p.jh, p.dffn - j0);
    __syncthreads();                                           // previous chunk (or fc^T) consumed
    stage_rows<C>(sA, p.w1 + (size_t)j0 * C, nj, tid);         // W1   [dffn][C]
    stage_rows<C>(sB, p.w2 + (size_t)j0 * C, nj, tid);         // W2^T [dffn][C]
    __syncthreads();
#pragma unroll 2
    for (int j = sub; j < nj; j += LPT) {
      const float4* w1r = reinterpret_cast<const float4*>(sA + (size_t)j * CS);
      // channel pairs as packed FFMA2 (sm_100): (h0, h1) += (w.x, w.y) * (x[4c], x[4c+1]) then (w.z, w.w) * (x[4c+2],
      // x[4c+3]) -- the same two partial sums in the same order as the scalar form; acc pairs += (w.x, w.y) * h with the
      // scalar h broadcast by the instruction.  Each half is an IEEE fma: results are bit-identical, issue slots halve.
      float2 hh = make_float2(__ldg(p.b1 + j0 + j), 0.f);
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 w = w1r[c4];
        hh = __ffma2_rn(make_float2(w.x, w.y), make_float2(x[4 * c4], x[4 * c4 + 1]), hh);
        hh = __ffma2_rn(make_float2(w.z, w.w), make_float2(x[4 * c4 + 2], x[4 * c4 + 3]), hh);
      }
      const float h = fmaxf(hh.x + hh.y, 0.f);
      const float2 h2 = make_float2(h, h);
      const float4* w2r = reinterpret_cast<const float4*>(sB + (size_t)j * CS);
#pragma unroll
      for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 w = w2r[c4];
        const float2 a0 = __ffma2_rn(make_float2(w.x, w.y), h2, make_float2(acc[4 * c4], acc[4 * c4 + 1]));
        const float2 a1 = __ffma2_rn(make_float2(w.z, w.w), h2, make_float2(acc[4 * c4 + 2], acc[4 * c4 + 3]));
        acc[4 * c4] = a0.x; acc[4 * c4 + 1] = a0.y; acc[4 * c4 + 2] = a1.x; acc[4 * c4 + 3] = a1.y;
      }
    }
  }
  allreduce_vec<C, LPT>(acc);
#pragma unroll
  for (int c = 0; c < C; ++c) x[c] += acc[c];
  layer_norm_inplace<C>(x, p.ln3_w, p.ln3_b, p.eps);
  if (live && sub == 0) {
    float4* dst = reinterpret_cast<float4*>(p.q_out + (size_t)t * C);
#pragma unroll
    for (int c4 = 0; c4 < C / 4; ++c4) dst[c4] = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
  }

  // ---- qp = w_q q2 (next layer's query projection, no bias): lane `sub` takes outputs k = sub, sub + LPT, ...
  if (p.wq) {
    __syncthreads();
    stage_rows<C>(sA, p.wq, p.hid, tid);                         // [hid][C] -> [hid][CS]
    __syncthreads();
    if (live) {
      for (int k = sub; k < p.hid; k += LPT) {
        const float4* wr = reinterpret_cast<const float4*>(sA + (size_t)k * CS);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
          const float4 w = wr[c4];
          s0 = fmaf(w.x, x[4 * c4], s0); s1 = fmaf(w.y, x[4 * c4 + 1], s1);
          s0 = fmaf(w.z, x[4 * c4 + 2], s0); s1 = fmaf(w.w, x[4 * c4 + 3], s1);
        }
        p.qp_out[(size_t)t * p.hid + k] = s0 + s1;
      }
    }
  }
}

template <int C, int LPT>
static int launch_token_mlp2(TokenMlpP& p, cudaStream_t st, int sms) {
  constexpr int CS = C + 4;
  // one wave, every SM busy: when the tokens do not fill `sms` full-size CTAs, shrink the CTAs so that each SM
  // gets ceil(T / sms) tokens instead of leaving SMs idle (level 1: 86 CTAs of 512 -> 138 CTAs of 320 threads)
  // Every CTA stages ALL the weights whatever its token count, so smaller CTAs only pay off when they bring many
  // more SMs in: measured 165 -> 126 us at 86 -> 138 CTAs, but 74 -> 111 us at 126 -> 144 CTAs (level 2).
  int threads = TmThreads<C>::value;
  const int full = cdiv((long long)p.T * LPT, threads);
  if (full * 4 < sms * 3) {
    const int per_sm = cdiv(p.T, sms) * LPT;
    threads = ((per_sm + 31) / 32) * 32;
    if (threads > TmThreads<C>::value) threads = TmThreads<C>::value;
  }
  const int TM_THREADS = threads;
  int jh = 16384 / C;                                  // ~2 * jh * C * 4 B = 128 KB of weights per chunk
  if (jh > p.dffn) jh = p.dffn;
  p.jh = jh;
  const size_t smem = sizeof(float) * ((size_t)(p.hid > jh ? p.hid : jh) * CS + (size_t)jh * CS);
  SGTA_REQUIRE(smem <= 220 * 1024, "sgta_token_mlp: weights do not fit shared memory");
  cudaFuncSetAttribute(token_mlp_kernel<C, LPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_k(token_mlp_kernel<C, LPT>, cdiv(p.T, TM_THREADS / LPT), TM_THREADS, smem, st, p);
  return check_launch("token_mlp_kernel");
}

template <int C>
static int launch_token_mlp(TokenMlpP& p, cudaStream_t st) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // lanes per token: the largest power of two that still fits ONE wave of CTAs (one CTA per SM: the
  // weight chunk fills shared memory), so no SM idles behind a second, mostly empty wave
  int lpt = 1;
  while (lpt < 32 && cdiv((long long)p.T * (lpt * 2), TmThreads<C>::value) <= sms) lpt <<= 1;
  switch (lpt) {
    case 1: return launch_token_mlp2<C, 1>(p, st, sms);
    case 2: return launch_token_mlp2<C, 2>(p, st, sms);
    case 4: return launch_token_mlp2<C, 4>(p, st, sms);
    case 8: return launch_token_mlp2<C, 8>(p, st, sms);
    case 16: return launch_token_mlp2<C, 16>(p, st, sms);
    default: return launch_token_mlp2<C, 32>(p, st, sms);
  }
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_token_mlp(const void* att, const void* q, const void* fc_w, const void* fc_b, const void* ln1_w,
                              const void* ln1_b, const void* w1, const void* b1, const void* w2, const void* b2,
                              const void* ln3_w, const void* ln3_b, const void* wq_next, void* q_out, void* qp_out,
                              int T, int C, int hid, int dffn, float eps, void* stream) {
  SGTA_REQUIRE(att && q && fc_w && fc_b && ln1_w && ln1_b && w1 && b1 && w2 && b2 && ln3_w && ln3_b && q_out,
               "sgta_token_mlp: null pointer");
  SGTA_REQUIRE(!wq_next || qp_out, "sgta_token_mlp: wq_next given without qp_out");
  SGTA_REQUIRE(T > 0 && hid > 0 && hid % 4 == 0 && dffn > 0, "sgta_token_mlp: bad shape");
  TokenMlpP p{(const float*)att, (const float*)q, (const float*)fc_w, (const float*)fc_b, (const float*)ln1_w,
              (const float*)ln1_b, (const float*)w1, (const float*)b1, (const float*)w2, (const float*)b2,
              (const float*)ln3_w, (const float*)ln3_b, (const float*)wq_next, (float*)q_out, (float*)qp_out,
              T, hid, dffn, 0, eps};
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 16: return launch_token_mlp<16>(p, st);
    case 32: return launch_token_mlp<32>(p, st);
    case 64: return launch_token_mlp<64>(p, st);
    default:
      set_error("sgta_token_mlp: C = %d not in {16, 32, 64}", C);
      return SGTA_EUNSUPPORTED;
  }
}
