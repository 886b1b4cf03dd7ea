// CUDA-core (fp32 FFMA) implementations of the path's operators.  These are the on-device
// reference for the tcgen05 kernels (engine option gemm=simt) and serve the GEMMs whose
// shapes do not fit a tensor-core tile (odd K such as 263/150 motion features, tiny M).
#include "ops.cuh"

#include <stdlib.h>

// ------------------------------------------------------------------------------------ GEMM
namespace {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__device__ __forceinline__ float load_a(const GemmArgs& a, int m, int k) {
  if (a.a_kind == A_SPLIT) {
    if (k < a.K1) {
      int64_t o = (int64_t)m * a.a1.cols + k;
      return join_f32(a.a1.hi[o], a.a1.lo()[o]);
    }
    int64_t o = (int64_t)m * a.a2.cols + (k - a.K1);
    return join_f32(a.a2.hi[o], a.a2.lo()[o]);
  }
  if (k >= a.lda) return 0.0f;               // weights packed with a zero-padded K (odd feature counts)
  float v = a.a_f32[(int64_t)m * a.lda + k];
  return a.a_kind == A_F32_RELU ? fmaxf(v, 0.0f) : v;
}

__global__ void __launch_bounds__(256) k_gemm_simt(const GemmArgs a) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int K = a.w.K, N = a.w.N, M = a.M;
  const __half* whi = a.w.w;
  const __half* wlo = a.w.w + a.w.plane_stride;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      const int m = m0 + lr, n = n0 + lr;
      As[lk + i][lr] = (m < M && k < K) ? load_a(a, m, k) : 0.0f;
      float wv = 0.0f;
      if (n < N && k < K) {
        int64_t o = (int64_t)n * K + k;
        wv = join_f32(whi[o], wlo[o]);
      }
      Ws[lk + i][lr] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int seq = m / a.in_group, pos = m - seq * a.in_group;
    const int64_t orow = (a.in_group >= M && a.out_group == 0)
                             ? (int64_t)m
                             : (int64_t)seq * a.out_group + a.out_off + pos;
    const bool zero = a.zero_lengths != nullptr && pos >= a.zero_lengths[seq];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] * a.w.inv_scale;
      if (a.w.bias) v += a.w.bias[n];
      if (a.addtab) v += a.addtab[(int64_t)(a.out_off + pos) * N + n];
      v = apply_act(v, a.act);
      if (zero) v = 0.0f;
      if (a.out.hi) {
        __half h, l;
        split_f32(v, h, l);
        int64_t o = orow * a.out.cols + a.out_col0 + n;
        a.out.hi[o] = h;
        a.out.lo()[o] = l;
      }
      if (a.out_f32) a.out_f32[orow * a.ldc + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------- LayerNorm
template <int VPL>
__global__ void __launch_bounds__(256) k_ln(const LnArgs a) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const int r = warp;
  const int64_t irow = (a.in_group == 0) ? (int64_t)r
                                         : (int64_t)(r / a.sel_group) * a.in_group + r % a.sel_group;
  float v[VPL];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int n = i * 32 + lane;
    float x = 0.0f;
    if (n < a.d) {
      if (a.c) x += a.c[irow * a.ldc + n];
      if (a.res.hi) {
        int64_t o = irow * a.res.cols + n;
        x += join_f32(a.res.hi[o], a.res.lo()[o]);
      }
      if (a.rowvec) x += a.rowvec[(int64_t)(irow / a.rv_group) * a.d + n];
    }
    v[i] = x;
    s += x;
  }
  const float inv_d = 1.0f / (float)a.d;
  float mean = warp_sum(s) * inv_d;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int n = i * 32 + lane;
    float dlt = (n < a.d) ? v[i] - mean : 0.0f;
    q += dlt * dlt;
  }
  float rstd = rsqrtf(warp_sum(q) * inv_d + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int n = i * 32 + lane;
    if (n < a.d) v[i] = (v[i] - mean) * rstd * a.gamma[n] + a.beta[n];
  }
  if (a.gamma2) {  // stack-final LayerNorm on top (cross_attention.py:62-63)
    s = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (i * 32 + lane < a.d) ? v[i] : 0.0f;
    mean = warp_sum(s) * inv_d;
    q = 0.0f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float dlt = (i * 32 + lane < a.d) ? v[i] - mean : 0.0f;
      q += dlt * dlt;
    }
    rstd = rsqrtf(warp_sum(q) * inv_d + 1e-5f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int n = i * 32 + lane;
      if (n < a.d) v[i] = (v[i] - mean) * rstd * a.gamma2[n] + a.beta2[n];
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int n = i * 32 + lane;
    if (n >= a.d) continue;
    if (a.out.hi) {
      __half h, l;
      split_f32(v[i], h, l);
      int64_t o = (int64_t)r * a.out.cols + n;
      a.out.hi[o] = h;
      a.out.lo()[o] = l;
    }
    if (a.out_f32) a.out_f32[(int64_t)r * a.ld_out + n] = v[i];
  }
}

// ------------------------------------------------------------------------------- attention
// One block per (sequence, head).  K (padded rows) and V live in shared memory as fp32; each
// warp owns query rows q = warp, warp + nwarps, ...  Softmax over the valid keys only (the
// reference masks padded keys with -inf: cross_attention.py:264-266, mld_vae.py:226-232).
constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32) k_attn_simt(const AttnArgs a) {
  extern __shared__ float sm[];
  const int s = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int hd = a.hd, Lk = a.Lk, Lq = a.Lq;
  int nk = Lk;
  if (a.lengths) {
    const int li = a.len_mod > 0 ? (a.seq0 + s) % a.len_mod : s;
    nk = min(Lk, a.kv_prefix + a.lengths[li]);
  }
  float* Ks = sm;                          // [Lk][hd+1]
  float* Vs = Ks + (size_t)Lk * (hd + 1);  // [Lk][hd]
  float* Qs = Vs + (size_t)Lk * hd;        // [ATT_WARPS][hd]
  float* Ps = Qs + ATT_WARPS * hd;         // [ATT_WARPS][Lk]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* khi = a.kv.hi;
  const __half* klo = a.kv.lo();
  for (int i = tid; i < nk * hd; i += blockDim.x) {
    const int t = i / hd, d = i - t * hd;
    const int64_t row = (int64_t)s * Lk + t;
    int64_t ok = row * a.kv.cols + a.k_col0 + h * hd + d;
    int64_t ov = row * a.kv.cols + a.v_col0 + h * hd + d;
    Ks[t * (hd + 1) + d] = join_f32(khi[ok], klo[ok]);
    Vs[t * hd + d] = join_f32(khi[ov], klo[ov]);
  }
  __syncthreads();
  const float scale = rsqrtf((float)hd);
  float* qv = Qs + warp * hd;
  float* pv = Ps + warp * Lk;
  for (int qi = warp; qi < Lq; qi += ATT_WARPS) {
    const int64_t qrow = (int64_t)s * Lq + qi;
    for (int d = lane; d < hd; d += 32) {
      int64_t o = qrow * a.q.cols + a.q_col0 + h * hd + d;
      qv[d] = join_f32(a.q.hi[o], a.q.lo()[o]) * scale;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int t = lane; t < nk; t += 32) {
      const float* kr = Ks + t * (hd + 1);
      float acc = 0.0f;
#pragma unroll 8
      for (int d = 0; d < hd; ++d) acc = fmaf(qv[d], kr[d], acc);
      pv[t] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int t = lane; t < nk; t += 32) {
      float e = expf(pv[t] - mx);
      pv[t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.0f;
      for (int t = 0; t < nk; ++t) acc = fmaf(pv[t], Vs[t * hd + d], acc);
      acc *= inv;
      __half hh, ll;
      split_f32(acc, hh, ll);
      int64_t o = qrow * a.out.cols + h * hd + d;
      a.out.hi[o] = hh;
      a.out.lo()[o] = ll;
    }
    __syncwarp();
  }
}

}  // namespace

void simt_gemm(const GemmArgs& a, cudaStream_t st) {
  dim3 grid((a.w.N + SG_BN - 1) / SG_BN, (a.M + SG_BM - 1) / SG_BM);
  k_gemm_simt<<<grid, 256, 0, st>>>(a);
}

void simt_ln(const LnArgs& a, cudaStream_t st) {
  const int rows_per_block = 8;
  dim3 grid((a.M + rows_per_block - 1) / rows_per_block);
  if (a.d <= 256) launch_pdl(k_ln<8>, grid, dim3(256), 0, st, a);
  else if (a.d <= 512) launch_pdl(k_ln<16>, grid, dim3(256), 0, st, a);
  else launch_pdl(k_ln<32>, grid, dim3(256), 0, st, a);
}

size_t simt_attention_smem(const AttnArgs& a) {
  return ((size_t)a.Lk * (2 * a.hd + 1) + (size_t)ATT_WARPS * (a.hd + a.Lk)) * sizeof(float);
}

int g_mldb_pdl = 1;

void simt_init() {
  if (const char* e = getenv("MLDB_PDL")) g_mldb_pdl = atoi(e) != 0;
  // opt in to the full 227 KB once (not during stream capture)
  cudaFuncSetAttribute(k_attn_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

void simt_attention(const AttnArgs& a, cudaStream_t st) {
  const size_t smem = simt_attention_smem(a);
  k_attn_simt<<<a.nseq * a.heads, ATT_WARPS * 32, smem, st>>>(a);
}
