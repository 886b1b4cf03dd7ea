// EXPERIMENTAL - compiles, but has NOT yet run on hardware; off by default (engine option `attn_tc`,
// env MLDB_ATTN_TC=1).  The product path uses attn_mma.cu.  First GPU task of the next round:
// tests/test_gpu_kernels.py::test_tc_attention_matches_mma (MLDB_EXPERIMENTAL=1).
//
// Multi-head attention core on tcgen05 (replaces k_attn_mma<64> for Lq <= 128, 16 <= Lk <= 128, hd = 64):
// one work item per (sequence, head), persistent CTAs, 6 warps:
//   warp 0  TMA producer: Q [128 x 64] (rows past the sequence are the next sequence's rows or zero
//           fill - their S / O rows are never stored), K [Np x 64], V [Np x 64], hi + lo planes,
//           SWIZZLE_128B boxes straight out of the split16 qkv buffer (column slices by coordinate)
//   warp 1  MMA issuer:  S = Q K^T   (M = 128, N = Np, K = 64: 4 k-steps x 3 products, K-major A and B)
//                        O = P V     (M = 128, N = 64, K = Np: Np/16 k-steps x 3 products, A = P from
//                                     shared memory (K-major), B = V as an MN-major operand: the
//                                     [keys x d] tile as TMA delivers it, no transpose)
//           issue order S(0), S(1), PV(0), S(2), PV(1), ... ; S and O double-buffered in TMEM
//   warps 2-5  softmax + epilogue, one query row per thread (TMEM lane = row):
//           S row -> registers, scale, key mask (>= nk -> -inf), max, exp2, sum (fp32), P re-split to
//           hi / lo fp16 and written to shared memory in the UMMA K-major SWIZZLE_128B layout (the
//           same trick as k_ffn_tc's hidden chunk), then O row * (1 / sum) -> split16 -> global.
// Numerics are those of attn_mma.cu: 3-product split MMAs with fp32 accumulation, fp32 softmax.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "ops.cuh"

namespace {

constexpr int ATC_THREADS = 192;
constexpr int HD = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins & 1023u) == 0) {         // a lost arrival must not hang the GPU
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor).
// K-major operand: rows = M/N index, 64 K-elements (128 B) per row; next 16-wide k-step = +32 B.
// MN-major operand (b_major = 1): rows = K index, 64 MN-elements (128 B) per row - the canonical
// ((8,n),(8,k)):((1,LBO),(8,SBO)) layout with one 64-wide MN block; next 16-deep k-step = +2048 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = F32 (bit 4), A = B = F16, b_major at bit 16, N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t make_idesc(int n, bool b_mn_major) {
  return (1u << 4) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct AtcParams {
  int nseq, heads, Lq, Lk, Np;          // Np = keys padded to a multiple of 16
  int q_col0, k_col0, v_col0;
  const int32_t* lengths; int kv_prefix, len_mod, seq0;
  __half* out_hi; __half* out_lo; int ld_out;
  float scale_log2e;                    // (1 / sqrt(hd)) * log2(e)
  int stages;
};

// shared memory: [stages] x {Qh, Ql (16 KB each), Kh, Kl, Vh, Vl (Np x 128 B each)}, then P:
// {hi k-block 0, hi k-block 1, lo k-block 0, lo k-block 1} (16 KB each), then barriers
__host__ __device__ inline int atc_stage_bytes(int Np) { return 2 * 16384 + 4 * Np * 128; }
constexpr int ATC_P_BYTES = 4 * 16384;

__global__ void __launch_bounds__(ATC_THREADS, 1)
k_attn_tc(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
          const __grid_constant__ CUtensorMap tmKVh, const __grid_constant__ CUtensorMap tmKVl, const AtcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Np = p.Np, stages = p.stages;
  const int stage_bytes = atc_stage_bytes(Np);          // a multiple of 1024 (Np % 16 == 0 -> 4*Np*128 % 8192 == 0)
  uint8_t* sP = smem + stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + ATC_P_BYTES);
  uint64_t* bar_full = bars;          // [2] Q/K/V of an item landed
  uint64_t* bar_empty = bars + 2;     // [2] ... and consumed (PV committed)
  uint64_t* bar_sfull = bars + 4;     // [2] S accumulated
  uint64_t* bar_sempty = bars + 6;    // [2] S read by the softmax warps
  uint64_t* bar_pfull = bars + 8;     // P written
  uint64_t* bar_pempty = bars + 9;    // P consumed (PV committed)
  uint64_t* bar_ofull = bars + 10;    // [2] O accumulated
  uint64_t* bar_oempty = bars + 12;   // [2] O read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int items = p.nseq * p.heads;
  const int nlocal = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
      mbar_init(smem_u32(&bar_sfull[s]), 1);
      mbar_init(smem_u32(&bar_sempty[s]), 4);
      mbar_init(smem_u32(&bar_ofull[s]), 1);
      mbar_init(smem_u32(&bar_oempty[s]), 4);
    }
    mbar_init(smem_u32(bar_pfull), 4);
    mbar_init(smem_u32(bar_pempty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // TMEM: S[2] at columns 0 / 128, O[2] at columns 256 / 320
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int j = 0; j < nlocal; ++j) {
      const int item = (int)blockIdx.x + j * (int)gridDim.x;
      const int s = item / p.heads, h = item - s * p.heads;
      const int st = j % stages;
      mbar_wait(smem_u32(&bar_empty[st]), (((uint32_t)(j / stages)) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bar_full[st]);
        mbar_expect_tx(full, (uint32_t)stage_bytes);
        const uint32_t base = smem_u32(smem + st * stage_bytes);
        const uint32_t kb = base + 32768, vb = kb + 2 * Np * 128;
        tma_load_2d(base, &tmQh, full, p.q_col0 + h * HD, s * p.Lq);
        tma_load_2d(base + 16384, &tmQl, full, p.q_col0 + h * HD, s * p.Lq);
        tma_load_2d(kb, &tmKVh, full, p.k_col0 + h * HD, s * p.Lk);
        tma_load_2d(kb + Np * 128, &tmKVl, full, p.k_col0 + h * HD, s * p.Lk);
        tma_load_2d(vb, &tmKVh, full, p.v_col0 + h * HD, s * p.Lk);
        tma_load_2d(vb + Np * 128, &tmKVl, full, p.v_col0 + h * HD, s * p.Lk);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = make_idesc(Np, false), idesc_o = make_idesc(HD, true);
    const uint32_t sP_u = smem_u32(sP);
    auto issue_s = [&](int j) {           // S(j) = Q K^T into S[j & 1]
      const int st = j % stages, b = j & 1;
      mbar_wait(smem_u32(&bar_full[st]), ((uint32_t)(j / stages)) & 1u);
      mbar_wait(smem_u32(&bar_sempty[b]), (((uint32_t)j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + st * stage_bytes);
        uint64_t qh = make_desc(base), ql = make_desc(base + 16384);
        uint64_t kh = make_desc(base + 32768), kl = make_desc(base + 32768 + Np * 128);
        const uint32_t tacc = tmem_base + (uint32_t)(b * 128);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          umma(tacc, ql, kh, idesc_s, kk != 0 ? 1u : 0u);
          umma(tacc, qh, kl, idesc_s, 1u);
          umma(tacc, qh, kh, idesc_s, 1u);
          qh += 2; ql += 2; kh += 2; kl += 2;          // next 16-wide slice of the head dimension: +32 B
        }
        umma_commit(smem_u32(&bar_sfull[b]));
      }
      __syncwarp();
    };
    auto issue_pv = [&](int j) {          // O(j) = P V into O[j & 1]
      const int st = j % stages, b = j & 1;
      mbar_wait(smem_u32(bar_pfull), (uint32_t)j & 1u);
      mbar_wait(smem_u32(&bar_oempty[b]), (((uint32_t)j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t base = smem_u32(smem + st * stage_bytes);
        const uint32_t vb = base + 32768 + 2 * Np * 128;
        uint64_t vh = make_desc(vb), vl = make_desc(vb + Np * 128);
        const uint32_t tacc = tmem_base + 256u + (uint32_t)(b * 64);
        for (int ks = 0; ks < Np / 16; ++ks) {
          // P: k-block (64 keys) ks / 4 of each plane, 16-key slice ks % 4 inside it
          const uint32_t po = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
          const uint64_t ph = make_desc(sP_u + po), pl = make_desc(sP_u + 32768 + po);
          umma(tacc, pl, vh, idesc_o, ks != 0 ? 1u : 0u);
          umma(tacc, ph, vl, idesc_o, 1u);
          umma(tacc, ph, vh, idesc_o, 1u);
          vh += 128; vl += 128;                        // next 16 keys of the MN-major V tile: +2048 B
        }
        umma_commit(smem_u32(&bar_ofull[b]));
        umma_commit(smem_u32(bar_pempty));
        umma_commit(smem_u32(&bar_empty[st]));         // Q / K / V of this item may be overwritten
      }
      __syncwarp();
    };
    if (nlocal > 0) issue_s(0);
    for (int j = 0; j < nlocal; ++j) {
      if (j + 1 < nlocal && stages > 1) issue_s(j + 1);   // S(j+1) overlaps softmax(j)
      issue_pv(j);
      if (j + 1 < nlocal && stages == 1) issue_s(j + 1);  // single stage: the slot is free only after PV(j)
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 2..5)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // query row of this thread
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int nch = Np / 16;                // 16-column chunks of S
    for (int j = 0; j < nlocal; ++j) {
      const int item = (int)blockIdx.x + j * (int)gridDim.x;
      const int s = item / p.heads, h = item - s * p.heads;
      const int b = j & 1;
      int nk = p.Lk;
      if (p.lengths) nk = min(p.Lk, p.kv_prefix + p.lengths[p.len_mod > 0 ? (p.seq0 + s) % p.len_mod : s]);
      // ---- S row -> registers (scaled into the log2 domain), masked, running max
      mbar_wait(smem_u32(&bar_sfull[b]), ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      float sv[128];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nch) {                      // warp-uniform
          uint32_t r[16];
          tmem_ld16(tmem_base + lane_addr + (uint32_t)(b * 128 + c * 16), r);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v = (c * 16 + i < nk) ? __uint_as_float(r[i]) * p.scale_log2e : -INFINITY;
            sv[c * 16 + i] = v;
            mx = fmaxf(mx, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_sempty[b]));
      // ---- P = exp2(s - max) (unnormalised), row sum in fp32
      float sum = 0.0f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nch) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(sv[c * 16 + i] - mx));   // exp2(-inf) = 0 for masked keys
            sv[c * 16 + i] = e;
            sum += e;
          }
        }
      }
      // ---- P -> shared memory (hi / lo planes, K-major SWIZZLE_128B: 16-B chunk index XOR (row & 7))
      mbar_wait(smem_u32(bar_pempty), ((uint32_t)j & 1u) ^ 1u);        // PV(j - 1) has read P
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < nch) {
          uint32_t ph[8], pl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const __half2 h2 = __floats2half2_rn(sv[c * 16 + 2 * i], sv[c * 16 + 2 * i + 1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(sv[c * 16 + 2 * i] - hf.x, sv[c * 16 + 2 * i + 1] - hf.y);
            ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
            pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          // chunk c covers keys c*16 .. c*16+15: k-block c / 4, 16-B slots 2*(c % 4) and 2*(c % 4) + 1 of the row
          uint8_t* const prow = sP + (c >> 2) * 16384 + row * 128;
          const int s0 = ((2 * (c & 3)) ^ (row & 7)) << 4, s1 = ((2 * (c & 3) + 1) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + s0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(prow + s1) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
          *reinterpret_cast<uint4*>(prow + 32768 + s0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          *reinterpret_cast<uint4*>(prow + 32768 + s1) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tcgen05.mma reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_pfull));
      // ---- O row / sum -> split16 -> global (each thread owns 128 contiguous bytes per plane)
      const float inv = 1.0f / sum;
      mbar_wait(smem_u32(&bar_ofull[b]), ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      uint32_t oh[32], ol[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[16];
        tmem_ld16(tmem_base + lane_addr + 256u + (uint32_t)(b * 64 + c * 16), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x0 = __uint_as_float(r[2 * i]) * inv, x1 = __uint_as_float(r[2 * i + 1]) * inv;
          const __half2 h2 = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          oh[c * 8 + i] = *reinterpret_cast<const uint32_t*>(&h2);
          ol[c * 8 + i] = *reinterpret_cast<const uint32_t*>(&l2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_oempty[b]));
      if (row < p.Lq) {
        const int64_t o = ((int64_t)s * p.Lq + row) * p.ld_out + h * HD;
        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + o);
        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dh[i] = make_uint4(oh[4 * i], oh[4 * i + 1], oh[4 * i + 2], oh[4 * i + 3]);
          dl[i] = make_uint4(ol[4 * i], ol[4 * i + 1], ol[4 * i + 2], ol[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled g_encode = nullptr;
int g_sm_count = 148;

// fp16 plane [rows, cols]; box = 64 columns x box_rows rows, SWIZZLE_128B, out-of-bounds rows zero-filled
bool make_map(CUtensorMap* m, const __half* base, int rows, int cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)HD, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int atc_smem(int Np, int stages) { return stages * atc_stage_bytes(Np) + ATC_P_BYTES + 256 + 1024; }

}  // namespace

bool tc_attention_init(int device) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn)
    return false;
  g_encode = (PFN_tmapEncodeTiled)fn;
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, device);
  return cudaFuncSetAttribute(k_attn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
}

bool tc_attention_supported(const AttnArgs& a) {
  // lazy: nothing of this file runs unless the experimental option is switched on
  static bool tried = false;
  if (!tried) {
    tried = true;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || !tc_attention_init(dev)) { cudaGetLastError(); g_encode = nullptr; }
  }
  if (!g_encode || a.hd != HD || a.Lq < 1 || a.Lq > 128 || a.Lk < 16 || a.Lk > 128) return false;
  if ((a.q.cols % 8) || (a.kv.cols % 8) || (a.q_col0 % 8) || (a.k_col0 % 8) || (a.v_col0 % 8)) return false;
  if ((a.out.cols % 8) || ((uintptr_t)a.q.hi & 15) || ((uintptr_t)a.kv.hi & 15) || ((uintptr_t)a.out.hi & 15)) return false;
  const int Np = (a.Lk + 15) & ~15;
  return atc_smem(Np, 1) <= 227 * 1024;
}

void tc_attention(const AttnArgs& a, cudaStream_t st) {
  const int Np = (a.Lk + 15) & ~15;
  const int stages = atc_smem(Np, 2) <= 227 * 1024 ? 2 : 1;
  CUtensorMap mQh, mQl, mKVh, mKVl;
  const bool ok = make_map(&mQh, a.q.hi, a.q.rows, a.q.cols, 128) && make_map(&mQl, a.q.lo(), a.q.rows, a.q.cols, 128) &&
                  make_map(&mKVh, a.kv.hi, a.kv.rows, a.kv.cols, Np) && make_map(&mKVl, a.kv.lo(), a.kv.rows, a.kv.cols, Np);
  if (!ok) {
    fprintf(stderr, "libmldb200: cuTensorMapEncodeTiled failed (attention Lq=%d Lk=%d)\n", a.Lq, a.Lk);
    return;
  }
  AtcParams p{};
  p.nseq = a.nseq; p.heads = a.heads; p.Lq = a.Lq; p.Lk = a.Lk; p.Np = Np;
  p.q_col0 = a.q_col0; p.k_col0 = a.k_col0; p.v_col0 = a.v_col0;
  p.lengths = a.lengths; p.kv_prefix = a.kv_prefix; p.len_mod = a.len_mod; p.seq0 = a.seq0;
  p.out_hi = a.out.hi; p.out_lo = a.out.lo(); p.ld_out = a.out.cols;
  p.scale_log2e = 1.4426950408889634f / sqrtf((float)a.hd);
  p.stages = stages;
  const int items = a.nseq * a.heads;
  const int grid = items < g_sm_count ? items : g_sm_count;
  launch_pdl(k_attn_tc, dim3(grid), dim3(ATC_THREADS), (size_t)atc_smem(Np, stages), st, mQh, mQl, mKVh, mKVl, p);
}
