// Microbenchmark: how fast can 148 persistent CTAs drain split16 output tiles to global memory?
// Variants: (a) the GEMM epilogue's pattern (4 lanes x 16 B per 64-B row segment), (b) 8 lanes x 16 B
// (128-B row segments), (c) fully contiguous 512 B per warp store, (d) TMA bulk tensor stores from
// shared memory (32 rows x 64 B boxes), (e) TMA stores of 128 rows x 128 B boxes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int ROWS_PER_TILE = 128;

// pattern a/b: LPR lanes per row, 16 B per lane -> row segment = LPR*16 bytes; tile = 128 rows x SEG
template <int LPR>
__global__ void k_store_seg(uint4* out, int64_t ld16 /* row pitch in uint4 */, int m_tiles, int n_segs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, hf = warp >> 2;           // 8 warps: row quarter, half of the segments
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  constexpr int RPW = 32 / LPR;                     // rows per warp instruction
  for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) {
    const int64_t row0 = (int64_t)t * ROWS_PER_TILE + q * 32;
    for (int s = hf; s < n_segs; s += 2) {
#pragma unroll
      for (int i = 0; i < 32 / RPW; ++i) {
        const int rr = i * RPW + lane / LPR, qq = lane % LPR;
        out[(row0 + rr) * ld16 + (int64_t)s * LPR + qq] = v;
      }
    }
  }
}
// contiguous: every warp instruction writes 512 contiguous bytes
__global__ void k_store_lin(uint4* out, int64_t n16) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) out[i] = v;
}
__global__ void k_read_lin(const uint4* in, int64_t n16, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = in[i]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
  }
  if (acc.x == 0x12345678u) sink[0] = acc;
}

// TMA store: box = BOXC fp16 columns x BOXR rows, one issuing thread per warp, staging in smem
template <int BOXR, int BOXB /* bytes per box row */, int NW /* issuing warps */>
__global__ void k_store_tma(const __grid_constant__ CUtensorMap tm, int m_tiles, int n_cols /* fp16 */) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BOX_BYTES = BOXR * BOXB;
  constexpr int BOXES_PER_TILE_ROWS = ROWS_PER_TILE / BOXR;       // 4 (32-row boxes) or 1
  if (warp >= NW) return;
  uint8_t* my = smem + warp * 2 * BOX_BYTES;                        // double buffer per warp
  for (int i = lane; i < 2 * BOX_BYTES / 16; i += 32) reinterpret_cast<uint4*>(my)[i] = make_uint4(i, warp, 1, 2);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int nbc = n_cols * 2 / BOXB;                                // boxes along a row
  int buf = 0;
  for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) {
    // 8 warps share the tile's boxes
    const int nboxes = BOXES_PER_TILE_ROWS * nbc;
    for (int b = warp; b < nboxes; b += NW) {
      const int br = b % BOXES_PER_TILE_ROWS, bc = b / BOXES_PER_TILE_ROWS;
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer we are about to reuse is free
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(my + buf * BOX_BYTES);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(src), "r"(bc * (BOXB / 2)), "r"(t * ROWS_PER_TILE + br * BOXR) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      buf ^= 1;
      __syncwarp();
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
static float time_it(F f, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  const int M = 40448, m_tiles = (M + 127) / 128;
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  PFN_enc enc = (PFN_enc)fn;
  for (int N : {768, 1024, 256}) {
    const int planes = 2;
    const size_t bytes = (size_t)planes * m_tiles * 128 * N * 2;
    uint4* buf; CK(cudaMalloc(&buf, bytes + (1 << 20)));
    const int64_t ld16 = (int64_t)N * 2 / 16;
    // treat the two planes as 2*m_tiles tiles of one [2*Mpad, N] matrix
    const int tiles = planes * m_tiles;
    printf("== N=%d: %d tiles, %.1f MB per pass\n", N, tiles, bytes / 1e6);
    float ms;
    ms = time_it([&] { k_store_seg<4><<<148, 256>>>(buf, ld16, tiles, N * 2 / 64); }, 20);
    printf("  st.global 64-B row segments : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
    ms = time_it([&] { k_store_seg<8><<<148, 256>>>(buf, ld16, tiles, N * 2 / 128); }, 20);
    printf("  st.global 128-B row segments: %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
    ms = time_it([&] { k_store_seg<32><<<148, 256>>>(buf, ld16, tiles, N * 2 / 512 ? N * 2 / 512 : 1); }, 20);
    printf("  st.global 512-B row segments: %7.1f us  %6.2f TB/s\n", ms * 1e3, (N * 2 >= 512 ? bytes : bytes) / ms / 1e9);
    ms = time_it([&] { k_store_lin<<<148 * 4, 256>>>(buf, bytes / 16); }, 20);
    printf("  st.global linear (592 CTAs) : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
    ms = time_it([&] { k_store_lin<<<148, 256>>>(buf, bytes / 16); }, 20);
    printf("  st.global linear (148 CTAs) : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
    ms = time_it([&] { k_read_lin<<<148 * 8, 256>>>(buf, bytes / 16, buf); }, 20);
    printf("  ld.global linear            : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
    {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)tiles * 128};
      cuuint64_t strides[1] = {(cuuint64_t)N * 2};
      cuuint32_t estr[2] = {1, 1};
      cuuint32_t box1[2] = {32, 32};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int smem1 = 8 * 2 * 32 * 64;
      ms = time_it([&] { k_store_tma<32, 64, 8><<<148, 256, smem1>>>(tm, tiles, N); }, 20);
      CK(cudaGetLastError());
      printf("  TMA store 32x64B boxes      : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
      cuuint32_t box2[2] = {64, 128};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int smem2 = 4 * 2 * 128 * 128;
      CK(cudaFuncSetAttribute(k_store_tma<128, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      ms = time_it([&] { k_store_tma<128, 128, 4><<<148, 256, smem2>>>(tm, tiles, N); }, 20);
      CK(cudaGetLastError());
      printf("  TMA store 128x128B boxes    : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
      cuuint32_t box3[2] = {64, 32};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box3, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const int smem3 = 8 * 2 * 32 * 128;
      CK(cudaFuncSetAttribute(k_store_tma<32, 128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
      ms = time_it([&] { k_store_tma<32, 128, 8><<<148, 256, smem3>>>(tm, tiles, N); }, 20);
      CK(cudaGetLastError());
      printf("  TMA store 32x128B boxes     : %7.1f us  %6.2f TB/s\n", ms * 1e3, bytes / ms / 1e9);
    }
    CK(cudaDeviceSynchronize());
    cudaFree(buf);
  }
  return 0;
}
