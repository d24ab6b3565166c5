// Hardware probe (B200): SWIZZLE_128B K-major UMMA descriptors whose start address / stride-byte-offset are NOT multiples of
// 1024 B -- needed to read the 9 taps of a conv from ONE halo tile in shared memory.  D = A * I (B = identity over 64
// channels) so D[m][c] reveals exactly which smem row the tensor core fetched for GEMM row m.
#include "../unsupervised_detection_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace cis;

__device__ __host__ inline float fval(int p, int c) { return (float)(((p * 7 + c * 3) % 13) - 6); }

__global__ void probe(int off_rows, int Wh, int bo_mode, int lbo, float* out /*[128][64]*/) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const int NP = 256;  // halo pixels
  // A halo: pixel p at row p (128 B), chunk j at (j ^ (p & 7))
  for (int i = threadIdx.x; i < NP * 64; i += blockDim.x) {
    int p = i / 64, c = i % 64;
    int j = c / 8, e = c % 8;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(gen + p * 128 + ((j ^ (p & 7)) << 4)) + e;
    *dst = __float2bfloat16(fval(p, c));
  }
  // B identity [64 rows n][64 k], K-major SW128, at offset 32 KB
  uint8_t* bgen = gen + 32768;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    int n = i / 64, c = i % 64;
    int j = c / 8, e = c % 8;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(bgen + n * 128 + ((j ^ (n & 7)) << 4)) + e;
    *dst = __float2bfloat16(n == c ? 1.f : 0.f);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<64>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      uint32_t sa = base + off_rows * 128 + k * 32;
      uint64_t da = make_smem_desc(sa, lbo, Wh * 128);
      if (bo_mode == 1) da |= (uint64_t)((sa >> 7) & 7) << 49;
      uint64_t db = make_smem_desc(base + 32768 + k * 32, 16, 1024);
      umma_bf16(tmem, da, db, idesc, k != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 64 + c0 + e] = v[e];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<64>(tmem);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  std::vector<float> h(128 * 64);
  int offs[] = {0, 1, 3, 8, 10, 11, 21};
  int whs[] = {8, 10, 11, 12, 16};
  for (int bo = 0; bo < 2; ++bo)
    for (int wh : whs)
      for (int off : offs) {
        cudaMemset(d, 0, 128 * 64 * 4);
        probe<<<1, 128, 65536>>>(off, wh, bo, 16, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bo=%d Wh=%d off=%d CUDA ERROR %s\n", bo, wh, off, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int c = 0; c < 64; ++c) {
            int p = off + (m / 8) * wh + (m % 8);
            if (h[m * 64 + c] != fval(p, c)) { if (first < 0) first = m * 64 + c; ++bad; }
          }
        printf("bo_mode=%d Wh=%2d off=%2d : %s (bad=%d first=%d)\n", bo, wh, off, bad ? "MISMATCH" : "ok", bad, first);
      }
  return 0;
}
