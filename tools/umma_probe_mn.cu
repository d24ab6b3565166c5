// Hardware probe (B200), prepared for DESIGN.md section 6 / E4: halo-resident weight gradient.  Can an MN-major SWIZZLE_128B UMMA
// operand be read IN PLACE from a TMA-written activation halo, i.e. with
//   * a descriptor start that is 128-byte (one pixel row) granular instead of 1024-byte aligned (tap shift dy*Wh+dx pixels),
//   * SBO (distance between 8-row K groups = between the 8-pixel rows of an 8x8 pixel tile) = Wh*128 B, not a multiple of 1024,
//   * LBO (distance between the two 64-channel MN atoms of an M=128 operand) = an arbitrary multiple of 128 B (two different taps)?
// D[m][n] = sum_k A[k][m] * B[k][n], A = x halo (K = pixels, M = channels of two taps), B = selector (K = pixels, N = 16): B[k][n] = (k == n)
// so D[m][n] = A[pixel n][channel m] and shows exactly which smem row the tensor core fetched for K index n, MN index m.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe_mn tools/umma_probe_mn.cu
#include "../unsupervised_detection_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace cis;

__device__ __host__ inline float fval(int p, int c) { return (float)(((p * 5 + c * 3) % 17) - 8); }

__global__ void probe(int off_rows, int Wh, int lbo_rows, float* out /*[128][16]*/) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const int NP = 256;   // halo pixels: pixel p = one 128-byte row (64 channels), 16-byte chunk j stored at (j ^ (p & 7)) like TMA SWIZZLE_128B
  for (int i = threadIdx.x; i < NP * 64; i += blockDim.x) {
    const int p = i / 64, c = i % 64, j = c / 8, e = c % 8;
    reinterpret_cast<__nv_bfloat16*>(gen + p * 128 + ((j ^ (p & 7)) << 4))[e] = __float2bfloat16(fval(p, c));
  }
  // B: 16 K rows (pixels) x 16 N (padded to one 64-wide MN atom), MN-major SW128, 1024-aligned at +32 KB: row k holds e_k
  uint8_t* bgen = gen + 32768;
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int k = i / 64, n = i % 64, j = n / 8, e = n % 8;
    reinterpret_cast<__nv_bfloat16*>(bgen + k * 128 + ((j ^ (k & 7)) << 4))[e] = __float2bfloat16(k == n ? 1.f : 0.f);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<32>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 16, 1, 1);                       // both operands MN-major
    const uint64_t da = make_smem_desc(base + off_rows * 128, lbo_rows * 128, Wh * 128);
    const uint64_t db = make_smem_desc(base + 32768, 8192, 1024);
    umma_bf16(tmem, da, db, idesc, 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v[16];
  tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 16 + e] = v[e];
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<32>(tmem);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  std::vector<float> h(128 * 16);
  const int offs[] = {0, 1, 3, 10, 11, 21};     // tap shift in pixels (dy*Wh + dx)
  const int whs[] = {8, 10, 12};                // halo row pitch in pixels: SBO = Wh * 128
  const int lbos[] = {64, 1, 2, 10, 11};        // distance between the two MN atoms in pixel rows (64 = the usual 8 KB)
  for (int wh : whs)
    for (int lbo : lbos)
      for (int off : offs) {
        cudaMemset(d, 0, 128 * 16 * 4);
        probe<<<1, 128, 65536>>>(off, wh, lbo, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("Wh=%d lbo=%d off=%d CUDA ERROR %s\n", wh, lbo, off, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int k = 0; k < 16; ++k) {
            // K index k = pixel (k/8 rows of the tile, k%8 within the row); MN index m: atom m/64 (second atom = +lbo rows), channel m%64
            const int p = off + (k / 8) * wh + (k % 8) + (m / 64) * lbo;
            if (p >= 256) continue;
            if (h[m * 16 + k] != fval(p, m % 64)) { if (first < 0) first = m * 16 + k; ++bad; }
          }
        printf("Wh=%2d LBO=%2d rows off=%2d : %s (bad=%d first=%d)\n", wh, lbo, off, bad ? "MISMATCH" : "ok", bad, first);
      }
  return 0;
}
