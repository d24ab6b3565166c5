// Hardware probe (B200), prepared for DESIGN.md section 6 / E1: can a SWIZZLE_NONE K-major UMMA A-descriptor read a COMPACT thin-channel
// halo (16 B = 8 bf16 channels per pixel, pixels contiguous) so that
//   * the 8 rows of a core matrix are 8 ADJACENT pixels (row stride 16 B is what the no-swizzle canonical layout prescribes),
//   * the second K core matrix of a K=16 MMA is the NEXT pixel (LBO = 16 B, i.e. core matrices overlap in memory) = the next tap,
//   * 8-row groups step by the halo row pitch (SBO = Wh * 16 B, not a multiple of 128)?
// D = A * I with B = 16x16 identity, so D[m][k] shows which (pixel, channel) the tensor core fetched for GEMM row m, K index k.
// Two passes: values encode the pixel index (<= 255, exact in bf16), then the channel.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe_noswz tools/umma_probe_noswz.cu
#include "../unsupervised_detection_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace cis;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
  return d;                        // layout type (bits 61..63) = 0: SWIZZLE_NONE / interleave
}

__global__ void probe(int off_pix, int lbo_pix, int sbo_pix, int pass, int swap_lbo_sbo, float* out /*[128][16]*/) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const int NP = 512;   // halo pixels, 16 B each
  for (int i = threadIdx.x; i < NP * 8; i += blockDim.x) {
    const int p = i / 8, c = i % 8;
    reinterpret_cast<__nv_bfloat16*>(gen)[i] = __float2bfloat16(pass == 0 ? (float)(p & 255) : (float)c);
  }
  // B = identity [16 n][16 k], K-major no swizzle: core matrix (ng, kh) = 8 n-rows x 16 B at ((kh * 2 + ng) * 128)
  uint8_t* bgen = gen + 16384;
  for (int i = threadIdx.x; i < 16 * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    const int ng = n / 8, nr = n % 8, kh = k / 8, ke = k % 8;
    reinterpret_cast<__nv_bfloat16*>(bgen + (kh * 2 + ng) * 128 + nr * 16)[ke] = __float2bfloat16(n == k ? 1.f : 0.f);
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
    const uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
    const uint32_t lbo = lbo_pix * 16, sbo = sbo_pix * 16;
    const uint64_t da = swap_lbo_sbo ? desc_noswz(base + off_pix * 16, sbo, lbo) : desc_noswz(base + off_pix * 16, lbo, sbo);
    const uint64_t db = swap_lbo_sbo ? desc_noswz(base + 16384, 128, 256) : desc_noswz(base + 16384, 256, 128);
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
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> hp(128 * 16), hc(128 * 16);
  const int offs[] = {0, 1, 5, 13};
  const int lbos[] = {1, 2, 8};          // next tap = +1 pixel (overlapping core matrices), +2, +8 (disjoint, the documented case)
  const int sbos[] = {8, 12, 13, 20};    // halo row pitch in pixels (8 + ex)
  for (int swap = 0; swap < 2; ++swap)
    for (int lbo : lbos)
      for (int sbo : sbos)
        for (int off : offs) {
          for (int pass = 0; pass < 2; ++pass) {
            cudaMemset(d, 0, 128 * 16 * 4);
            probe<<<1, 128, 32768>>>(off, lbo, sbo, pass, swap, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("swap=%d lbo=%d sbo=%d off=%d CUDA ERROR %s\n", swap, lbo, sbo, off, cudaGetErrorString(e)); return 1; }
            cudaMemcpy((pass ? hc : hp).data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost);
          }
          int bad = 0, first = -1;
          for (int m = 0; m < 128; ++m)
            for (int k = 0; k < 16; ++k) {
              const int p = off + (m / 8) * sbo + (m % 8) + (k / 8) * lbo;
              if (hp[m * 16 + k] != (float)(p & 255) || hc[m * 16 + k] != (float)(k % 8)) { if (first < 0) first = m * 16 + k; ++bad; }
            }
          printf("swap_fields=%d LBO=%2d px SBO=%2d px off=%2d : %s (bad=%d", swap, lbo, sbo, off, bad ? "MISMATCH" : "ok", bad);
          if (bad) printf(" first m=%d k=%d got pixel %g ch %g want pixel %d", first / 16, first % 16, hp[first], hc[first],
                          (off + (first / 16 / 8) * sbo + (first / 16 % 8) + (first % 16 / 8) * lbo) & 255);
          printf(")\n");
        }
  return 0;
}
