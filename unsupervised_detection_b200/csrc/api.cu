// Error plumbing + version of libcis_b200.so (C ABI declared in include/cis_b200.h).
#include "../../include/cis_b200.h"
#include "common.cuh"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

int cis_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int cis_set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return CIS_ERR_CUDA;
}
int cis_check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cis_set_cuda_error(e, where);
  return CIS_OK;
}
extern "C" const char* cis_last_error(void) { return g_err; }
extern "C" int cis_version(void) { return 100; }

// ---- host CRC-32C, slicing-by-8 (tables built once) ----------------------------------------------------------------------
static uint32_t g_crc_tab[8][256];
static bool g_crc_init = false;
static void crc_init() {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
  g_crc_init = true;
}
extern "C" uint32_t cis_crc32c(uint32_t crc, const void* data, size_t n) {
  if (!g_crc_init) crc_init();
  const unsigned char* p = (const unsigned char*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7)) { c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8); n--; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= c;
    c = g_crc_tab[7][v & 0xff] ^ g_crc_tab[6][(v >> 8) & 0xff] ^ g_crc_tab[5][(v >> 16) & 0xff] ^ g_crc_tab[4][(v >> 24) & 0xff] ^
        g_crc_tab[3][(v >> 32) & 0xff] ^ g_crc_tab[2][(v >> 40) & 0xff] ^ g_crc_tab[1][(v >> 48) & 0xff] ^ g_crc_tab[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = g_crc_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}
