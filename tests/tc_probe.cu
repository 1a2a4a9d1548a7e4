// tc_probe.cu -- bring-up probe of the tcgen05 building blocks in csrc/ikr_tc.cuh (test
// infrastructure, run on a B200):  D[128 x N] = A[128 x K] . B[N x K]^T with A (bf16) written to TMEM
// by tcgen05.st, B (bf16) staged in shared memory in the no-swizzle K-major core-matrix layout by one
// bulk copy, K / 16 single-thread MMAs, D read back by tcgen05.ld and compared with a host reference.
// Usage: tc_probe [variant] [N] [KS]     variant bit0: swap LBO/SBO, bit1: swap A half-words,
//                                         bit2: k-half-major B image (LBO = N * 16, SBO = 128)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/_tc_probe tests/tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../neural-ode-ion-channels_b200/csrc/ikr_device.cuh"
#include "../neural-ode-ion-channels_b200/csrc/ikr_tc.cuh"

using namespace ikr;

struct ProbeParams {
  const uint32_t* a_words;  // [128][KS * 8]
  const void* b_img;        // KS k-steps x (N x 32 bytes)
  float* d_out;             // [128][N]
  int KS, N;
  uint32_t lbo, sbo, idesc, kstep_bytes;
  int iters;                // repeat the MMA sequence (timing)
  long long* cycles;
};

__global__ void __launch_bounds__(192, 1) probe_kernel(const ProbeParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);          // [0] B landed, [1] D ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32);
  unsigned char* bs = smem + 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 128) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    mbar_expect_tx(&bar[0], (unsigned)p.KS * p.kstep_bytes);
    bulk_g2s(bs, p.b_img, (unsigned)p.KS * p.kstep_bytes, &bar[0]);
  }
  if (warp == 5) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const uint32_t a_col = (uint32_t)((p.N + 15) / 16 * 16);
  if (tid < 128) {
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int ks = 0; ks < p.KS; ++ks) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = p.a_words[(size_t)tid * p.KS * 8 + ks * 8 + j];
      tc::st8(lane_addr + a_col + ks * 8, v);
    }
    tc::wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 160) {
    mbar_wait(&bar[0], 0);
    tc::fence_after_sync();
    long long c0 = clock64();
    for (int it = 0; it < p.iters; ++it)
      for (int ks = 0; ks < p.KS; ++ks) {
        const uint64_t bd = tc::smem_desc(smem_u32(bs) + ks * p.kstep_bytes, p.lbo, p.sbo);
        tc::mma_ts(tbase, tbase + a_col + ks * 8, bd, p.idesc, ks > 0 ? 1u : 0u);
      }
    tc::commit(smem_u32(&bar[1]));
    mbar_wait(&bar[1], 0);
    long long c1 = clock64();
    if (p.cycles) p.cycles[blockIdx.x] = c1 - c0;
  }
  if (tid < 128) {
    mbar_wait(&bar[1], 0);
    tc::fence_after_sync();
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < p.N; c += 16) {
      uint32_t v[16];
      tc::ld16(lane_addr + c, v);
      tc::wait_ld();
      for (int j = 0; j < 16; ++j) p.d_out[(size_t)tid * p.N + c + j] = __uint_as_float(v[j]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

static uint16_t f2bf(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float x;
  memcpy(&x, &u, 4);
  return x;
}

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e = (x);                                                            \
    if (e != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = argc > 2 ? atoi(argv[2]) : 208;
  const int KS = argc > 3 ? atoi(argv[3]) : 13;
  const int iters = argc > 4 ? atoi(argv[4]) : 1;
  const int K = KS * 16, M = 128;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(1);
  for (auto& x : A) x = bf2f(f2bf((float)rand() / RAND_MAX - 0.5f));
  for (auto& x : B) x = bf2f(f2bf((float)rand() / RAND_MAX - 0.5f));
  std::vector<uint32_t> aw((size_t)M * KS * 8);
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < KS * 8; ++j) {
      uint16_t e = f2bf(A[(size_t)m * K + 2 * j]), o = f2bf(A[(size_t)m * K + 2 * j + 1]);
      aw[(size_t)m * KS * 8 + j] = (variant & 2) ? ((uint32_t)e << 16 | o) : ((uint32_t)o << 16 | e);
    }
  const uint32_t kstep_bytes = (uint32_t)N * 32;
  std::vector<uint16_t> bimg((size_t)KS * N * 16);
  uint32_t lbo, sbo;
  for (int ks = 0; ks < KS; ++ks)
    for (int n = 0; n < N; ++n)
      for (int kk = 0; kk < 16; ++kk) {
        size_t byte;
        if (variant & 4) byte = (size_t)(kk / 8) * (N * 16) + (size_t)(n / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2;
        else byte = (size_t)(n / 8) * 256 + (size_t)(kk / 8) * 128 + (n % 8) * 16 + (kk % 8) * 2;
        bimg[((size_t)ks * kstep_bytes + byte) / 2] = f2bf(B[(size_t)n * K + ks * 16 + kk]);
      }
  if (variant & 4) { lbo = (uint32_t)N * 16; sbo = 128; } else { lbo = 128; sbo = 256; }
  if (variant & 1) { uint32_t t = lbo; lbo = sbo; sbo = t; }

  ProbeParams p;
  uint32_t* d_a; void* d_b; float* d_d; long long* d_c;
  CK(cudaMalloc(&d_a, aw.size() * 4));
  CK(cudaMalloc(&d_b, bimg.size() * 2));
  CK(cudaMalloc(&d_d, (size_t)M * N * 4));
  CK(cudaMalloc(&d_c, 8));
  CK(cudaMemcpy(d_a, aw.data(), aw.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, bimg.data(), bimg.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_d, 0xFF, (size_t)M * N * 4));
  p.a_words = d_a; p.b_img = d_b; p.d_out = d_d; p.KS = KS; p.N = N;
  p.lbo = lbo; p.sbo = sbo; p.idesc = tc::idesc_bf16_f32(128, N); p.kstep_bytes = kstep_bytes;
  p.iters = iters; p.cycles = d_c;
  const size_t smem = 128 + (size_t)KS * kstep_bytes;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 192, smem>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)M * N);
  long long cyc = 0;
  CK(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, d_c, 8, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
      double err = fabs(ref - (double)D[(size_t)m * N + n]);
      if (!(err <= 1e-3)) ++bad;
      if (err > max_err || err != err) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
  printf("variant %d N %d KS %d idesc 0x%08x lbo %u sbo %u: max_err %.3e (max |ref| %.3f) bad %d/%d  "
         "cycles %lld (%d MMAs => %.1f cyc/MMA)  D[0][0..3] = %g %g %g %g\n",
         variant, N, KS, p.idesc, lbo, sbo, max_err, max_ref, bad, M * N, cyc, iters * KS,
         (double)cyc / (iters * KS), D[0], D[1], D[2], D[3]);
  return bad == 0 ? 0 : 1;
}
