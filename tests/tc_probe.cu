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

// ---- mode "stream": how fast can every SM stream `chunk` byte bulk copies out of an L2-resident
// image of `img_bytes` through a `stages` deep ring (no MMA)?  One thread per CTA.
__global__ void __launch_bounds__(32, 1) stream_kernel(const unsigned char* img, unsigned img_bytes,
                                                       unsigned chunk, int stages, int n_chunks,
                                                       long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  unsigned char* ring = smem + 128;
  if (threadIdx.x != 0) return;
  for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
  mbar_fence_init();
  const unsigned per_cycle = img_bytes / chunk;
  long long c0 = clock64();
  int issued = 0;
  for (; issued < stages && issued < n_chunks; ++issued) {
    mbar_expect_tx(&full[issued % stages], chunk);
    bulk_g2s(ring + (size_t)(issued % stages) * chunk, img + (size_t)(issued % per_cycle) * chunk, chunk,
             &full[issued % stages]);
  }
  for (int q = 0; q < n_chunks; ++q) {
    const int s = q % stages;
    mbar_wait(&full[s], (q / stages) & 1);
    if (issued < n_chunks) {
      mbar_expect_tx(&full[s], chunk);
      bulk_g2s(ring + (size_t)s * chunk, img + (size_t)(issued % per_cycle) * chunk, chunk, &full[s]);
      ++issued;
    }
  }
  cycles[blockIdx.x] = clock64() - c0;
}

// ---- mode "pattern": MMA issue pattern of the forward kernel (6 MMAs per k-step over 3 A terms and
// 3 B blocks, ring of `stages` k-steps in shared memory), timing only.  flags bit0: commit to an
// mbarrier after every k-step; bit1: a second thread streams bulk copies into the ring meanwhile;
// bit2: wait for the k-step commit of step q-2 before issuing step q+1 (the kernel's refill wait).
__global__ void __launch_bounds__(192, 1) pattern_kernel(const unsigned char* img, int NP, int KS, int stages,
                                                         int layers, int flags, long long* cycles) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);   // [0] done, [1..] per-stage commit, [32..] stream
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 1024);
  unsigned char* ring = smem + 2048;
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned block_bytes = NP * 32, stage_bytes = 3 * block_bytes;
  if (tid == 0) {
    for (int i = 0; i < 64; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 5) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 1028);
  if (tid == 0) *stop = 0;
  __syncthreads();
  if (tid == 128 && (flags & 2)) {
    // background streamer: keeps overwriting the ring (timing only, contents irrelevant)
    int q = 0;
    while (!*stop) {
      const int s = q % stages;
      mbar_expect_tx(&bar[32 + s], stage_bytes);
      bulk_g2s(ring + (size_t)s * stage_bytes, img + (size_t)(q % 65) * stage_bytes, stage_bytes, &bar[32 + s]);
      if (q >= 4) { const int qq = q - 4; mbar_wait(&bar[32 + qq % stages], (qq / stages) & 1); }
      ++q;
    }
    for (int qq = q > 4 ? q - 4 : 0; qq < q; ++qq) mbar_wait(&bar[32 + qq % stages], (qq / stages) & 1);
  }
  if (tid == 160) {
    const uint32_t idesc = tc::idesc_bf16_f32(128, NP);
    const uint32_t a1 = tbase + NP, a2 = a1 + 8 * KS, a3 = a2 + 8 * KS;
    long long c0 = clock64();
    unsigned q = 0;
    for (int l = 0; l < layers; ++l)
      for (int j = 0; j < KS; ++j, ++q) {
        const uint32_t sb = smem_u32(ring + (size_t)(q % stages) * stage_bytes);
        const uint64_t b1 = tc::smem_desc(sb, 128, 256), b2 = tc::smem_desc(sb + block_bytes, 128, 256),
                       b3 = tc::smem_desc(sb + 2 * block_bytes, 128, 256);
        const int order = (flags >> 4) & 3;
        if (order == 0) {
          tc::mma_ts(tbase, a1 + 8 * j, b1, idesc, j > 0);
          tc::mma_ts(tbase, a2 + 8 * j, b1, idesc, 1);
          tc::mma_ts(tbase, a3 + 8 * j, b1, idesc, 1);
          tc::mma_ts(tbase, a1 + 8 * j, b2, idesc, 1);
          tc::mma_ts(tbase, a2 + 8 * j, b2, idesc, 1);
          tc::mma_ts(tbase, a1 + 8 * j, b3, idesc, 1);
        } else if (order == 1) {   // same A consecutively
          tc::mma_ts(tbase, a1 + 8 * j, b1, idesc, j > 0);
          tc::mma_ts(tbase, a1 + 8 * j, b2, idesc, 1);
          tc::mma_ts(tbase, a1 + 8 * j, b3, idesc, 1);
          tc::mma_ts(tbase, a2 + 8 * j, b1, idesc, 1);
          tc::mma_ts(tbase, a2 + 8 * j, b2, idesc, 1);
          tc::mma_ts(tbase, a3 + 8 * j, b1, idesc, 1);
        } else if (order == 2) {   // never the same A or B twice in a row
          tc::mma_ts(tbase, a1 + 8 * j, b1, idesc, j > 0);
          tc::mma_ts(tbase, a2 + 8 * j, b2, idesc, 1);
          tc::mma_ts(tbase, a3 + 8 * j, b1, idesc, 1);
          tc::mma_ts(tbase, a1 + 8 * j, b2, idesc, 1);
          tc::mma_ts(tbase, a2 + 8 * j, b1, idesc, 1);
          tc::mma_ts(tbase, a1 + 8 * j, b3, idesc, 1);
        } else {                   // one MMA per k-step only (6x fewer): per-MMA floor
          tc::mma_ts(tbase, a1 + 8 * j, b1, idesc, j > 0);
        }
        if (flags & 1) tc::commit(smem_u32(&bar[1 + q % stages]));
        if ((flags & 4) && q >= 2) { const unsigned qp = q - 2; mbar_wait(&bar[1 + qp % stages], (qp / stages) & 1); }
      }
    tc::commit(smem_u32(&bar[0]));
    mbar_wait(&bar[0], 0);
    cycles[blockIdx.x] = clock64() - c0;
    *stop = 1;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tbase, tc::kTmemCols);
}


// tcgen05.st 32x32b with 16 / 32 / 64 registers (probe only)
#define TC_R4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
__device__ __forceinline__ void st16p(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), TC_R4(v, 0), TC_R4(v, 4), TC_R4(v, 8), TC_R4(v, 12) : "memory");
}
__device__ __forceinline__ void st32p(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               ::"r"(taddr), TC_R4(v, 0), TC_R4(v, 4), TC_R4(v, 8), TC_R4(v, 12), TC_R4(v, 16), TC_R4(v, 20), TC_R4(v, 24), TC_R4(v, 28) : "memory");
}

// ---- mode "tmem": tcgen05.ld / st throughput with `warps` warps (warp w -> lane quarter w % 4)
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, int mode, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(smem_u32(&tmem_slot), tc::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t taddr = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[16];
  for (int i = 0; i < 16; ++i) v[i] = tid + i;
  float acc = 0.f;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 16 + (warp >> 2) * 48) % 496);
    if (mode == 0) {          // load 16 columns, wait, consume
      tc::ld16(taddr + col, v);
      tc::wait_ld();
      acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
    } else if (mode == 1) {   // two loads in flight
      uint32_t w[16];
      tc::ld16(taddr + col, v);
      tc::ld16(taddr + ((col + 16) % 496), w);
      tc::wait_ld();
      acc += __uint_as_float(v[0]) + __uint_as_float(w[15]);
    } else if (mode == 2) {   // store 8 columns
      uint32_t w[8];
      for (int i = 0; i < 8; ++i) w[i] = v[i] + it;
      tc::st8(taddr + col, w);
      if ((it & 7) == 7) tc::wait_st();
    } else if (mode == 3) {   // store 8 columns, back to back
      uint32_t w[8];
      for (int i = 0; i < 8; ++i) w[i] = v[i];
      tc::st8(taddr + col, w);
    } else if (mode == 4) {   // store 16 columns, back to back
      st16p(taddr + col, v);
    } else {                  // store 32 columns, back to back
      uint32_t w[32];
      for (int i = 0; i < 32; ++i) w[i] = v[i & 15];
      st32p(taddr + (col % 480), w);
    }
  }
  tc::wait_st();
  const long long c1 = clock64();
  if ((tid & 31) == 0) cycles[blockIdx.x * 16 + warp] = c1 - c0;
  if (acc == 123.456f) sink[0] = acc;
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_slot, tc::kTmemCols);
}

static int run_tmem(int argc, char** argv) {
  const int warps = argc > 2 ? atoi(argv[2]) : 4;
  const int mode = argc > 3 ? atoi(argv[3]) : 0;
  const int iters = 20000;
  long long* d_c; float* d_s;
  cudaMalloc(&d_c, 148 * 16 * 8); cudaMalloc(&d_s, 4);
  cudaMemset(d_c, 0, 148 * 16 * 8);
  tmem_kernel<<<148, warps * 32, 0>>>(iters, mode, d_c, d_s);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("tmem kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  long long c[16];
  cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < warps; ++i) if (c[i] > mx) mx = c[i];
  const double bytes = (double)warps * iters * 32 * (mode == 5 ? 32 : mode == 4 ? 16 : mode >= 2 ? 8 : (mode == 1 ? 32 : 16)) * 4;
  printf("tmem mode %d warps %d: %.1f cycles per iteration per warp, %.1f B/cycle per SM\n", mode, warps,
         (double)mx / iters, bytes / mx);
  return 0;
}

// ---- mode "mn": D[128 x N] = A[128 x K] . B[N x K]^T with BOTH operands MN-major in shared memory
// (the weight-gradient GEMM: K = samples).  Image of one K = 16 step: [k / 8][mn / 8][k % 8][mn % 8] bf16.
__global__ void __launch_bounds__(192, 1) mn_kernel(const void* a_img, const void* b_img, float* d_out, int N,
                                                    int KS, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b,
                                                    uint32_t sbo_b) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32);
  unsigned char* as = smem + 128;
  const unsigned a_step = 2 * 16 * 128, b_step = 2 * (N / 8) * 128;
  unsigned char* bs = as + KS * a_step;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 128) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    mbar_expect_tx(&bar[0], KS * (a_step + b_step));
    bulk_g2s(as, a_img, KS * a_step, &bar[0]);
    bulk_g2s(bs, b_img, KS * b_step, &bar[0]);
  }
  if (warp == 5) tc::tmem_alloc(smem_u32(tmem_slot), tc::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  if (tid == 160) {
    mbar_wait(&bar[0], 0);
    tc::fence_after_sync();
    const uint32_t idesc = tc::idesc_bf16_f32_mn(128, N);
    for (int ks = 0; ks < KS; ++ks)
      tc::mma_ss(tbase, tc::smem_desc(smem_u32(as) + ks * a_step, lbo_a, sbo_a),
                 tc::smem_desc(smem_u32(bs) + ks * b_step, lbo_b, sbo_b), idesc, ks > 0);
    tc::commit(smem_u32(&bar[1]));
  }
  if (tid < 128) {
    mbar_wait(&bar[1], 0);
    tc::fence_after_sync();
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; c += 16) {
      uint32_t v[16];
      tc::ld16(lane_addr + c, v);
      tc::wait_ld();
      for (int j = 0; j < 16; ++j) d_out[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 5) tc::tmem_dealloc(tbase, tc::kTmemCols);
}

static uint16_t f2bf(float x);
static float bf2f(uint16_t h);

static int run_mn(int argc, char** argv) {
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  const int N = argc > 3 ? atoi(argv[3]) : 208;
  const int KS = 8, K = KS * 16, M = 128;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(2);
  for (auto& x : A) x = bf2f(f2bf((float)rand() / RAND_MAX - 0.5f));
  for (auto& x : B) x = bf2f(f2bf((float)rand() / RAND_MAX - 0.5f));
  auto image = [&](const std::vector<float>& X, int rows) {
    const int groups = rows / 8;
    std::vector<uint16_t> img((size_t)KS * 2 * groups * 64);
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < K; ++k) {
        const size_t e = ((((size_t)(k / 16) * 2 + (k % 16) / 8) * groups + r / 8) * 8 + k % 8) * 8 + r % 8;
        img[e] = f2bf(X[(size_t)r * K + k]);
      }
    return img;
  };
  std::vector<uint16_t> ai = image(A, M), bi = image(B, N);
  uint32_t lbo_a = 16 * 128, sbo_a = 128, lbo_b = (N / 8) * 128, sbo_b = 128;
  if (variant & 1) { uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t; t = lbo_b; lbo_b = sbo_b; sbo_b = t; }
  void *d_a, *d_b; float* d_d;
  cudaMalloc(&d_a, ai.size() * 2); cudaMalloc(&d_b, bi.size() * 2); cudaMalloc(&d_d, (size_t)M * N * 4);
  cudaMemcpy(d_a, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(d_d, 0xFF, (size_t)M * N * 4);
  const size_t smem = 128 + ai.size() * 2 + bi.size() * 2;
  cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mn_kernel<<<1, 192, smem>>>(d_a, d_b, d_d, N, KS, lbo_a, sbo_a, lbo_b, sbo_b);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("mn kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  std::vector<float> D((size_t)M * N);
  cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
      const double err = fabs(ref - (double)D[(size_t)m * N + n]);
      if (!(err <= 1e-3)) ++bad;
      if (err > max_err || err != err) max_err = err;
    }
  printf("mn variant %d N %d: lbo_a %u sbo_a %u lbo_b %u sbo_b %u: max_err %.3e bad %d/%d D[0][0..1] %g %g\n", variant, N,
         lbo_a, sbo_a, lbo_b, sbo_b, max_err, bad, M * N, D[0], D[1]);
  return bad == 0 ? 0 : 1;
}

static int run_pattern(int argc, char** argv) {
  const int flags = argc > 2 ? atoi(argv[2]) : 0;
  const int NP = argc > 3 ? atoi(argv[3]) : 208;
  const int KS = argc > 4 ? atoi(argv[4]) : 12;
  const int stages = argc > 5 ? atoi(argv[5]) : 9;
  const int grid = argc > 6 ? atoi(argv[6]) : 148;
  const int layers = 2000;
  const unsigned stage_bytes = 3 * NP * 32;
  unsigned char* d_img; long long* d_c;
  cudaMalloc(&d_img, 65 * stage_bytes);
  cudaMemset(d_img, 0, 65 * stage_bytes);
  cudaMalloc(&d_c, grid * 8);
  const size_t smem = 2048 + (size_t)stages * stage_bytes;
  cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pattern_kernel<<<grid, 192, smem>>>(d_img, NP, KS, stages, layers, flags, d_c);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("pattern kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  std::vector<long long> c(grid);
  cudaMemcpy(c.data(), d_c, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 62;
  for (auto v : c) { if (v > mx) mx = v; if (v < mn) mn = v; }
  printf("pattern flags %d NP %d KS %d stages %d grid %d: cycles per k-step (6 MMAs) min %.1f max %.1f\n", flags,
         NP, KS, stages, grid, (double)mn / (layers * KS), (double)mx / (layers * KS));
  return 0;
}

static int run_stream(int argc, char** argv) {
  const unsigned chunk = argc > 2 ? atoi(argv[2]) : 19968;
  const int stages = argc > 3 ? atoi(argv[3]) : 9;
  const int grid = argc > 4 ? atoi(argv[4]) : 148;
  const unsigned img_bytes = (argc > 5 ? atoi(argv[5]) : 65) * chunk;
  const int n_chunks = 20000;
  unsigned char* d_img; long long* d_c;
  cudaMalloc(&d_img, img_bytes);
  cudaMemset(d_img, 0, img_bytes);
  cudaMalloc(&d_c, grid * 8);
  const size_t smem = 128 + (size_t)stages * chunk;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    stream_kernel<<<grid, 32, smem>>>(d_img, img_bytes, chunk, stages, n_chunks, d_c);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("stream kernel failed\n"); return 2; }
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(grid);
  cudaMemcpy(c.data(), d_c, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 62;
  for (auto v : c) { if (v > mx) mx = v; if (v < mn) mn = v; }
  printf("stream: chunk %u B x %d stages, grid %d, image %u B: %.3f ms => %.2f TB/s aggregate, "
         "cycles/chunk min %.0f max %.0f\n", chunk, stages, grid, img_bytes, ms,
         (double)grid * n_chunks * chunk / (ms * 1e-3) / 1e12, (double)mn / n_chunks, (double)mx / n_chunks);
  return 0;
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
  if (argc > 1 && !strcmp(argv[1], "stream")) return run_stream(argc, argv);
  if (argc > 1 && !strcmp(argv[1], "pattern")) return run_pattern(argc, argv);
  if (argc > 1 && !strcmp(argv[1], "tmem")) return run_tmem(argc, argv);
  if (argc > 1 && !strcmp(argv[1], "mn")) return run_mn(argc, argv);
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
