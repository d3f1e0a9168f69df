// Stand-alone probe of the tcgen05 building blocks in i-dqn_b200/csrc/tc_core.cuh: one CTA computes
// C[128 x N] = A[128 x K] * B[K x N] (bf16x3 split, fp32 accumulate in TMEM) for every A/B major combination and
// for both readings of the descriptor's LBO/SBO fields, and reports the error against a double reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu ; run under `timeout`.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../i-dqn_b200/csrc/tc_core.cuh"

using namespace tc;

template <bool A_MN, bool B_MN, int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ C, int K, int swap_lbo_sbo, int nprod) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using AT_K = KMajorTile<128>;
  using AT_MN = MNMajorTile<128>;
  using BT_K = KMajorTile<N>;
  using BT_MN = MNMajorTile<N>;
  constexpr uint32_t A_BYTES = 128 * BK * 2, B_BYTES = N * BK * 2;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + A_BYTES;
  uint8_t* b_hi = a_lo + A_BYTES;
  uint8_t* b_lo = b_hi + B_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, N < 32 ? 32 : N);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
  uint32_t phase = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    // produce A tile
    for (int u = tid; u < 128 * 4; u += 128) {
      float x[8];
      uint32_t off;
      if (!A_MN) {
        int r = u % 128, ku = u / 128;
        for (int i = 0; i < 8; ++i) x[i] = A[(size_t)r * K + k0 + 8 * ku + i];
        off = AT_K::unit_off(r, ku);
      } else {
        int k = u % 32, g = u / 32;
        for (int i = 0; i < 8; ++i) x[i] = A[(size_t)(8 * g + i) * K + k0 + k];
        off = AT_MN::unit_off(k, g);
      }
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(a_hi + off) = hi;
      *reinterpret_cast<uint4*>(a_lo + off) = lo;
    }
    for (int u = tid; u < N * 4; u += 128) {
      float x[8];
      uint32_t off;
      if (!B_MN) {
        int n = u % N, ku = u / N;
        for (int i = 0; i < 8; ++i) x[i] = B[(size_t)(k0 + 8 * ku + i) * N + n];
        off = BT_K::unit_off(n, ku);
      } else {
        int k = u % 32, g = u / 32;
        for (int i = 0; i < 8; ++i) x[i] = B[(size_t)(k0 + k) * N + 8 * g + i];
        off = BT_MN::unit_off(k, g);
      }
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(b_hi + off) = hi;
      *reinterpret_cast<uint4*>(b_lo + off) = lo;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tcgen05_after_sync();
      uint32_t a_lbo = A_MN ? AT_MN::LBO : AT_K::LBO, a_sbo = 128;
      uint32_t b_lbo = B_MN ? BT_MN::LBO : BT_K::LBO, b_sbo = 128;
      if (swap_lbo_sbo) {
        uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t;
        t = b_lbo; b_lbo = b_sbo; b_sbo = t;
      }
      for (int j = 0; j < BK / 16; ++j) {
        uint32_t aoff = A_MN ? AT_MN::k16_off(j) : AT_K::k16_off(j);
        uint32_t boff = B_MN ? BT_MN::k16_off(j) : BT_K::k16_off(j);
        uint64_t dah = make_smem_desc(smem_u32(a_hi) + aoff, a_lbo, a_sbo);
        uint64_t dal = make_smem_desc(smem_u32(a_lo) + aoff, a_lbo, a_sbo);
        uint64_t dbh = make_smem_desc(smem_u32(b_hi) + boff, b_lbo, b_sbo);
        uint64_t dbl = make_smem_desc(smem_u32(b_lo) + boff, b_lbo, b_sbo);
        mma_bf16(tmem, dah, dbh, idesc, (k0 > 0 || j > 0) ? 1u : 0u);
        if (nprod >= 3) {
          mma_bf16(tmem, dah, dbl, idesc, 1u);
          mma_bf16(tmem, dal, dbh, idesc, 1u);
        }
      }
      mma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tcgen05_after_sync();
  }
  // epilogue: warp w owns TMEM lanes [32w, 32w+32)
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) C[(size_t)row * N + c0 + i] = v[i];
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, N < 32 ? 32 : N);
}

template <bool A_MN, bool B_MN, int N>
static int run(int K, int swap, int nprod) {
  std::vector<float> A(128 * K), B((size_t)K * N), C(128 * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 4), cudaMalloc(&dB, B.size() * 4), cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0xff, C.size() * 4);
  size_t smem = 2 * (128 * BK * 2) + 2 * (N * BK * 2) + 1024;
  cudaFuncSetAttribute(probe_kernel<A_MN, B_MN, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<A_MN, B_MN, N><<<1, 128, smem>>>(dA, dB, dC, K, swap, nprod);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("A_%s B_%s N=%d K=%d swap=%d nprod=%d : CUDA ERROR %s\n", A_MN ? "MN" : "K", B_MN ? "MN" : "K", N, K, swap,
           nprod, cudaGetErrorString(e));
    return 2;
  }
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * B[(size_t)k * N + n];
      maxerr = fmax(maxerr, fabs(r - C[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(r));
    }
  printf("A_%s B_%s N=%d K=%d swap=%d nprod=%d : max|err|=%.3e max|ref|=%.3e rel=%.3e %s\n", A_MN ? "MN" : "K",
         B_MN ? "MN" : "K", N, K, swap, nprod, maxerr, maxref, maxerr / maxref, maxerr / maxref < 1e-4 ? "OK" : "WRONG");
  cudaFree(dA), cudaFree(dB), cudaFree(dC);
  return maxerr / maxref < 1e-4 ? 0 : 1;
}

int main(int argc, char** argv) {
  int combo = argc > 1 ? atoi(argv[1]) : 0, swap = argc > 2 ? atoi(argv[2]) : 0, nprod = argc > 3 ? atoi(argv[3]) : 3;
  int K = argc > 4 ? atoi(argv[4]) : 64;
  switch (combo) {
    case 0: return run<false, false, 64>(K, swap, nprod);
    case 1: return run<false, true, 64>(K, swap, nprod);
    case 2: return run<true, false, 64>(K, swap, nprod);
    case 3: return run<true, true, 64>(K, swap, nprod);
    case 4: return run<false, true, 32>(K, swap, nprod);
    case 5: return run<false, true, 160>(K, swap, nprod);
    case 6: return run<true, true, 256>(K, swap, nprod);
    case 7: return run<false, false, 32>(K, swap, nprod);
  }
  return 3;
}
