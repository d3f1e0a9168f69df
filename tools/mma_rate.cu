// Issue-to-completion rate of tcgen05.mma (kind::f16, bf16, cta_group::1, M = 128) for the operand layouts the conv /
// dense kernels use: K-major vs MN-major operands, SWIZZLE_128B vs 64B, descriptors whose start address is shifted by
// whole rows (tap shifts), and the MN-major "second M group aliases the first one row apart" trick of conv_wgrad_kernel.
// One CTA per SM is enough: the numbers are cycles per MMA of ONE SM's tensor pipe (operands are zeros in shared memory).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu ; run under `timeout`.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../i-dqn_b200/csrc/tc_core.cuh"
#include "../i-dqn_b200/csrc/tma_core.cuh"

using namespace tc;

struct Cfg {
  const char* name;
  int a_mn, b_mn;        // operand majors (1 = MN-major)
  int N;
  uint32_t a_start;      // byte offset of the A descriptor start inside the A region (row shift * row bytes)
  uint32_t a_lbo, a_sbo, a_lt, a_kstep;  // bytes, layout type, bytes per K = 16 step
  uint32_t b_start, b_lbo, b_sbo, b_lt, b_kstep;
  int ksteps;            // K = 16 steps per "tile" (descriptor offsets wrap after this many)
  int a2;                // 1: a second MMA per step with N/2 columns (the A_lo * B_hi term)
  int alt;               // accumulators the stream of MMAs rotates over (1: every MMA accumulates into the same columns)
};

// KS / A2 / ALT are compile-time so that the timed loop is straight-line code with every descriptor in a register: the
// issuing thread must not be what is measured (a first version with run-time loop bounds and a division per MMA
// measured ~100-170 cycles of ITS OWN overhead per iteration)
template <int KS, int A2, int ALT>
__global__ void __launch_bounds__(128) rate_kernel(const Cfg c, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn != 0, c.b_mn != 0);
    const uint32_t idesc_half = make_idesc_bf16(128, c.N / 2 < 16 ? 16 : c.N / 2, c.a_mn != 0, c.b_mn != 0);
    const uint32_t a_hi32 = tma::desc_hi32(c.a_sbo, c.a_lt), b_hi32 = tma::desc_hi32(c.b_sbo, c.b_lt);
    const uint32_t a0 = tma::desc_lo32(base + c.a_start, c.a_lbo), b0 = tma::desc_lo32(base + 100 * 1024 + c.b_start, c.b_lbo);
    // warm-up
    for (int i = 0; i < 8; ++i) tma::mma_bf16_split<true>(tmem, a0, a_hi32, b0, b_hi32, idesc);
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    tcgen05_after_sync();
    const long long t0 = clock64();
    uint32_t ad[KS], bd[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) ad[j] = a0 + j * (c.a_kstep >> 4), bd[j] = b0 + j * (c.b_kstep >> 4);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        // ALT > 1: the same step for ALT independent accumulators back to back (M tiles sharing a weight tile)
#pragma unroll
        for (int q = 0; q < ALT; ++q) tma::mma_bf16_split<true>(tmem + q * (512 / ALT), ad[j] + q * 1024, a_hi32, bd[j], b_hi32, idesc);
        if (A2) {
#pragma unroll
          for (int q = 0; q < ALT; ++q) tma::mma_bf16_split<true>(tmem + q * (512 / ALT), ad[j] + 64 + q * 1024, a_hi32, bd[j], b_hi32, idesc_half);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 1);
    tcgen05_after_sync();
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  const uint32_t SW128 = tma::LT_SW128, SW64 = tma::LT_SW64;
  std::vector<Cfg> cfgs = {
      // name                                      a_mn b_mn  N   a_start a_lbo a_sbo a_lt a_kstep  b_start b_lbo b_sbo b_lt b_kstep ks a2
      {"A K-major, B K-major, N=64", 0, 0, 64, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0},
      {"A K-major, B K-major, N=128", 0, 0, 128, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0},
      {"A K-major, B K-major, N=256", 0, 0, 256, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0},
      {"A K-major shifted 3 rows, B K-major, N=128", 0, 0, 128, 3 * 128, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0},
      {"A K-major, B MN-major (2 x 64 cols, LBO 8 KB), N=128 [conv fwd]", 0, 1, 128, 0, 16, 1024, SW128, 32, 0, 8192, 1024, SW128, 2048, 4, 0},
      {"A K-major, B MN-major, N=64", 0, 1, 64, 0, 16, 1024, SW128, 32, 0, 8192, 1024, SW128, 2048, 4, 0},
      {"A K-major, B MN-major N=128 + second MMA N=64 [conv fwd pair]", 0, 1, 128, 0, 16, 1024, SW128, 32, 0, 8192, 1024, SW128, 2048, 4, 1},
      {"A K-major, B K-major N=256 + second MMA N=128 [conv dgrad pair]", 0, 0, 256, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 1},
      {"A MN-major (2 x 64 rows, LBO 8 KB), B MN-major, N=128 [wgrad aligned]", 1, 1, 128, 0, 8192, 1024, SW128, 2048, 0, 8192, 1024, SW128, 2048, 8, 0},
      {"A MN-major, start shifted 1 row, B MN-major, N=128", 1, 1, 128, 128, 8192, 1024, SW128, 2048, 0, 8192, 1024, SW128, 2048, 8, 0},
      {"A MN-major, second group 1 row apart (LBO 128), N=128 [wgrad L0/L2 tiles]", 1, 1, 128, 0, 128, 1024, SW128, 2048, 0, 8192, 1024, SW128, 2048, 8, 0},
      {"A MN-major, second group 13 rows apart (LBO 1664), N=128", 1, 1, 128, 0, 1664, 1024, SW128, 2048, 0, 8192, 1024, SW128, 2048, 8, 0},
      {"A MN-major, second group 8 rows apart (LBO 1024), N=128", 1, 1, 128, 0, 1024, 1024, SW128, 2048, 0, 8192, 1024, SW128, 2048, 8, 0},
      {"A MN-major LBO 128, B MN-major SW64 (64-byte rows), N=64 [wgrad L0]", 1, 1, 64, 0, 128, 1024, SW128, 2048, 0, 4096, 512, SW64, 1024, 8, 0},
      {"A MN-major LBO 8 KB, B MN-major SW64, N=64", 1, 1, 64, 0, 8192, 1024, SW128, 2048, 0, 4096, 512, SW64, 1024, 8, 0},
      {"A MN-major LBO 8 KB, B K-major, N=128", 1, 0, 128, 0, 8192, 1024, SW128, 2048, 0, 16, 1024, SW128, 32, 4, 0},
      {"A MN-major LBO 8 KB, B MN-major, N=256", 1, 1, 256, 0, 8192, 1024, SW128, 2048, 0, 8192, 1024, SW128, 2048, 8, 0},
      {"2 accumulators alternating: A K-major, B K-major, N=64 (per MMA)", 0, 0, 64, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0, 2},
      {"2 accumulators alternating: A K-major, B K-major, N=128 (per MMA)", 0, 0, 128, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0, 2},
      {"2 accumulators alternating: A K-major, B K-major, N=256 (per MMA)", 0, 0, 256, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0, 2},
      {"4 accumulators alternating: N=128 (per MMA)", 0, 0, 128, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0, 4},
      {"4 accumulators alternating: N=64 (per MMA)", 0, 0, 64, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 0, 4},
      {"2 accumulators alternating, conv fwd pair (N=128 then N=64) [per step and tile]", 0, 1, 128, 0, 16, 1024, SW128, 32, 0, 8192, 1024, SW128, 2048, 4, 1, 2},
      {"2 accumulators alternating, conv dgrad pair (N=256 then N=128) [per step and tile]", 0, 0, 256, 0, 16, 1024, SW128, 32, 0, 16, 1024, SW128, 32, 4, 1, 2},
  };
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 4);
  const int iters = 64;
  printf("%-78s  clk/MMA-step   (ideal tensor time M128 x N x K16: N/2 clk)\n", "configuration");
  for (const Cfg& c : cfgs) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      const int alt = c.alt > 0 ? c.alt : 1;
#define RUN(KS, A2, ALT)                                                                                         \
  if (c.ksteps == KS && c.a2 == A2 && alt == ALT) {                                                              \
    cudaFuncSetAttribute(rate_kernel<KS, A2, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024); \
    rate_kernel<KS, A2, ALT><<<1, 128, 201 * 1024 + 1024>>>(c, iters, d_out);                                     \
  }
      RUN(4, 0, 1) RUN(4, 1, 1) RUN(8, 0, 1) RUN(8, 1, 1) RUN(4, 0, 2) RUN(4, 1, 2) RUN(4, 0, 4) RUN(8, 0, 2)
#undef RUN
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%-78s  FAILED: %s\n", c.name, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    }
    const double per = (double)h / (iters * c.ksteps * (c.alt > 0 ? c.alt : 1));
    printf("%-78s  %8.1f       (%d%s)\n", c.name, per, c.N / 2, c.a2 ? " + half" : "");
  }
  return 0;
}
