// Stand-alone probe of the TMA + swizzled UMMA operand layouts the image-resident conv kernels rely on:
//   * cuTensorMapEncodeTiled boxes (SWIZZLE_128B / 64B) landing in shared memory,
//   * tcgen05.mma reading them through K-major and MN-major SW128/SW64 descriptors,
//   * descriptors whose start address is shifted by whole rows (tap shifts) with / without base_offset,
//   * an MN-major M=128 operand whose second 64-row group aliases the first at a one-row offset (LBO = row pitch).
// One CTA, bf16 inputs exactly representable, fp32 accumulate; compares against a double reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu ; run under `timeout`.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../i-dqn_b200/csrc/tc_core.cuh"

using namespace tc;

struct ProbeArgs {
  int x_rows, y_rows;       // rows loaded by TMA (each row = `row_bytes`)
  int row_bytes;            // 128 (SW128) or 64 (SW64)
  int layout_type;          // UMMA layout type: 2 = SW128, 4 = SW64
  int a_mn, b_mn;           // operand majors
  int N;                    // UMMA N
  int nk;                   // number of K=16 MMAs
  uint32_t a_start, a_kstep, a_lbo, a_sbo;  // bytes
  uint32_t b_start, b_kstep, b_lbo, b_sbo;
  int use_base_offset;
};

__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout_type, int use_bo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  if (use_bo) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapX,
                                                    const __grid_constant__ CUtensorMap mapY, const ProbeArgs p,
                                                    float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* xs = smem;
  uint8_t* ys = smem + ((p.x_rows * p.row_bytes + 2047) / 1024) * 1024 + 1024;  // extra slack rows stay zero
  for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  fence_proxy_async_smem();
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_load, (uint32_t)(p.x_rows + p.y_rows) * p.row_bytes);
    tma_load_2d(xs, &mapX, &bar_load, 0, 0);
    tma_load_2d(ys, &mapY, &bar_load, 0, 0);
    mbar_wait(&bar_load, 0);
    tcgen05_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, p.N, p.a_mn, p.b_mn);
    for (int j = 0; j < p.nk; ++j) {
      const uint64_t da = make_desc_sw(smem_u32(xs) + p.a_start + j * p.a_kstep, p.a_lbo, p.a_sbo, p.layout_type, p.use_base_offset);
      const uint64_t db = make_desc_sw(smem_u32(ys) + p.b_start + j * p.b_kstep, p.b_lbo, p.b_sbo, p.layout_type, p.use_base_offset);
      mma_bf16(tmem, da, db, idesc, j > 0);
    }
    mma_commit(&bar_mma);
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tcgen05_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) out[(size_t)tid * p.N + c0 + i] = v[i];
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    exit(1);
  }
  return (EncodeTiled)fn;
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);
}

struct Case {
  const char* name;
  int kind;  // 0: A K-major shifted, B K-major; 1: A K-major shifted, B MN-major; 2: A MN (alias LBO) shifted, B MN
  int sw;    // 128 or 64
  int shift;
  int use_bo;
};

int main() {
  EncodeTiled enc = get_encode();
  const int XR = 176, YR = 64;  // rows in global tensors
  std::vector<Case> cases;
  for (int sw : {128, 64})
    for (int kind : {0, 1, 2})
      for (int shift : {0, 1, 3, 8, 13})
        for (int bo : {0, 1}) cases.push_back({"", kind, sw, shift, bo});
  int bad = 0;
  for (const Case& cs : cases) {
    const int W = cs.sw / 2;  // elements per row
    std::vector<float> X((size_t)XR * W), Y((size_t)YR * W);
    srand(1234);
    for (auto& v : X) v = (float)((rand() % 15) - 7);
    for (auto& v : Y) v = (float)((rand() % 9) - 4) * 0.5f;
    std::vector<uint16_t> Xb(X.size()), Yb(Y.size());
    for (size_t i = 0; i < X.size(); ++i) Xb[i] = f2bf(X[i]);
    for (size_t i = 0; i < Y.size(); ++i) Yb[i] = f2bf(Y[i]);
    uint16_t *dX, *dY;
    float* dOut;
    cudaMalloc(&dX, Xb.size() * 2);
    cudaMalloc(&dY, Yb.size() * 2);
    cudaMemcpy(dX, Xb.data(), Xb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dY, Yb.data(), Yb.size() * 2, cudaMemcpyHostToDevice);
    ProbeArgs p{};
    p.row_bytes = cs.sw;
    p.layout_type = cs.sw == 128 ? 2 : 4;
    p.use_base_offset = cs.use_bo;
    const uint32_t group = 8 * cs.sw;  // 8 rows
    int M_valid = 128, N = W;          // N = elements per Y row (MN-major B) or Y rows (K-major B)
    if (cs.kind == 0) {                // D[m][n] = sum_k X[m+s][k] Y[n][k], K = W
      p.x_rows = 160, p.y_rows = 64;
      p.a_mn = 0, p.b_mn = 0, p.N = 64, N = 64;
      p.nk = W / 16;
      p.a_start = cs.shift * cs.sw, p.a_kstep = 32, p.a_lbo = 16, p.a_sbo = group;
      p.b_start = 0, p.b_kstep = 32, p.b_lbo = 16, p.b_sbo = group;
    } else if (cs.kind == 1) {         // D[m][n] = sum_k X[m+s][k] Y[k][n], K = W rows of Y, N = W
      p.x_rows = 160, p.y_rows = W;
      p.a_mn = 0, p.b_mn = 1, p.N = W;
      p.nk = W / 16;
      p.a_start = cs.shift * cs.sw, p.a_kstep = 32, p.a_lbo = 16, p.a_sbo = group;
      p.b_start = 0, p.b_kstep = 2 * group, p.b_lbo = 16, p.b_sbo = group;  // single N group (N = W = one atom)
    } else {                           // D[m][n] = sum_k X[k+s+(m>=W)][m%W] Y[k][n], K = 64 rows; M groups alias via LBO
      p.x_rows = 96, p.y_rows = 64;
      p.a_mn = 1, p.b_mn = 1, p.N = W;
      p.nk = 4;
      p.a_start = cs.shift * cs.sw, p.a_kstep = 2 * group, p.a_lbo = cs.sw, p.a_sbo = group;
      p.b_start = 0, p.b_kstep = 2 * group, p.b_lbo = 16, p.b_sbo = group;
      M_valid = 2 * W;  // sw128: 128 rows = 2 groups; sw64: 128 rows = 4 groups of 32 -> groups g alias at g rows
    }
    CUtensorMap mx, my;
    cuuint64_t gdx[2] = {(cuuint64_t)W, (cuuint64_t)XR}, gdy[2] = {(cuuint64_t)W, (cuuint64_t)YR};
    cuuint64_t gs[1] = {(cuuint64_t)cs.sw};
    cuuint32_t bx[2] = {(cuuint32_t)W, (cuuint32_t)p.x_rows}, by[2] = {(cuuint32_t)W, (cuuint32_t)p.y_rows};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle swz = cs.sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r1 = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gdx, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&my, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dY, gdy, gs, by, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      printf("encode failed %d %d\n", (int)r1, (int)r2);
      return 1;
    }
    cudaMalloc(&dOut, 128 * 256 * 4);
    cudaMemset(dOut, 0, 128 * 256 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe_kernel<<<1, 128, 98 * 1024>>>(mx, my, p, dOut);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("kind %d sw %d shift %d bo %d: CUDA error %s\n", cs.kind, cs.sw, cs.shift, cs.use_bo, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> out((size_t)128 * p.N);
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int nbad = 0;
    const int K = cs.kind == 2 ? 64 : W;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < p.N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) {
          double a, b;
          if (cs.kind == 0) a = X[(size_t)(m + cs.shift) * W + k], b = Y[(size_t)n * W + k];
          else if (cs.kind == 1) a = X[(size_t)(m + cs.shift) * W + k], b = Y[(size_t)k * W + n];
          else a = X[(size_t)(k + cs.shift + m / W) * W + (m % W)], b = Y[(size_t)k * W + n];
          ref += a * b;
        }
        const double err = fabs(ref - out[(size_t)m * p.N + n]);
        if (err > 1e-3) ++nbad;
        maxerr = std::max(maxerr, err);
      }
    printf("kind %d sw %3d shift %2d base_offset %d : maxerr %.3g bad %d/%d %s\n", cs.kind, cs.sw, cs.shift, cs.use_bo,
           maxerr, nbad, 128 * p.N, nbad ? "FAIL" : "ok");
    bad += nbad != 0;
    cudaFree(dX), cudaFree(dY), cudaFree(dOut);
    (void)M_valid;
    (void)N;
  }
  printf("%d failing configurations\n", bad);
  return 0;
}
