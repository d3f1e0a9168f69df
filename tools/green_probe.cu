// Probe: CUDA green contexts (SM partitions) driven through the runtime API and captured into a CUDA graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/green_probe tools/green_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
#define CU(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { const char* s; cuGetErrorString(r, &s); printf("CU error %s at %s:%d\n", s, __FILE__, __LINE__); return 1; } } while (0)
#define RT(x) do { cudaError_t r = (x); if (r != cudaSuccess) { printf("RT error %s at %s:%d\n", cudaGetErrorString(r), __FILE__, __LINE__); return 1; } } while (0)

__global__ void spin(int* hist, long long cycles) {
  extern __shared__ char sm[];
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) atomicAdd(&hist[smid], 1);
  long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

static int report(const char* tag, int* d_hist, int n) {
  std::vector<int> h(n);
  cudaMemcpy(h.data(), d_hist, n * sizeof(int), cudaMemcpyDeviceToHost);
  int used = 0, lo = 1 << 30, hi = -1;
  for (int i = 0; i < n; ++i) if (h[i]) { ++used; if (i < lo) lo = i; if (i > hi) hi = i; }
  printf("%s: SMs used %d (smid %d..%d)\n", tag, used, lo, hi);
  cudaMemset(d_hist, 0, n * sizeof(int));
  return 0;
}

int main() {
  RT(cudaSetDevice(0));
  RT(cudaFree(0));
  CUdevice dev; CU(cuDeviceGet(&dev, 0));
  CUdevResource all; CU(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs %u\n", all.sm.smCount);
  for (unsigned want : {112u, 104u, 96u}) {
    CUdevResource grp, rem; unsigned nb = 1;
    CUresult r = cuDevSmResourceSplitByCount(&grp, &nb, &all, &rem, 0, want);
    printf("split want %u -> rc %d groups %u grp %u rem %u\n", want, (int)r, nb, grp.sm.smCount, rem.sm.smCount);
  }
  CUdevResource grp, rem; unsigned nb = 1;
  CU(cuDevSmResourceSplitByCount(&grp, &nb, &all, &rem, 0, 112));
  CUdevResourceDesc dA, dB;
  CU(cuDevResourceGenerateDesc(&dA, &grp, 1));
  CU(cuDevResourceGenerateDesc(&dB, &rem, 1));
  CUgreenCtx gA, gB;
  CU(cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CU(cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sA, sB;
  CU(cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
  CU(cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
  cudaStream_t main_s; RT(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
  int* hist; RT(cudaMalloc(&hist, 256 * sizeof(int))); RT(cudaMemset(hist, 0, 256 * sizeof(int)));
  RT(cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const long long cyc = 200000;  // ~100 us
  // 1. runtime launches on green streams
  spin<<<400, 128, 200 * 1024, sA>>>(hist, 1000); RT(cudaGetLastError()); RT(cudaStreamSynchronize(sA)); report("green A (112 asked)", hist, 256);
  spin<<<400, 128, 200 * 1024, sB>>>(hist, 1000); RT(cudaGetLastError()); RT(cudaStreamSynchronize(sB)); report("green B (remainder)", hist, 256);
  spin<<<400, 128, 200 * 1024, main_s>>>(hist, 1000); RT(cudaGetLastError()); RT(cudaStreamSynchronize(main_s)); report("primary", hist, 256);
  // 2. concurrency, direct: A and B each one wave of 100 us
  cudaEvent_t e0, e1, ef, ja, jb; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventCreateWithFlags(&ef, cudaEventDisableTiming); cudaEventCreateWithFlags(&ja, cudaEventDisableTiming); cudaEventCreateWithFlags(&jb, cudaEventDisableTiming);
  auto enqueue = [&]() {
    cudaEventRecord(ef, main_s);
    cudaStreamWaitEvent(sA, ef, 0); cudaStreamWaitEvent(sB, ef, 0);
    spin<<<grp.sm.smCount, 128, 200 * 1024, sA>>>(hist, cyc);
    spin<<<rem.sm.smCount, 128, 200 * 1024, sB>>>(hist, cyc);
    cudaEventRecord(ja, sA); cudaEventRecord(jb, sB);
    cudaStreamWaitEvent(main_s, ja, 0); cudaStreamWaitEvent(main_s, jb, 0);
    spin<<<148, 128, 200 * 1024, main_s>>>(hist, cyc);
  };
  for (int rep = 0; rep < 2; ++rep) {
    RT(cudaEventRecord(e0, main_s));
    enqueue();
    RT(cudaEventRecord(e1, main_s)); RT(cudaStreamSynchronize(main_s));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("direct fork/join: %.1f us (serial would be ~%.0f, overlapped ~%.0f)\n", ms * 1e3, 3 * cyc / 1.9e3, 2 * cyc / 1.9e3);
  }
  report("direct", hist, 256);
  // 3. the same captured into a graph
  cudaGraph_t g; cudaGraphExec_t ge;
  RT(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal));
  enqueue();
  RT(cudaStreamEndCapture(main_s, &g));
  RT(cudaGraphInstantiate(&ge, g, 0));
  for (int rep = 0; rep < 3; ++rep) {
    RT(cudaEventRecord(e0, main_s));
    RT(cudaGraphLaunch(ge, main_s));
    RT(cudaEventRecord(e1, main_s)); RT(cudaStreamSynchronize(main_s));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("graph fork/join: %.1f us\n", ms * 1e3);
  }
  report("graph", hist, 256);
  // 4. does the partition hold inside the graph?  only the A kernel
  RT(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal));
  cudaEventRecord(ef, main_s); cudaStreamWaitEvent(sA, ef, 0);
  spin<<<400, 128, 200 * 1024, sA>>>(hist, 1000);
  cudaEventRecord(ja, sA); cudaStreamWaitEvent(main_s, ja, 0);
  RT(cudaStreamEndCapture(main_s, &g));
  RT(cudaGraphInstantiate(&ge, g, 0));
  RT(cudaGraphLaunch(ge, main_s)); RT(cudaStreamSynchronize(main_s));
  report("graph, A kernel only", hist, 256);
  printf("done\n");
  return 0;
}
