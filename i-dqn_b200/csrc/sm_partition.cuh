// SM partitions (CUDA green contexts) for the backward pass of the step -- an opt-in experiment (IDQN_F_PARTITION).
//
// After the Dense_0 data gradient the step has two independent pieces of work: the conv backward chain (tensor /
// shared-memory bound, one ~220 KB CTA per SM, grids of <= 112 CTAs) and the Dense_0 wgrad+Adam kernel (HBM bound).
// Their CTAs cannot share an SM (shared memory), and plain stream concurrency serialises them (whichever kernel gets
// the SMs first keeps refilling them; a zero-shared-memory Adam stream co-resides only if its carve-out preference
// and its registers PER SM SUB-PARTITION leave room for the conv CTA).  A green context owns a fixed set of SMs, so
// the two pieces do run side by side: the conv chain on `conv_sms` SMs, the HBM stream on the rest, the remaining
// tiles of the HBM kernel on the whole machine after the join.  Streams created from a green context are ordinary
// cudaStream_t: runtime launches, events and CUDA-graph capture keep the partition (tools/green_probe.cu).
// Measured on B200 (K = 5): 112 | 36 SMs with 60% of the tiles inside the partition: 0.412 ms/step against 0.381 ms
// for the single-stream order -- the HBM kernel is bound by the bytes one SM can keep in flight (~50 GB/s per SM
// with 160 KB of staging), so it scales with the SM count just like the conv chain does, and splitting the machine
// buys nothing.  Kept because the mechanism is correct and cheap; off by default.
// The driver entry points are fetched with cudaGetDriverEntryPoint: the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace smpart {

struct Api {
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
  CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                        unsigned int);
  CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);
  CUresult (*GreenCtxDestroy)(CUgreenCtx);
  bool ok;
};

static inline const Api& api() {
  static Api a = [] {
    Api x;
    memset(&x, 0, sizeof(x));
    struct { const char* name; void** slot; } syms[] = {
        {"cuDeviceGet", (void**)&x.DeviceGet},
        {"cuDeviceGetDevResource", (void**)&x.DeviceGetDevResource},
        {"cuDevSmResourceSplitByCount", (void**)&x.DevSmResourceSplitByCount},
        {"cuDevResourceGenerateDesc", (void**)&x.DevResourceGenerateDesc},
        {"cuGreenCtxCreate", (void**)&x.GreenCtxCreate},
        {"cuGreenCtxStreamCreate", (void**)&x.GreenCtxStreamCreate},
        {"cuGreenCtxDestroy", (void**)&x.GreenCtxDestroy},
    };
    x.ok = true;
    for (auto& s : syms) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q) != cudaSuccess || !*s.slot) x.ok = false;
    }
    return x;
  }();
  return a;
}

struct Partition {
  CUgreenCtx ctx[2];
  cudaStream_t stream[2];  // [0]: conv chain, [1]: HBM stream
  int sms[2];
};

// split the device into `conv_sms` SMs (rounded by the driver to its granularity) and the remainder
static inline bool create(int device, int conv_sms, Partition* out) {
  memset(out, 0, sizeof(*out));
  const Api& a = api();
  if (!a.ok) return false;
  CUdevice dev;
  CUdevResource all, grp, rem;
  if (a.DeviceGet(&dev, device) != CUDA_SUCCESS) return false;
  if (a.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
  if (conv_sms <= 0 || conv_sms >= (int)all.sm.smCount) return false;
  unsigned nb = 1;
  if (a.DevSmResourceSplitByCount(&grp, &nb, &all, &rem, 0, (unsigned)conv_sms) != CUDA_SUCCESS || nb != 1) return false;
  if (rem.sm.smCount == 0) return false;
  CUdevResource* res[2] = {&grp, &rem};
  for (int i = 0; i < 2; ++i) {
    CUdevResourceDesc desc;
    if (a.DevResourceGenerateDesc(&desc, res[i], 1) != CUDA_SUCCESS) return false;
    if (a.GreenCtxCreate(&out->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream st;
    if (a.GreenCtxStreamCreate(&st, out->ctx[i], CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    out->stream[i] = (cudaStream_t)st;
    out->sms[i] = (int)res[i]->sm.smCount;
  }
  return true;
}

static inline void destroy(Partition* p) {
  const Api& a = api();
  for (int i = 0; i < 2; ++i) {
    if (p->stream[i]) cudaStreamDestroy(p->stream[i]);
    if (p->ctx[i] && a.ok) a.GreenCtxDestroy(p->ctx[i]);
  }
  memset(p, 0, sizeof(*p));
}

}  // namespace smpart
