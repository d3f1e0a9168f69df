// Neighbour exchange of the head-sharded i-DQN chain over NVLink peer memory (one process per GPU).
//
// Reference semantics: slimdqn/networks/idqn.py:13-24 (shift_params / sync_target_params over the GLOBAL head index)
// at the target events of idqn.py:74-94.  With the K heads split in contiguous blocks over the GPUs of one box the only
// cross-GPU traffic is one head's parameters (16.2 MB for the NatureCNN) per event and boundary:
//   D-sync   target[0] of rank r+1  <-  online[last] of rank r
//   T-shift  online[last] of rank r <-  online[0] of rank r+1 (its value BEFORE that rank's own shift)
// Each rank maps its neighbours' arenas with CUDA IPC and PUSHES its boundary head with its own kernel: 16-byte stores
// straight into the neighbour's slot over NVLink, the arrival flag released by the last CTA of the same kernel.  All
// ordering is on the device -- a "slot is free" flag from the receiver, an "arrived" flag from the sender, both plain
// words in peer-mapped device memory written with st.release.sys / polled with ld.acquire.sys -- so no host thread,
// no NCCL rendezvous and no allocation sits on the event path.  The receiver rebuilds the bf16 operand planes of the one
// head that arrived (half the wire bytes of sending them).
#include "common.cuh"

#include <algorithm>

struct PeerFlags {  // one block per rank, in ITS device memory, written by its neighbours
  unsigned int next_ready_d;    // written by rank r+1: "my target[0] may be overwritten for D-event e"
  unsigned int prev_arrived_d;  // written by rank r-1: "D-event e has landed in your target[0]"
  unsigned int prev_ready_t;    // written by rank r-1: "my online[last] may be overwritten for T-event e"
  unsigned int next_arrived_t;  // written by rank r+1: "T-event e has landed in your online[last]"
  unsigned int done_ctas[2];    // local: CTA tickets of the push kernels (D, T)
  unsigned int pad[2];
};

struct idqn_peer {
  idqn_handle* h;
  int has_prev, has_next;
  PeerFlags* flags;                    // mine
  float* stage;                        // [stride] pre-shift copy of my online[0] (T-shift)
  // peer mappings (cudaIpcOpenMemHandle)
  float* next_target;                  // rank r+1: target arena
  PeerFlags* next_flags;
  float* prev_online;                  // rank r-1: online arena
  PeerFlags* prev_flags;
  int prev_heads;                      // heads held by rank r-1 (its last slot = prev_heads - 1)
  unsigned int epoch_d, epoch_t;       // events executed so far (every rank runs the same schedule)
  cudaStream_t push_stream;            // the push runs NEXT to the in-shard copies of the event
  cudaEvent_t ev_fork, ev_pushed;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// signal: *flag = value (release: everything this stream did before is visible to the peer that acquires it)
__global__ void peer_signal_kernel(unsigned int* flag, unsigned int value) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(flag, value);
  }
}
// wait until *flag >= value (bounded: a protocol bug traps instead of hanging the GPU)
__global__ void peer_wait_kernel(const unsigned int* flag, unsigned int value) {
  if (threadIdx.x == 0) {
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(flag) - value) < 0) {
      __nanosleep(64);
      if (++spins > (1ull << 26)) __trap();
    }
  }
}

// push n16 16-byte vectors src -> dst (peer memory) once *ready >= epoch; the last CTA to finish releases *arrived = epoch
__global__ void __launch_bounds__(512) peer_push_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16,
                                                         const unsigned int* ready, unsigned int* arrived,
                                                         unsigned int* ticket, unsigned int epoch) {
  if (threadIdx.x == 0) {
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(ready) - epoch) < 0) {
      __nanosleep(32);
      if (++spins > (1ull << 26)) __trap();
    }
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    dst[i] = a, dst[i + stride] = b, dst[i + 2 * stride] = c, dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = __ldcs(src + i);
  __threadfence_system();  // this thread's peer stores are performed system-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0;
      __threadfence_system();
      st_release_sys(arrived, epoch);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// what a rank publishes: IPC handles of its online arena, target arena and flag block
struct PeerExport {
  cudaIpcMemHandle_t online, target, flags;
  int32_t heads;
  int32_t pad[3];
};

extern "C" int idqn_peer_export_size(void) { return (int)sizeof(PeerExport); }

extern "C" int idqn_peer_create(idqn_handle* h, idqn_peer** out, void* export_blob) {
  REQUIRE(h && out && export_blob, "null argument");
  CK(cudaSetDevice(h->cfg.device));
  idqn_peer* p = new idqn_peer();
  memset(p, 0, sizeof(*p));
  p->h = h;
  CK(cudaMalloc(&p->flags, sizeof(PeerFlags)));
  CK(cudaMemset(p->flags, 0, sizeof(PeerFlags)));
  CK(cudaMalloc(&p->stage, sizeof(float) * h->stride));
  CK(cudaStreamCreateWithFlags(&p->push_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&p->ev_pushed, cudaEventDisableTiming));
  PeerExport* e = (PeerExport*)export_blob;
  memset(e, 0, sizeof(*e));
  CK(cudaIpcGetMemHandle(&e->online, h->online));
  CK(cudaIpcGetMemHandle(&e->target, h->target));
  CK(cudaIpcGetMemHandle(&e->flags, p->flags));
  e->heads = h->K;
  *out = p;
  return IDQN_OK;
}

// prev_blob / next_blob: the export of rank r-1 / r+1 (null at the ends of the chain)
extern "C" int idqn_peer_connect(idqn_peer* p, const void* prev_blob, const void* next_blob) {
  REQUIRE(p, "null argument");
  CK(cudaSetDevice(p->h->cfg.device));
  if (prev_blob) {
    const PeerExport* e = (const PeerExport*)prev_blob;
    CK(cudaIpcOpenMemHandle((void**)&p->prev_online, e->online, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle((void**)&p->prev_flags, e->flags, cudaIpcMemLazyEnablePeerAccess));
    p->prev_heads = e->heads, p->has_prev = 1;
  }
  if (next_blob) {
    const PeerExport* e = (const PeerExport*)next_blob;
    CK(cudaIpcOpenMemHandle((void**)&p->next_target, e->target, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle((void**)&p->next_flags, e->flags, cudaIpcMemLazyEnablePeerAccess));
    p->has_next = 1;
  }
  return IDQN_OK;
}

extern "C" int idqn_peer_destroy(idqn_peer* p) {
  if (!p) return IDQN_OK;
  cudaSetDevice(p->h->cfg.device);
  cudaStreamSynchronize(p->h->stream);
  void* maps[] = {p->prev_online, p->prev_flags, p->next_target, p->next_flags};
  for (void* m : maps)
    if (m) cudaIpcCloseMemHandle(m);
  if (p->flags) cudaFree(p->flags);
  if (p->stage) cudaFree(p->stage);
  if (p->push_stream) {
    cudaStreamSynchronize(p->push_stream);
    cudaStreamDestroy(p->push_stream);
  }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_pushed) cudaEventDestroy(p->ev_pushed);
  delete p;
  return IDQN_OK;
}

static int push_grid(const idqn_handle* h) { return std::min(h->sm_count, 96); }  // enough CTAs to fill the NVLink write path

// sync_target_params (idqn.py:20-24) over the global head index, enqueued on the learner's stream
extern "C" int idqn_peer_sync_target(idqn_peer* p) {
  REQUIRE(p, "null argument");
  idqn_handle* h = p->h;
  CK(cudaSetDevice(h->cfg.device));
  const unsigned int e = ++p->epoch_d;
  cudaStream_t st = h->stream;
  const int64_t n16 = h->stride / 4;  // the head stride is a multiple of 128 floats
  // my target[0] is free for the left neighbour: everything that read it precedes this signal on the stream.  The
  // signal goes out BEFORE my own push (which may have to wait for the right neighbour), so the pushes of all ranks
  // run concurrently instead of chaining through the box
  if (p->has_prev) {
    peer_signal_kernel<<<1, 32, 0, st>>>(&p->prev_flags->next_ready_d, e);
    CK(cudaGetLastError());
  }
  // my online[last] is final (the step that produced it precedes us on the stream): push it into the right neighbour's
  // target[0] -- on a second stream, next to the in-shard copies below (both only READ the online arena)
  if (p->has_next) {
    const float* src = h->online + (int64_t)(h->K - 1) * h->stride;
    CK(cudaEventRecord(p->ev_fork, st));
    CK(cudaStreamWaitEvent(p->push_stream, p->ev_fork, 0));
    peer_push_kernel<<<push_grid(h), 512, 0, p->push_stream>>>((const uint4*)src, (uint4*)p->next_target, n16,
                                                                &p->flags->next_ready_d, &p->next_flags->prev_arrived_d,
                                                                &p->flags->done_ctas[0], e);
    CK(cudaGetLastError());
    CK(cudaEventRecord(p->ev_pushed, p->push_stream));
  }
  int rc = idqn_sync_target(h);  // in-shard part, target[1:] <- online[:-1], fp32 + planes
  if (rc) return rc;
  if (p->has_next) CK(cudaStreamWaitEvent(st, p->ev_pushed, 0));  // the next step rewrites online[last]
  if (p->has_prev) {
    peer_wait_kernel<<<1, 32, 0, st>>>(&p->flags->prev_arrived_d, e);
    CK(cudaGetLastError());
    rc = idqn_mark_head_planes_dirty(h, IDQN_TARGET, 0);  // rebuilt by the next step's plane refresh, on this stream
    if (rc) return rc;
  }
  return IDQN_OK;
}

// target <- online, then shift_params (idqn.py:75-80, 13-17) over the global head index
extern "C" int idqn_peer_shift_params(idqn_peer* p) {
  REQUIRE(p, "null argument");
  idqn_handle* h = p->h;
  CK(cudaSetDevice(h->cfg.device));
  const unsigned int e = ++p->epoch_t;
  cudaStream_t st = h->stream;
  const int64_t n16 = h->stride / 4;
  int rc = idqn_copy_online_to_target(h);
  if (rc) return rc;
  if (p->has_prev)  // the value my left neighbour needs is my online[0] BEFORE my own shift overwrites it
    CK(cudaMemcpyAsync(p->stage, h->online, sizeof(float) * h->stride, cudaMemcpyDeviceToDevice, st));
  rc = idqn_shift_params(h);  // in-shard part (reads online[last], which becomes free afterwards)
  if (rc) return rc;
  if (p->has_next) {
    peer_signal_kernel<<<1, 32, 0, st>>>(&p->next_flags->prev_ready_t, e);
    CK(cudaGetLastError());
  }
  if (p->has_prev) {
    float* dst = p->prev_online + (int64_t)(p->prev_heads - 1) * h->stride;
    peer_push_kernel<<<push_grid(h), 512, 0, st>>>((const uint4*)p->stage, (uint4*)dst, n16, &p->flags->prev_ready_t,
                                                   &p->prev_flags->next_arrived_t, &p->flags->done_ctas[1], e);
    CK(cudaGetLastError());
  }
  if (p->has_next) {
    peer_wait_kernel<<<1, 32, 0, st>>>(&p->flags->next_arrived_t, e);
    CK(cudaGetLastError());
    rc = idqn_mark_head_planes_dirty(h, IDQN_ONLINE, h->K - 1);
    if (rc) return rc;
  }
  return IDQN_OK;
}
