// Device-resident replay store + batch gather (slimdqn/sample_collection/replay_buffer.py:202-230).
//
// The reference keeps snappy-compressed ReplayElements in a host OrderedDict and rebuilds a batch with
// dict lookups + decompress + np.stack + H2D every step.  Here elements live uncompressed in fixed HBM slots
// (1M Atari elements = 56.4 GB of the 180 GB) and `sample` is one coalesced 16-byte-vector gather straight
// into the learner's batch staging: no host round trip on the step path.
#include "common.cuh"

#include <algorithm>

struct GatherArgs {
  const uint8_t* state;
  const uint8_t* next_state;
  const int32_t* action;
  const double* reward;
  const uint8_t* terminal;
  const uint8_t* episode_end;
  const int64_t* slots;
  int64_t state_bytes;
  int n;
  uint8_t* o_state;
  uint8_t* o_next;
  int32_t* o_action;
  float* o_reward_f32;   // learner staging (f32, like the jit boundary of the reference) or null
  double* o_reward_f64;  // host-facing (np.stack of python floats) or null
  uint8_t* o_terminal;
  uint8_t* o_episode_end;  // may be null
};

// grid: (chunks, 2n) — blockIdx.y < n copies state of sample y, otherwise next_state of sample y-n
__global__ void __launch_bounds__(256) replay_gather_kernel(const GatherArgs a) {
  const int which = blockIdx.y >= a.n;
  const int i = blockIdx.y - which * a.n;
  const int64_t slot = a.slots[i];
  const uint8_t* src = (which ? a.next_state : a.state) + slot * a.state_bytes;
  uint8_t* dst = (which ? a.o_next : a.o_state) + (int64_t)i * a.state_bytes;
  if ((a.state_bytes & 15) == 0) {
    const int64_t n16 = a.state_bytes >> 4;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n16; j += (int64_t)gridDim.x * blockDim.x)
      d4[j] = __ldcs(s4 + j);  // streamed once: keep it out of the way of the weights in L2
  } else {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.state_bytes;
         j += (int64_t)gridDim.x * blockDim.x)
      dst[j] = src[j];
  }
  if (blockIdx.x == 0 && which == 0 && threadIdx.x == 0) {
    a.o_action[i] = a.action[slot];
    if (a.o_reward_f32) a.o_reward_f32[i] = __double2float_rn(a.reward[slot]);
    if (a.o_reward_f64) a.o_reward_f64[i] = a.reward[slot];
    a.o_terminal[i] = a.terminal[slot];
    if (a.o_episode_end) a.o_episode_end[i] = a.episode_end[slot];
  }
}

static int ensure_slots(idqn_replay* r, int n) {
  if (n <= r->cap_slots) return IDQN_OK;
  if (r->d_slots) cudaFree(r->d_slots);
  r->cap_slots = std::max(n, 256);
  CK(cudaMalloc(&r->d_slots, sizeof(int64_t) * r->cap_slots));
  return IDQN_OK;
}

static int launch_gather(idqn_replay* r, GatherArgs& a, cudaStream_t st, const int64_t* d_slots) {
  a.state = r->state, a.next_state = r->next_state, a.action = r->action, a.reward = r->reward;
  a.terminal = r->terminal, a.episode_end = r->episode_end, a.slots = d_slots, a.state_bytes = r->state_bytes;
  int64_t per = (r->state_bytes & 15) == 0 ? r->state_bytes / 16 : r->state_bytes;
  int chunks = (int)std::min<int64_t>(std::max<int64_t>((per + 255) / 256, 1), 64);
  replay_gather_kernel<<<dim3(chunks, 2 * a.n), 256, 0, st>>>(a);
  CK(cudaGetLastError());
  return IDQN_OK;
}

extern "C" int idqn_replay_create(int64_t n_slots, int64_t state_bytes, int device, idqn_replay** out) {
  REQUIRE(out && n_slots > 0 && state_bytes > 0, "bad argument");
  CK(cudaSetDevice(device));
  idqn_replay* r = new idqn_replay();
  memset(r, 0, sizeof(*r));
  r->device = device, r->n_slots = n_slots, r->state_bytes = state_bytes;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  size_t need = (size_t)n_slots * (2 * state_bytes + 16);
  if (need > free_b) {
    idqn_set_error("replay store needs %zu bytes of HBM, only %zu free", need, free_b);
    delete r;
    return IDQN_ENOMEM;
  }
  CK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&r->state, (size_t)n_slots * state_bytes));
  CK(cudaMalloc(&r->next_state, (size_t)n_slots * state_bytes));
  CK(cudaMalloc(&r->action, sizeof(int32_t) * n_slots));
  CK(cudaMalloc(&r->reward, sizeof(double) * n_slots));
  CK(cudaMalloc(&r->terminal, n_slots));
  CK(cudaMalloc(&r->episode_end, n_slots));
  int rc = ensure_slots(r, 256);
  if (rc) return rc;
  CK(cudaEventCreateWithFlags(&r->ev_learner, cudaEventDisableTiming));
  r->learner_slots = new std::vector<int64_t>();
  *out = r;
  return IDQN_OK;
}

extern "C" int idqn_replay_destroy(idqn_replay* r) {
  if (!r) return IDQN_OK;
  cudaSetDevice(r->device);
  cudaStreamSynchronize(r->stream);
  if (r->ev_learner) {
    cudaEventSynchronize(r->ev_learner);
    cudaEventDestroy(r->ev_learner);
  }
  delete r->learner_slots;
  void* ptrs[] = {r->state, r->next_state, r->action, r->reward, r->terminal, r->episode_end, r->d_slots, r->d_gslots};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (r->h_stage) cudaFreeHost(r->h_stage);
  cudaStreamDestroy(r->stream);
  delete r;
  return IDQN_OK;
}

extern "C" int idqn_replay_put(idqn_replay* r, int64_t slot, const void* state, const void* next_state, int32_t action,
                               double reward, uint8_t is_terminal, uint8_t episode_end) {
  REQUIRE(r && state && next_state && slot >= 0 && slot < r->n_slots, "bad argument");
  CK(cudaSetDevice(r->device));
  if (r->learner_pending) {
    // a learner step enqueued by idqn_learn_from_replay may still be reading its sampled slots in place
    if (cudaEventQuery(r->ev_learner) == cudaSuccess) r->learner_pending = 0;
    else if (std::find(r->learner_slots->begin(), r->learner_slots->end(), slot) != r->learner_slots->end())
      CK(cudaStreamWaitEvent(r->stream, r->ev_learner, 0));
  }
  CK(cudaMemcpyAsync(r->state + slot * r->state_bytes, state, r->state_bytes, cudaMemcpyHostToDevice, r->stream));
  CK(cudaMemcpyAsync(r->next_state + slot * r->state_bytes, next_state, r->state_bytes, cudaMemcpyHostToDevice,
                     r->stream));
  CK(cudaMemcpyAsync(r->action + slot, &action, sizeof(int32_t), cudaMemcpyHostToDevice, r->stream));
  CK(cudaMemcpyAsync(r->reward + slot, &reward, sizeof(double), cudaMemcpyHostToDevice, r->stream));
  CK(cudaMemcpyAsync(r->terminal + slot, &is_terminal, 1, cudaMemcpyHostToDevice, r->stream));
  CK(cudaMemcpyAsync(r->episode_end + slot, &episode_end, 1, cudaMemcpyHostToDevice, r->stream));
  CK(cudaStreamSynchronize(r->stream));
  return IDQN_OK;
}

extern "C" int idqn_replay_gather_host(idqn_replay* r, const int64_t* slots, int n, void* state, void* next_state,
                                       int32_t* action, double* reward, uint8_t* terminal, uint8_t* episode_end) {
  REQUIRE(r && slots && n > 0 && state && next_state && action && reward && terminal && episode_end, "bad argument");
  for (int i = 0; i < n; ++i) REQUIRE(slots[i] >= 0 && slots[i] < r->n_slots, "slot %lld out of range", (long long)slots[i]);
  CK(cudaSetDevice(r->device));
  if (n > r->cap_gslots) {
    if (r->d_gslots) CK(cudaFree(r->d_gslots));
    r->cap_gslots = std::max(n, 256);
    CK(cudaMalloc(&r->d_gslots, sizeof(int64_t) * r->cap_gslots));
  }
  int rc = IDQN_OK;
  // device-side staging for the packed batch
  const size_t sb = (size_t)n * r->state_bytes;
  const size_t off_next = (sb + 255) / 256 * 256;
  const size_t off_act = off_next + (sb + 255) / 256 * 256;
  const size_t off_rew = off_act + ((size_t)n * 4 + 255) / 256 * 256;
  const size_t off_term = off_rew + ((size_t)n * 8 + 255) / 256 * 256;
  const size_t off_end = off_term + ((size_t)n + 255) / 256 * 256;
  const size_t total = off_end + ((size_t)n + 255) / 256 * 256;
  uint8_t* d_stage = nullptr;
  CK(cudaMallocAsync((void**)&d_stage, total, r->stream));
  CK(cudaMemcpyAsync(r->d_gslots, slots, sizeof(int64_t) * n, cudaMemcpyHostToDevice, r->stream));
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n;
  a.o_state = d_stage, a.o_next = d_stage + off_next, a.o_action = (int32_t*)(d_stage + off_act);
  a.o_reward_f64 = (double*)(d_stage + off_rew), a.o_terminal = d_stage + off_term, a.o_episode_end = d_stage + off_end;
  rc = launch_gather(r, a, r->stream, r->d_gslots);
  if (rc) return rc;
  CK(cudaMemcpyAsync(state, a.o_state, sb, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaMemcpyAsync(next_state, a.o_next, sb, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaMemcpyAsync(action, a.o_action, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaMemcpyAsync(reward, a.o_reward_f64, sizeof(double) * n, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaMemcpyAsync(terminal, a.o_terminal, n, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaMemcpyAsync(episode_end, a.o_episode_end, n, cudaMemcpyDeviceToHost, r->stream));
  CK(cudaFreeAsync(d_stage, r->stream));
  CK(cudaStreamSynchronize(r->stream));
  return IDQN_OK;
}

// the step just enqueued on the learner's stream reads `slots` of the store asynchronously: remember them (and an event
// behind the step) so that idqn_replay_put orders a write into one of them after the read
static int note_learner_read(idqn_handle* h, idqn_replay* r, const int64_t* slots, int n) {
  CK(cudaEventRecord(r->ev_learner, h->stream));
  r->learner_slots->assign(slots, slots + n);
  r->learner_pending = 1;
  return IDQN_OK;
}

extern "C" int idqn_learn_from_replay(idqn_handle* h, idqn_replay* r, const int64_t* slots, int n, int u8,
                                      float* losses) {
  REQUIRE(h && r && slots, "null argument");
  REQUIRE(n == h->B, "sample size %d != learner batch size %d", n, h->B);
  REQUIRE(r->device == h->cfg.device, "replay store and learner live on different devices");
  REQUIRE(r->state_bytes == h->in_elems * (u8 ? 1 : 4), "element size mismatch: store %lld bytes, learner %lld",
          (long long)r->state_bytes, (long long)(h->in_elems * (u8 ? 1 : 4)));
  for (int i = 0; i < n; ++i) REQUIRE(slots[i] >= 0 && slots[i] < r->n_slots, "slot %lld out of range", (long long)slots[i]);
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_slots(r, n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(r->d_slots, slots, sizeof(int64_t) * n, cudaMemcpyHostToDevice, h->stream));
  if (h->img_on && u8) {
    // image path: its first kernel (space-to-depth) reads the frames from the replay slots in place and gathers the
    // scalars; the graph of this staging set is re-captured if the store or its index buffer changed
    if (h->rsrc.state != r->state || h->rsrc.slots != r->d_slots) {
      for (int i = 6; i < 8; ++i)
        if (h->graph[i]) {
          CK(cudaGraphExecDestroy(h->graph[i]));
          h->graph[i] = nullptr;
        }
    }
    h->rsrc.state = r->state, h->rsrc.next_state = r->next_state, h->rsrc.slots = r->d_slots;
    h->rsrc.state_bytes = r->state_bytes, h->rsrc.action = r->action, h->rsrc.reward = r->reward, h->rsrc.terminal = r->terminal;
    h->rsrc_on = 1, h->graph_set = 3;
    rc = idqn_learn_step_resident(h, u8, losses);
    h->rsrc_on = 0, h->graph_set = 0;
    if (rc) return rc;
    return note_learner_read(h, r, slots, n);
  }
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n;
  a.o_state = (uint8_t*)h->s, a.o_next = (uint8_t*)h->s2, a.o_action = h->action, a.o_reward_f32 = h->reward;
  a.o_terminal = h->terminal;
  rc = launch_gather(r, a, h->stream, r->d_slots);
  if (rc) return rc;
  rc = idqn_learn_step_resident(h, u8, losses);
  if (rc) return rc;
  return note_learner_read(h, r, slots, n);
}
