// Weight-streaming tcgen05 kernels for the big Dense layer (Dense_0: 7744 x 512) at batch 32 (sm_100a).
//
// At batch 32 the layer is bound by reading its 15.9 MB kernel once per net: the weights are the M-side operand
// (128 rows per tile), the batch sits on N.  The kernel reads the fp32 MASTER weights -- there are no bf16 copies of
// this layer in HBM (they cost 79 MB of extra writes per K = 5 step in the wgrad+Adam kernel and doubled every target
// event): a persistent CTA streams 64-deep K blocks of fp32 weights and of the small bf16 batch operand through a TMA
// ring, four converter warps split every fp32 tile into bf16 hi/lo operand tiles IN SHARED MEMORY, written straight in
// the SWIZZLE_128B layout tcgen05.mma expects, and per K = 16 step
//     W_hi * [x_hi | x_lo]   (one MMA, N = 64: the batch planes sit next to each other in the stage)
//   + W_lo * x_hi            (one MMA, N = 32)
// accumulate in TMEM ("bf16x3" split, fp32-faithful to 2^-17); the epilogue adds the two column sets.
//   MODE 0  forward   y[b][o]  = relu(sum_i x[b][i] W[i][o] + bias[o])      A = W^T: MN-major, rows = i, M = o
//           split-K over units; the partial tiles are summed in split order by the consumer (head_q_kernel)
//   MODE 1  dgrad     dx[b][i] = relu'(x[b][i]) sum_o dy[b][o] W[i][o]       A = W:   K-major, rows = i = M, K = o
// Warp roles: 0 = TMA producer, 1 = MMA issue, 2..5 = epilogue, 6..9 = fp32 -> bf16 hi/lo converters.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_core.cuh"
#include "tma_core.cuh"
#include "conv_img.cuh"  // tl_stamp

namespace dense {
using img::tl_stamp;
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int NTHREADS = 320;  // warp 0: TMA, warp 1: MMA, warps 2..5: epilogue, warps 6..9: converters
constexpr int CONV_T0 = 192, CONV_THREADS = 128;
constexpr int BKD = 64;        // K depth of a stage
constexpr int NB = 32;         // batch columns
constexpr uint32_t W_BYTES = 128 * BKD * 4;   // fp32 weight tile of a stage: 32 KB
constexpr uint32_t B_BYTES = NB * BKD * 2;    // one plane of the batch tile: 4 KB
constexpr uint32_t STAGE_BYTES = W_BYTES + 2 * B_BYTES;  // 40 KB
constexpr uint32_t OP_PLANE = 128 * BKD * 2;  // one bf16 plane of the converted weight tile: 16 KB
constexpr uint32_t OP_BYTES = 2 * OP_PLANE;   // hi + lo
constexpr int N_OP = 2;                       // converted operand buffers

struct Args {
  int n_units, tiles, splits, kb_per_unit;  // unit = (net, tile, split); kb_per_unit K blocks each
  int unit0;                                // first unit of this launch (best_action: the units of one net)
  int nets;
  int heads;            // K: nets < heads read the online arena, the others the target arena (forward)
  int stages;
  // forward
  int I, O;
  NetPtr w;             // fp32 arenas (bias)
  int64_t b_off;
  float* y;             // [nets][ystride]  (fwd)  /  dx [heads][xstride] (dgrad)
  bf16 *yh, *yl;        // planes of the output
  int64_t ystride;
  const float* xact;    // dgrad: layer input (relu output), same indexing as dx; or
  const bf16* xmask_hi; // its bf16 hi plane (hi > 0 <=> x > 0)
  float* part;          // fwd: [nets*tiles][splits][32][128]
  int* tickets;         // fwd: [nets*tiles]
  // L2 policy of the weight loads: nets < keep_heads are kept (evict_last: the data gradient re-reads them ~20 us after the
  // forward), every other net's weights are streamed evict_first
  int keep_heads;
  // dgrad planes destination: dyZ layout of the preceding conv layer (zP > 0) or plain
  int zP, zW, zC, zOff;
  int64_t zRows, zstride;
  int debug;
  int tl_id;
};

struct Smem {
  uint32_t op_off, bar_off, total;
};
__host__ __device__ inline Smem smem_layout(const Args& p) {
  Smem s;
  s.op_off = p.stages * STAGE_BYTES;
  s.bar_off = s.op_off + N_OP * OP_BYTES;
  s.total = s.bar_off + 256 + 1024;
  return s;
}

__device__ __forceinline__ void conv_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
dense_stream_kernel(const __grid_constant__ CUtensorMap mapW_on, const __grid_constant__ CUtensorMap mapW_tg,
                    const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                    const Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const Smem L = smem_layout(p);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full = bars;             // [stages]  TMA landed
  uint64_t* empty = bars + 8;        // [stages]  count 2: converters have read the fp32 tile + MMAs have read the batch tile
  uint64_t* acc_full = bars + 16;    // [2]
  uint64_t* acc_empty = bars + 18;   // [2]
  uint64_t* op_full = bars + 20;     // [N_OP]  converted operand written
  uint64_t* op_empty = bars + 22;    // [N_OP]  MMAs reading it complete
  __shared__ uint32_t tmem_base_s;

  pdl_trigger();
  ktl_begin(p.tl_id);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < p.stages; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 2);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 4);
    for (int i = 0; i < N_OP; ++i) mbar_init(&op_full[i], 1), mbar_init(&op_empty[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 128);  // 2 accumulator buffers x 64 columns
  fence_proxy_async_smem();
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      // The weights do not depend on the kernels before this one in the step (only the batch operand does): the weight
      // boxes of the first `stages` K blocks are requested BEFORE griddepcontrol.wait, i.e. while the preceding kernel
      // (conv2 / the loss kernels) is still running -- the HBM stream starts a few microseconds early.
      int st = 0, ui = 0, blk = 0;
      uint32_t ph = 0;
      bool waited = false;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ui) {
        const int uu = u + p.unit0;
        const int sp = uu % p.splits, nt = uu / p.splits, tile = nt % p.tiles, net = nt / p.tiles;
        const CUtensorMap* mapW = net < p.heads ? &mapW_on : &mapW_tg;
        const int head = net < p.heads ? net : net - p.heads;
        for (int kb = 0; kb < p.kb_per_unit; ++kb, ++blk) {
          if (!waited && blk >= p.stages) {
            // every stage holds its weight tile: now the batch operands of those stages, then the ordinary pipeline
            pdl_wait();
            waited = true;
            int u2 = blockIdx.x, kb2 = 0;
            for (int b2 = 0; b2 < blk; ++b2) {
              const int uu2 = u2 + p.unit0, sp2 = uu2 % p.splits, net2 = (uu2 / p.splits) / p.tiles;
              const int k02 = (sp2 * p.kb_per_unit + kb2) * BKD;
              const uint32_t s2 = base + b2 * STAGE_BYTES;
              tma::load_3d(s2 + W_BYTES, &mapB_hi, &full[b2], k02, 0, net2);
              tma::load_3d(s2 + W_BYTES + B_BYTES, &mapB_lo, &full[b2], k02, 0, net2);
              if (++kb2 == p.kb_per_unit) kb2 = 0, u2 += gridDim.x;
            }
          }
          mbar_wait(&empty[st], ph ^ 1);
          tl_stamp(p.debug, 2000 + ui * 16 + kb);
          tma::expect_tx(&full[st], STAGE_BYTES);
          const uint32_t s0 = base + st * STAGE_BYTES;
          const int k0 = (sp * p.kb_per_unit + kb) * BKD;
          const uint64_t pol = net < p.keep_heads ? tma::L2_EVICT_LAST : tma::L2_EVICT_FIRST;
          if (MODE == 0) tma::load_3d_hint(s0, mapW, &full[st], tile * 128, k0, head, pol);  // [64 i][128 o] fp32
          else tma::load_3d_hint(s0, mapW, &full[st], k0, tile * 128, head, pol);            // [128 i][64 o] fp32
          if (waited) {
            tma::load_3d(s0 + W_BYTES, &mapB_hi, &full[st], k0, 0, net);
            tma::load_3d(s0 + W_BYTES + B_BYTES, &mapB_lo, &full[st], k0, 0, net);
          }
          if (++st == p.stages) st = 0, ph ^= 1;
        }
      }
      if (!waited) {  // fewer K blocks than stages in this CTA: the batch operands of all of them
        pdl_wait();
        int u2 = blockIdx.x, kb2 = 0;
        for (int b2 = 0; b2 < blk; ++b2) {
          const int uu2 = u2 + p.unit0, sp2 = uu2 % p.splits, net2 = (uu2 / p.splits) / p.tiles;
          const int k02 = (sp2 * p.kb_per_unit + kb2) * BKD;
          const uint32_t s2 = base + b2 * STAGE_BYTES;
          tma::load_3d(s2 + W_BYTES, &mapB_hi, &full[b2], k02, 0, net2);
          tma::load_3d(s2 + W_BYTES + B_BYTES, &mapB_lo, &full[b2], k02, 0, net2);
          if (++kb2 == p.kb_per_unit) kb2 = 0, u2 += gridDim.x;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc2 = make_idesc_bf16(128, 2 * NB, MODE == 0, false);
      const uint32_t idesc1 = make_idesc_bf16(128, NB, MODE == 0, false);
      const uint32_t hi32 = tma::desc_hi32(1024, tma::LT_SW128);
      const uint32_t astep = MODE == 0 ? 128u : 2u;    // K = 16: 16 rows of 128 B (MN-major) or 32 bytes (K-major)
      const uint32_t a_lbo = MODE == 0 ? 8192u : 16u;  // MN-major: second 64-wide o group; K-major: unused
      int st = 0, ai = 0, ob = 0;
      uint32_t oph = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ai) {
        const int ab = ai & 1;
        mbar_wait(&acc_empty[ab], ((ai >> 1) & 1) ^ 1);
        const uint32_t d = tmem + (uint32_t)ab * 2 * NB;
        for (int kb = 0; kb < p.kb_per_unit; ++kb) {
          mbar_wait(&op_full[ob], oph);  // implies full[st]: the converters waited for the stage before writing
          tcgen05_after_sync();
          tl_stamp(p.debug, 3000 + ai * 32 + kb);
          const uint32_t a0 = base + L.op_off + ob * OP_BYTES;
          const uint32_t ah = tma::desc_lo32(a0, a_lbo), al = tma::desc_lo32(a0 + OP_PLANE, a_lbo);
          const uint32_t bh = tma::desc_lo32(base + st * STAGE_BYTES + W_BYTES, 16);
#pragma unroll
          for (int j = 0; j < BKD / 16; ++j) {
            if (kb == 0 && j == 0) tma::mma_bf16_split<false>(d, ah, hi32, bh, hi32, idesc2);
            else tma::mma_bf16_split<true>(d, ah + j * astep, hi32, bh + 2 * j, hi32, idesc2);
            tma::mma_bf16_split<true>(d, al + j * astep, hi32, bh + 2 * j, hi32, idesc1);
          }
          mma_commit(&op_empty[ob]);
          mma_commit(&empty[st]);
          tl_stamp(p.debug, 3000 + ai * 32 + 16 + kb);
          if (++st == p.stages) st = 0;
          if (++ob == N_OP) ob = 0, oph ^= 1;
        }
        mma_commit(&acc_full[ab]);
      }
    }
  } else if (warp >= 6) {
    // ===== converters: fp32 weight tile -> bf16 hi / lo operand tiles in the SWIZZLE_128B UMMA layout =====
    // MODE 0: tile [64 rows i][128 o] (512-byte rows); operand = two 64-wide o groups (8 KB apart) of [64 rows i][64 o]
    // MODE 1: tile [128 rows i][64 o] (256-byte rows); operand = [128 rows i][64 o]
    // a thread converts 4 consecutive floats: one conflict-free LDS.128, one 8-byte store per plane (half a 16-byte chunk;
    // chunk c of row r sits at c ^ (r & 7), the swizzle TMA would have applied)
    const int ct = tid - CONV_T0;
    constexpr int F4_PER_ROW = MODE == 0 ? 32 : 16;
    int st = 0, ob = 0;
    uint32_t ph = 0, oph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      for (int kb = 0; kb < p.kb_per_unit; ++kb) {
        mbar_wait(&full[st], ph);
        mbar_wait(&op_empty[ob], oph ^ 1);
        const float4* src = reinterpret_cast<const float4*>(smem + st * STAGE_BYTES);
        uint8_t* dst = smem + L.op_off + ob * OP_BYTES;
        float4 v[16];
#pragma unroll
        for (int it = 0; it < 16; ++it) v[it] = src[it * CONV_THREADS + ct];
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int idx = it * CONV_THREADS + ct;
          const int r = idx / F4_PER_ROW, c4 = idx % F4_PER_ROW;
          uint32_t off;
          if (MODE == 0) off = (uint32_t)(c4 >> 4) * 8192u + (uint32_t)r * 128u + ((uint32_t)(((c4 & 15) >> 1) ^ (r & 7)) << 4) + (uint32_t)(c4 & 1) * 8u;
          else off = (uint32_t)r * 128u + ((uint32_t)((c4 >> 1) ^ (r & 7)) << 4) + (uint32_t)(c4 & 1) * 8u;
          uint2 h2, l2;
          split4(v[it], h2, l2);
          *reinterpret_cast<uint2*>(dst + off) = h2;
          *reinterpret_cast<uint2*>(dst + OP_PLANE + off) = l2;
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        conv_bar_sync();
        if (ct == 0) {
          tma::arrive(&op_full[ob]);
          tma::arrive(&empty[st]);  // the fp32 tile has been read (the MMAs still own the batch tile of the stage)
        }
        if (++st == p.stages) st = 0, ph ^= 1;
        if (++ob == N_OP) ob = 0, oph ^= 1;
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4; thread = accumulator row (o or i) =====
    const int q = warp & 3, r = q * 32 + lane;
    const int et = tid - 64;  // 0..127
    int ai = 0;
    pdl_wait();
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ai) {
      const int ab = ai & 1;
      const int uu = u + p.unit0;
      const int sp = uu % p.splits, nt = uu / p.splits, tile = nt % p.tiles, net = nt / p.tiles;
      mbar_wait(&acc_full[ab], (ai >> 1) & 1);
      tcgen05_after_sync();
      if (et == 0) tl_stamp(p.debug, 1200 + 2 * ai);
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ab * 2 * NB;
      float v[NB];
      {
        float a0[16], a1[16], b0[16], b1[16];
        tmem_ld16_nowait(taddr, a0);
        tmem_ld16_nowait(taddr + 16, a1);
        tmem_ld16_nowait(taddr + 32, b0);
        tmem_ld16_nowait(taddr + 48, b1);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = a0[e] + b0[e], v[16 + e] = a1[e] + b1[e];
      }
      tcgen05_before_sync();
      __syncwarp();
      if (lane == 0) tma::arrive(&acc_empty[ab]);  // accumulator drained into registers
      const int row = tile * 128 + r;              // o (fwd) / i (dgrad)
      if (MODE == 0) {
        const bool reduce = p.splits == 1;
        if (p.splits > 1) {
          // split-K: store the partial tile; the consumer (head_q_kernel) sums the splits in a fixed order together
          // with bias and relu -- no fence / ticket handshake on this kernel's critical path
          float* mine = p.part + ((int64_t)nt * p.splits + sp) * (NB * 128);
#pragma unroll
          for (int b = 0; b < NB; ++b) __stcg(mine + b * 128 + r, v[b]);
        }
        if (reduce && row < p.O) {
          const float bb = __ldg(p.w.get<float>(net) + p.b_off + row);
          const int64_t o0 = (int64_t)net * p.ystride + row;
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const float o = fmaxf(v[b] + bb, 0.f);
            p.y[o0 + (int64_t)b * p.O] = o;
            st1_planes(p.yh + o0 + (int64_t)b * p.O, p.yl + o0 + (int64_t)b * p.O, o);
          }
        }
      } else {
        if (row < p.I) {
          int64_t zrow = 0;
          if (p.zP > 0) {
            const int pix = row / p.zC, ch = row - pix * p.zC, yy = pix / p.zW, xx = pix - yy * p.zW;
            zrow = (int64_t)net * p.zstride + ((int64_t)(yy + p.zOff) * p.zP + xx + p.zOff) * p.zC + ch;
          }
          const int64_t x0 = (int64_t)net * p.ystride + row;
          // relu' mask of all 32 samples first (independent loads in flight together), then the stores
          uint32_t posmask = 0;
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const int64_t idx = x0 + (int64_t)b * p.I;
            bool pos;
            if (p.xmask_hi) {
              const unsigned short bits = __ldg(reinterpret_cast<const unsigned short*>(p.xmask_hi) + idx);
              pos = bits != 0 && !(bits & 0x8000u);  // bf16 > 0
            } else {
              pos = __ldg(p.xact + idx) > 0.f;
            }
            posmask |= (pos ? 1u : 0u) << b;
          }
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const int64_t idx = x0 + (int64_t)b * p.I;
            const float o = ((posmask >> b) & 1u) ? v[b] : 0.f;
            if (p.y) p.y[idx] = o;
            const int64_t pi = p.zP > 0 ? zrow + (int64_t)b * p.zRows * p.zC : idx;
            st1_planes(p.yh + pi, p.yl + pi, o);
          }
        }
      }
      if (et == 0) tl_stamp(p.debug, 1201 + 2 * ai);
    }
  }
  tcgen05_before_sync();
  __syncthreads();
  ktl_end(p.tl_id);
  if (warp == 1) tmem_dealloc(tmem, 128);
}

}  // namespace dense
