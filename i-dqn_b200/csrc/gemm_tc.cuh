// tcgen05 implicit-GEMM kernels for the i-DQN step (sm_100a).
//
// Every GEMM operand (input frames, activations, activation gradients, weights) is kept in HBM as two bf16
// planes hi = bf16(x), lo = bf16(x - hi) next to its fp32 master, written once by whoever produces the tensor
// (forward / dgrad epilogues, the Adam kernels, the input prep kernel).  A CTA (256 threads) owns a 128 x NT fp32
// accumulator in TMEM and streams 32-wide K blocks through a 4-stage shared-memory ring with 16-byte cp.async
// copies (zero-fill handles SAME padding and ragged edges), written directly in the UMMA canonical no-swizzle
// layouts of tc_core.cuh: the main loop is address arithmetic + LDGSTS only, three k-blocks of loads stay in
// flight, and no transposes are materialised (both K-major and MN-major operands are used).  One thread issues
// hi*hi + hi*lo + lo*hi tcgen05.mma per 16-wide slice (fp32-faithful "bf16x3"); tcgen05.commit ->
// mbarrier releases the stage.  Epilogues read TMEM with tcgen05.ld and are fused per problem: bias+relu (+ the
// output's planes), relu' mask, deterministic split-K fix-up, or Adam (+ the new weight planes).
//
// Problems (all dims runtime):                         A operand             B operand
//   TcFwdConv    y = relu(conv(x)*s + b)               im2col, K-major       weights [k][n], MN-major  (heads concat on N)
//   TcDgradConv  dx = convT(dy, W) * relu'(x)          dy gather, K-major    weights, K-major          (per stride class)
//   TcWgradConv  dW = im2col(x)^T dy                   im2col^T, MN-major    dy [k][n], MN-major       (split-K)
//   TcFwdDenseT  y^T[o][b] = W^T x^T                   W [k][o], MN-major    x [b][k], K-major         (split-K, batch on N)
//   TcDgradDenseT dx^T[i][b] = W dy^T * relu'(x)       W [i][o], K-major     dy [b][o], K-major
//   TcWgradDenseAdam dW[i][o] = x^T dy, then Adam      x [b][i], MN-major    dy [b][o], MN-major       (K = batch)
#pragma once
#include <algorithm>

#include "common.cuh"
#include "gemm_simt.cuh"  // NetPtr
#include "tc_core.cuh"

namespace tcg {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int NS = 4;      // smem ring depth (max)
constexpr int NTHR = 256;  // threads per CTA

struct Src {  // source of one 16-byte unit (8 elements) in both planes
  const bf16* hi;
  const bf16* lo;
  bool ok;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;  // src-size 0 -> the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void st16(float* __restrict__ dst, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// 16 consecutive elements into both planes (32 bytes each)
__device__ __forceinline__ void st16_planes(bf16* __restrict__ hi, bf16* __restrict__ lo, const float* v) {
  uint4 h0, l0, h1, l1;
  split8(v, h0, l0);
  split8(v + 8, h1, l1);
  reinterpret_cast<uint4*>(hi)[0] = h0, reinterpret_cast<uint4*>(hi)[1] = h1;
  reinterpret_cast<uint4*>(lo)[0] = l0, reinterpret_cast<uint4*>(lo)[1] = l1;
}

// ==================================================================================================
// problem definitions.  Every problem provides
//   Ctx ctx(bx, by, bz)          per-CTA constants (m0, k range, plane pointers)
//   Row rowA(ctx, m)             decode of an operand / epilogue row
//   srcA / srcB                  global source of one 8-element unit (K-major: (row|n, k8); MN-major: (k, mn8))
//   epi(ctx, m|row, n0, v[16])   16 consecutive accumulator columns of one row, or tile_epi for TILE_EPI problems
// ==================================================================================================

struct TcFwdConv {
  static constexpr bool TILE_EPI = false;
  ConvGeom g;
  NetPtr xh, xl;    // input planes (per net / per group)
  NetPtr wh, wl;    // weight planes (online nets, then target nets)
  NetPtr w;         // fp32 arenas (bias)
  int64_t w_off, b_off;
  float* y;         // [nets][ystride] fp32 output
  bf16 *yh, *yl;    // output planes, same indexing
  int64_t ystride;
  float scale;
  int relu;
  const float* skip;  // optional residual input (same indexing as y): y += skip after the activation (impala blocks)
  int planes_relu;    // the consumer reads relu(y) (pre-activation block / trunk): the PLANES hold relu(y), y stays linear
  int nh;           // heads concatenated along N inside one group (they share the input)
  int hpt;          // heads per N tile
  int M, K, NT;     // NT = UMMA N (multiple of 16)
  int nstage;

  struct Ctx {
    int m0, kbeg, kend, group, head0, nheads;
    const bf16 *ah, *al;
  };
  __device__ __forceinline__ Ctx ctx(int bx, int by, int bz) const {
    Ctx c;
    c.m0 = bx * 128;
    c.kbeg = 0, c.kend = K;
    c.group = bz;
    c.head0 = by * hpt;
    c.nheads = min(hpt, nh - c.head0);
    c.ah = xh.get<bf16>(bz * nh);
    c.al = xl.get<bf16>(bz * nh);
    return c;
  }
  struct Row {
    int iy0, ix0, valid;
    int64_t base;  // element offset of sample b
  };
  __device__ __forceinline__ Row rowA(const Ctx& c, int m) const {
    Row r;
    r.valid = m < M;
    uint32_t b, rem, oy, ox;
    g.d_ohow.divmod(r.valid ? m : 0, b, rem);
    g.d_ow.divmod(rem, oy, ox);
    r.iy0 = (int)oy * g.S - g.PH;
    r.ix0 = (int)ox * g.S - g.PW;
    r.base = (int64_t)b * g.IH * g.IW * g.IC;
    return r;
  }
  // A unit: row = output pixel, 8 consecutive k = (ky, kx, c..c+7); when IC == 4 a unit is two taps of four channels and
  // `half` selects one of them: each is decoded on its own, so the pair may straddle a kernel row (odd KW) or the end of K
  __device__ __forceinline__ Src srcA(const Ctx& c, const Row& r, int k, int half) const {
    uint32_t ky, rem, kx, ch;
    const int kk = k + 4 * half;
    g.d_kwic.divmod(kk < K ? kk : 0, ky, rem);
    g.d_ic.divmod(rem, kx, ch);
    const int iy = r.iy0 + (int)ky, ix = r.ix0 + (int)kx;
    const bool ok = r.valid && kk < c.kend && (unsigned)iy < (unsigned)g.IH && (unsigned)ix < (unsigned)g.IW;
    const int64_t idx = ok ? r.base + ((int64_t)iy * g.IW + ix) * g.IC + ch : 0;
    return Src{c.ah + idx, c.al + idx, ok};
  }
  // B unit (MN-major): one k, 8 consecutive n = head*OC + oc..oc+7
  __device__ __forceinline__ Src srcB(const Ctx& c, int k, int n) const {
    uint32_t hl, oc;
    g.d_oc.divmod(n, hl, oc);
    const bool ok = k < c.kend && (int)hl < c.nheads;
    const int net = c.group * nh + c.head0 + (ok ? (int)hl : 0);
    const int64_t idx = w_off + (int64_t)(ok ? k : 0) * g.OC + oc;
    return Src{wh.get<bf16>(net) + idx, wl.get<bf16>(net) + idx, ok};
  }
  __device__ __forceinline__ void epi(const Ctx& c, int m, int n0, const float* v) const {
    if (m >= M) return;
    uint32_t hl, oc;
    g.d_oc.divmod(n0, hl, oc);
    if ((g.OC & 15) == 0) {  // the 16 columns stay inside one head: vector bias loads and stores
      if ((int)hl >= c.nheads) return;
      const int net = c.group * nh + c.head0 + (int)hl;
      const float* bias = w.get<float>(net) + b_off + oc;
      float r[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + q);
        r[4 * q] = v[4 * q] * scale + bb.x, r[4 * q + 1] = v[4 * q + 1] * scale + bb.y;
        r[4 * q + 2] = v[4 * q + 2] * scale + bb.z, r[4 * q + 3] = v[4 * q + 3] * scale + bb.w;
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaxf(r[i], 0.f);
      }
      const int64_t o = (int64_t)net * ystride + (int64_t)m * g.OC + oc;
      if (skip) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 sk = __ldg(reinterpret_cast<const float4*>(skip + o) + q);
          r[4 * q] += sk.x, r[4 * q + 1] += sk.y, r[4 * q + 2] += sk.z, r[4 * q + 3] += sk.w;
        }
      }
      st16(y + o, r);
      if (yh) {
        if (planes_relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = fmaxf(r[i], 0.f);
        }
        st16_planes(yh + o, yl + o, r);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      g.d_oc.divmod(n0 + i, hl, oc);
      if ((int)hl >= c.nheads) break;
      const int net = c.group * nh + c.head0 + (int)hl;
      float r = v[i] * scale + __ldg(w.get<float>(net) + b_off + oc);
      if (relu) r = fmaxf(r, 0.f);
      const int64_t o = (int64_t)net * ystride + (int64_t)m * g.OC + oc;
      if (skip) r += __ldg(skip + o);
      y[o] = r;
      if (yh) st1_planes(yh + o, yl + o, planes_relu ? fmaxf(r, 0.f) : r);
    }
  }
};

// --------------------------------------------------------------------------------------------------
struct TcDgradConv {
  static constexpr bool TILE_EPI = false;
  ConvGeom g;
  const bf16 *dyh, *dyl;  // [z][dystride]
  int64_t dystride;
  NetPtr wh, wl;          // online weight planes
  int64_t w_off;
  const float* xact;      // layer input (relu output) fp32, for the mask
  int mask;               // 1: dx *= (xact > 0); 0: the input entered the layer linearly (impala: first conv of a Stack)
  const float* add;       // optional gradient reaching the same activation through a residual connection: dx += add
  float* dx;
  bf16 *dxh, *dxl;
  int64_t xstride;
  int ncls, JH, JW, K, NT, nstage;
  FastDiv d_jwoc;
  int cls_niy[IDQN_MAX_CLASSES], cls_nix[IDQN_MAX_CLASSES];
  FastDiv cls_d_n[IDQN_MAX_CLASSES], cls_d_nix[IDQN_MAX_CLASSES];

  struct Ctx {
    int m0, M, kbeg, kend, z, py, px, ky0, kx0;
    FastDiv d_n, d_nix;
    const bf16 *ah, *al, *bh, *bl;
  };
  __device__ __forceinline__ Ctx ctx(int bx, int by, int bz) const {
    Ctx c;
    const int z = bz / ncls, cl = bz - z * ncls;
    c.z = z;
    c.m0 = bx * 128;
    c.py = cl / g.S, c.px = cl - c.py * g.S;
    c.ky0 = (c.py + g.PH) % g.S, c.kx0 = (c.px + g.PW) % g.S;
    c.d_n = cls_d_n[cl], c.d_nix = cls_d_nix[cl];
    c.M = g.B * cls_niy[cl] * cls_nix[cl];
    c.kbeg = 0, c.kend = K;
    c.ah = dyh + (int64_t)z * dystride, c.al = dyl + (int64_t)z * dystride;
    c.bh = wh.get<bf16>(z) + w_off, c.bl = wl.get<bf16>(z) + w_off;
    return c;
  }
  struct Row {
    int b, iy, ix, valid;
  };
  __device__ __forceinline__ Row rowA(const Ctx& c, int m) const {
    Row r;
    r.valid = m < c.M;
    uint32_t b, rem, iyp, ixp;
    c.d_n.divmod(r.valid ? m : 0, b, rem);
    c.d_nix.divmod(rem, iyp, ixp);
    r.b = b, r.iy = iyp * g.S + c.py, r.ix = ixp * g.S + c.px;
    return r;
  }
  // A unit: row = input pixel, k = (jy, jx, co..co+7) -> dy[b, oy, ox, co..]
  __device__ __forceinline__ Src srcA(const Ctx& c, const Row& r, int k, int) const {
    uint32_t jy, rem, jx, co;
    d_jwoc.divmod(k < K ? k : 0, jy, rem);
    g.d_oc.divmod(rem, jx, co);
    const int ky = c.ky0 + jy * g.S, kx = c.kx0 + jx * g.S;
    const int ny = r.iy + g.PH - ky, nx = r.ix + g.PW - kx;
    const int oy = ny / g.S, ox = nx / g.S;  // exact for this stride class when ny, nx >= 0
    const bool ok = r.valid && k < c.kend && ky < g.KH && kx < g.KW && ny >= 0 && nx >= 0 && oy < g.OH && ox < g.OW;
    const int64_t idx = ok ? (((int64_t)r.b * g.OH + oy) * g.OW + ox) * g.OC + co : 0;
    return Src{c.ah + idx, c.al + idx, ok};
  }
  // B unit (K-major): row n = input channel c, k = (jy, jx, co..co+7) -> W[ky,kx,c,co..]
  __device__ __forceinline__ Src srcB(const Ctx& c, int n, int k) const {
    uint32_t jy, rem, jx, co;
    d_jwoc.divmod(k < K ? k : 0, jy, rem);
    g.d_oc.divmod(rem, jx, co);
    const int ky = c.ky0 + jy * g.S, kx = c.kx0 + jx * g.S;
    const bool ok = n < g.IC && k < c.kend && ky < g.KH && kx < g.KW;
    const int64_t idx = ok ? ((int64_t)(ky * g.KW + kx) * g.IC + n) * g.OC + co : 0;
    return Src{c.bh + idx, c.bl + idx, ok};
  }
  __device__ __forceinline__ void epi(const Ctx& c, const Row& r, int n0, const float* v) const {
    if (!r.valid) return;
    const int64_t base = (int64_t)c.z * xstride + (((int64_t)r.b * g.IH + r.iy) * g.IW + r.ix) * g.IC;
    if ((g.IC & 15) == 0) {
      if (n0 >= g.IC) return;
      float o[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 xa = mask ? __ldg(reinterpret_cast<const float4*>(xact + base + n0) + q) : make_float4(1.f, 1.f, 1.f, 1.f);
        o[4 * q] = xa.x > 0.f ? v[4 * q] : 0.f, o[4 * q + 1] = xa.y > 0.f ? v[4 * q + 1] : 0.f;
        o[4 * q + 2] = xa.z > 0.f ? v[4 * q + 2] : 0.f, o[4 * q + 3] = xa.w > 0.f ? v[4 * q + 3] : 0.f;
        if (add) {
          const float4 ad = __ldg(reinterpret_cast<const float4*>(add + base + n0) + q);
          o[4 * q] += ad.x, o[4 * q + 1] += ad.y, o[4 * q + 2] += ad.z, o[4 * q + 3] += ad.w;
        }
      }
      st16(dx + base + n0, o);
      st16_planes(dxh + base + n0, dxl + base + n0, o);
      return;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= g.IC) break;
      float o = (!mask || xact[base + n] > 0.f) ? v[i] : 0.f;  // relu'(0) = 0 as in jax
      if (add) o += add[base + n];
      dx[base + n] = o;
      st1_planes(dxh + base + n, dxl + base + n, o);
    }
  }
};

// --------------------------------------------------------------------------------------------------
struct TcWgradConv {
  static constexpr bool TILE_EPI = false;
  ConvGeom g;
  NetPtr xh, xl;          // layer input planes
  const bf16 *dyh, *dyl;
  int64_t dystride;
  const bf16* ones;       // 16 B {1,0,...} followed by 16 B of zeros: the bias-gradient row
  float* gout;
  int64_t gstride, w_off;
  float scale;
  int S, kchunk;          // split-K
  int M, K, NT, nstage;

  struct Ctx {
    int m0, kbeg, kend, z, sp;
    const bf16 *ah, *al, *bh, *bl;
    float* out;
  };
  __device__ __forceinline__ Ctx ctx(int bx, int by, int bz) const {
    Ctx c;
    c.z = bz / S, c.sp = bz - c.z * S;
    c.m0 = bx * 128;
    c.kbeg = c.sp * kchunk;
    c.kend = min(K, c.kbeg + kchunk);
    c.ah = xh.get<bf16>(c.z), c.al = xl.get<bf16>(c.z);
    c.bh = dyh + (int64_t)c.z * dystride, c.bl = dyl + (int64_t)c.z * dystride;
    c.out = gout + (int64_t)c.z * gstride + w_off;
    return c;
  }
  struct Row {};
  __device__ __forceinline__ Row rowA(const Ctx&, int) const { return Row{}; }
  // A unit (MN-major): one k = output pixel, 8 consecutive m = (ky, kx, c..c+7); row Kd is the ones row whose
  // product with dy is the bias gradient (it lands on the bias slot right behind the kernel in the arena)
  // (IC == 4: a unit is two 4-channel taps, `half` selects one; each is decoded on its own -- with Kd = 36 (3 x 3 x 4) the ones
  // row is the second half of the unit that starts at row 32)
  __device__ __forceinline__ Src srcA(const Ctx& c, int k, int m, int half) const {
    const int mm = m + 4 * half;
    if (mm == g.Kd) return Src{ones, ones + 8, k < c.kend};
    const bool inr = k < c.kend && mm < g.Kd;
    uint32_t b, rem, oy, ox, ky, rem2, kx, ch;
    g.d_ohow.divmod(inr ? k : 0, b, rem);
    g.d_ow.divmod(rem, oy, ox);
    g.d_kwic.divmod(inr ? mm : 0, ky, rem2);
    g.d_ic.divmod(rem2, kx, ch);
    const int iy = (int)(oy * g.S + ky) - g.PH, ix = (int)(ox * g.S + kx) - g.PW;
    const bool ok = inr && (unsigned)iy < (unsigned)g.IH && (unsigned)ix < (unsigned)g.IW;
    const int64_t idx = ok ? (((int64_t)b * g.IH + iy) * g.IW + ix) * g.IC + ch : 0;
    return Src{c.ah + idx, c.al + idx, ok};
  }
  // B unit (MN-major): one k = output pixel, 8 consecutive n = oc
  __device__ __forceinline__ Src srcB(const Ctx& c, int k, int n) const {
    const bool ok = k < c.kend && n < g.OC;
    const int64_t idx = ok ? (int64_t)k * g.OC + n : 0;
    return Src{c.bh + idx, c.bl + idx, ok};
  }
  __device__ __forceinline__ void epi(const Ctx& c, int m, int n0, const float* v) const {
    if (m >= M) return;
    const float sc = m < g.Kd ? scale : 1.f;
    if ((g.OC & 15) == 0) {
      if (n0 >= g.OC) return;
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = v[i] * sc;
      st16(c.out + (int64_t)m * g.OC + n0, o);
      return;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= g.OC) break;
      c.out[(int64_t)m * g.OC + n] = v[i] * sc;
    }
  }
};

// --------------------------------------------------------------------------------------------------
// Dense layers with the batch on the N side ("weights are the M operand"): no padding of B=32 to 128 rows.
struct TcFwdDenseT {  // y[net][b][o] = relu(sum_i x[b][i] W[i][o] + bias[o]);  M = O, N = B, K = I
  static constexpr bool TILE_EPI = false;
  NetPtr xh, xl;      // [nets][B][I] planes
  NetPtr wh, wl, w;   // weight planes, fp32 arenas (bias)
  int64_t w_off, b_off;
  float* y;
  bf16 *yh, *yl;
  int64_t ystride;
  int relu;
  int I, O, B;        // K, M, N
  int S, kchunk, NT, nstage;

  struct Ctx {
    int m0, kbeg, kend, z, sp;
    const bf16 *ah, *al, *bh, *bl;
    const float* bias;
  };
  __device__ __forceinline__ Ctx ctx(int bx, int by, int bz) const {
    Ctx c;
    c.z = bz / S, c.sp = bz - c.z * S;
    c.m0 = bx * 128;
    c.kbeg = c.sp * kchunk;
    c.kend = min(I, c.kbeg + kchunk);
    c.ah = wh.get<bf16>(c.z) + w_off, c.al = wl.get<bf16>(c.z) + w_off;
    c.bh = xh.get<bf16>(c.z), c.bl = xl.get<bf16>(c.z);
    c.bias = w.get<float>(c.z) + b_off;
    return c;
  }
  struct Row {};
  __device__ __forceinline__ Row rowA(const Ctx&, int) const { return Row{}; }
  // A unit (MN-major): one k = input feature i, 8 consecutive m = o
  __device__ __forceinline__ Src srcA(const Ctx& c, int k, int m, int) const {
    const bool ok = k < c.kend && m < O;
    const int64_t idx = ok ? (int64_t)k * O + m : 0;
    return Src{c.ah + idx, c.al + idx, ok};
  }
  // B unit (K-major): row n = sample b, 8 consecutive k = i
  __device__ __forceinline__ Src srcB(const Ctx& c, int n, int k) const {
    const bool ok = n < B && k < c.kend;
    const int64_t idx = ok ? (int64_t)n * I + k : 0;
    return Src{c.bh + idx, c.bl + idx, ok};
  }
  __device__ __forceinline__ void epi(const Ctx& c, int m, int n0, const float* v) const {
    if (m >= O) return;
    const float bb = __ldg(c.bias + m);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= B) break;
      float r = v[i] + bb;
      if (relu) r = fmaxf(r, 0.f);
      const int64_t o = (int64_t)c.z * ystride + (int64_t)n * O + m;
      y[o] = r;
      if (yh) st1_planes(yh + o, yl + o, r);
    }
  }
};

struct TcDgradDenseT {  // dx[z][b][i] = relu'(x[b][i]) * sum_o dy[b][o] W[i][o];  M = I, N = B, K = O
  static constexpr bool TILE_EPI = false;
  const bf16 *dyh, *dyl;  // [z][B][O]
  int64_t dystride;
  NetPtr wh, wl;
  int64_t w_off;
  const float* xact;      // [z][B][I]
  float* dx;
  bf16 *dxh, *dxl;
  int64_t xstride;
  int I, O, B, NT, nstage;
  // optional: planes go to the zero-embedded pitched layout dyZ of the preceding conv layer (conv_img.cuh):
  // feature i = (y*zW + x)*zC + c  ->  element ((b*zRows + (y+zOff)*zP + x+zOff)*zC + c), head stride zstride
  int zP, zW, zC, zOff;
  int64_t zRows, zstride;

  struct Ctx {
    int m0, kbeg, kend, z;
    const bf16 *ah, *al, *bh, *bl;
  };
  __device__ __forceinline__ Ctx ctx(int bx, int by, int bz) const {
    Ctx c;
    c.z = bz;
    c.m0 = bx * 128;
    c.kbeg = 0, c.kend = O;
    c.ah = wh.get<bf16>(bz) + w_off, c.al = wl.get<bf16>(bz) + w_off;
    c.bh = dyh + (int64_t)bz * dystride, c.bl = dyl + (int64_t)bz * dystride;
    return c;
  }
  struct Row {
    int m, valid;
  };
  __device__ __forceinline__ Row rowA(const Ctx& c, int m) const { return Row{m, m < I}; }
  // A unit (K-major): row m = input feature i, 8 consecutive k = o
  __device__ __forceinline__ Src srcA(const Ctx& c, const Row& r, int k, int) const {
    const bool ok = r.valid && k < c.kend;
    const int64_t idx = ok ? (int64_t)r.m * O + k : 0;
    return Src{c.ah + idx, c.al + idx, ok};
  }
  // B unit (K-major): row n = sample b, 8 consecutive k = o
  __device__ __forceinline__ Src srcB(const Ctx& c, int n, int k) const {
    const bool ok = n < B && k < c.kend;
    const int64_t idx = ok ? (int64_t)n * O + k : 0;
    return Src{c.bh + idx, c.bl + idx, ok};
  }
  __device__ __forceinline__ void epi(const Ctx& c, const Row& r, int n0, const float* v) const {
    if (!r.valid) return;
    int64_t zrow = 0;
    if (zP > 0) {
      const int pix = r.m / zC, ch = r.m - pix * zC, y = pix / zW, x = pix - y * zW;
      zrow = (int64_t)c.z * zstride + ((int64_t)(y + zOff) * zP + x + zOff) * zC + ch;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= B) break;
      const int64_t idx = (int64_t)c.z * xstride + (int64_t)n * I + r.m;
      const float o = xact[idx] > 0.f ? v[i] : 0.f;
      dx[idx] = o;
      const int64_t pi = zP > 0 ? zrow + (int64_t)n * zRows * zC : idx;
      st1_planes(dxh + pi, dxl + pi, o);
    }
  }
};

struct TcWgradDenseAdam {  // dW[i][o] = sum_b x[b][i] dy[b][o] (K = B), fused optax.adam on W, mu, nu
  static constexpr bool TILE_EPI = true;
  NetPtr xh, xl;           // [z][B][I] planes
  const bf16 *dyh, *dyl;   // [z][B][O] planes
  int64_t dystride;
  const bf16* ones;
  float *W, *mu, *nu, *grad;  // arenas (grad may be null: gradient never materialised)
  bf16 *Wh, *Wl;              // weight planes, refreshed with the new W
  int64_t stride, w_off;
  const int32_t* count;
  float lr, b1, b2, eps;
  int I, O, B, NT, nstage;
  int adam;                // 0: only write the gradient
  int bx0;                 // first 128-row tile of this launch (the step may split the layer over two launches)

  struct Ctx {
    int m0, n0, kbeg, kend, z;
    const bf16 *ah, *al, *bh, *bl;
  };
  __device__ __forceinline__ Ctx ctx(int bx, int by, int bz) const {
    Ctx c;
    c.z = bz;
    c.m0 = (bx + bx0) * 128;
    c.n0 = by * NT;
    c.kbeg = 0, c.kend = B;
    c.ah = xh.get<bf16>(bz), c.al = xl.get<bf16>(bz);
    c.bh = dyh + (int64_t)bz * dystride, c.bl = dyl + (int64_t)bz * dystride;
    return c;
  }
  struct Row {};
  __device__ __forceinline__ Row rowA(const Ctx&, int) const { return Row{}; }
  // A unit (MN-major): one k = sample b, 8 consecutive m = i; row I is the ones row (bias gradient)
  __device__ __forceinline__ Src srcA(const Ctx& c, int k, int m, int) const {
    if (m == I) return Src{ones, ones + 8, k < c.kend};
    const bool ok = k < c.kend && m < I;
    const int64_t idx = ok ? (int64_t)k * I + m : 0;
    return Src{c.ah + idx, c.al + idx, ok};
  }
  // B unit (MN-major): one k = sample b, 8 consecutive n = o (tile-relative n + n0)
  __device__ __forceinline__ Src srcB(const Ctx& c, int k, int n) const {
    const bool ok = k < c.kend && c.n0 + n < O;
    const int64_t idx = ok ? (int64_t)k * O + c.n0 + n : 0;
    return Src{c.bh + idx, c.bl + idx, ok};
  }
  __device__ __forceinline__ void epi(const Ctx&, int, int, const float*) const {}
  // Tile epilogue: the 128 x NT gradient tile sits in shared memory (row stride ld); the CTA walks it row-major
  // so that every warp touches 512 contiguous bytes of W / mu / nu per access (the thread-per-row TMEM layout
  // would stride by a whole 2 KB weight row), four independent float4 triples in flight per thread.
  __device__ __forceinline__ void tile_epi(const Ctx& c, const float* __restrict__ tile, int ld, int tid,
                                           int nthreads) const {
    const int nf4 = NT >> 2;
    const int total = 128 * nf4;
    const AdamCoef ac = adam_coef(b1, b2, lr, eps, count[c.z]);
    const int64_t hb = (int64_t)c.z * stride + w_off;
    float* __restrict__ Wp = W + hb;
    float* __restrict__ Mp = mu + hb;
    float* __restrict__ Vp = nu + hb;
    float* __restrict__ Gp = grad ? grad + hb : nullptr;
    constexpr int U = 4;
    for (int i0 = tid; i0 < total; i0 += U * nthreads) {
      float4 Pv[U], Mv[U], Vv[U], Gv[U];
      int64_t off[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * nthreads;
        const int row = i / nf4, c4 = i - row * nf4;
        const int m = c.m0 + row, n = c.n0 + 4 * c4;
        ok[u] = i < total && m <= I && n < O;  // row I = bias (arena slot b_off = w_off + I*O)
        off[u] = ok[u] ? (int64_t)m * O + n : 0;
        Gv[u] = ok[u] ? *reinterpret_cast<const float4*>(tile + row * ld + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (adam) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (ok[u]) {
            Pv[u] = __ldcs(reinterpret_cast<const float4*>(Wp + off[u]));  // touched once per step: streaming
            Mv[u] = __ldcs(reinterpret_cast<const float4*>(Mp + off[u]));
            Vv[u] = __ldcs(reinterpret_cast<const float4*>(Vp + off[u]));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (ok[u]) {
            adam_elem(ac, Gv[u].x, Pv[u].x, Mv[u].x, Vv[u].x);
            adam_elem(ac, Gv[u].y, Pv[u].y, Mv[u].y, Vv[u].y);
            adam_elem(ac, Gv[u].z, Pv[u].z, Mv[u].z, Vv[u].z);
            adam_elem(ac, Gv[u].w, Pv[u].w, Mv[u].w, Vv[u].w);
            __stcs(reinterpret_cast<float4*>(Wp + off[u]), Pv[u]);
            __stcs(reinterpret_cast<float4*>(Mp + off[u]), Mv[u]);
            __stcs(reinterpret_cast<float4*>(Vp + off[u]), Vv[u]);
            uint2 h2, l2;
            split4(Pv[u], h2, l2);
            *reinterpret_cast<uint2*>(Wh + hb + off[u]) = h2;
            *reinterpret_cast<uint2*>(Wl + hb + off[u]) = l2;
          }
        }
      }
      if (Gp) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (ok[u]) *reinterpret_cast<float4*>(Gp + off[u]) = Gv[u];
      }
    }
  }
};

// ==================================================================================================
// the kernel
// ==================================================================================================
// A_MN / B_MN: operand is MN-major (srcX(ctx, k, mn8)) instead of K-major (srcA(ctx, row, k) / srcB(ctx, n, k)).
// A_HALF: A units are two 8-byte halves (first conv layer, IC == 4).  A_PLANES == 1: the A operand is exact in
// bf16 (uint8 frames), its lo plane is not read.  EPI_ROW: the epilogue takes the Row context (dgrad problems).
// 256 threads: all 8 warps issue copies; warp w reads TMEM lanes 32*(w%4).. and the column chunks of parity w/4.
template <bool A_MN, bool B_MN, int A_PLANES, bool A_HALF, bool EPI_ROW, bool SPLITK, class P>
__global__ void __launch_bounds__(NTHR) tc_gemm_kernel(const P p, float* __restrict__ part, int* __restrict__ tickets) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t mma_done[NS];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_last;

  pdl_trigger();
  pdl_wait();  // this kernel reads its operands right away
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row_t = tid & 127, half = tid >> 7;
  const int NT = p.NT;
  const int nst = p.nstage;
  const uint32_t a_bytes = 128 * BK * 2, b_bytes = (uint32_t)NT * BK * 2;
  const uint32_t stage_bytes = A_PLANES * a_bytes + 2 * b_bytes;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < NT) tmem_cols <<= 1;

  const typename P::Ctx c = p.ctx(blockIdx.x, blockIdx.y, blockIdx.z);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&mma_done[s], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, NT, A_MN, B_MN);
  const uint32_t smem_base = smem_u32(smem);

  // epilogue row context (dgrad problems) and the two gather rows of a K-major A operand
  typename P::Row rowctx;
  if constexpr (EPI_ROW) rowctx = p.rowA(c, c.m0 + row_t);
  typename P::Row rowg[2];
  if constexpr (!A_MN) {
    rowg[0] = p.rowA(c, c.m0 + (tid >> 2));
    rowg[1] = p.rowA(c, c.m0 + (tid >> 2) + 64);
  }

  // Unit -> thread mapping: lanes run along the direction that is contiguous in global memory
  //   K-major  tile [R x 32]: unit e -> kunit = e & 3, row = e >> 2          (4 lanes = 64 contiguous bytes per plane)
  //   MN-major tile [32 x C]: unit e -> group = (e & 7) + 8 * (e >> 8), k = (e >> 3) & 31   (8 lanes = 128 bytes)
  constexpr int MAXB = 4;  // NT <= 256 -> at most 4 B units per thread
  uint32_t offa[2], offb[MAXB];
  bool actb[MAXB];
  const int nbu = B_MN ? ((NT + 63) / 64) * 256 : NT * 4;  // B unit slots
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int e = tid + NTHR * j;
    if constexpr (!A_MN) offa[j] = (uint32_t)(e & 3) * (128 * 16) + (uint32_t)(e >> 2) * 16;
    else {
      const int grp = (e & 7) + 8 * (e >> 8), k = (e >> 3) & 31;
      offa[j] = (uint32_t)(k >> 3) * (128 * 16) + (uint32_t)grp * 128 + (uint32_t)(k & 7) * 16;
    }
  }
#pragma unroll
  for (int j = 0; j < MAXB; ++j) {
    const int e = tid + NTHR * j;
    if constexpr (!B_MN) {
      actb[j] = e < nbu;
      offb[j] = (uint32_t)(e & 3) * ((uint32_t)NT * 16) + (uint32_t)(e >> 2) * 16;
    } else {
      const int grp = (e & 7) + 8 * (e >> 8), k = (e >> 3) & 31;
      actb[j] = e < nbu && grp < (NT >> 3);
      offb[j] = (uint32_t)(k >> 3) * ((uint32_t)NT * 16) + (uint32_t)grp * 128 + (uint32_t)(k & 7) * 16;
    }
  }
  // issue the cp.async copies of k-block kb into ring stage s
  auto issue = [&](int kb, int s) {
    const uint32_t a_hi = smem_base + (uint32_t)s * stage_bytes;
    const uint32_t a_lo = a_hi + a_bytes;
    const uint32_t b_hi = a_hi + A_PLANES * a_bytes;
    const uint32_t b_lo = b_hi + b_bytes;
    const int k0 = c.kbeg + kb * BK;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int e = tid + NTHR * j;
      if constexpr (A_HALF) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          Src s2;
          if constexpr (!A_MN) s2 = p.srcA(c, rowg[j], k0 + 8 * (e & 3), h);
          else s2 = p.srcA(c, k0 + ((e >> 3) & 31), c.m0 + 8 * ((e & 7) + 8 * (e >> 8)), h);
          cp_async8(a_hi + offa[j] + 8 * h, s2.hi, s2.ok);
          if constexpr (A_PLANES == 2) cp_async8(a_lo + offa[j] + 8 * h, s2.lo, s2.ok);
        }
      } else {
        Src s2;
        if constexpr (!A_MN) s2 = p.srcA(c, rowg[j], k0 + 8 * (e & 3), 0);
        else s2 = p.srcA(c, k0 + ((e >> 3) & 31), c.m0 + 8 * ((e & 7) + 8 * (e >> 8)), 0);
        cp_async16(a_hi + offa[j], s2.hi, s2.ok);
        if constexpr (A_PLANES == 2) cp_async16(a_lo + offa[j], s2.lo, s2.ok);
      }
    }
#pragma unroll
    for (int j = 0; j < MAXB; ++j) {
      const int e = tid + NTHR * j;
      if (j * NTHR < nbu && actb[j]) {
        Src s2;
        if constexpr (!B_MN) s2 = p.srcB(c, e >> 2, k0 + 8 * (e & 3));
        else s2 = p.srcB(c, k0 + ((e >> 3) & 31), 8 * ((e & 7) + 8 * (e >> 8)));
        cp_async16(b_hi + offb[j], s2.hi, s2.ok);
        cp_async16(b_lo + offb[j], s2.lo, s2.ok);
      }
    }
  };

  const int nkb = c.kend > c.kbeg ? (c.kend - c.kbeg + BK - 1) / BK : 0;
  // prologue: fill nst-1 stages (one commit group per k-block, possibly empty, so the group count is uniform)
  for (int kb = 0; kb < nst - 1; ++kb) {
    if (kb < nkb) issue(kb, kb);
    cp_async_commit();
  }
  for (int it = 0; it < nkb; ++it) {
    const int s = it % nst;
    // this thread's copies of k-block `it` have landed once at most nst-2 younger groups are pending
    if (nst >= 4) cp_async_wait<2>();
    else if (nst == 3) cp_async_wait<1>();
    else cp_async_wait<0>();
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (warp == 0 && elect_one()) {  // one elected lane: no per-instruction ELECT/BRA.U.ANY waterfall (tc_core.cuh)
      tcgen05_after_sync();
      const uint32_t a_hi = smem_base + (uint32_t)s * stage_bytes;
      const uint32_t a_lo = a_hi + a_bytes;
      const uint32_t b_hi = a_hi + A_PLANES * a_bytes;
      const uint32_t b_lo = b_hi + b_bytes;
      const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)NT * 16;
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {
        const uint64_t dah = make_smem_desc(a_hi + 2 * j * a_lbo, a_lbo, 128);
        const uint64_t dbh = make_smem_desc(b_hi + 2 * j * b_lbo, b_lbo, 128);
        const uint64_t dbl = make_smem_desc(b_lo + 2 * j * b_lbo, b_lbo, 128);
        mma_bf16(tmem, dah, dbh, idesc, (it > 0 || j > 0) ? 1u : 0u);
        mma_bf16(tmem, dah, dbl, idesc, 1u);
        if constexpr (A_PLANES == 2) {
          const uint64_t dal = make_smem_desc(a_lo + 2 * j * a_lbo, a_lbo, 128);
          mma_bf16(tmem, dal, dbh, idesc, 1u);
        }
      }
      mma_commit(&mma_done[s]);
    }
    // refill: k-block it+nst-1 goes into the stage the MMAs of iteration it-1 were reading
    const int nxt = it + nst - 1;
    if (nxt < nkb) {
      if (it >= 1) mbar_wait(&mma_done[(it - 1) % nst], ((it - 1) / nst) & 1);
      issue(nxt, nxt % nst);
    }
    cp_async_commit();
  }
  if (nkb > 0) {
    const int last = nkb - 1;
    mbar_wait(&mma_done[last % nst], (last / nst) & 1);
    tcgen05_after_sync();
  }

  // ---- epilogue: warp w reads TMEM lanes [32(w%4), +32) -> accumulator row m0 + row_t; the two warps sharing a
  // lane quadrant take the even / odd 16-column chunks
  const int m = c.m0 + row_t;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  auto load_chunk = [&](int c0, float* v) {
    if (nkb > 0) tmem_ld16(lane_base + c0, v);
    else
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
  };

  if constexpr (P::TILE_EPI) {
    // all MMAs are complete, the operand ring is dead: stage the tile in shared memory, then walk it row-major
    float* tile = reinterpret_cast<float*>(smem);
    const int ld = NT + 4;
    for (int c0 = 16 * half; c0 < NT; c0 += 32) {
      float v[16];
      load_chunk(c0, v);
      st16(tile + row_t * ld + c0, v);
    }
    tcgen05_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, tmem_cols);
    p.tile_epi(c, tile, ld, tid, NTHR);
    return;
  }

  if constexpr (SPLITK) {
    const int S = p.S;
    if (S > 1) {
      const int z = blockIdx.z / S, sp = blockIdx.z - z * S;
      const int tiles = gridDim.x * gridDim.y;
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      float* mine = part + ((int64_t)(z * tiles + tile) * S + sp) * ((int64_t)NT * 128);
      for (int c0 = 16 * half; c0 < NT; c0 += 32) {
        float v[16];
        load_chunk(c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) __stcg(mine + (int64_t)(c0 + i) * 128 + row_t, v[i]);
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const int t = atomicAdd(&tickets[z * tiles + tile], 1);
        s_last = (t == S - 1);
        if (s_last) tickets[z * tiles + tile] = 0;
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        const float* base = part + (int64_t)(z * tiles + tile) * S * ((int64_t)NT * 128);
        for (int c0 = 16 * half; c0 < NT; c0 += 32) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
          for (int q = 0; q < S; ++q) {  // fixed order: deterministic
            const float* src = base + (int64_t)q * NT * 128 + (int64_t)c0 * 128 + row_t;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldcg(src + i * 128);
          }
          p.epi(c, m, c0, v);
        }
      }
      tcgen05_before_sync();
      __syncthreads();
      if (warp == 0) tmem_dealloc(tmem, tmem_cols);
      return;
    }
  }
  for (int c0 = 16 * half; c0 < NT; c0 += 32) {
    float v[16];
    load_chunk(c0, v);
    if constexpr (EPI_ROW) p.epi(c, rowctx, c0, v);
    else p.epi(c, m, c0, v);
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <bool A_MN, bool B_MN, int A_PLANES, bool A_HALF, bool EPI_ROW, bool SPLITK, class P>
static inline cudaError_t launch_tc(const P& p, dim3 grid, float* part, int* tickets, cudaStream_t st, bool pdl = false) {
  size_t smem = (size_t)p.nstage * (A_PLANES * 128 * BK * 2 + 2 * (size_t)p.NT * BK * 2);
  if (P::TILE_EPI) smem = std::max(smem, (size_t)128 * (p.NT + 4) * sizeof(float));
  auto kern = tc_gemm_kernel<A_MN, B_MN, A_PLANES, A_HALF, EPI_ROW, SPLITK, P>;
  static bool configured = false;  // one flag per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return launch_pdl(pdl, kern, grid, dim3(NTHR), smem, st, p, part, tickets);
}

// ring depth: at most 56 KB of stages per CTA, i.e. three or four CTAs resident per SM (impala K = 5 step, shared-memory limit
// per CTA 100 / 72 / 54 / 40 KB: 8.27 / 7.74 / 7.67 / 7.66 ms -- these kernels hide their gather latency with resident CTAs, not
// with ring depth; the k-block order, hence every rounding, does not depend on the depth).  IDQN_TC_SMEM_KB overrides.
static inline int pick_stages(int a_planes, int NT, int kblocks) {
  const size_t stage = (size_t)a_planes * 128 * BK * 2 + 2 * (size_t)NT * BK * 2;
  static const size_t limit = (getenv("IDQN_TC_SMEM_KB") ? (size_t)atoi(getenv("IDQN_TC_SMEM_KB")) : 56) * 1024;
  int n = NS;
  while (n > 2 && n * stage > limit) --n;
  return std::max(2, std::min(n, kblocks + 1));  // the ring needs >= 2 stages
}

// ---- plane maintenance ------------------------------------------------------------------------------------
// fp32 -> (hi, lo) planes, 8 elements per thread
__global__ void __launch_bounds__(256) to_planes_kernel(const float* __restrict__ src, bf16* __restrict__ hi,
                                                        bf16* __restrict__ lo, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 h, l;
    split8(x, h, l);
    reinterpret_cast<uint4*>(hi)[i] = h;
    reinterpret_cast<uint4*>(lo)[i] = l;
  }
}
// uint8 frames -> hi plane (exact), 8 pixels per thread
__global__ void __launch_bounds__(256) u8_to_plane_kernel(const uint8_t* __restrict__ src, bf16* __restrict__ hi,
                                                          int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src) + i);
    const uint32_t w[2] = {raw.x, raw.y};
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (float)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
    reinterpret_cast<uint4*>(hi)[i] = pack8_exact(x);
  }
}

}  // namespace tcg
