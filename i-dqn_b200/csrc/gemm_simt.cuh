// fp32 CUDA-core implicit GEMM used (a) as the exact-fp32 path for every layer shape the tensor-core
// path does not cover and (b) as the on-device cross-check of the tcgen05 kernels.
//
// C[z][m][n] = sum_k A_z(m,k) * B_z(k,n), with A/B produced on the fly by a "problem" object:
//   FwdProb    y = relu(conv_SAME(x) * scale + bias)          (architectures/dqn.py:42-52,68)
//   WgradProb  dW[(ky,kx,c),o] = sum_{b,oy,ox} im2col(x) * dY  (+ bias row), flax kernel layout
//   DgradProb  dX = conv_transpose(dY, W) masked by relu'(x)   (one stride-parity class per grid.z slice)
// Split-K partials are combined by the last-arriving CTA in a fixed order (deterministic results).
#pragma once
#include "common.cuh"
#include "tc_core.cuh"

struct NetPtr {  // per-net pointer selection: nets [0,zsplit) use p0, the rest p1
  const void* p0;
  const void* p1;
  int64_t stride0, stride1;  // in elements of the pointee
  int zsplit;
  template <class T>
  __device__ __forceinline__ const T* get(int z) const {
    return z < zsplit ? (const T*)p0 + (int64_t)z * stride0 : (const T*)p1 + (int64_t)(z - zsplit) * stride1;
  }
};

// --------------------------------------------------------------------------------------------
struct FwdProb {
  ConvGeom g;
  NetPtr x;  // input activations (u8 or f32)
  NetPtr w;  // head arena base
  int x_u8;
  int64_t w_off, b_off;
  float* y;
  __nv_bfloat16 *yh, *yl;  // optional bf16 hi/lo planes of the output (consumed by the tensor-core kernels)
  int64_t ystride;
  float scale;
  int relu;
  int relu_in;        // the conv reads relu(x) (impala's pre-activation blocks)
  int planes_relu;    // the consumer reads relu(y): the planes hold relu(y), y stays linear
  const float* skip;  // optional residual input, same layout / net stride as y: y += skip (after the activation)
  int nz, S;  // nets, split-K factor
  int M, N, K, kchunk;

  struct Ctx {
    const FwdProb* p;
    const uint8_t* xu;
    const float* xf;
    const float* wk;
    const float* bias;
    const float* skip;
    float* y;
    __nv_bfloat16 *yh, *yl;
    int M, kbeg, kend, z;
    __device__ __forceinline__ float loadA(int m, int k) const {
      if (m >= M || k >= kend) return 0.f;
      const ConvGeom& g = p->g;
      uint32_t b, r, oy, ox, ky, r2, kx, c;
      g.d_ohow.divmod(m, b, r);
      g.d_ow.divmod(r, oy, ox);
      g.d_kwic.divmod(k, ky, r2);
      g.d_ic.divmod(r2, kx, c);
      int iy = (int)(oy * g.S + ky) - g.PH, ix = (int)(ox * g.S + kx) - g.PW;
      if ((unsigned)iy >= (unsigned)g.IH || (unsigned)ix >= (unsigned)g.IW) return 0.f;
      int64_t idx = (((int64_t)b * g.IH + iy) * g.IW + ix) * g.IC + c;
      const float v = xu ? (float)__ldg(xu + idx) : __ldg(xf + idx);
      return p->relu_in ? fmaxf(v, 0.f) : v;
    }
    __device__ __forceinline__ float loadB(int k, int n) const {
      if (k >= kend || n >= p->N) return 0.f;
      return __ldg(wk + (int64_t)k * p->N + n);
    }
    __device__ __forceinline__ void store(int m, int n, float acc) const {
      if (m >= M || n >= p->N) return;
      float v = acc * p->scale + __ldg(bias + n);
      if (p->relu) v = fmaxf(v, 0.f);
      if (skip) v += __ldg(skip + (int64_t)m * p->N + n);
      y[(int64_t)m * p->N + n] = v;
      if (yh) tc::st1_planes(yh + (int64_t)m * p->N + n, yl + (int64_t)m * p->N + n, p->planes_relu ? fmaxf(v, 0.f) : v);
    }
  };
  __device__ __forceinline__ Ctx ctx(int zz) const {
    Ctx c;
    int z = zz / S, sp = zz - z * S;
    c.p = this;
    c.z = z;
    const float* base = w.get<float>(z);
    c.wk = base + w_off;
    c.bias = base + b_off;
    c.xu = x_u8 ? x.get<uint8_t>(z) : nullptr;
    c.xf = x_u8 ? nullptr : x.get<float>(z);
    c.y = y + (int64_t)z * ystride;
    c.skip = skip ? skip + (int64_t)z * ystride : nullptr;
    c.yh = yh ? yh + (int64_t)z * ystride : nullptr;
    c.yl = yl ? yl + (int64_t)z * ystride : nullptr;
    c.M = M;
    c.kbeg = sp * kchunk;
    c.kend = min(K, c.kbeg + kchunk);
    return c;
  }
};

// --------------------------------------------------------------------------------------------
struct WgradProb {
  ConvGeom g;
  NetPtr x;   // layer input (u8 / f32)
  int x_u8;
  const float* dy;  // [nz][B*OH*OW][OC]
  int64_t dystride;
  float* gout;      // grad arena base; row m of the [Kd+1, OC] result goes to gout[z*gstride + w_off + m*OC + n]
  int64_t gstride, w_off;
  float scale;      // 1/255 for the u8/255 first layer (applied to kernel rows only)
  int relu_in;      // the layer read relu(x)
  int nz, S;
  int M, N, K, kchunk;

  struct Ctx {
    const WgradProb* p;
    const uint8_t* xu;
    const float* xf;
    const float* dy;
    float* out;
    int M, kbeg, kend, z;
    __device__ __forceinline__ float loadA(int m, int k) const {  // = im2col(x)[row k][col m]; row Kd = ones
      if (m >= M || k >= kend) return 0.f;
      const ConvGeom& g = p->g;
      if (m == g.Kd) return 1.f;
      uint32_t b, r, oy, ox, ky, r2, kx, c;
      g.d_ohow.divmod(k, b, r);
      g.d_ow.divmod(r, oy, ox);
      g.d_kwic.divmod(m, ky, r2);
      g.d_ic.divmod(r2, kx, c);
      int iy = (int)(oy * g.S + ky) - g.PH, ix = (int)(ox * g.S + kx) - g.PW;
      if ((unsigned)iy >= (unsigned)g.IH || (unsigned)ix >= (unsigned)g.IW) return 0.f;
      int64_t idx = (((int64_t)b * g.IH + iy) * g.IW + ix) * g.IC + c;
      const float v = xu ? (float)__ldg(xu + idx) : __ldg(xf + idx);
      return p->relu_in ? fmaxf(v, 0.f) : v;
    }
    __device__ __forceinline__ float loadB(int k, int n) const {
      if (k >= kend || n >= p->N) return 0.f;
      return __ldg(dy + (int64_t)k * p->N + n);
    }
    __device__ __forceinline__ void store(int m, int n, float acc) const {
      if (m >= M || n >= p->N) return;
      out[(int64_t)m * p->N + n] = (m < p->g.Kd) ? acc * p->scale : acc;
    }
  };
  __device__ __forceinline__ Ctx ctx(int zz) const {
    Ctx c;
    int z = zz / S, sp = zz - z * S;
    c.p = this;
    c.z = z;
    c.xu = x_u8 ? x.get<uint8_t>(z) : nullptr;
    c.xf = x_u8 ? nullptr : x.get<float>(z);
    c.dy = dy + (int64_t)z * dystride;
    c.out = gout + (int64_t)z * gstride + w_off;
    c.M = M;
    c.kbeg = sp * kchunk;
    c.kend = min(K, c.kbeg + kchunk);
    return c;
  }
};

// --------------------------------------------------------------------------------------------
#define IDQN_MAX_CLASSES 16
struct DgradProb {
  ConvGeom g;
  const float* dy;  // [nz][B*OH*OW][OC]
  int64_t dystride;
  NetPtr w;         // head arena base (online)
  int64_t w_off;
  const float* xact;  // layer input activations (relu outputs) [nz][B*IH*IW*IC] for the mask
  int mask;           // 1: dx *= (xact > 0) (the input went through a relu on its way into this layer); 0: linear input
  const float* add;   // optional [nz][..] gradient that reaches the same activation through a residual connection: dx += add
  float* dx;          // same shape
  __nv_bfloat16 *dxh, *dxl;  // optional planes of dx
  int64_t xstride;
  int nz, S;          // S: split-K (always 1 here)
  int ncls;           // S_conv^2 stride-parity classes
  int JH, JW;         // taps per class per axis: ceil(KH/S), ceil(KW/S)
  int N, K, kchunk, M;  // M = max rows over classes (grid sizing)
  FastDiv d_jwoc;       // JW*OC
  int cls_niy[IDQN_MAX_CLASSES], cls_nix[IDQN_MAX_CLASSES];
  FastDiv cls_d_n[IDQN_MAX_CLASSES], cls_d_nix[IDQN_MAX_CLASSES];  // niy*nix, nix

  struct Ctx {
    const DgradProb* p;
    const float* dy;
    const float* wk;
    const float* xact;
    const float* add;
    float* dx;
    __nv_bfloat16 *dxh, *dxl;
    int M, kbeg, kend, z;
    int py, px, ky0, kx0;
    FastDiv d_n, d_nix;
    __device__ __forceinline__ void rowdec(int m, uint32_t& b, int& iy, int& ix) const {
      uint32_t r, iyp, ixp;
      d_n.divmod(m, b, r);
      d_nix.divmod(r, iyp, ixp);
      iy = iyp * p->g.S + py;
      ix = ixp * p->g.S + px;
    }
    __device__ __forceinline__ float loadA(int m, int k) const {
      if (m >= M || k >= kend) return 0.f;
      const ConvGeom& g = p->g;
      uint32_t b, jy, r2, jx, co;
      int iy, ix;
      rowdec(m, b, iy, ix);
      p->d_jwoc.divmod(k, jy, r2);
      g.d_oc.divmod(r2, jx, co);
      int ky = ky0 + jy * g.S, kx = kx0 + jx * g.S;
      if (ky >= g.KH || kx >= g.KW) return 0.f;
      int ny = iy + g.PH - ky, nx = ix + g.PW - kx;
      if (ny < 0 || nx < 0) return 0.f;
      int oy = ny / g.S, ox = nx / g.S;  // exact by construction of the class
      if (oy >= g.OH || ox >= g.OW) return 0.f;
      return __ldg(dy + (((int64_t)b * g.OH + oy) * g.OW + ox) * g.OC + co);
    }
    __device__ __forceinline__ float loadB(int k, int n) const {
      if (k >= kend || n >= p->N) return 0.f;
      const ConvGeom& g = p->g;
      uint32_t jy, r2, jx, co;
      p->d_jwoc.divmod(k, jy, r2);
      g.d_oc.divmod(r2, jx, co);
      int ky = ky0 + jy * g.S, kx = kx0 + jx * g.S;
      if (ky >= g.KH || kx >= g.KW) return 0.f;
      return __ldg(wk + ((int64_t)(ky * g.KW + kx) * g.IC + n) * g.OC + co);
    }
    __device__ __forceinline__ void store(int m, int n, float acc) const {
      if (m >= M || n >= p->N) return;
      const ConvGeom& g = p->g;
      uint32_t b;
      int iy, ix;
      rowdec(m, b, iy, ix);
      int64_t idx = (((int64_t)b * g.IH + iy) * g.IW + ix) * g.IC + n;
      float o = (!p->mask || xact[idx] > 0.f) ? acc : 0.f;  // relu'(0) = 0 as in jax
      if (add) o += add[idx];
      dx[idx] = o;
      if (dxh) tc::st1_planes(dxh + idx, dxl + idx, o);
    }
  };
  __device__ __forceinline__ Ctx ctx(int zz) const {
    Ctx c;
    int z = zz / ncls, cl = zz - z * ncls;
    c.p = this;
    c.z = z;
    c.py = cl / g.S;
    c.px = cl - c.py * g.S;
    c.ky0 = (c.py + g.PH) % g.S;
    c.kx0 = (c.px + g.PW) % g.S;
    c.d_n = cls_d_n[cl];
    c.d_nix = cls_d_nix[cl];
    c.M = g.B * cls_niy[cl] * cls_nix[cl];
    c.dy = dy + (int64_t)z * dystride;
    c.wk = w.get<float>(z) + w_off;
    c.xact = xact + (int64_t)z * xstride;
    c.add = add ? add + (int64_t)z * xstride : nullptr;
    c.dx = dx + (int64_t)z * xstride;
    c.dxh = dxh ? dxh + (int64_t)z * xstride : nullptr;
    c.dxl = dxl ? dxl + (int64_t)z * xstride : nullptr;
    c.kbeg = 0;
    c.kend = K;
    return c;
  }
};

// --------------------------------------------------------------------------------------------
// The kernel: BMxBN tile, BK=16, 4x4 register tile per thread, register-prefetched single smem stage.
template <int BM, int BN, bool A_KFAST, bool B_KFAST, class P>
__global__ void __launch_bounds__((BM / 4) * (BN / 4)) gemm_simt_kernel(const P p, float* __restrict__ part,
                                                                       int* __restrict__ tickets) {
  constexpr int BK = 16;
  constexpr int NT = (BM / 4) * (BN / 4);
  constexpr int EA = BM * BK / NT;
  constexpr int EB = BK * BN / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const typename P::Ctx c = p.ctx(blockIdx.z);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= c.M) return;  // dgrad classes have fewer rows than the grid was sized for (uniform per CTA)

  const int ty = tid / (BN / 4), tx = tid % (BN / 4);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[EA], rb[EB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < EA; ++i) {
      int e = tid + i * NT;
      int ak = A_KFAST ? e % BK : e / BM;
      int am = A_KFAST ? e / BK : e % BM;
      ra[i] = c.loadA(m0 + am, k0 + ak);
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
      int e = tid + i * NT;
      int bk = B_KFAST ? e % BK : e / BN;
      int bn = B_KFAST ? e / BK : e % BN;
      rb[i] = c.loadB(k0 + bk, n0 + bn);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < EA; ++i) {
      int e = tid + i * NT;
      int ak = A_KFAST ? e % BK : e / BM;
      int am = A_KFAST ? e / BK : e % BM;
      As[ak][am] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
      int e = tid + i * NT;
      int bk = B_KFAST ? e % BK : e / BN;
      int bn = B_KFAST ? e / BK : e % BN;
      Bs[bk][bn] = rb[i];
    }
  };

  if (c.kbeg < c.kend) {
    fetch(c.kbeg);
    for (int k0 = c.kbeg; k0 < c.kend; k0 += BK) {
      stash();
      __syncthreads();
      if (k0 + BK < c.kend) fetch(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  if (p.S > 1) {
    // deterministic split-K: publish the partial, the last CTA of this (net, tile) adds all S of them in order
    const int z = blockIdx.z / p.S, sp = blockIdx.z - z * p.S;
    const int tiles = gridDim.x * gridDim.y;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    float* mine = part + ((int64_t)(z * tiles + tile) * p.S + sp) * (16 * NT);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) __stcg(mine + (i * 4 + j) * NT + tid, acc[i][j]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      int t = atomicAdd(&tickets[z * tiles + tile], 1);
      s_last = (t == p.S - 1);
      if (s_last) tickets[z * tiles + tile] = 0;  // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = part + (int64_t)(z * tiles + tile) * p.S * (16 * NT);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s = 0.f;
        for (int q = 0; q < p.S; ++q) s += __ldcg(base + (int64_t)q * (16 * NT) + (i * 4 + j) * NT + tid);
        acc[i][j] = s;
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c.store(m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
}

// host-side launcher: picks the tile shape from N and returns the launch geometry it used
template <bool A_KFAST, bool B_KFAST, class P>
static inline cudaError_t launch_gemm_simt(const P& p, int gridz, float* part, int* tickets, cudaStream_t st) {
  if (p.N <= 32) {
    dim3 grid((p.M + 63) / 64, (p.N + 31) / 32, gridz);
    gemm_simt_kernel<64, 32, A_KFAST, B_KFAST, P><<<grid, 128, 0, st>>>(p, part, tickets);
  } else {
    dim3 grid((p.M + 63) / 64, (p.N + 63) / 64, gridz);
    gemm_simt_kernel<64, 64, A_KFAST, B_KFAST, P><<<grid, 256, 0, st>>>(p, part, tickets);
  }
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------------
// max-pool k x k / s, SAME (-inf) padding (flax nn.max_pool, architectures/dqn.py:20), NHWC fp32, nz nets.
// The backward pass routes dy to the FIRST maximum of its window in (ky, kx) row-major order (what a strict > scan finds).
struct PoolArgs {
  int nz, B, IH, IW, C, OH, OW, PH, PW, K, S;
  const float* x;   // [nz][B][IH][IW][C]
  float* y;         // [nz][B][OH][OW][C]   (forward)
  const float* dy;  // backward
  float* dx;
  __nv_bfloat16 *ph, *pl;    // optional bf16 hi / lo planes of the output (forward: of y, relu'd if planes_relu; backward: of dx)
  int planes_relu;
  // optional record of the forward's choices, one byte per output element (ky * 3 + kx), indexed like y for nets < arg_nets:
  // written by the learning step's forward, read by its backward instead of re-scanning nine inputs per window (3 x 3 / 2 kernels)
  uint8_t* arg;
  int arg_nets;
  int64_t xstride, ystride;  // floats between nets
};
__device__ __forceinline__ int pool_argmax(const PoolArgs& a, const float* xb, int oy, int ox, int c, float* best_out) {
  float best = -INFINITY;
  int arg = -1;
  for (int ky = 0; ky < a.K; ++ky) {
    const int iy = oy * a.S + ky - a.PH;
    if ((unsigned)iy >= (unsigned)a.IH) continue;
    for (int kx = 0; kx < a.K; ++kx) {
      const int ix = ox * a.S + kx - a.PW;
      if ((unsigned)ix >= (unsigned)a.IW) continue;
      const float v = __ldg(xb + ((int64_t)iy * a.IW + ix) * a.C + c);
      if (arg < 0 || v > best) best = v, arg = iy * a.IW + ix;
    }
  }
  if (best_out) *best_out = best;
  return arg;
}
// ---- the pool the reference uses (3 x 3 / 2), four channels per thread ---------------------------------------------------
// grid = (chunks of 256 threads inside one image, images = nz * B); thread = (row, col, group of 4 channels): 16-byte loads,
// 32-bit index arithmetic, the nine window loads unrolled.  Same decisions as the generic kernels below (first maximum of a
// window in row-major order, -inf padding never selected); used when C % 4 == 0 and every pointer / stride is 16-byte aligned.
__device__ __forceinline__ float4 pool_ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// first maximum of window (oy, ox), per channel of the group: value and window element ky * 3 + kx
__device__ __forceinline__ void pool3_window(const PoolArgs& a, const float* xb, int oy, int ox, int c, float4& best, int4& arg) {
  best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  arg = make_int4(-1, -1, -1, -1);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - a.PH;
    if ((unsigned)iy >= (unsigned)a.IH) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - a.PW;
      if ((unsigned)ix >= (unsigned)a.IW) continue;
      const int code = ky * 3 + kx;
      const float4 v = pool_ld4(xb + (uint32_t)(iy * a.IW + ix) * a.C + c);
      if (arg.x < 0 || v.x > best.x) best.x = v.x, arg.x = code;
      if (arg.y < 0 || v.y > best.y) best.y = v.y, arg.y = code;
      if (arg.z < 0 || v.z > best.z) best.z = v.z, arg.z = code;
      if (arg.w < 0 || v.w > best.w) best.w = v.w, arg.w = code;
    }
  }
}
__global__ void __launch_bounds__(256) maxpool3_fwd_v4_kernel(const PoolArgs a) {
  const int C4 = a.C >> 2;
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;  // (oy, ox, c4) inside image blockIdx.y
  if (i >= (uint32_t)(a.OH * a.OW * C4)) return;
  const int c = (int)(i % C4) * 4, q = (int)(i / C4), ox = q % a.OW, oy = q / a.OW;
  const int z = blockIdx.y / a.B, b = blockIdx.y - z * a.B;
  const float* xb = a.x + (int64_t)z * a.xstride + (int64_t)b * a.IH * a.IW * a.C;
  float4 best;
  int4 arg;
  pool3_window(a, xb, oy, ox, c, best, arg);
  const int64_t o = (int64_t)z * a.ystride + ((int64_t)b * a.OH * a.OW + q) * a.C + c;
  *reinterpret_cast<float4*>(a.y + o) = best;
  if (a.arg && z < a.arg_nets) *reinterpret_cast<uint32_t*>(a.arg + o) = (uint32_t)arg.x | ((uint32_t)arg.y << 8) | ((uint32_t)arg.z << 16) | ((uint32_t)arg.w << 24);
  if (a.ph) {
    if (a.planes_relu) best = make_float4(fmaxf(best.x, 0.f), fmaxf(best.y, 0.f), fmaxf(best.z, 0.f), fmaxf(best.w, 0.f));
    uint2 hi, lo;
    tc::split4(best, hi, lo);
    *reinterpret_cast<uint2*>(a.ph + o) = hi, *reinterpret_cast<uint2*>(a.pl + o) = lo;
  }
}
__global__ void __launch_bounds__(256) maxpool3_bwd_v4_kernel(const PoolArgs a) {
  const int C4 = a.C >> 2;
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;  // (iy, ix, c4) inside image blockIdx.y
  if (i >= (uint32_t)(a.IH * a.IW * C4)) return;
  const int c = (int)(i % C4) * 4, pos = (int)(i / C4), ix = pos % a.IW, iy = pos / a.IW;
  const int z = blockIdx.y / a.B, b = blockIdx.y - z * a.B;
  const float* xb = a.x + (int64_t)z * a.xstride + (int64_t)b * a.IH * a.IW * a.C;
  const int64_t yoff = (int64_t)z * a.ystride + (int64_t)b * a.OH * a.OW * a.C;
  const float* dyb = a.dy + yoff;
  // windows oy with 2 oy - PH <= iy <= 2 oy - PH + 2, ascending (oy, ox): the order the generic kernel sums in
  const int oy_lo = max(0, (iy + a.PH - 1) / 2), oy_hi = min(a.OH - 1, (iy + a.PH) / 2);
  const int ox_lo = max(0, (ix + a.PW - 1) / 2), ox_hi = min(a.OW - 1, (ix + a.PW) / 2);
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int oy = oy_lo; oy <= oy_hi; ++oy)
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const int me = (iy - (oy * 2 - a.PH)) * 3 + (ix - (ox * 2 - a.PW));  // this input as an element of window (oy, ox)
      const uint32_t w = (uint32_t)(oy * a.OW + ox) * a.C + c;
      int4 arg;
      if (a.arg) {
        const uint32_t packed = __ldg(reinterpret_cast<const uint32_t*>(a.arg + yoff + w));
        arg = make_int4(packed & 0xff, (packed >> 8) & 0xff, (packed >> 16) & 0xff, packed >> 24);
      } else {
        float4 best;
        pool3_window(a, xb, oy, ox, c, best, arg);
      }
      const float4 d = pool_ld4(dyb + w);
      if (arg.x == me) g.x += d.x;
      if (arg.y == me) g.y += d.y;
      if (arg.z == me) g.z += d.z;
      if (arg.w == me) g.w += d.w;
    }
  const int64_t o = (int64_t)z * a.xstride + ((int64_t)b * a.IH * a.IW + pos) * a.C + c;
  *reinterpret_cast<float4*>(a.dx + o) = g;
  if (a.ph) {
    uint2 hi, lo;
    tc::split4(g, hi, lo);
    *reinterpret_cast<uint2*>(a.ph + o) = hi, *reinterpret_cast<uint2*>(a.pl + o) = lo;
  }
}
// the fast kernels apply: 3 x 3 / 2, channel groups of four, 16-byte aligned tensors, one image within 2^31 elements
inline bool pool3_v4_ok(const PoolArgs& a, const void* p0, const void* p1, const void* p2) {
  auto al = [](const void* p, size_t n) { return (reinterpret_cast<uintptr_t>(p) % n) == 0; };
  return a.K == 3 && a.S == 2 && a.C % 4 == 0 && a.xstride % 4 == 0 && a.ystride % 4 == 0 && al(p0, 16) && al(p1, 16) && al(p2, 16) &&
         (!a.ph || (al(a.ph, 8) && al(a.pl, 8))) && (int64_t)a.IH * a.IW * a.C < (1ll << 31) && (int64_t)a.nz * a.B <= 65535;
}

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const PoolArgs a) {
  const int64_t per = (int64_t)a.B * a.OH * a.OW * a.C, total = per * a.nz;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i / per);
    int64_t r = i - (int64_t)z * per;
    const int c = (int)(r % a.C);
    r /= a.C;
    const int ox = (int)(r % a.OW);
    r /= a.OW;
    const int oy = (int)(r % a.OH), b = (int)(r / a.OH);
    float best;
    pool_argmax(a, a.x + (int64_t)z * a.xstride + (int64_t)b * a.IH * a.IW * a.C, oy, ox, c, &best);
    const int64_t o = (int64_t)z * a.ystride + (i - (int64_t)z * per);
    a.y[o] = best;
    if (a.ph) tc::st1_planes(a.ph + o, a.pl + o, a.planes_relu ? fmaxf(best, 0.f) : best);
  }
}
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const PoolArgs a) {
  // one thread per INPUT element: the (at most ceil(K/S)^2) windows that contain it, gathered -- no atomics, fixed order
  const int64_t per = (int64_t)a.B * a.IH * a.IW * a.C, total = per * a.nz;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i / per);
    int64_t r = i - (int64_t)z * per;
    const int c = (int)(r % a.C);
    r /= a.C;
    const int ix = (int)(r % a.IW);
    r /= a.IW;
    const int iy = (int)(r % a.IH), b = (int)(r / a.IH);
    const float* xb = a.x + (int64_t)z * a.xstride + (int64_t)b * a.IH * a.IW * a.C;
    const float* dyb = a.dy + (int64_t)z * a.ystride + (int64_t)b * a.OH * a.OW * a.C;
    // windows oy with oy*S - PH <= iy <= oy*S - PH + K - 1
    const int oy_lo = max(0, (iy + a.PH - a.K + a.S) / a.S), oy_hi = min(a.OH - 1, (iy + a.PH) / a.S);
    const int ox_lo = max(0, (ix + a.PW - a.K + a.S) / a.S), ox_hi = min(a.OW - 1, (ix + a.PW) / a.S);
    float g = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
      for (int ox = ox_lo; ox <= ox_hi; ++ox)
        if (pool_argmax(a, xb, oy, ox, c, nullptr) == iy * a.IW + ix) g += __ldg(dyb + ((int64_t)oy * a.OW + ox) * a.C + c);
    const int64_t o = (int64_t)z * a.xstride + (i - (int64_t)z * per);
    a.dx[o] = g;
    if (a.ph) tc::st1_planes(a.ph + o, a.pl + o, g);
  }
}
