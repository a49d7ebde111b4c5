// Shared device/host helpers for the vsx kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vsx.h"

namespace vsx {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VSX_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      vsx::set_error(__VA_ARGS__);             \
      return VSX_ERR_ARG;                      \
    }                                          \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every converted kernel starts with pdl_launch_dependents() (the NEXT kernel of the stream may be scheduled onto SMs as this grid's CTAs
// retire, instead of after the whole grid has drained and a fresh launch has travelled through the front end) and calls pdl_wait()
// before its first global-memory access (blocks until the PREVIOUS grid has completed and its writes are visible).  Because every kernel
// launched with the attribute waits before touching memory, and a grid cannot complete before its wait has returned, completion order
// stays the stream order transitively; kernels launched without the attribute (torch's own, unconverted ones) serialise as usual.
// VSX_PDL=0 in the environment launches everything without the attribute (the device-side instructions are then no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------- per-segment extents inside ONE launch (multi-architecture batches)
// A batch that carries several sub-architectures is a sequence of row ranges ("segments": consecutive samples that share one set of
// extents).  Row kernels take the whole batch in one launch and look their row's extents up here instead of being launched once per
// segment.  Same layout as vsx_row_segments in include/vsx.h.  count == 0: uniform extents (the scalar kernel arguments apply).
struct RowSegs {
  int count;
  int row_end[VSX_MAX_SEGMENTS];   // exclusive end row of segment i (relative to the first row of the launch)
  int keep[VSX_MAX_SEGMENTS];      // kept channels of segment i; 0: its rows are skipped
  int keep2[VSX_MAX_SEGMENTS];     // second extent where the operation has one
};
#ifdef __CUDACC__
__device__ __forceinline__ int seg_of_row(const RowSegs& s, long r) {
  int i = 0;
  while (i < s.count - 1 && r >= s.row_end[i]) ++i;
  return i;
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }
int num_sms();

// ---------------------------------------------------------------- storage type helpers
template <typename T> struct Store;
template <> struct Store<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Store<bf16> {
  static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 4 consecutive elements as float4 (16-byte aligned for float, 8-byte for bf16)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x), b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&a);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Column-sum flush: colsum[c..c+3] += t for the columns below `keep`.  One 16-byte vector reduction (red.global.add.v4.f32,
// sm_90+) instead of four scalar atomics: thousands of CTAs flush the same few hundred addresses, and the L2 atomic units
// serialise per operation, so the op count -- not the bytes -- bounds the tail of these kernels.
__device__ __forceinline__ void red_add4(float* p, float4 t, int c, int keep) {
  if (c + 3 < keep && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  } else {
    if (c < keep) atomicAdd(p, t.x);
    if (c + 1 < keep) atomicAdd(p + 1, t.y);
    if (c + 2 < keep) atomicAdd(p + 2, t.z);
    if (c + 3 < keep) atomicAdd(p + 3, t.w);
  }
}

// exact (erf) GELU and its derivative -- torch.nn.GELU default used by nets/supernet_blocks.py:17
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Cheap variants for the bf16 training path (GEMM epilogues are CUDA-core / SFU bound at K <= 256): erfc by
// Abramowitz-Stegun 7.1.26 (|abs err| < 4.3e-7 on gelu and gelu' in fp32, i.e. ~1e-4 of a bf16 ulp at 1.0) -- one
// MUFU.RCP + one MUFU.EX2 + 8 FMA-pipe ops instead of erff's ~30; gelu' re-uses the same exponential.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float half_erfc_abs(float x, float& e) {   // 0.5*erfc(|x|/sqrt2); e = exp(-x^2/2)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  e = ex2_approx(z * z * -1.4426950408889634f);      // one MUFU, no range fix-up: underflow to zero is the right answer here
  float p = fmaf(t, 1.061405429f * 0.5f, -1.453152027f * 0.5f);
  p = fmaf(t, p, 1.421413741f * 0.5f);
  p = fmaf(t, p, -0.284496736f * 0.5f);
  p = fmaf(t, p, 0.254829592f * 0.5f);
  return p * t * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float e;
  const float q = half_erfc_abs(x, e);
  return x * (x >= 0.f ? 1.0f - q : q);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float e;
  const float q = half_erfc_abs(x, e);
  return (x >= 0.f ? 1.0f - q : q) + x * 0.39894228040143268f * e;
}
// ---- packed fp32x2 versions (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issued instruction).  The GELU / GELU' epilogues
// of the K <= 256 GEMMs are bound by CUDA-core instruction issue, so halving the FMA-pipe instruction count is a direct win; the two
// SFU operations per element (rcp, ex2) stay scalar.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// (cdf(x0), cdf(x1)) of the standard normal and e = exp(-x^2/2) for a pair, same A-S 7.1.26 evaluation as half_erfc_abs
__device__ __forceinline__ f32x2 normal_cdf2(float x0, float x1, f32x2& e2) {
  const f32x2 x = pk2(x0, x1), ax = pk2(fabsf(x0), fabsf(x1));
  const f32x2 d = fma2(pk2(0.3275911f * 0.70710678118654752f, 0.3275911f * 0.70710678118654752f), ax, pk2(1.0f, 1.0f));
  float d0, d1;
  unpk2(d, d0, d1);
  const f32x2 t = pk2(rcp_approx(d0), rcp_approx(d1));
  const f32x2 zz = mul2(mul2(x, x), pk2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  float z0, z1;
  unpk2(zz, z0, z1);
  e2 = pk2(ex2_approx(z0), ex2_approx(z1));
  f32x2 p = fma2(t, pk2(1.061405429f * 0.5f, 1.061405429f * 0.5f), pk2(-1.453152027f * 0.5f, -1.453152027f * 0.5f));
  p = fma2(t, p, pk2(1.421413741f * 0.5f, 1.421413741f * 0.5f));
  p = fma2(t, p, pk2(-0.284496736f * 0.5f, -0.284496736f * 0.5f));
  p = fma2(t, p, pk2(0.254829592f * 0.5f, 0.254829592f * 0.5f));
  const f32x2 q = mul2(mul2(p, t), e2);                 // 0.5 * erfc(|x| / sqrt 2)
  // cdf = 0.5 + copysign(0.5 - q, x)
  float h0, h1;
  unpk2(add2(pk2(0.5f, 0.5f), mul2(q, pk2(-1.0f, -1.0f))), h0, h1);
  return add2(pk2(0.5f, 0.5f), pk2(copysignf(h0, x0), copysignf(h1, x1)));
}
// ---- MUFU-free GELU / GELU' for the bf16 GEMM epilogues.  The A-S evaluation above needs two SFU operations per element (rcp, ex2);
// the SFU issues one warp instruction per 8 clocks and scheduler, so a 128 x 128 epilogue unit paid 2048 clocks of SFU time -- more
// than its tensor-core time at K <= 256 (tools/gemm_timeline.py: 3300 clocks of epilogue math per unit).  Both functions have the form
// 0.5 + u * P(u^2) with P smooth, so P is evaluated as a polynomial in s = u^2 * (2 / 4.5^2) - 1 on |u| <= 4.5 (u is clamped: beyond
// 4.5 both functions are within 4e-6 of their limits) with packed fp32x2 Horner steps on the FMA pipe only:
//   Phi(u)   = 0.5 + u * P10(s)   |abs err| < 7e-6   (fp32 Horner, Chebyshev least-squares fit; tools/fit_gelu_poly.py)
//   gelu'(u) = 0.5 + u * P11(s)   |abs err| < 5e-5
// i.e. below 1/40 of a bf16 ulp of the values they scale.  The fp32 parity mode keeps erff (gelu_f / gelu_grad_f).
__device__ __forceinline__ f32x2 poly_cdf2(float x0, float x1, f32x2& xc) {
  xc = pk2(fminf(fmaxf(x0, -4.5f), 4.5f), fminf(fmaxf(x1, -4.5f), 4.5f));
  const f32x2 s = fma2(mul2(xc, xc), pk2(0.0987654321f, 0.0987654321f), pk2(-1.0f, -1.0f));
  f32x2 p = pk2(1.074221019e-03f, 1.074221019e-03f);
  p = fma2(p, s, pk2(-2.304612824e-03f, -2.304612824e-03f));
  p = fma2(p, s, pk2(2.497475973e-03f, 2.497475973e-03f));
  p = fma2(p, s, pk2(-5.324486247e-03f, -5.324486247e-03f));
  p = fma2(p, s, pk2(1.161688181e-02f, 1.161688181e-02f));
  p = fma2(p, s, pk2(-1.897149445e-02f, -1.897149445e-02f));
  p = fma2(p, s, pk2(2.822568407e-02f, 2.822568407e-02f));
  p = fma2(p, s, pk2(-4.012141573e-02f, -4.012141573e-02f));
  p = fma2(p, s, pk2(5.470778012e-02f, 5.470778012e-02f));
  p = fma2(p, s, pk2(-7.719306673e-02f, -7.719306673e-02f));
  p = fma2(p, s, pk2(1.569048135e-01f, 1.569048135e-01f));
  return fma2(xc, p, pk2(0.5f, 0.5f));
}
__device__ __forceinline__ f32x2 poly_gelu_grad2(float x0, float x1) {
  const f32x2 xc = pk2(fminf(fmaxf(x0, -4.5f), 4.5f), fminf(fmaxf(x1, -4.5f), 4.5f));
  const f32x2 s = fma2(mul2(xc, xc), pk2(0.0987654321f, 0.0987654321f), pk2(-1.0f, -1.0f));
  f32x2 p = pk2(-6.985080961e-03f, -6.985080961e-03f);
  p = fma2(p, s, pk2(1.433847597e-02f, 1.433847597e-02f));
  p = fma2(p, s, pk2(-1.157771896e-02f, -1.157771896e-02f));
  p = fma2(p, s, pk2(2.306988463e-02f, 2.306988463e-02f));
  p = fma2(p, s, pk2(-5.256838405e-02f, -5.256838405e-02f));
  p = fma2(p, s, pk2(7.393936118e-02f, 7.393936118e-02f));
  p = fma2(p, s, pk2(-8.730810965e-02f, -8.730810965e-02f));
  p = fma2(p, s, pk2(9.659937234e-02f, 9.659937234e-02f));
  p = fma2(p, s, pk2(-9.497554799e-02f, -9.497554799e-02f));
  p = fma2(p, s, pk2(8.712603100e-02f, 8.712603100e-02f));
  p = fma2(p, s, pk2(-8.996616928e-02f, -8.996616928e-02f));
  p = fma2(p, s, pk2(1.594292470e-01f, 1.594292470e-01f));
  return fma2(xc, p, pk2(0.5f, 0.5f));
}
// in place on 32 values: d = gelu'(v), v = gelu(v): both polynomials share the clamp and s, and their two Horner chains interleave
__device__ __forceinline__ void gelu_and_grad32(float (&v)[32], float (&d)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const f32x2 xc = pk2(fminf(fmaxf(v[j], -4.5f), 4.5f), fminf(fmaxf(v[j + 1], -4.5f), 4.5f));
    const f32x2 s = fma2(mul2(xc, xc), pk2(0.0987654321f, 0.0987654321f), pk2(-1.0f, -1.0f));
    f32x2 p = pk2(1.074221019e-03f, 1.074221019e-03f), q = pk2(-6.985080961e-03f, -6.985080961e-03f);
    q = fma2(q, s, pk2(1.433847597e-02f, 1.433847597e-02f));
    p = fma2(p, s, pk2(-2.304612824e-03f, -2.304612824e-03f)), q = fma2(q, s, pk2(-1.157771896e-02f, -1.157771896e-02f));
    p = fma2(p, s, pk2(2.497475973e-03f, 2.497475973e-03f)), q = fma2(q, s, pk2(2.306988463e-02f, 2.306988463e-02f));
    p = fma2(p, s, pk2(-5.324486247e-03f, -5.324486247e-03f)), q = fma2(q, s, pk2(-5.256838405e-02f, -5.256838405e-02f));
    p = fma2(p, s, pk2(1.161688181e-02f, 1.161688181e-02f)), q = fma2(q, s, pk2(7.393936118e-02f, 7.393936118e-02f));
    p = fma2(p, s, pk2(-1.897149445e-02f, -1.897149445e-02f)), q = fma2(q, s, pk2(-8.730810965e-02f, -8.730810965e-02f));
    p = fma2(p, s, pk2(2.822568407e-02f, 2.822568407e-02f)), q = fma2(q, s, pk2(9.659937234e-02f, 9.659937234e-02f));
    p = fma2(p, s, pk2(-4.012141573e-02f, -4.012141573e-02f)), q = fma2(q, s, pk2(-9.497554799e-02f, -9.497554799e-02f));
    p = fma2(p, s, pk2(5.470778012e-02f, 5.470778012e-02f)), q = fma2(q, s, pk2(8.712603100e-02f, 8.712603100e-02f));
    p = fma2(p, s, pk2(-7.719306673e-02f, -7.719306673e-02f)), q = fma2(q, s, pk2(-8.996616928e-02f, -8.996616928e-02f));
    p = fma2(p, s, pk2(1.569048135e-01f, 1.569048135e-01f)), q = fma2(q, s, pk2(1.594292470e-01f, 1.594292470e-01f));
    unpk2(fma2(xc, q, pk2(0.5f, 0.5f)), d[j], d[j + 1]);
    unpk2(mul2(pk2(v[j], v[j + 1]), fma2(xc, p, pk2(0.5f, 0.5f))), v[j], v[j + 1]);
  }
}
// in place on 32 values: v = gelu(v)   /   v = v * gelu'(u)
__device__ __forceinline__ void gelu_fast32(float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    f32x2 xc;
    const f32x2 cdf = poly_cdf2(v[j], v[j + 1], xc);
    unpk2(mul2(pk2(v[j], v[j + 1]), cdf), v[j], v[j + 1]);
  }
}
__device__ __forceinline__ void gelu_grad_fast32(float (&v)[32], const float (&u)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) unpk2(mul2(pk2(v[j], v[j + 1]), poly_gelu_grad2(u[j], u[j + 1])), v[j], v[j + 1]);
}

template <typename OutT> __device__ __forceinline__ float gelu_sel(float x) { return gelu_fast(x); }
template <> __device__ __forceinline__ float gelu_sel<float>(float x) { return gelu_f(x); }
template <typename OutT> __device__ __forceinline__ float gelu_grad_sel(float x) { return gelu_grad_fast(x); }
template <> __device__ __forceinline__ float gelu_grad_sel<float>(float x) { return gelu_grad_f(x); }

// ---------------------------------------------------------------- PTX wrappers: mbarrier / TMA / tcgen05
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (visible as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 3000000000LL) {
      printf("vsx: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier: the row prefetch of the
// HBM-bound row kernels (LayerNorm, gradient casts)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}

// shared -> global tile store / reduce-add (bulk async group); OOB parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---------------------------------------------------------------- CTA pairs (cta_group::2): cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier that may live in the peer CTA (shared::cluster address)
// (.relaxed: the callers order their own tensor-memory reads with tcgen05.wait::ld + fence::before_thread_sync; a .release at cluster scope
// costs ~1800 clocks here because it drains every outstanding shared-memory store of the thread first -- tools/gemm_timeline.py)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA tile load of a CTA pair: data lands in THIS CTA's shared memory (dst: own shared::cta address, valid in the cluster window), the
// bytes are counted on `cluster_bar`, which may be the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.cta_group::2 [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows from each CTA's smem] * B[N x 16: N/2 columns from each CTA's smem]; issued by ONE thread
// of the leader CTA
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the barrier at this shared-memory offset in BOTH CTAs of the pair receives one arrival when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base+i), v[j] = column (col+j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- host: TMA descriptors
// 2-D bf16 tensor [rows, cols] with `ld` elements between rows; box = box_cols x box_rows, 128B swizzle.
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                 uint32_t box_rows, int dtype = VSX_BF16);
// 4-D channels-last bf16 map [B][H][W][C], box = C x box_w x box_h x 1, no swizzle, OOB zero fill / store clipping.
int make_tmap_nhwc(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_w, uint32_t box_h);

}  // namespace vsx
