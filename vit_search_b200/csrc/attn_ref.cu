// Exact-softmax multi-head attention, forward and backward, fp32 math on CUDA cores.
//
// This is the HIGH-PRECISION attention used by the fp32 parity path (dtype VSX_F32) and for head dims the
// tensor-core kernel (attn_mma.cu) does not cover.  One CTA per (sample, head); masked heads (h >= heads_keep)
// are never computed -- their output / gradient slices are zero-filled, which is exactly what the reference's
// head ChannelDrop produces (nets/supernet_blocks.py:111-112) after computing them densely.
//
// Restates nets/supernet_blocks.py:102-109: qkv features ordered (3, H, D); scores = q k^T * scale; softmax over
// keys; out = P v, heads concatenated.  Nothing of size N x N ever goes to HBM: K and V of the head live in shared
// memory (N <= 257, D <= 64), the backward recomputes P from the saved log-sum-exp.
#include "common.cuh"

namespace vsx {
namespace {

constexpr int AT_WARPS = 8;
constexpr int MAX_D = 64;

template <typename T>
__device__ __forceinline__ void load_head_matrix(float* dst, int ldd, const T* src, long ld_src, int N, int D) {
  // dst[j][d] = src[j*ld_src + d]; ldd = D + 1 (bank-conflict padding)
  for (int idx = threadIdx.x; idx < N * (D / 4); idx += blockDim.x) {
    const int j = idx / (D / 4), d4 = (idx % (D / 4)) * 4;
    const float4 v = ld4(src + (long)j * ld_src + d4);
    float* o = dst + j * ldd + d4;
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
  }
}

template <typename T>
__device__ __forceinline__ void zero_head_slice(T* dst, long ld, int N, int D) {
  for (int idx = threadIdx.x; idx < N * (D / 4); idx += blockDim.x) {
    const int j = idx / (D / 4), d4 = (idx % (D / 4)) * 4;
    st4(dst + (long)j * ld + d4, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// grid = (H, B).  qkv rows of sample b start at b*N; row pitch 3*H*D.
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_fwd_ref_kernel(const T* __restrict__ qkv, T* __restrict__ o,
                                                                      float* __restrict__ lse, int N, int H, int D, int Hk,
                                                                      float scale) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const long ldq = 3L * H * D, ldo = (long)H * D;
  T* ob = o + (long)b * N * ldo + h * D;
  if (h >= Hk) {
    zero_head_slice(ob, ldo, N, D);
    return;
  }
  const int ldk = D + 1;
  float* Ks = sm;
  float* Vs = Ks + N * ldk;
  float* wbuf = Vs + N * ldk;                          // per warp: q[MAX_D] + p[Npad]
  const int Npad = (N + 31) / 32 * 32;
  const T* base = qkv + (long)b * N * ldq + h * D;
  load_head_matrix(Ks, ldk, base + (long)H * D, ldq, N, D);
  load_head_matrix(Vs, ldk, base + 2L * H * D, ldq, N, D);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = wbuf + warp * (MAX_D + Npad);
  float* ps = qs + MAX_D;
  for (int i = warp; i < N; i += AT_WARPS) {
    for (int d = lane; d < D; d += 32) qs[d] = Store<T>::ld(base + (long)i * ldq + d) * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
      const float* kr = Ks + j * ldk;
#pragma unroll 8
      for (int d = 0; d < D; ++d) s += qs[d] * kr[d];
      ps[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = __expf(ps[j] - mx);
      ps[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc += ps[j] * Vs[j * ldk + d];
      Store<T>::st(ob + (long)i * ldo + d, acc * inv);
    }
    if (lane == 0) lse[((long)b * H + h) * N + i] = mx + __logf(sum);
    __syncwarp();
  }
}

// Backward.  dq_i = scale * sum_j ds_ij k_j ; dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i p_ij do_i,
// ds_ij = p_ij (do_i . v_j - delta_i), delta_i = do_i . o_i, p_ij = exp(scale q_i.k_j - lse_i).
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_ref_kernel(const T* __restrict__ qkv, const T* __restrict__ o,
                                                                      const T* __restrict__ d_o, const float* __restrict__ lse,
                                                                      T* __restrict__ dqkv, int N, int H, int D, int Hk, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y;
  const long ldq = 3L * H * D, ldo = (long)H * D;
  T* dbase = dqkv + (long)b * N * ldq + h * D;
  if (h >= Hk) {
    zero_head_slice(dbase, ldq, N, D);
    zero_head_slice(dbase + (long)H * D, ldq, N, D);
    zero_head_slice(dbase + 2L * H * D, ldq, N, D);
    return;
  }
  const int ldk = D + 1;
  const int Npad = (N + 31) / 32 * 32;
  float* M0 = sm;                    // phase A: K      phase B: Q
  float* M1 = M0 + N * ldk;          // phase A: V      phase B: dO
  float* lse_s = M1 + N * ldk;
  float* delta_s = lse_s + Npad;
  float* wbuf = delta_s + Npad;      // per warp: a[MAX_D] b[MAX_D] p[Npad] ds[Npad]
  const T* base = qkv + (long)b * N * ldq + h * D;
  const T* ob = o + (long)b * N * ldo + h * D;
  const T* dob = d_o + (long)b * N * ldo + h * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* va = wbuf + warp * (2 * MAX_D + 2 * Npad);
  float* vb = va + MAX_D;
  float* ps = vb + MAX_D;
  float* dss = ps + Npad;

  for (int i = warp; i < N; i += AT_WARPS) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc += Store<T>::ld(dob + (long)i * ldo + d) * Store<T>::ld(ob + (long)i * ldo + d);
    acc = warp_sum(acc);
    if (lane == 0) {
      delta_s[i] = acc;
      lse_s[i] = lse[((long)b * H + h) * N + i];
    }
  }
  // ---- phase A: dq (K, V resident)
  load_head_matrix(M0, ldk, base + (long)H * D, ldq, N, D);
  load_head_matrix(M1, ldk, base + 2L * H * D, ldq, N, D);
  __syncthreads();
  for (int i = warp; i < N; i += AT_WARPS) {
    for (int d = lane; d < D; d += 32) {
      va[d] = Store<T>::ld(base + (long)i * ldq + d);
      vb[d] = Store<T>::ld(dob + (long)i * ldo + d);
    }
    __syncwarp();
    const float l = lse_s[i], dl = delta_s[i];
    for (int j = lane; j < N; j += 32) {
      float s = 0.f, dp = 0.f;
      const float* kr = M0 + j * ldk;
      const float* vr = M1 + j * ldk;
#pragma unroll 8
      for (int d = 0; d < D; ++d) {
        s += va[d] * kr[d];
        dp += vb[d] * vr[d];
      }
      const float p = __expf(s * scale - l);
      dss[j] = p * (dp - dl) * scale;
    }
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc += dss[j] * M0[j * ldk + d];
      Store<T>::st(dbase + (long)i * ldq + d, acc);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- phase B: dk, dv (Q, dO resident)
  load_head_matrix(M0, ldk, base, ldq, N, D);
  load_head_matrix(M1, ldk, dob, ldo, N, D);
  __syncthreads();
  for (int j = warp; j < N; j += AT_WARPS) {
    for (int d = lane; d < D; d += 32) {
      va[d] = Store<T>::ld(base + (long)H * D + (long)j * ldq + d);       // k_j
      vb[d] = Store<T>::ld(base + 2L * H * D + (long)j * ldq + d);        // v_j
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      float s = 0.f, dp = 0.f;
      const float* qr = M0 + i * ldk;
      const float* dr = M1 + i * ldk;
#pragma unroll 8
      for (int d = 0; d < D; ++d) {
        s += qr[d] * va[d];
        dp += dr[d] * vb[d];
      }
      const float p = __expf(s * scale - lse_s[i]);
      ps[i] = p;
      dss[i] = p * (dp - delta_s[i]) * scale;
    }
    __syncwarp();
    for (int d = lane; d < D; d += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < N; ++i) {
        ak += dss[i] * M0[i * ldk + d];
        av += ps[i] * M1[i * ldk + d];
      }
      Store<T>::st(dbase + (long)H * D + (long)j * ldq + d, ak);
      Store<T>::st(dbase + 2L * H * D + (long)j * ldq + d, av);
    }
    __syncwarp();
  }
}

size_t fwd_smem(int N, int D) {
  const int Npad = (N + 31) / 32 * 32;
  return sizeof(float) * (2 * N * (D + 1) + AT_WARPS * (MAX_D + Npad));
}
size_t bwd_smem(int N, int D) {
  const int Npad = (N + 31) / 32 * 32;
  return sizeof(float) * (2 * N * (D + 1) + 2 * Npad + AT_WARPS * (2 * MAX_D + 2 * Npad));
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) {
    set_error("%s: needs %zu bytes of shared memory (> 227 KB); tokens x head_dim too large", what, bytes);
    return VSX_ERR_ARG;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

}  // namespace

template <typename T>
int attn_fwd_ref(const void* qkv, void* o, float* lse, int B, int N, int H, int D, int Hk, float scale, cudaStream_t st) {
  const size_t smem = fwd_smem(N, D);
  int rc = set_smem(attn_fwd_ref_kernel<T>, smem, "vsx_attn_fwd");
  if (rc) return rc;
  launch_pdl(attn_fwd_ref_kernel<T>, dim3(dim3(H, B)), dim3(AT_WARPS * 32), smem, st, (const T*)qkv, (T*)o, lse, N, H, D, Hk, scale);
  return check_launch("vsx_attn_fwd");
}
template <typename T>
int attn_bwd_ref(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int N, int H, int D, int Hk,
                 float scale, cudaStream_t st) {
  const size_t smem = bwd_smem(N, D);
  int rc = set_smem(attn_bwd_ref_kernel<T>, smem, "vsx_attn_bwd");
  if (rc) return rc;
  launch_pdl(attn_bwd_ref_kernel<T>, dim3(dim3(H, B)), dim3(AT_WARPS * 32), smem, st, (const T*)qkv, (const T*)o, (const T*)d_o, lse, (T*)dqkv, N, H, D, Hk, scale);
  return check_launch("vsx_attn_bwd");
}

template int attn_fwd_ref<float>(const void*, void*, float*, int, int, int, int, int, float, cudaStream_t);
template int attn_fwd_ref<bf16>(const void*, void*, float*, int, int, int, int, int, float, cudaStream_t);
template int attn_bwd_ref<float>(const void*, const void*, const void*, const float*, void*, int, int, int, int, int, float, cudaStream_t);
template int attn_bwd_ref<bf16>(const void*, const void*, const void*, const float*, void*, int, int, int, int, int, float, cudaStream_t);

}  // namespace vsx
