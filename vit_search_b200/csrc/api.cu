// C-ABI plumbing shared by all kernels: error text, launch checks, TMA descriptor encoding.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vsx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long g_launches = 0;     // kernels launched through the library (every launcher ends in check_launch)

int check_launch(const char* what) {
  __atomic_fetch_add(&g_launches, 1L, __ATOMIC_RELAXED);     // forward and autograd threads both launch
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

bool pdl_enabled() {
  static const bool on = !(getenv("VSX_PDL") != nullptr && atoi(getenv("VSX_PDL")) == 0);
  return on;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box_cols,
                 uint32_t box_rows, int dtype) {
  const uint64_t es = dtype == VSX_F32 ? 4 : 2;
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VSX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems * es) % 16 != 0 || cols == 0 || rows == 0) {
    set_error("make_tmap_2d: operand must be 16-byte aligned with a 16-byte-multiple pitch (base=%p ld=%llu cols=%llu rows=%llu)",
              base, (unsigned long long)ld_elems, (unsigned long long)cols, (unsigned long long)rows);
    return VSX_ERR_ARG;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * es};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, dtype == VSX_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld_elems, box_cols, box_rows);
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

// 3-D bf16 tensor [batch][rows][cols] (row pitch ld, batch pitch batch_stride elements); box = box_cols x box_rows x 1, 128B
// swizzle.  Coordinates beyond `rows` inside a batch entry are zero-filled, which is what lets the attention kernels load
// 128-row tiles of a 257-token sample without touching the next sample.
int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld_elems,
                 uint64_t batch_stride_elems, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VSX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld_elems * 2) % 16 != 0 || (batch_stride_elems * 2) % 16 != 0 || cols == 0 || rows == 0 ||
      batch == 0) {
    set_error("make_tmap_3d: operand must be 16-byte aligned with 16-byte-multiple pitches (base=%p ld=%llu batch_stride=%llu)", base,
              (unsigned long long)ld_elems, (unsigned long long)batch_stride_elems);
    return VSX_ERR_ARG;
  }
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (cols=%llu rows=%llu batch=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch, (unsigned long long)ld_elems, box_cols, box_rows);
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

// 4-D bf16 map over a [B][rows][heads][head_dim] tensor (the qkv / attention-output layout with heads as their own dimension);
// box = 64 x 1 x box_rows x 1, 128B swizzle.  A box is 64 columns wide whatever head_dim is: columns >= head_dim are out of bounds in
// dimension 0 and therefore zero-filled, as are rows >= `rows` -- a head of width 32 or 48 lands in shared memory as the zero-padded
// 128-byte-row tile the head_dim-64 kernels expect.
int make_tmap_heads(CUtensorMap* m, const void* base, uint64_t head_dim, uint64_t heads, uint64_t rows, uint64_t batch, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VSX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (head_dim * 2) % 16 != 0 || head_dim == 0 || head_dim > 64 || heads == 0 || rows == 0 || batch == 0) {
    set_error("make_tmap_heads: need a 16-byte aligned base and a head_dim that is a multiple of 8, <= 64 (base=%p head_dim=%llu)", base,
              (unsigned long long)head_dim);
    return VSX_ERR_ARG;
  }
  cuuint64_t dims[4] = {head_dim, heads, rows, batch};
  cuuint64_t strides[3] = {head_dim * 2, heads * head_dim * 2, rows * heads * head_dim * 2};
  cuuint32_t box[4] = {64, 1, box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (heads) failed with CUresult %d (head_dim=%llu heads=%llu rows=%llu batch=%llu box_rows=%u)", (int)r,
              (unsigned long long)head_dim, (unsigned long long)heads, (unsigned long long)rows, (unsigned long long)batch, box_rows);
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

// 4-D channels-last bf16 map [B][H][W][C]; box = C x box_w x box_h x 1, no swizzle (dense [box_h][box_w][C] in shared memory).
// Out-of-range coordinates (negative included) are zero-filled on loads and clipped on stores: the halo of a convolution tile.
int make_tmap_nhwc(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint32_t box_w, uint32_t box_h) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VSX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (C * 2) % 16 != 0 || C == 0 || W == 0 || H == 0 || B == 0) {
    set_error("make_tmap_nhwc: map must be 16-byte aligned with a 16-byte-multiple pixel pitch (base=%p C=%llu)", base, (unsigned long long)C);
    return VSX_ERR_ARG;
  }
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (NHWC) failed with CUresult %d (C=%llu W=%llu H=%llu B=%llu box=%ux%u)", (int)r, (unsigned long long)C,
              (unsigned long long)W, (unsigned long long)H, (unsigned long long)B, box_w, box_h);
    return VSX_ERR_CUDA;
  }
  return VSX_OK;
}

}  // namespace vsx

extern "C" const char* vsx_last_error(void) { return vsx::g_err; }
extern "C" long vsx_launch_count(void) { return vsx::g_launches; }
extern "C" int vsx_abi_version(void) { return VSX_ABI_VERSION; }
extern "C" int vsx_device_ok(int dev) {
  int major = 0, count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || dev >= count) {
    vsx::set_error("vsx_device_ok: no CUDA device %d", dev);
    return VSX_ERR_NO_GPU;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    vsx::set_error("vsx_device_ok: device %d has compute capability %d.x; libvsx is built for sm_100a only", dev, major);
    return VSX_ERR_NO_GPU;
  }
  return VSX_OK;
}
