// SwitchTokenMix on the GPU (token_mixup.py:101-162): the batch augmentation that feeds the train step.
//   first half of the batch : a box of patches is pasted from a permuted image (:101-109), the per-patch soft targets switch with it
//                             (:117-123) and the image-level target is mixed with lam = 1 - box area (:125-126)
//   second half             : image-level mixup x*lam + x[perm]*(1-lam) (:131-137), per-patch targets = the mixed target (:141-143)
// The reference spends ~10 full passes over the batch (fancy-index copies, in-place mul/add, repeat, scatter); here the samples are
// read once / twice and written once, and the targets are produced by one small kernel.  All random draws (two permutations, the box,
// two lambdas) stay on the host with the reference's RNG protocol (vit_search_b200/token_mixup.py); arithmetic is fp32 in the
// reference's operation order (x*lam, x'*(1-lam), then the add: no fused multiply-add), so results are bit-identical.
#include "common.cuh"

namespace vsx {
namespace {

__global__ void __launch_bounds__(256) token_mix_samples_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ perm1,
                                                                 const int* __restrict__ perm2, int B, int n1, int CHW, int HW, int W, int py0, int py1,
                                                                 int px0, int px1, float lam2, float one_minus_lam2) {
  pdl_launch_dependents();
  pdl_wait();
  const long total4 = (long)B * CHW / 4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
    const long e = i * 4;
    const int b = (int)(e / CHW);
    const int r = (int)(e - (long)b * CHW);
    const float4 x = ld4(in + e);
    float4 o = x;
    if (b < n1) {
      const int yx = r % HW, y = yx / W, x0 = yx - y * W;      // W % 4 == 0: the four elements share the row
      if (y >= py0 && y < py1 && x0 + 3 >= px0 && x0 < px1) {
        const float4 s = ld4(in + (long)perm1[b] * CHW + r);
        if (x0 >= px0 && x0 < px1) o.x = s.x;
        if (x0 + 1 >= px0 && x0 + 1 < px1) o.y = s.y;
        if (x0 + 2 >= px0 && x0 + 2 < px1) o.z = s.z;
        if (x0 + 3 >= px0 && x0 + 3 < px1) o.w = s.w;
      }
    } else {
      const float4 s = ld4(in + (long)(n1 + perm2[b - n1]) * CHW + r);
      o.x = __fadd_rn(__fmul_rn(x.x, lam2), __fmul_rn(s.x, one_minus_lam2));
      o.y = __fadd_rn(__fmul_rn(x.y, lam2), __fmul_rn(s.y, one_minus_lam2));
      o.z = __fadd_rn(__fmul_rn(x.z, lam2), __fmul_rn(s.z, one_minus_lam2));
      o.w = __fadd_rn(__fmul_rn(x.w, lam2), __fmul_rn(s.w, one_minus_lam2));
    }
    st4(out + e, o);
  }
}

// one thread per (sample, class): the smoothed one-hot of the sample and of its partner, the mixed target, P per-patch targets
__global__ void __launch_bounds__(256) token_mix_targets_kernel(const long* __restrict__ labels, const int* __restrict__ perm1, const int* __restrict__ perm2,
                                                                 float* __restrict__ targets, float* __restrict__ ptargets, int B, int n1, int K, int PL,
                                                                 int by0, int by1, int bx0, int bx1, float on, float off, float lam1,
                                                                 float one_minus_lam1, float lam2, float one_minus_lam2) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)B * K;
  const int P = PL * PL;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i / K), k = (int)(i - (long)b * K);
    const int partner = b < n1 ? perm1[b] : n1 + perm2[b - n1];
    const float y = labels[b] == k ? on : off, yp = labels[partner] == k ? on : off;
    const bool first = b < n1;
    const float t = first ? __fadd_rn(__fmul_rn(y, lam1), __fmul_rn(yp, one_minus_lam1)) : __fadd_rn(__fmul_rn(y, lam2), __fmul_rn(yp, one_minus_lam2));
    targets[i] = t;
    float* pt = ptargets + ((long)b * P) * K + k;
    for (int p = 0; p < P; ++p) {
      const int py = p / PL, px = p - py * PL;
      const bool inbox = py >= by0 && py < by1 && px >= bx0 && px < bx1;
      pt[(long)p * K] = first ? (inbox ? yp : y) : t;
    }
  }
}

}  // namespace
}  // namespace vsx

using namespace vsx;

extern "C" int vsx_token_mix(const float* samples, float* out, const long* labels, const int* perm_patch, const int* perm_image, float* targets,
                             float* patch_targets, int batch, int channels, int height, int width, int patch_len, int num_classes, int box_y0,
                             int box_y1, int box_x0, int box_x1, float on_value, float off_value, double lam_patch, double lam_image, void* stream) {
  VSX_REQUIRE(batch >= 2 && channels > 0 && patch_len > 0 && height % patch_len == 0 && width % patch_len == 0 && width % 4 == 0,
              "vsx_token_mix: need batch >= 2, image sides divisible by patch_len and width %% 4 == 0 (batch=%d %dx%d patch_len=%d)", batch, height,
              width, patch_len);
  VSX_REQUIRE(0 <= box_y0 && box_y0 <= box_y1 && box_y1 <= patch_len && 0 <= box_x0 && box_x0 <= box_x1 && box_x1 <= patch_len,
              "vsx_token_mix: box [%d,%d)x[%d,%d) outside the %dx%d patch grid", box_y0, box_y1, box_x0, box_x1, patch_len, patch_len);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n1 = batch / 2, ph = height / patch_len, pw = width / patch_len;
  const long CHW = (long)channels * height * width;
  VSX_REQUIRE(CHW < (1L << 31), "vsx_token_mix: sample too large");
  // the reference holds lam as a Python float (double), forms 1 - lam in double, and each is rounded to fp32 when it multiplies the fp32
  // tensor: the lambdas cross the ABI as doubles so that (float)(1 - lam) is not formed from an already rounded lam
  const float oml1 = (float)(1.0 - lam_patch), oml2 = (float)(1.0 - lam_image);
  const float lam1f = (float)lam_patch, lam2f = (float)lam_image;
  long blocks = ((long)batch * CHW / 4 + 255) / 256;
  const long cap = (long)num_sms() * 16;
  launch_pdl(token_mix_samples_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, st, samples, out, perm_patch, perm_image, batch, n1, (int)CHW, height * width,
                                                                               width, box_y0 * ph, box_y1 * ph, box_x0 * pw, box_x1 * pw, lam2f, oml2);
  int rc = check_launch("vsx_token_mix");
  if (rc) return rc;
  blocks = ((long)batch * num_classes + 255) / 256;
  launch_pdl(token_mix_targets_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, st, labels, perm_patch, perm_image, targets, patch_targets, batch, n1,
                                                                               num_classes, patch_len, box_y0, box_y1, box_x0, box_x1, on_value, off_value,
                                                                               lam1f, oml1, lam2f, oml2);
  return check_launch("vsx_token_mix");
}
