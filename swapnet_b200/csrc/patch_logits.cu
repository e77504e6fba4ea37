// swapnet_b200 — helpers of the one-output-channel conv (PatchGAN logits, modules/discriminators.py:131:
// Conv2d(ndf*8, 1, kernel_size=4, stride=1, padding=1)).
//
// A conv with ONE output channel is an HBM-bound op that the generic tap GEMM serves badly (the 512-channel
// input tile is re-fetched once per tap for a single output column).  It is evaluated instead as
//   P[px, t]  = sum_c x[px, c] * W[0, c, t]                (ONE 1-tap GEMM, N = 16: x is read once)
//   y[o]      = bias + sum_t P[o + off_t, t]               (tap_sum_fwd, below)
// and backward through the adjoint of the same factorisation:
//   dP[px, t] = dy[px - off_t]                             (tap_shift_pack, below)
//   dW[0,c,t] = sum_px x[px, c] dP[px, t],   dx[px, c] = sum_t dP[px, t] W[0, c, t]     (1-tap wgrad / tap GEMMs)
#include "common.cuh"
#include "../../include/swapnet_b200.h"

void sn_count_launch(int n);

namespace {

constexpr int kThreads = 256;

inline int grid_for(long long total) {
  long long g = (total + kThreads - 1) / kThreads;
  if (g > 148 * 16) g = 148 * 16;
  return g < 1 ? 1 : (int)g;
}

__global__ void tap_sum_fwd_kernel(const float* __restrict__ P, int ppitch, int N, int H, int W, int K, int pad,
                                   const float* __restrict__ bias, float* __restrict__ y, int ypitch) {
  const int OH = H + 2 * pad - K + 1, OW = W + 2 * pad - K + 1;
  const long long total = (long long)N * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW), oh = (int)((i / OW) % OH);
    const long long n = i / ((long long)OW * OH);
    float acc = 0.f;
    for (int kh = 0; kh < K; ++kh) {
      const int h = oh + kh - pad;
      if (h < 0 || h >= H) continue;
      for (int kw = 0; kw < K; ++kw) {
        const int w = ow + kw - pad;
        if (w < 0 || w >= W) continue;
        acc += P[((n * H + h) * W + w) * ppitch + kh * K + kw];
      }
    }
    y[i * ypitch] = acc + (bias ? bias[0] : 0.f);
  }
}

// dP[n,h,w,t] = dy[n, h - kh + pad, w - kw + pad] (0 outside); dy given as split planes (channel 0)
__global__ void tap_shift_pack_kernel(const uint16_t* __restrict__ dy_hi, const uint16_t* __restrict__ dy_lo,
                                      int dypitch, int dyfmt, int N, int H, int W, int K, int pad,
                                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int pitch, int coff,
                                      int fmt) {
  const int OH = H + 2 * pad - K + 1, OW = W + 2 * pad - K + 1;
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    for (int kh = 0; kh < K; ++kh)
      for (int kw = 0; kw < K; ++kw) {
        const int oh = h - kh + pad, ow = w - kw + pad;
        float v = 0.f;
        if (oh >= 0 && oh < OH && ow >= 0 && ow < OW) {
          const long long o = ((n * OH + oh) * OW + ow) * dypitch;
          v = decode16(dy_hi[o], dyfmt) + (dy_lo ? decode16(dy_lo[o], dyfmt) : 0.f);
        }
        uint16_t a, b;
        split16(v, fmt, a, b);
        hi[i * pitch + coff + kh * K + kw] = a;
        lo[i * pitch + coff + kh * K + kw] = b;
      }
  }
}

}  // namespace

#define LAUNCH_CHECK()                         \
  do {                                         \
    sn_count_launch(1);                        \
    SN_CHECK_CUDA(cudaGetLastError());         \
  } while (0)

extern "C" {

int sn_tap_sum_fwd(const float* p, int p_pitch, int n, int h, int w, int k, int pad, const float* bias, float* y,
                   int y_pitch, void* stream) {
  SN_REQUIRE(p && y && k >= 1 && k * k <= p_pitch, "tap_sum: bad arguments");
  const long long total = (long long)n * (h + 2 * pad - k + 1) * (w + 2 * pad - k + 1);
  tap_sum_fwd_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(p, p_pitch, n, h, w, k, pad, bias, y,
                                                                             y_pitch);
  LAUNCH_CHECK();
  return SN_OK;
}

int sn_tap_shift_pack(const void* dy_hi, const void* dy_lo, int dy_pitch, int dy_fmt, int n, int h, int w, int k,
                      int pad, void* dst_hi, void* dst_lo, int dst_pitch, int dst_coff, int fmt, void* stream) {
  SN_REQUIRE(dy_hi && dst_hi && dst_lo && k >= 1 && k * k <= dst_pitch - dst_coff, "tap_shift_pack: bad arguments");
  tap_shift_pack_kernel<<<grid_for((long long)n * h * w), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)dy_hi, (const uint16_t*)dy_lo, dy_pitch, dy_fmt, n, h, w, k, pad, (uint16_t*)dst_hi,
      (uint16_t*)dst_lo, dst_pitch, dst_coff, fmt);
  LAUNCH_CHECK();
  return SN_OK;
}

}  // extern "C"
