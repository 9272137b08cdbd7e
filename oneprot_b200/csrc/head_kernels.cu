// HBM-bound row kernels of the projection heads in front of the ClipLoss path (SURVEY.md section 8f,
// reference src/models/components/base_encoder.py:107-194): masked mean pooling, LayerNorm and GELU,
// forward and backward.  The Linear layers between them run on the tcgen05 GEMM of clip_kernels.cu
// (oneprot_gemm_bf16_ex); the L2-normalise / logit-scale epilogue after them is there as well.
//
// All kernels: 16-byte vector loads (8 bf16 or 2 x 4 fp32 per lane), fp32 arithmetic, warp-shuffle
// reductions, one warp per row where a row reduction is needed; column reductions (d gamma, d beta)
// go through per-chunk partial slots that are summed in a fixed order (deterministic).
#if !defined(ONEPROT_KERNEL_EMULATION) || defined(ONEPROT_HOST_EMULATION)   // tests/emu: the kernel bodies below are also compiled for the CPU
#include "host_trace.h"
#include "../../include/oneprot_clip.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

namespace opint {
int fail(int code, const std::string& msg);
void count_launch(int n);
}  // namespace opint
#endif

#include "row_regs.cuh"

namespace oph {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <bool FP32>
__device__ __forceinline__ void load8(const void* base, size_t idx, float (&f)[8]) {
  if (FP32) {
    const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + idx);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}

template <bool FP32>
__device__ __forceinline__ void store8(void* base, size_t idx, const float (&f)[8]) {
  if (FP32) {
    *reinterpret_cast<float4*>(static_cast<float*>(base) + idx) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(static_cast<float*>(base) + idx + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + idx) =
        make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
}

constexpr int LN_MAXC = 8;   // 8 chunks x 32 lanes x 8 elements = rows of up to 2048 elements held in registers

using oprow::Vec8;

// ---- LayerNorm forward: y = (x - mean) * rstd * gamma + beta, one warp per row, one read of x.
// C = row chunks of 256 elements held in registers (ceil(d / 256) <= C); the loads of a row are all issued first.
template <bool FP32, int C>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ gamma,
                                                            const void* __restrict__ beta, void* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int rows, int d,
                                                            float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const size_t base = static_cast<size_t>(row) * d;
    Vec8<FP32> v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane * 8 + c * 256;
      if (k < d) v[c].load(x, base + k);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (lane * 8 + c * 256 < d) {
        float f[8];
        v[c].unpack(f);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += f[u];
      }
    }
    const float mu = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (lane * 8 + c * 256 < d) {
        float f[8];
        v[c].unpack(f);
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float t = f[u] - mu; q = fmaf(t, t, q); }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);     // biased variance, eps inside the root (torch.nn.LayerNorm)
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane * 8 + c * 256;
      if (k < d) {
        float f[8], g[8], b[8], o[8];
        v[c].unpack(f);
        load8<FP32>(gamma, k, g);
        load8<FP32>(beta, k, b);
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = fmaf((f[u] - mu) * rs, g[u], b[u]);
        store8<FP32>(y, base + k, o);
      }
    }
  }
}

// ---- LayerNorm backward in ONE pass over x and gy (3 d bytes per row: read x, read gy, write gx):
//   gx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = gy * gamma        (row part, one warp per row)
//   d gamma = sum_rows gy * xhat, d beta = sum_rows gy                         (column part)
// The column sums of the rows a warp handles live in that warp's PRIVATE slice of shared memory (no atomics: a lane
// only ever touches its own addresses, rows are dealt to warps in a fixed order), the 8 warps of a block are added in
// warp order into one partial slot per block, and sum_slots_f32_kernel adds the slots in block order - deterministic.
// Dynamic shared memory: 8 warps x 2 x C x 256 floats.  gx may be null (only the column part) and so may dg_part.
#ifdef ONEPROT_KERNEL_EMULATION
#define HD_DYNAMIC_SMEM(name) static float name[8 * 2 * LN_MAXC * 256]
#else
#define HD_DYNAMIC_SMEM(name) extern __shared__ float name[]
#endif
template <bool FP32, int C>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* __restrict__ x, const void* __restrict__ gy,
                                                            const void* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, void* __restrict__ gx,
                                                            float* __restrict__ dg_part, float* __restrict__ db_part, int ld,
                                                            int rows, int d) {
  HD_DYNAMIC_SMEM(acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);
  // warp-private accumulators, laid out so that the float4 accesses of a warp are conflict-free:
  // [warp][dg | db][chunk][half 0 | 1][lane][4]
  float* my = acc + static_cast<size_t>(warp) * (2 * C * 256);
  if (dg_part) {
#pragma unroll
    for (int i = 0; i < 2 * C * 2; ++i) *reinterpret_cast<float4*>(my + (i * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * wpb + warp; row < rows; row += gridDim.x * wpb) {
    const size_t base = static_cast<size_t>(row) * d;
    Vec8<FP32> vx[C], vg[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane * 8 + c * 256;
      if (k < d) { vx[c].load(x, base + k); vg[c].load(gy, base + k); }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int k = lane * 8 + c * 256;
      if (k < d) {
        float fx[8], fg[8], gm[8];
        vx[c].unpack(fx);
        vg[c].unpack(fg);
        load8<FP32>(gamma, k, gm);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float xh = (fx[u] - mu) * rs, g = fg[u] * gm[u];
          s1 += g;
          s2 = fmaf(g, xh, s2);
        }
        if (dg_part) {
          float* pg = my + ((c * 2) * 32 + lane) * 4;
          float* pb = my + ((C * 2 + c * 2) * 32 + lane) * 4;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 ag = *reinterpret_cast<float4*>(pg + h * 128), ab = *reinterpret_cast<float4*>(pb + h * 128);
            ag.x = fmaf(fg[4 * h + 0], (fx[4 * h + 0] - mu) * rs, ag.x); ab.x += fg[4 * h + 0];
            ag.y = fmaf(fg[4 * h + 1], (fx[4 * h + 1] - mu) * rs, ag.y); ab.y += fg[4 * h + 1];
            ag.z = fmaf(fg[4 * h + 2], (fx[4 * h + 2] - mu) * rs, ag.z); ab.z += fg[4 * h + 2];
            ag.w = fmaf(fg[4 * h + 3], (fx[4 * h + 3] - mu) * rs, ag.w); ab.w += fg[4 * h + 3];
            *reinterpret_cast<float4*>(pg + h * 128) = ag;
            *reinterpret_cast<float4*>(pb + h * 128) = ab;
          }
        }
      }
    }
    if (gx) {
      const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int k = lane * 8 + c * 256;
        if (k < d) {
          float fx[8], fg[8], gm[8], o[8];
          vx[c].unpack(fx);
          vg[c].unpack(fg);
          load8<FP32>(gamma, k, gm);
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = rs * (fg[u] * gm[u] - c1 - (fx[u] - mu) * rs * c2);
          store8<FP32>(gx, base + k, o);
        }
      }
    }
  }
  if (dg_part) {
    __syncthreads();
    // thread t adds column t of every 256-column chunk over the 8 warps (warp order), one slot per block
    const int t = threadIdx.x;
    const int l = t >> 3, u = t & 7;                         // column t of a chunk = lane l, element u
    const int off = ((u >> 2) * 32 + l) * 4 + (u & 3);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int col = c * 256 + t;
      if (col < d) {
        float sg = 0.f, sb = 0.f;
        for (int w = 0; w < wpb; ++w) {
          const float* base_w = acc + static_cast<size_t>(w) * (2 * C * 256);
          sg += base_w[(c * 2) * 128 + off];
          sb += base_w[(C * 2 + c * 2) * 128 + off];
        }
        dg_part[static_cast<size_t>(blockIdx.x) * ld + col] = sg;
        db_part[static_cast<size_t>(blockIdx.x) * ld + col] = sb;
      }
    }
  }
}

// ---- rows longer than 2048 elements (ESM-2 3B / 15B: 2560 / 5120): same math, the row is re-read from
// L1 / L2 instead of being held in registers (two extra passes forward, one backward)
template <bool FP32>
__global__ void layernorm_fwd_long_kernel(const void* __restrict__ x, const void* __restrict__ gamma, const void* __restrict__ beta,
                                          void* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int rows, int d,
                                          float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const size_t base = static_cast<size_t>(row) * d;
    float s = 0.f;
    for (int k = lane * 8; k < d; k += 256) {
      float f[8];
      load8<FP32>(x, base + k, f);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += f[u];
    }
    const float mu = warp_sum(s) * inv_d;
    float q = 0.f;
    for (int k = lane * 8; k < d; k += 256) {
      float f[8];
      load8<FP32>(x, base + k, f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { const float t = f[u] - mu; q = fmaf(t, t, q); }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    for (int k = lane * 8; k < d; k += 256) {
      float f[8], g[8], b[8], o[8];
      load8<FP32>(x, base + k, f);
      load8<FP32>(gamma, k, g);
      load8<FP32>(beta, k, b);
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = fmaf((f[u] - mu) * rs, g[u], b[u]);
      store8<FP32>(y, base + k, o);
    }
  }
}

template <bool FP32>
__global__ void layernorm_bwd_rows_long_kernel(const void* __restrict__ x, const void* __restrict__ gy, const void* __restrict__ gamma,
                                               const float* __restrict__ mean, const float* __restrict__ rstd, void* __restrict__ gx,
                                               int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float inv_d = 1.f / static_cast<float>(d);
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const size_t base = static_cast<size_t>(row) * d;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane * 8; k < d; k += 256) {
      float fx[8], fg[8], gm[8];
      load8<FP32>(x, base + k, fx);
      load8<FP32>(gy, base + k, fg);
      load8<FP32>(gamma, k, gm);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float g = fg[u] * gm[u];
        s1 += g;
        s2 = fmaf(g, (fx[u] - mu) * rs, s2);
      }
    }
    const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
    for (int k = lane * 8; k < d; k += 256) {
      float fx[8], fg[8], gm[8], o[8];
      load8<FP32>(x, base + k, fx);
      load8<FP32>(gy, base + k, fg);
      load8<FP32>(gamma, k, gm);
#pragma unroll
      for (int u = 0; u < 8; ++u) o[u] = rs * (fg[u] * gm[u] - c1 - (fx[u] - mu) * rs * c2);
      store8<FP32>(gx, base + k, o);
    }
  }
}

// ---- LayerNorm backward, column part: partial d gamma = sum_rows gy * xhat, d beta = sum_rows gy.
// grid (ceil(d / 256), row chunks); block = 8 warps x (32 lanes x 8 columns); slot = blockIdx.y.
template <bool FP32>
__global__ void layernorm_bwd_cols_kernel(const void* __restrict__ x, const void* __restrict__ gy, const float* __restrict__ mean,
                                          const float* __restrict__ rstd, int rows, int d, int rows_per_chunk,
                                          float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, int ld) {
  __shared__ float red[2][8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float ag[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ab[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (k < d) {
    for (int row = r0 + warp; row < r1; row += 8) {
      const size_t base = static_cast<size_t>(row) * d + k;
      const float mu = mean[row], rs = rstd[row];
      float fx[8], fg[8];
      load8<FP32>(x, base, fx);
      load8<FP32>(gy, base, fg);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        ag[u] = fmaf(fg[u], (fx[u] - mu) * rs, ag[u]);
        ab[u] += fg[u];
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) { red[0][warp][lane * 8 + u] = ag[u]; red[1][warp][lane * 8 + u] = ab[u]; }
  __syncthreads();
  // 256 threads: thread t sums column t of the tile over the 8 warps (fixed order)
  const int t = threadIdx.x, col = blockIdx.x * 256 + t;
  if (col < d) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { sg += red[0][w][t]; sb += red[1][w][t]; }
    dgamma_part[static_cast<size_t>(blockIdx.y) * ld + col] = sg;
    dbeta_part[static_cast<size_t>(blockIdx.y) * ld + col] = sb;
  }
}

// out[k] = sum_s part[s * ld + k] in slot order (deterministic); one thread per column
__global__ void sum_slots_f32_kernel(const float* __restrict__ part, int slots, int ld, int count, float* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  float acc = 0.f;
  for (int s = 0; s < slots; ++s) acc += part[static_cast<size_t>(s) * ld + k];
  out[k] = acc;
}

// ---- L1 regulariser of the training step: mean |x| over all elements (oneprot_module.py:43-44, 99-101) ----
// Every block adds |x| over a fixed grid-stride slice and writes one partial; the last kernel adds the partials in
// block order and divides (deterministic).  Algorithmic bytes: one read of x forward, one read + one write backward.
constexpr int ABS_MAX_BLOCKS = 4096;
template <bool FP32>
__global__ void __launch_bounds__(256) abs_sum_kernel(const void* __restrict__ x, size_t total8, float* __restrict__ part) {
  __shared__ float red[8];
  float acc = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[8];
    load8<FP32>(x, i * 8, f);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += fabsf(f[u]);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    part[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) abs_mean_final_kernel(const float* __restrict__ part, int blocks, float inv_count, float* __restrict__ out) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < blocks; i += 256) acc += part[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += red[i];
    out[0] = t * inv_count;
  }
}

// gx = g[0] * coef * sign(x)  (sign(0) = 0, the gradient torch.abs uses)
template <bool FP32>
__global__ void abs_mean_bwd_kernel(const void* __restrict__ x, const float* __restrict__ g, float coef, void* __restrict__ gx, size_t total8) {
  const float c = g[0] * coef;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[8], o[8];
    load8<FP32>(x, i * 8, f);
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = f[u] > 0.f ? c : (f[u] < 0.f ? -c : 0.f);
    store8<FP32>(gx, i * 8, o);
  }
}

// ---- GELU (exact, erf) -------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
  return fmaf(v, pdf, cdf);
}

// 4 independent 16-byte vectors per thread and step: all loads are issued before the first erf (bytes in flight)
template <bool FP32, bool BWD>
__global__ void __launch_bounds__(256) gelu_kernel(const void* __restrict__ x, const void* __restrict__ gy, void* __restrict__ out,
                                                   size_t total8) {
  constexpr int U = 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i0 = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i0 < total8; i0 += U * stride) {
    Vec8<FP32> vx[U], vg[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const size_t i = i0 + j * stride;
      if (i < total8) {
        vx[j].load(x, i * 8);
        if (BWD) vg[j].load(gy, i * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const size_t i = i0 + j * stride;
      if (i < total8) {
        float f[8], o[8];
        vx[j].unpack(f);
        if (BWD) {
          float g[8];
          vg[j].unpack(g);
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = g[u] * gelu_grad_f(f[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = gelu_f(f[u]);
        }
        store8<FP32>(out, i * 8, o);
      }
    }
  }
}

// ---- masked mean pooling over the token axis: y[b] = sum_l mask[b,l] x[b,l] / sum_l mask[b,l] ----
// grid (B, ceil(D / 256)); 8 warps stride over the tokens, lane = 8 columns; mask == nullptr: plain mean
template <bool FP32>
__global__ void meanpool_fwd_kernel(const void* __restrict__ x, const float* __restrict__ mask, void* __restrict__ y,
                                    float* __restrict__ inv_count, int L, int D, int normalize) {
  __shared__ float red[8][256 + 8];
  __shared__ float cnt_s[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = blockIdx.y * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float cnt = 0.f;
  // 8 tokens per warp and step: their loads are all issued before the first accumulation (bytes in flight; the tokens
  // are still added in increasing order, so the result does not depend on the unrolling)
  constexpr int U = 8;
  for (int l0 = warp; l0 < L; l0 += 8 * U) {
    Vec8<FP32> v[U];
    float m[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int l = l0 + 8 * j;
      m[j] = (l < L) ? (mask ? mask[static_cast<size_t>(b) * L + l] : 1.f) : 0.f;
      if (m[j] != 0.f && k < D) v[j].load(x, (static_cast<size_t>(b) * L + l) * D + k);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      cnt += m[j];
      if (m[j] != 0.f && k < D) {
        float f[8];
        v[j].unpack(f);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(m[j], f[u], acc[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[warp][lane * 8 + u] = acc[u];
  if (lane == 0) cnt_s[warp] = cnt;
  __syncthreads();
  float total = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) total += cnt_s[w];
  // normalize = 0: plain weighted sum (attention pooling, d weight of its score layer)
  const float inv = normalize ? 1.f / total : 1.f;   // all-masked rows give inf / nan like the reference's 0 / 0
  if (threadIdx.x == 0 && blockIdx.y == 0 && inv_count) inv_count[b] = inv;
  if (warp == 0 && k < D) {
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][lane * 8 + u];
      o[u] = s * inv;
    }
    store8<FP32>(y, static_cast<size_t>(b) * D + k, o);
  }
}

template <bool FP32>
__global__ void meanpool_bwd_kernel(const void* __restrict__ gy, const float* __restrict__ mask, const float* __restrict__ inv_count,
                                    void* __restrict__ gx, int L, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = blockIdx.y * 256 + lane * 8;
  if (k >= D) return;
  float g[8];
  load8<FP32>(gy, static_cast<size_t>(b) * D + k, g);
  const float inv = inv_count[b];
  for (int l = warp; l < L; l += 8) {
    const float m = (mask ? mask[static_cast<size_t>(b) * L + l] : 1.f) * inv;
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = m * g[u];
    store8<FP32>(gx, (static_cast<size_t>(b) * L + l) * D + k, o);
  }
}

// ---- Attention1dPooling (base_encoder.py:84-104): score_l = <w, x_l> + bias, masked softmax over the
// tokens, weighted sum.  token_dot: one warp per token; vec_stride = 0: one vector for all batches
// (the score layer), D: one vector per batch row (d p = <g_b, x_l> in the backward).
template <bool FP32>
__global__ void token_dot_kernel(const void* __restrict__ x, const void* __restrict__ vec, size_t vec_stride,
                                 const float* __restrict__ bias, const float* __restrict__ mask, float* __restrict__ out,
                                 int B, int L, int D) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int tokens = B * L;
  const float bs = bias ? *bias : 0.f;
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < tokens; t += gridDim.x * wpb) {
    if (mask && mask[t] == 0.f) {          // masked_fill_(~mask, -inf), base_encoder.py:97-100
      if (lane == 0) out[t] = -INFINITY;
      continue;
    }
    const size_t vb = static_cast<size_t>(t / L) * vec_stride;
    float acc = 0.f;
    for (int k = lane * 8; k < D; k += 256) {
      float fx[8], fv[8];
      load8<FP32>(x, static_cast<size_t>(t) * D + k, fx);
      load8<FP32>(vec, vb + k, fv);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fmaf(fx[u], fv[u], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[t] = acc + bs;
  }
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();               // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// one block per batch row: p = softmax(s) over L (fp32, any L)
__global__ void softmax_rows_kernel(const float* __restrict__ s, float* __restrict__ p, int L) {
  __shared__ float red[32];
  const float* sr = s + static_cast<size_t>(blockIdx.x) * L;
  float* pr = p + static_cast<size_t>(blockIdx.x) * L;
  float m = -INFINITY;
  for (int l = threadIdx.x; l < L; l += blockDim.x) m = fmaxf(m, sr[l]);
  m = block_reduce(m, true, red);
  float z = 0.f;
  for (int l = threadIdx.x; l < L; l += blockDim.x) z += __expf(sr[l] - m);
  z = block_reduce(z, false, red);
  const float inv = 1.f / z;
  for (int l = threadIdx.x; l < L; l += blockDim.x) pr[l] = __expf(sr[l] - m) * inv;
}

// ds = p * (dp - sum_k p_k dp_k)
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ ds, int L) {
  __shared__ float red[32];
  const size_t base = static_cast<size_t>(blockIdx.x) * L;
  float acc = 0.f;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float pl = p[base + l];
    if (pl != 0.f) acc = fmaf(pl, dp[base + l], acc);     // masked tokens: p = 0, dp may be anything
  }
  acc = block_reduce(acc, false, red);
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float pl = p[base + l];
    ds[base + l] = (pl != 0.f) ? pl * (dp[base + l] - acc) : 0.f;
  }
}

// gx[b,l,:] = p[b,l] * g[b,:] + ds[b,l] * w[:]
template <bool FP32>
__global__ void attnpool_bwd_x_kernel(const void* __restrict__ g, const float* __restrict__ p, const float* __restrict__ ds,
                                      const void* __restrict__ w, void* __restrict__ gx, int L, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = blockIdx.y * 256 + lane * 8;
  if (k >= D) return;
  float fg[8], fw[8];
  load8<FP32>(g, static_cast<size_t>(b) * D + k, fg);
  load8<FP32>(w, k, fw);
  for (int l = warp; l < L; l += 8) {
    const float pl = p[static_cast<size_t>(b) * L + l], dl = ds[static_cast<size_t>(b) * L + l];
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = fmaf(pl, fg[u], dl * fw[u]);
    store8<FP32>(gx, (static_cast<size_t>(b) * L + l) * D + k, o);
  }
}

}  // namespace oph

#if !defined(ONEPROT_KERNEL_EMULATION) || defined(ONEPROT_HOST_EMULATION)
namespace {
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
#define HD_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return opint::fail(ONEPROT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Resident blocks per SM of a row kernel (occupancy API, cached per instantiation): the grid is ONE wave of persistent
// blocks whose warps stride over the rows, so no tail wave runs at a fraction of the machine.
template <typename Kern>
int resident_blocks(Kern kern, int& cache, size_t smem) {
  if (!cache) {
    int nb = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, smem) != cudaSuccess || nb < 1) nb = 1;
    cache = nb;
  }
  return cache;
}

template <bool FP32, int C>
void launch_ln_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd, int rows, int d, float eps,
                   cudaStream_t st) {
  static int nb = 0;
  const int blocks = std::min(cdiv(rows, 8), oneprot_num_sms() * resident_blocks(oph::layernorm_fwd_kernel<FP32, C>, nb, 0));
  oph::layernorm_fwd_kernel<FP32, C><<<blocks, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, d, eps);
}

constexpr int LN_BWD_MAX_BLOCKS_PER_SM = 8;
inline int ln_bwd_max_blocks(int rows) { return std::max(1, std::min(cdiv(rows, 8), oneprot_num_sms() * LN_BWD_MAX_BLOCKS_PER_SM)); }

// -> number of partial slots written (= blocks) when dgamma is wanted
template <bool FP32, int C>
int launch_ln_bwd(const void* x, const void* gy, const void* gamma, const float* mean, const float* rstd, void* gx, float* pg, float* pb,
                  int ld, int rows, int d, cudaStream_t st) {
  static int nb = 0;
  const size_t smem = pg ? sizeof(float) * 8 * 2 * C * 256 : 0;
  static int nb_nosmem = 0;
  const int per_sm = std::min(LN_BWD_MAX_BLOCKS_PER_SM, resident_blocks(oph::layernorm_bwd_kernel<FP32, C>, pg ? nb : nb_nosmem, smem));
  const int blocks = std::max(1, std::min(cdiv(rows, 8), oneprot_num_sms() * per_sm));
  oph::layernorm_bwd_kernel<FP32, C><<<blocks, 256, smem, st>>>(x, gy, gamma, mean, rstd, gx, pg, pb, ld, rows, d);
  return blocks;
}

#define LN_SWITCH_C(cneed, CALL)                                            \
  switch (cneed) {                                                          \
    case 1: CALL(1); break;                                                 \
    case 2: CALL(2); break;                                                 \
    case 3: CALL(3); break;                                                 \
    case 4: CALL(4); break;                                                 \
    case 5: CALL(5); break;                                                 \
    case 6: CALL(6); break;                                                 \
    default: CALL(8); break;                                                \
  }

int ln_row_chunks(int rows, int d) {
  // enough (column tile, row chunk) blocks for ~2 per SM, at least 64 rows per chunk
  const int tiles = cdiv(d, 256);
  const int want = std::max(1, 2 * oneprot_num_sms() / tiles);
  return std::max(1, std::min(cdiv(rows, 64), want));
}
}  // namespace

extern "C" {

int oneprot_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd, int rows, int d,
                          int is_fp32, float eps, void* stream) {
  if (!x || !gamma || !beta || !y || !mean || !rstd) return opint::fail(ONEPROT_ERR_ARG, "layernorm_fwd: null pointer");
  if (rows <= 0 || d <= 0 || d % 8) return opint::fail(ONEPROT_ERR_ARG, "layernorm_fwd: need d a positive multiple of 8");
  if (!al16(x) || !al16(gamma) || !al16(beta) || !al16(y)) return opint::fail(ONEPROT_ERR_ARG, "layernorm_fwd: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("layernorm_fwd x=%p gamma=%p beta=%p y=%p mean=%p rstd=%p rows=%d d=%d fp32=%d st=%p", x, gamma, beta, y, (void*)mean, (void*)rstd, rows, d, is_fp32, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const int blocks = std::min(cdiv(rows, 8), oneprot_num_sms() * 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d > 256 * oph::LN_MAXC) {       // row does not fit the register-resident kernel
    if (is_fp32) oph::layernorm_fwd_long_kernel<true><<<blocks, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, d, eps);
    else oph::layernorm_fwd_long_kernel<false><<<blocks, 256, 0, st>>>(x, gamma, beta, y, mean, rstd, rows, d, eps);
  } else if (is_fp32) {
#define LN_CALL(CC) launch_ln_fwd<true, CC>(x, gamma, beta, y, mean, rstd, rows, d, eps, st)
    LN_SWITCH_C(cdiv(d, 256), LN_CALL)
#undef LN_CALL
  } else {
#define LN_CALL(CC) launch_ln_fwd<false, CC>(x, gamma, beta, y, mean, rstd, rows, d, eps, st)
    LN_SWITCH_C(cdiv(d, 256), LN_CALL)
#undef LN_CALL
  }
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

size_t oneprot_layernorm_bwd_scratch_bytes(int rows, int d) {
  if (rows <= 0 || d <= 0) return 0;
  const size_t ld = static_cast<size_t>(cdiv(d, 256)) * 256;
  const int slots = d > 256 * oph::LN_MAXC ? ln_row_chunks(rows, d) : ln_bwd_max_blocks(rows);     // one partial slot per block
  return 2 * static_cast<size_t>(slots) * ld * sizeof(float);
}

int oneprot_layernorm_bwd(const void* x, const void* gy, const void* gamma, const float* mean, const float* rstd, void* gx,
                          float* dgamma, float* dbeta, void* scratch, size_t scratch_bytes, int rows, int d, int is_fp32,
                          void* stream) {
  if (!x || !gy || !gamma || !mean || !rstd) return opint::fail(ONEPROT_ERR_ARG, "layernorm_bwd: null pointer");
  if (rows <= 0 || d <= 0 || d % 8) return opint::fail(ONEPROT_ERR_ARG, "layernorm_bwd: need d a positive multiple of 8");
  if ((dgamma != nullptr) != (dbeta != nullptr)) return opint::fail(ONEPROT_ERR_ARG, "layernorm_bwd: dgamma and dbeta go together");
  if (!gx && !dgamma) return opint::fail(ONEPROT_ERR_ARG, "layernorm_bwd: nothing to compute");
  if (dgamma && (!scratch || scratch_bytes < oneprot_layernorm_bwd_scratch_bytes(rows, d)))
    return opint::fail(ONEPROT_ERR_ARG, "layernorm_bwd: scratch too small");
  if (!al16(x) || !al16(gy) || !al16(gamma) || (gx && !al16(gx))) return opint::fail(ONEPROT_ERR_ARG, "layernorm_bwd: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("layernorm_bwd x=%p gy=%p gamma=%p mean=%p rstd=%p gx=%p dgamma=%p dbeta=%p scratch=%p rows=%d d=%d fp32=%d st=%p", x, gy, gamma, (const void*)mean, (const void*)rstd, gx, (void*)dgamma, (void*)dbeta, scratch, rows, d, is_fp32, stream);
  const bool fused = d <= 256 * oph::LN_MAXC;      // one pass over x and gy; longer rows: row kernel + column kernel
  opint::count_launch(fused ? (dgamma ? 3 : 1) : ((gx ? 1 : 0) + (dgamma ? 3 : 0)));
  if (optrace::dry()) return ONEPROT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ld = cdiv(d, 256) * 256;
  if (fused) {
    float* pg = dgamma ? static_cast<float*>(scratch) : nullptr;
    float* pb = dgamma ? pg + static_cast<size_t>(ln_bwd_max_blocks(rows)) * ld : nullptr;
    int slots = 0;
    if (is_fp32) {
#define LN_CALL(CC) slots = launch_ln_bwd<true, CC>(x, gy, gamma, mean, rstd, gx, pg, pb, ld, rows, d, st)
      LN_SWITCH_C(cdiv(d, 256), LN_CALL)
#undef LN_CALL
    } else {
#define LN_CALL(CC) slots = launch_ln_bwd<false, CC>(x, gy, gamma, mean, rstd, gx, pg, pb, ld, rows, d, st)
      LN_SWITCH_C(cdiv(d, 256), LN_CALL)
#undef LN_CALL
    }
    HD_CUDA(cudaGetLastError());
    if (dgamma) {
      oph::sum_slots_f32_kernel<<<cdiv(d, 256), 256, 0, st>>>(pg, slots, ld, d, dgamma);
      oph::sum_slots_f32_kernel<<<cdiv(d, 256), 256, 0, st>>>(pb, slots, ld, d, dbeta);
      HD_CUDA(cudaGetLastError());
    }
    return ONEPROT_OK;
  }
  if (gx) {
    const int blocks = std::min(cdiv(rows, 8), oneprot_num_sms() * 8);
    if (is_fp32) oph::layernorm_bwd_rows_long_kernel<true><<<blocks, 256, 0, st>>>(x, gy, gamma, mean, rstd, gx, rows, d);
    else oph::layernorm_bwd_rows_long_kernel<false><<<blocks, 256, 0, st>>>(x, gy, gamma, mean, rstd, gx, rows, d);
    HD_CUDA(cudaGetLastError());
  }
  if (dgamma) {
    const int chunks = ln_row_chunks(rows, d);
    const int rpc = cdiv(rows, chunks);
    float* pg = static_cast<float*>(scratch);
    float* pb = pg + static_cast<size_t>(chunks) * ld;
    const dim3 grid(cdiv(d, 256), chunks);
    if (is_fp32) oph::layernorm_bwd_cols_kernel<true><<<grid, 256, 0, st>>>(x, gy, mean, rstd, rows, d, rpc, pg, pb, ld);
    else oph::layernorm_bwd_cols_kernel<false><<<grid, 256, 0, st>>>(x, gy, mean, rstd, rows, d, rpc, pg, pb, ld);
    oph::sum_slots_f32_kernel<<<cdiv(d, 256), 256, 0, st>>>(pg, chunks, ld, d, dgamma);
    oph::sum_slots_f32_kernel<<<cdiv(d, 256), 256, 0, st>>>(pb, chunks, ld, d, dbeta);
    HD_CUDA(cudaGetLastError());
  }
  return ONEPROT_OK;
}

int oneprot_gelu(const void* x, const void* gy, void* out, size_t count, int is_fp32, void* stream) {
  if (!x || !out || count == 0 || count % 8) return opint::fail(ONEPROT_ERR_ARG, "gelu: need a positive element count that is a multiple of 8");
  if (!al16(x) || !al16(out) || (gy && !al16(gy))) return opint::fail(ONEPROT_ERR_ARG, "gelu: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("gelu x=%p gy=%p out=%p count=%zu fp32=%d st=%p", x, gy, out, count, is_fp32, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const size_t total8 = count / 8;
  const int blocks = static_cast<int>(std::min<size_t>((total8 + 255) / 256, static_cast<size_t>(oneprot_num_sms()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gy) {
    if (is_fp32) oph::gelu_kernel<true, true><<<blocks, 256, 0, st>>>(x, gy, out, total8);
    else oph::gelu_kernel<false, true><<<blocks, 256, 0, st>>>(x, gy, out, total8);
  } else {
    if (is_fp32) oph::gelu_kernel<true, false><<<blocks, 256, 0, st>>>(x, nullptr, out, total8);
    else oph::gelu_kernel<false, false><<<blocks, 256, 0, st>>>(x, nullptr, out, total8);
  }
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_meanpool_fwd(const void* x, const float* mask, void* y, float* inv_count, int B, int L, int D, int is_fp32,
                         int normalize, void* stream) {
  if (!x || !y || (normalize && !inv_count) || B <= 0 || L <= 0 || D <= 0 || D % 8) return opint::fail(ONEPROT_ERR_ARG, "meanpool_fwd: need D a positive multiple of 8");
  if (!al16(x) || !al16(y)) return opint::fail(ONEPROT_ERR_ARG, "meanpool_fwd: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("meanpool_fwd x=%p mask=%p y=%p inv_count=%p B=%d L=%d D=%d fp32=%d normalize=%d st=%p", x, (const void*)mask, y, (void*)inv_count, B, L, D, is_fp32, normalize, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const dim3 grid(B, cdiv(D, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) oph::meanpool_fwd_kernel<true><<<grid, 256, 0, st>>>(x, mask, y, inv_count, L, D, normalize);
  else oph::meanpool_fwd_kernel<false><<<grid, 256, 0, st>>>(x, mask, y, inv_count, L, D, normalize);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_meanpool_bwd(const void* gy, const float* mask, const float* inv_count, void* gx, int B, int L, int D, int is_fp32,
                         void* stream) {
  if (!gy || !inv_count || !gx || B <= 0 || L <= 0 || D <= 0 || D % 8) return opint::fail(ONEPROT_ERR_ARG, "meanpool_bwd: need D a positive multiple of 8");
  if (!al16(gy) || !al16(gx)) return opint::fail(ONEPROT_ERR_ARG, "meanpool_bwd: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("meanpool_bwd gy=%p mask=%p inv_count=%p gx=%p B=%d L=%d D=%d fp32=%d st=%p", gy, (const void*)mask, (const void*)inv_count, gx, B, L, D, is_fp32, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const dim3 grid(B, cdiv(D, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) oph::meanpool_bwd_kernel<true><<<grid, 256, 0, st>>>(gy, mask, inv_count, gx, L, D);
  else oph::meanpool_bwd_kernel<false><<<grid, 256, 0, st>>>(gy, mask, inv_count, gx, L, D);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_token_dot(const void* x, const void* vec, int vec_per_batch, const float* bias, const float* mask, float* out, int B,
                      int L, int D, int is_fp32, void* stream) {
  if (!x || !vec || !out || B <= 0 || L <= 0 || D <= 0 || D % 8) return opint::fail(ONEPROT_ERR_ARG, "token_dot: need D a positive multiple of 8");
  if (!al16(x) || !al16(vec)) return opint::fail(ONEPROT_ERR_ARG, "token_dot: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("token_dot x=%p vec=%p per_batch=%d bias=%p mask=%p out=%p B=%d L=%d D=%d fp32=%d st=%p", x, vec, vec_per_batch, (const void*)bias, (const void*)mask, (void*)out, B, L, D, is_fp32, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const long long tokens = static_cast<long long>(B) * L;
  const int blocks = static_cast<int>(std::min<long long>((tokens + 7) / 8, static_cast<long long>(oneprot_num_sms()) * 16));
  const size_t stride = vec_per_batch ? static_cast<size_t>(D) : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) oph::token_dot_kernel<true><<<blocks, 256, 0, st>>>(x, vec, stride, bias, mask, out, B, L, D);
  else oph::token_dot_kernel<false><<<blocks, 256, 0, st>>>(x, vec, stride, bias, mask, out, B, L, D);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_softmax_rows(const float* s, float* p, int B, int L, void* stream) {
  if (!s || !p || B <= 0 || L <= 0) return opint::fail(ONEPROT_ERR_ARG, "softmax_rows: bad argument");
  if (optrace::recording()) optrace::add("softmax_rows s=%p p=%p B=%d L=%d st=%p", (const void*)s, (void*)p, B, L, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  oph::softmax_rows_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(s, p, L);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_softmax_rows_bwd(const float* p, const float* dp, float* ds, int B, int L, void* stream) {
  if (!p || !dp || !ds || B <= 0 || L <= 0) return opint::fail(ONEPROT_ERR_ARG, "softmax_rows_bwd: bad argument");
  if (optrace::recording()) optrace::add("softmax_rows_bwd p=%p dp=%p ds=%p B=%d L=%d st=%p", (const void*)p, (const void*)dp, (void*)ds, B, L, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  oph::softmax_rows_bwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, dp, ds, L);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_attnpool_bwd_x(const void* g, const float* p, const float* ds, const void* w, void* gx, int B, int L, int D, int is_fp32,
                           void* stream) {
  if (!g || !p || !ds || !w || !gx || B <= 0 || L <= 0 || D <= 0 || D % 8) return opint::fail(ONEPROT_ERR_ARG, "attnpool_bwd_x: need D a positive multiple of 8");
  if (!al16(g) || !al16(w) || !al16(gx)) return opint::fail(ONEPROT_ERR_ARG, "attnpool_bwd_x: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("attnpool_bwd_x g=%p p=%p ds=%p w=%p gx=%p B=%d L=%d D=%d fp32=%d st=%p", g, (const void*)p, (const void*)ds, w, gx, B, L, D, is_fp32, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const dim3 grid(B, cdiv(D, 256));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) oph::attnpool_bwd_x_kernel<true><<<grid, 256, 0, st>>>(g, p, ds, w, gx, L, D);
  else oph::attnpool_bwd_x_kernel<false><<<grid, 256, 0, st>>>(g, p, ds, w, gx, L, D);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

static int abs_blocks(size_t total8) {
  return static_cast<int>(std::min<size_t>(std::min<size_t>((total8 + 255) / 256, static_cast<size_t>(oneprot_num_sms()) * 16), oph::ABS_MAX_BLOCKS));
}

size_t oneprot_abs_mean_scratch_bytes(size_t /*count*/) { return sizeof(float) * static_cast<size_t>(oph::ABS_MAX_BLOCKS); }   // one partial per block, at most

int oneprot_abs_mean_fwd(const void* x, size_t count, size_t true_count, int is_fp32, float* out, void* scratch, size_t scratch_bytes,
                         void* stream) {
  if (!x || !out || !scratch || count == 0 || count % 8 || true_count == 0 || true_count > count)
    return opint::fail(ONEPROT_ERR_ARG, "abs_mean_fwd: need a positive element count that is a multiple of 8 and 0 < true_count <= count");
  if (!al16(x) || scratch_bytes < oneprot_abs_mean_scratch_bytes(count)) return opint::fail(ONEPROT_ERR_ARG, "abs_mean_fwd: x must be 16-byte aligned, scratch of oneprot_abs_mean_scratch_bytes");
  if (optrace::recording()) optrace::add("abs_mean_fwd x=%p count=%zu true_count=%zu fp32=%d out=%p scratch=%p st=%p", x, count, true_count, is_fp32, (void*)out, scratch, stream);
  opint::count_launch(2);
  if (optrace::dry()) return ONEPROT_OK;
  const size_t total8 = count / 8;
  const int blocks = abs_blocks(total8);
  float* part = static_cast<float*>(scratch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) oph::abs_sum_kernel<true><<<blocks, 256, 0, st>>>(x, total8, part);
  else oph::abs_sum_kernel<false><<<blocks, 256, 0, st>>>(x, total8, part);
  oph::abs_mean_final_kernel<<<1, 256, 0, st>>>(part, blocks, 1.0f / static_cast<float>(true_count), out);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_abs_mean_bwd(const void* x, const float* g, size_t count, size_t true_count, int is_fp32, void* gx, void* stream) {
  if (!x || !g || !gx || count == 0 || count % 8 || true_count == 0 || true_count > count)
    return opint::fail(ONEPROT_ERR_ARG, "abs_mean_bwd: need a positive element count that is a multiple of 8 and 0 < true_count <= count");
  if (!al16(x) || !al16(gx)) return opint::fail(ONEPROT_ERR_ARG, "abs_mean_bwd: pointers must be 16-byte aligned");
  if (optrace::recording()) optrace::add("abs_mean_bwd x=%p g=%p count=%zu true_count=%zu fp32=%d gx=%p st=%p", x, (const void*)g, count, true_count, is_fp32, gx, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  const size_t total8 = count / 8;
  const int blocks = abs_blocks(total8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float coef = 1.0f / static_cast<float>(true_count);
  if (is_fp32) oph::abs_mean_bwd_kernel<true><<<blocks, 256, 0, st>>>(x, g, coef, gx, total8);
  else oph::abs_mean_bwd_kernel<false><<<blocks, 256, 0, st>>>(x, g, coef, gx, total8);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_sum_slots_f32(const float* part, int slots, int ld, int count, float* out, void* stream) {
  if (!part || !out || slots <= 0 || count <= 0 || ld < count) return opint::fail(ONEPROT_ERR_ARG, "sum_slots_f32: bad argument");
  if (optrace::recording()) optrace::add("sum_slots_f32 part=%p slots=%d ld=%d count=%d out=%p st=%p", (const void*)part, slots, ld, count, (void*)out, stream);
  opint::count_launch(1);
  if (optrace::dry()) return ONEPROT_OK;
  oph::sum_slots_f32_kernel<<<cdiv(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(part, slots, ld, count, out);
  HD_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

}  // extern "C"
#endif  // ONEPROT_KERNEL_EMULATION
