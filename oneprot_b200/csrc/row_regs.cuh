// A row chunk of 8 elements held in registers IN ITS STORAGE FORMAT (one 16-byte vector for bf16, two for fp32) and
// unpacked where it is used.  The register-resident row kernels (LayerNorm, L2-normalise; head_kernels.cu /
// vector_kernels.cuh) keep whole rows this way: half the registers for bf16 against fp32 copies, and the occupancy
// (warps per SM x bytes in flight per warp) is what carries an HBM-bound row kernel.
#pragma once
#include <cstddef>
#include <cstdint>

namespace oprow {

template <bool FP32> struct Vec8;
template <> struct Vec8<false> {
  uint4 u;
  __device__ __forceinline__ void load(const void* base, size_t idx) {
    u = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + idx);
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct Vec8<true> {
  float4 a, b;
  __device__ __forceinline__ void load(const void* base, size_t idx) {
    a = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
    b = *reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx + 4);
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

constexpr int MAXC = 8;   // 8 chunks x 32 lanes x 8 elements = rows of up to 2048 elements held in registers

}  // namespace oprow
