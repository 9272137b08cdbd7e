// Minimal SIMT emulation to run the SOURCE of the plain (non-tensor-core) CUDA kernels on the CPU in
// the tests: one OS thread per CUDA thread of a block (blocks run one after the other), pthread
// barriers for __syncthreads and for the warp-synchronous shuffles.  Test infrastructure only - it
// exists so that indexing, reductions and arithmetic of csrc/head_kernels.cu can be checked against the
// numpy oracle on a machine without a GPU.  Not a performance model and not part of the product.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace emu {

struct Ctx {
  uint3 tid, bid;
  dim3 bdim, gdim;
  int lin;                 // linear thread id inside the block
  float* wslot_f;          // 32 floats of this thread's warp
  double* wslot_d;
  pthread_barrier_t* wbar; // barrier of this thread's warp
  pthread_barrier_t* bbar; // barrier of the block
};
inline thread_local Ctx tls;

inline void sync_block() { pthread_barrier_wait(tls.bbar); }

inline float shfl_xor(float v, int m) {
  const int lane = tls.lin & 31;
  tls.wslot_f[lane] = v;
  pthread_barrier_wait(tls.wbar);
  const float r = tls.wslot_f[lane ^ m];
  pthread_barrier_wait(tls.wbar);
  return r;
}
inline double shfl_xor(double v, int m) {
  const int lane = tls.lin & 31;
  tls.wslot_d[lane] = v;
  pthread_barrier_wait(tls.wbar);
  const double r = tls.wslot_d[lane ^ m];
  pthread_barrier_wait(tls.wbar);
  return r;
}

// runs body() once per CUDA thread; block.x must be a multiple of 32 (1-D blocks, grid up to 2-D)
inline void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  const int nt = static_cast<int>(block.x), nw = nt / 32;
  std::vector<pthread_barrier_t> wbars(nw);
  pthread_barrier_t bbar;
  for (auto& b : wbars) pthread_barrier_init(&b, nullptr, 32);
  pthread_barrier_init(&bbar, nullptr, nt);
  std::vector<float> wf(static_cast<size_t>(nw) * 32);
  std::vector<double> wd(static_cast<size_t>(nw) * 32);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    th.emplace_back([&, t] {
      Ctx& c = tls;
      c.lin = t;
      c.tid = make_uint3(t, 0, 0);
      c.bdim = block;
      c.gdim = grid;
      c.wslot_f = wf.data() + (t / 32) * 32;
      c.wslot_d = wd.data() + (t / 32) * 32;
      c.wbar = &wbars[t / 32];
      c.bbar = &bbar;
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          c.bid = make_uint3(bx, by, 0);
          body();
          pthread_barrier_wait(&bbar);   // blocks run one after the other (static __shared__ storage is reused)
        }
    });
  }
  for (auto& x : th) x.join();
  for (auto& b : wbars) pthread_barrier_destroy(&b);
  pthread_barrier_destroy(&bbar);
}

inline float uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

}  // namespace emu

// ---- CUDA spellings -> emulation -----------------------------------------------------------------
#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __restrict__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define threadIdx (emu::tls.tid)
#define blockIdx (emu::tls.bid)
#define blockDim (emu::tls.bdim)
#define gridDim (emu::tls.gdim)
#define __syncthreads() emu::sync_block()
#define __shfl_xor_sync(mask, v, m) emu::shfl_xor((v), (m))
#define __uint_as_float(u) emu::uint_as_float(u)
#define __float_as_uint(f) emu::float_as_uint(f)
#define __int_as_float(i) emu::uint_as_float(static_cast<uint32_t>(i))
#define __expf(x) expf(x)
#define __log2f(x) log2f(x)
#define rsqrtf(x) (1.0f / sqrtf(x))
#define __ldg(p) (*(p))
#define __ldcs(p) (*(p))
#define __stcs(p, v) (*(p) = (v))
using std::min;
using std::max;

// ---- atomics / fences used by the vector kernels ---------------------------------------------------
inline unsigned int atomicMax(unsigned int* p, unsigned int v) {
  unsigned int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
#define __threadfence() __atomic_thread_fence(__ATOMIC_SEQ_CST)
