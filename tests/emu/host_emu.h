// CUDA runtime / driver calls of the library's HOST code, stubbed for the CPU emulation: memsets and
// copies become memset / memcpy, events and attributes are no-ops, the SM count is small (persistent CTAs
// loop over several work items), and cuTensorMapEncodeTiled fills the EmuMap the emulated TMA reads.
// Test infrastructure only (see cuda_emu.h / ptx_emu.h).
#pragma once
#include "ptx_emu.h"

#include <cstdlib>

namespace emu {
inline int sms() {
  const char* e = getenv("ONEPROT_EMU_SMS");
  const int v = e ? atoi(e) : 3;
  return v > 0 ? v : 3;
}
inline CUresult encode_tiled(CUtensorMap* m, CUtensorMapDataType, cuuint32_t rank, void* ptr, const cuuint64_t* dims, const cuuint64_t* strides,
                             const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill) {
  if (rank != 2 || box[0] != 64) return CUDA_ERROR_INVALID_VALUE;
  std::memset(m, 0, sizeof(*m));
  op::EmuMap e{static_cast<const uint16_t*>(ptr), dims[0], dims[1], strides[0] / 2, box[1]};
  std::memcpy(m, &e, sizeof(e));
  return CUDA_SUCCESS;
}
inline cudaError_t driver_entry(const char*, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  *fn = reinterpret_cast<void*>(&encode_tiled);
  if (q) *q = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}
inline cudaError_t ok(...) { return cudaSuccess; }
inline cudaError_t event_create(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(0x1); return cudaSuccess; }
inline cudaError_t get_attr(int* v, cudaDeviceAttr, int) { *v = sms(); return cudaSuccess; }
inline cudaError_t get_device(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t get_props(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); p->major = 10; return cudaSuccess; }
}  // namespace emu

#define cudaGetLastError() cudaSuccess
#define cudaGetErrorString(e) "emulated CUDA call"
#define cudaFuncSetAttribute(k, a, v) emu::ok((void*)(k), (a), (v))
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor(p, k, t, s) (*(p) = 2, (void)(k), cudaSuccess)
#define cudaGetDevice(p) emu::get_device(p)
#define cudaDeviceGetAttribute(p, a, d) emu::get_attr((p), (a), (d))
#define cudaGetDeviceProperties(p, d) emu::get_props((p), (d))
#define cudaGetDriverEntryPoint(name, fn, flags, q) emu::driver_entry((name), (fn), (flags), (q))
#define cudaMemsetAsync(p, v, n, s) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaMemcpyAsync(d, s, n, kind, st) (std::memcpy((d), (s), (n)), cudaSuccess)
#define cudaEventCreateWithFlags(e, f) emu::event_create((e), (f))
#define cudaEventRecord(e, s) emu::ok((e), (s))
#define cudaStreamWaitEvent(s, e, f) emu::ok((s), (e), (f))
#define cudaEventDestroy(e) emu::ok(e)
