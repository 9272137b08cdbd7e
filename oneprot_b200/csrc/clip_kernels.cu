// B200 (sm_100a) kernels + C ABI for the OneProt ClipLoss hot path.  See include/oneprot_clip.h for
// the contract of every entry point and DESIGN.md for the data layout and rooflines.
//
// One warp-specialised mainloop serves every tensor-core kernel here (one persistent CTA per SM):
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 4-stage smem ring, SWIZZLE_128B)
//   warp 1      UMMA issuer    (tcgen05.mma cta_group::1, M=128 x N=256 x K=16, fp32 accum in TMEM)
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators, double buffered)
//   warp 3      spare; in the fused forward it pushes this rank's rows through the NVLink multicast
//   warps 4-11  epilogue       (tcgen05.ld 32x32b -> registers; thread = one accumulator row)
// The epilogues differ:
//   clip_s_kernel<FWD>  exp-sums of the logits (row sums thread-local, column sums in registers)
//   clip_s_kernel<FWD_E> the same + the exponentials kept as a bf16 panel (stored-exponentials backward, opt-in)
//   clip_s_kernel<MAX>  exact maximum logit (robust tier, early-exits when the norm bound suffices)
//   clip_s_kernel<DZ>   dL/dZ panel, bf16, staged through swizzled smem and written by TMA stores
//   gemm_kernel         dA = Wz.B / dB = Wz^T.A with row-scale, row-dot, fp32 accumulate
// plus the HBM-bound vector kernels and the multimem (NVLS) exchange kernels at the end.
// ONEPROT_KERNEL_EMULATION (tests/emu): the kernel bodies of this file are also compiled for the CPU, with
// ptx_emu.h supplying functional stand-ins for the TMA / mbarrier / tcgen05 / TMEM primitives of ptx.cuh.
// With ONEPROT_HOST_EMULATION on top, the C ABI functions at the end are compiled for the CPU as well
// (tests/emu/translate.py rewrites their <<<...>>> launches, host_emu.h stubs the CUDA runtime calls).
#ifndef ONEPROT_KERNEL_EMULATION
#include "ptx.cuh"
#define OP_DYNAMIC_SMEM(name) extern __shared__ uint8_t name[]
#else
#include "ptx_emu.h"
#endif
#include "host_trace.h"
#include "../../include/oneprot_clip.h"

#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <atomic>
#include <algorithm>

#include "row_regs.cuh"

namespace op {

constexpr int BM = 128;            // accumulator rows  (TMEM lanes)
constexpr int BN = 256;            // accumulator cols  (TMEM columns per buffer)
constexpr int BK = 64;             // bf16 elements per smem row = 128 B = one swizzle atom
constexpr int MAX_STAGES = 6;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float G_MARGIN = 100.0f;   // e_ij <= 2^100 guaranteed by the Cauchy-Schwarz bound

struct __align__(16) SmemTail {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad[3];
  alignas(16) float colvec[BN];   // read back as float4
};
// dynamic smem = [NS operand stages | optional bf16 staging for TMA stores | barriers etc.]
constexpr int STORE_STAGING_BYTES = 2 * 16384;     // one 16 KiB TMA-store box {64 cols, 128 rows} per epilogue warp group
constexpr int smem_bytes(int ns, int staging) {
  return ns * STAGE_BYTES + staging + 1024 /*alignment slack*/ + static_cast<int>(sizeof(SmemTail));
}

// ------------------------------------------------------------------------------------------
// mainloop pieces
// ------------------------------------------------------------------------------------------
template <int A_MN, int B_MN>
__device__ __forceinline__ void load_stage(uint8_t* sa, uint8_t* sb, const CUtensorMap* mapA, const CUtensorMap* mapB,
                                           uint64_t* bar, int m0, int n0, int k0) {
  mbar_arrive_expect_tx(bar, STAGE_BYTES);
  if (A_MN == 0) {
    tma_load_2d(sa, mapA, bar, k0, m0);                       // box {64 k, 128 rows}
  } else {
#pragma unroll
    for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, mapA, bar, m0 + c * 64, k0);  // box {64 m, 64 k}
  }
  if (B_MN == 0) {
    tma_load_2d(sb, mapB, bar, k0, n0);                       // box {64 k, 256 rows}
  } else {
#pragma unroll
    for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, mapB, bar, n0 + c * 64, k0);
  }
}

// The four UMMAs of one 64-deep stage.  Only the 14-bit start-address field of the shared-memory descriptors moves
// from one K = 16 step to the next (+32 B K-major, +2048 B MN-major; no carry: the stage lies below 256 KiB), so the
// descriptors of a stage are built once and stepped by one 32-bit add each.  The MMA warp is the pacing warp of the
// S-kernels (ncu, round 2: ~105 instructions per stage took ~700 cycles beside two busy epilogue warps on its
// scheduler against 512 cycles of tensor work; the warp never waited for operands or accumulators) - every
// instruction taken out of this loop is tensor-pipe time.
template <int A_MN, int B_MN>
__device__ __forceinline__ void mma_stage(uint32_t sa, uint32_t sb, uint32_t tmem_d, bool first) {
  constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
  constexpr uint32_t step_a = (A_MN ? 2048 : 32) >> 4, step_b = (B_MN ? 2048 : 32) >> 4;
  const uint64_t da = A_MN ? make_smem_desc(sa, 8192, 1024) : make_smem_desc(sa, 16, 1024);
  const uint64_t db = B_MN ? make_smem_desc(sb, 8192, 1024) : make_smem_desc(sb, 16, 1024);
#pragma unroll
  for (int k = 0; k < BK / 16; ++k)
    umma_bf16(tmem_d, da + static_cast<uint64_t>(k * step_a), db + static_cast<uint64_t>(k * step_b), idesc,
              (first && k == 0) ? 0u : 1u);
}

template <int NS>
struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == NS) { stage = 0; phase ^= 1; }
  }
};

struct Smem {
  uint8_t* stages;
  uint8_t* staging;   // valid only when the kernel was launched with staging bytes
  SmemTail* tail;
};

template <int NS, int STAGING>
__device__ __forceinline__ Smem carve_smem(uint8_t* raw) {
  uintptr_t p = (reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023);
  Smem s;
  s.stages = reinterpret_cast<uint8_t*>(p);
  s.staging = reinterpret_cast<uint8_t*>(p + NS * STAGE_BYTES);
  s.tail = reinterpret_cast<SmemTail*>(p + NS * STAGE_BYTES + STAGING);
  return s;
}

// Common prologue: barrier init, TMEM alloc.  Returns the TMEM base address.
__device__ __forceinline__ uint32_t kernel_prologue(const Smem& s, const CUtensorMap* mapA, const CUtensorMap* mapB) {
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && lane_id() == 0) {
    prefetch_tmap(mapA);
    prefetch_tmap(mapB);
  }
  if (warp == 1 && lane_id() == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&s.tail->full[i], 1);
      mbar_init(&s.tail->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tail->tfull[i], 1);
      mbar_init(&s.tail->tempty[i], EPI_THREADS / 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&s.tail->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return s.tail->tmem_base;
}

__device__ __forceinline__ void kernel_epilogue_dealloc(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Butterfly transposing reduction: every lane holds v[0..31] (one value per column, for its
// own row); afterwards lane l holds in v[0] the sum over the 32 lanes of column l.
__device__ __forceinline__ void warp_transpose_sum32(float (&v)[32]) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < o; ++k) {
      const float keep = up ? v[k + o] : v[k];
      const float send = up ? v[k] : v[k + o];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

// same butterfly with max: lane l ends with the maximum over the 32 lanes of column l in v[0]
__device__ __forceinline__ void warp_transpose_max32(float (&v)[32]) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < o; ++k) {
      const float keep = up ? v[k + o] : v[k];
      const float send = up ? v[k] : v[k + o];
      v[k] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}

// ------------------------------------------------------------------------------------------
// S-kernels: logits tile = A_rows(128 x d) . B_all(256 x d)^T, both K-major.
// Work item = (column block j, chunk of CI row blocks); rows sweep inside an item so that the
// per-column accumulators stay in registers.
// ------------------------------------------------------------------------------------------
struct SParams {
  int rows;        // rows of A in this launch
  int N;           // columns = rows of B_all
  int nK;          // ceil(d / 64)
  int nI, nJ;      // row blocks (128), column blocks (256)
  int CI;          // row blocks per work item
  int nChunks;     // ceil(nI / CI)
  int grow0;       // global row index of row 0 (diagonal)
  const float* scale;   // device scalar
  const float* stats;   // [maxA2, maxB2]
  // forward
  float* rowpart;  // [2*nJ][ldr]
  float* colpart;  // [4*nChunks][ldc]
  int ldr, ldc;
  // dz
  const float* wr;
  const float* wc;
  const float* dg;
  __nv_bfloat16* Wz;
  int ldw;
  // fused all-gather of the second operand (forward only; ag_src == nullptr: off).  The spare warp
  // of every CTA pushes a slice of this rank's rows through the NVLink multicast alias and raises
  // per-chunk flags; the TMA producer waits for the flag of a column block before loading it.
  const uint4* ag_src;            // this rank's rows (local)
  uint4* ag_dst_mc;               // multicast alias of this rank's rows inside the gathered operand
  unsigned int* ag_counters;      // local, [ag_chunks], zero between launches
  unsigned int* ag_flags_mc;      // multicast alias of flags[W][ag_chunks + 1]
  const unsigned int* ag_flags;   // local copy of the flags
  float* ag_stats_mc;             // multicast alias of stats_all[W][4]
  const float* ag_stats;          // local copy
  float* stats_out;               // global maxima written for the later kernels (2 floats)
  unsigned int ag_epoch;
  int ag_rank, ag_world, ag_chunks, ag_chunk16, ag_rows, ag_bpc;   // chunk16: 16-byte vectors per chunk; bpc: column blocks per chunk
};

#ifdef ONEPROT_KERNEL_EMULATION
__device__ __forceinline__ void wait_flag(const unsigned int*, unsigned int) {}   // single-GPU emulation: no exchange
#else
__device__ __forceinline__ void wait_flag(const unsigned int* f, unsigned int epoch) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  if (v == epoch) return;
  const long long t0 = clock64();
  do {
    __nanosleep(64);
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (clock64() - t0 > ONEPROT_WAIT_TRAP_CYCLES) __trap();
  } while (v != epoch);
}
#endif

// Work item -> (logical column block, row chunk): column-major - the ~#SM items in flight cover ~#SM / nChunks column
// blocks x ALL row chunks, so every row-operand tile is shared by ~9 CTAs at a time (N = 32768).  A "square" order
// (SC row chunks x #SM / SC column blocks per wave, to keep a smaller working set hot in L2) was measured on B200 in
// round 2 and is monotonically SLOWER the more CTAs share a row tile: FWD_E 1.745 / 1.785 / 1.828 / 1.875 / 1.908 ms for
// 9 / 18 / 37 / 74 / 148 sharers, FWD and DZ alike - although ncu shows the plain order re-streaming the row operand
// once per wave (0.97 GB of DRAM reads for 0.13 GB of operands).  DRAM is not the limiter; lock-step sharers are.
__device__ __forceinline__ void item_to_jc(const SParams& p, int item, int& jl, int& ch) {
  jl = item / p.nChunks;
  ch = item - jl * p.nChunks;
}

// logical column-block index -> actual column block: chunk-major (arrival order), own rank first
__device__ __forceinline__ int map_jb(const SParams& p, int jl) {
  if (!p.ag_src) return jl;
  const int per = p.ag_world * p.ag_bpc;
  const int c = jl / per, rem = jl % per;
  const int r = (rem / p.ag_bpc + p.ag_rank) % p.ag_world;
  return (r * p.ag_rows) / BN + c * p.ag_bpc + rem % p.ag_bpc;
}

__device__ __forceinline__ void load_c_and_G(const SParams& p, float& c, float& negG) {
  const float s = __ldg(p.scale);
  c = s * LOG2E;
  float ma, mb;
  if (p.ag_src) {
    ma = 0.f; mb = 0.f;
    for (int r = 0; r < p.ag_world; ++r) {
      wait_flag(p.ag_flags + r * (p.ag_chunks + 1) + p.ag_chunks, p.ag_epoch);
      ma = fmaxf(ma, *(volatile const float*)(p.ag_stats + r * 4));
      mb = fmaxf(mb, *(volatile const float*)(p.ag_stats + r * 4 + 1));
    }
  } else {
    ma = __ldg(p.stats);
    mb = __ldg(p.stats + 1);
  }
  const float U = fabsf(c) * sqrtf(ma * mb);
  negG = -fmaxf(0.f, U - G_MARGIN);
  // exact maximum logit from the max pass (single-GPU robust tier): a tighter reference than the
  // Cauchy-Schwarz bound when the bound alone would push every term below 2^-126
  if (!p.ag_src && p.stats[3] != 0.f) negG = -fmaxf(0.f, p.stats[2] - G_MARGIN);
  if (p.ag_src && p.stats_out && blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32) {
    p.stats_out[0] = ma;
    p.stats_out[1] = mb;
  }
}

enum { EPI_FWD = 0, EPI_DZ = 1, EPI_MAX = 2, EPI_RCMAX = 3, EPI_RANK = 4,    // RCMAX: row/col maxima; RANK: retrieval ranks
       EPI_SFWD = 5, EPI_SDZ = 6,     // SigLIP: row sums of softplus(z) / dL/dZ panel sigma(z) wr_i - [i == j] dg_i
       EPI_FWD_E = 8,                 // EPI_FWD that also keeps the exponentials e_ij as a bf16 panel: the backward then
                                      // rescales them in place (dz_from_exp_kernel) instead of recomputing the logits
       EPI_SFWD_K = 10 };             // EPI_SFWD that also keeps sigma(z_ij) - [i == j] as a bf16 panel: dL/dz up to the constant
                                      // g / n, so the SigLIP backward needs neither a recompute nor a rescale pass

template <int EPI>
struct SCfg {
  static constexpr int NS = 4;                                             // operand ring depth
  static constexpr bool PANEL = (EPI == EPI_DZ || EPI == EPI_SDZ);   // writes a bf16 dL/dZ panel by TMA stores
  // L2 hints of the two panel writers of the ClipLoss step: panel stores evict-first (written once, read once later),
  // operand loads evict-last (re-read by every CTA sweeping the same blocks).  Measured on B200 in round 2, kernels
  // alone at N = 32768: DZ 0.852 vs 0.865 ms per 16384 rows, FWD_E 1.912 vs 1.935 ms; results bit-identical.
  static constexpr bool L2_HINTS = (EPI == EPI_DZ || EPI == EPI_FWD_E);
  static constexpr bool KEEP_E = (EPI == EPI_FWD_E);
  static constexpr bool SUMS = (EPI == EPI_FWD || KEEP_E);                    // row / column exp-sums (+ the fused all-gather)
  static constexpr bool KEEP_S = (EPI == EPI_SFWD_K);
  static constexpr int STAGING = (PANEL || KEEP_E || KEEP_S) ? STORE_STAGING_BYTES : 0;   // FWD / MAX / RCMAX / RANK / SFWD need none
  static constexpr int SMEM = smem_bytes(NS, STAGING);
};

// Kept-panel forwards (FWD_E, SFWD_K): 32 columns of this thread's row -> bf16 -> the warp group's SWIZZLE_128B staging
// box {64 cols, 128 rows}; every second call the box goes out by one TMA store (the box protocol of the dL/dZ panel).
template <bool HINT>
__device__ __forceinline__ void keep_stage_store(const float (&v)[32], int cc, int r, int h, uint32_t stage_s, const uint8_t* box,
                                                 bool store_issuer, const CUtensorMap* map, int col0, int row0, uint64_t policy) {
  if ((cc & 1) == 0) {
    if (store_issuer) bulk_wait_read<0>();       // the previous TMA store must have finished reading the staging box
    named_bar_sync(2 + h, 128);
  }
  const uint32_t line = stage_s + r * 128;       // row r of the box: 128-byte line, 16-byte slots XOR-swizzled by (r % 8)
#pragma unroll
  for (int v4 = 0; v4 < 4; ++v4) {
    const uint32_t slot = static_cast<uint32_t>(((cc & 1) * 4 + v4) ^ (r & 7));
    st_shared_v4(line + slot * 16, pack_bf16x2(v[v4 * 8], v[v4 * 8 + 1]), pack_bf16x2(v[v4 * 8 + 2], v[v4 * 8 + 3]),
                 pack_bf16x2(v[v4 * 8 + 4], v[v4 * 8 + 5]), pack_bf16x2(v[v4 * 8 + 6], v[v4 * 8 + 7]));
  }
  if (cc & 1) {
    fence_proxy_async();                         // generic-proxy smem writes -> visible to the TMA engine
    named_bar_sync(2 + h, 128);
    if (store_issuer) {
      if (HINT) tma_store_2d_hint(map, box, col0, row0, policy);
      else tma_store_2d(map, box, col0, row0);
      bulk_commit();
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
clip_s_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
              const __grid_constant__ CUtensorMap mapW, const SParams p) {
  constexpr int NS = SCfg<EPI>::NS;
  if (EPI == EPI_MAX) {
    // the max pass is always enqueued but does work only when the norm bound is not rigorous
    const float U = fabsf(__ldg(p.scale) * LOG2E) * sqrtf(__ldg(p.stats) * __ldg(p.stats + 1));
    if (U <= G_MARGIN) return;     // uniform over the grid, before any barrier / TMEM allocation
  }
  OP_DYNAMIC_SMEM(smem_raw);
  const Smem s = carve_smem<NS, SCfg<EPI>::STAGING>(smem_raw);
  const uint32_t tmem_base = kernel_prologue(s, &mapA, &mapB);
  const int warp = threadIdx.x >> 5;
  const int items = p.nJ * p.nChunks;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    reg_dealloc<56>();   // the 4 control warps hand registers to the 8 epilogue warps
    if (lane_id() == 0) {
      PipeState<NS> ps;
      const uint64_t pol_keep = SCfg<EPI>::L2_HINTS ? l2_policy_evict_last() : 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int jl, ch;
        item_to_jc(p, item, jl, ch);
        const int jb = map_jb(p, jl);
        const int ib1 = min(p.nI, (ch + 1) * p.CI);
        if (SCfg<EPI>::SUMS && p.ag_src) {
          // columns [256 jb, 256 jb + 256) belong to one chunk of one rank: wait until it has landed
          const int col0 = jb * BN;
          const int r = col0 / p.ag_rows;
          const int c = ((col0 - r * p.ag_rows) / BN) / p.ag_bpc;
          wait_flag(p.ag_flags + r * (p.ag_chunks + 1) + c, p.ag_epoch);
        }
        for (int ib = ch * p.CI; ib < ib1; ++ib) {
          for (int kb = 0; kb < p.nK; ++kb) {
            mbar_wait(&s.tail->empty[ps.stage], ps.phase ^ 1);
            uint8_t* sa = s.stages + ps.stage * STAGE_BYTES;
            if (SCfg<EPI>::L2_HINTS) {
              // operands are re-read by every CTA sweeping the same blocks: keep them in L2 against the streamed panel
              mbar_arrive_expect_tx(&s.tail->full[ps.stage], STAGE_BYTES);
              tma_load_2d_hint(sa, &mapA, &s.tail->full[ps.stage], kb * BK, ib * BM, pol_keep);
              tma_load_2d_hint(sa + A_STAGE_BYTES, &mapB, &s.tail->full[ps.stage], kb * BK, jb * BN, pol_keep);
            } else {
              load_stage<0, 0>(sa, sa + A_STAGE_BYTES, &mapA, &mapB, &s.tail->full[ps.stage], ib * BM, jb * BN, kb * BK);
            }
            ps.advance();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ UMMA issuer
    // All 32 lanes walk the loop in lock step (uniform control flow keeps the descriptors in uniform registers); the
    // lane elect.sync picks - the same one every time - issues the UMMAs and their commits.
    reg_dealloc<56>();
    {
      PipeState<NS> ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t stages_s = smem_u32(s.stages);
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int jl, ch;
        item_to_jc(p, item, jl, ch);
        const int ib1 = min(p.nI, (ch + 1) * p.CI);
        for (int ib = ch * p.CI; ib < ib1; ++ib) {
          mbar_wait_warp(&s.tail->tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = 0; kb < p.nK; ++kb) {
            mbar_wait_warp(&s.tail->full[ps.stage], ps.phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = stages_s + ps.stage * STAGE_BYTES;
              mma_stage<0, 0>(sa, sa + A_STAGE_BYTES, tmem_d, kb == 0);
              umma_commit(&s.tail->empty[ps.stage]);
            }
            __syncwarp();
            ps.advance();
          }
          if (elect_one()) umma_commit(&s.tail->tfull[acc]);
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp < EPI_WARP0) {
    reg_dealloc<56>();
#ifndef ONEPROT_KERNEL_EMULATION
    if (SCfg<EPI>::SUMS && warp == 3 && p.ag_src) {
      // ---------------------------------------------- all-gather push (spare warp)
      const int lane = lane_id();
      const int fl = p.ag_rank * (p.ag_chunks + 1);
      if (blockIdx.x == 0 && lane == 0) {   // this rank's norm maxima, then their flag
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.ag_stats_mc + p.ag_rank * 4), "f"(p.stats[0]) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.ag_stats_mc + p.ag_rank * 4 + 1), "f"(p.stats[1]) : "memory");
        __threadfence_system();
        asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.ag_flags_mc + fl + p.ag_chunks), "r"(p.ag_epoch) : "memory");
      }
      const int per = (p.ag_chunk16 + gridDim.x - 1) / gridDim.x;
      const int lo = min(p.ag_chunk16, static_cast<int>(blockIdx.x) * per), hi = min(p.ag_chunk16, lo + per);
      for (int c = 0; c < p.ag_chunks; ++c) {
        const size_t base = static_cast<size_t>(c) * p.ag_chunk16;
        for (int i = lo + lane; i < hi; i += 32) {
          const uint4 v = p.ag_src[base + i];
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.ag_dst_mc + base + i), "r"(v.x),
                       "r"(v.y), "r"(v.z), "r"(v.w)
                       : "memory");
        }
        __threadfence_system();             // this lane's stores are performed system-wide
        __syncwarp();
        if (lane == 0) {
          const unsigned int prev = atomicAdd(p.ag_counters + c, 1u);
          if (prev == gridDim.x - 1) {      // every CTA has pushed (and fenced) its slice of chunk c
            p.ag_counters[c] = 0;
            __threadfence_system();
            asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.ag_flags_mc + fl + c), "r"(p.ag_epoch) : "memory");
          }
        }
      }
    }
#endif
  } else {
    // ------------------------------------------------ epilogue (8 warps)
    reg_alloc<224>();
    const int ew = warp - EPI_WARP0;
    const int q = ew & 3;        // TMEM lane quadrant (== warp % 4)
    const int h = ew >> 2;       // column half of the 256-wide tile
    const int lane = lane_id();
    const int r = q * 32 + lane; // row inside the tile
    float c, negG;
    constexpr bool PANEL = SCfg<EPI>::PANEL;
    constexpr bool SUMS = SCfg<EPI>::SUMS, KEEP_E = SCfg<EPI>::KEEP_E, KEEP_S = SCfg<EPI>::KEEP_S;
    if (EPI == EPI_MAX || EPI == EPI_RCMAX) { c = __ldg(p.scale) * LOG2E; negG = 0.f; }
    else if (EPI == EPI_RANK) { c = 1.f; negG = 0.f; }
    else if (EPI == EPI_SFWD || EPI == EPI_SDZ || KEEP_S) {
      // SigLIP: x = log2(e) (logit_scale <a, b> + logit_bias); the bias pointer travels in p.wc (no column weights here)
      c = __ldg(p.scale) * LOG2E;
      negG = p.wc ? __ldg(p.wc) * LOG2E : 0.f;
    }
    else load_c_and_G(p, c, negG);
    int acc = 0;
    uint32_t acc_phase = 0;
    float xmax = 0.f;               // EPI_MAX: running max(0, x) of this thread

    float colacc[(SUMS || EPI == EPI_RCMAX || EPI == EPI_RANK) ? 128 : 1];
    // DZ: bf16 staging of 64 columns of this warp group's half tile = one SWIZZLE_128B box {64 cols, 128 rows}
    const uint32_t colvec_s = smem_u32(&s.tail->colvec[0]);
    const uint32_t stage_s = smem_u32(s.staging) + h * 16384;
    const bool store_issuer = (q == 0) && (lane == 0);
    const uint64_t pol_stream = SCfg<EPI>::L2_HINTS ? l2_policy_evict_first() : 0;   // the panel is written once, read once later
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int jl, ch;
      item_to_jc(p, item, jl, ch);
      const int jb = map_jb(p, jl);
      const int ib1 = min(p.nI, (ch + 1) * p.CI);
      const int j0 = jb * BN + h * 128;   // first column this thread sees
      if (SUMS || EPI == EPI_RCMAX || EPI == EPI_RANK) {
#pragma unroll
        for (int k = 0; k < 128; ++k) colacc[k] = (EPI == EPI_RCMAX) ? -INFINITY : 0.f;
      }
      if (EPI == EPI_DZ || EPI == EPI_RANK) {
        // stage the per-column weights of this item's 256 columns
        named_bar_sync(1, EPI_THREADS);
        const int t = threadIdx.x - EPI_WARP0 * 32;
        const int jj = jb * BN + t;
        st_shared_f32(colvec_s + t * 4, (jj < p.N) ? __ldg(p.wc + jj) : 0.f);
        named_bar_sync(1, EPI_THREADS);
      }
      for (int ib = ch * p.CI; ib < ib1; ++ib) {
        const int i = ib * BM + r;
        const bool rowok = i < p.rows;
        const bool edge = (ib * BM + BM > p.rows) || (jb * BN + BN > p.N);
        float wr_i = 0.f, dg_i = 0.f;
        bool diag_tile = false;
        if (PANEL) {
          if (rowok) { wr_i = __ldg(p.wr + i); dg_i = __ldg(p.dg + i); }
        }
        if (PANEL || KEEP_S) {
          const int g0 = p.grow0 + ib * BM;       // global rows [g0, g0+128)
          diag_tile = (g0 < j0 + 128) && (j0 < g0 + BM);
        }
        mbar_wait(&s.tail->tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + h * 128;
        float rsum = 0.f;
        float rsum2 = 0.f;                        // SFWD_K: row sum of sigma(z) (for d logit_bias)
        // DZ keeps no per-column accumulators, so the whole 128-column row slice fits in registers:
        // read it at once and hand the TMEM buffer back before the (store-paced) rest of the epilogue.
        float vall[PANEL ? 4 : 1][32];
        if (PANEL) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32(taddr + cc * 32, vall[cc]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.tail->tempty[acc]);
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float vbuf[32];
          if (!PANEL) {
            tmem_ld_32x32(taddr + cc * 32, vbuf);
            tmem_ld_wait();
            if (cc == 3) {
              // accumulator fully read: hand the TMEM buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s.tail->tempty[acc]);
            }
          }
          float (&v)[32] = PANEL ? vall[cc] : vbuf;
          if (EPI == EPI_MAX) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const bool ok = rowok && (j0 + cc * 32 + k < p.N);
              xmax = fmaxf(xmax, ok ? v[k] * c : 0.f);
            }
            continue;
          }
          if (EPI == EPI_RANK) {
            // retrieval ranks: how many logits of this row (column) beat the label logit d_i (d_j).
            // wr = label dot of the row, colvec = label dots of the columns (raw dot products).
            if (cc == 0) rsum = 0.f;
            const float d_i = rowok ? __ldg(p.wr + i) : INFINITY;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 w4 = ld_shared_f4(colvec_s + (h * 128 + cc * 32 + k4 * 4) * 4);
              const float dj[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int k = k4 * 4 + u;
                const int col = j0 + cc * 32 + k;
                const bool ok = rowok && col < p.N && (p.grow0 + i != col);   // the label itself never counts
                rsum += (ok && v[k] > d_i) ? 1.f : 0.f;
                colacc[cc * 32 + k] += (ok && v[k] > dj[u]) ? 1.f : 0.f;
              }
            }
            continue;
          }
          if (EPI == EPI_RCMAX) {
            // rsum doubles as the running row maximum of this 128-column slice (log2 units)
            if (cc == 0) rsum = -INFINITY;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const bool ok = rowok && (j0 + cc * 32 + k < p.N);
              const float x = ok ? v[k] * c : -INFINITY;
              rsum = fmaxf(rsum, x);
              colacc[cc * 32 + k] = fmaxf(colacc[cc * 32 + k], x);
            }
            continue;
          }
          if (EPI == EPI_SFWD || KEEP_S) {
            // softplus in log2 units, overflow-free: log2(1 + 2^x) = max(x, 0) + log2(1 + 2^-|x|)
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const bool ok = rowok && (j0 + cc * 32 + k < p.N);
              const float x = fmaf(v[k], c, negG);
              // log2(1 + e), e = 2^-|x|: below 2^-10 the rounding of 1 + e would cost up to eps / e relative
              // (and bias the sum of N tiny terms), so use e log2(e) (1 - e / 2) there
              const float e = ex2(-fabsf(x));
              const float l = (e < 9.765625e-4f) ? e * fmaf(-0.72134752f, e, LOG2E) : __log2f(1.f + e);
              const float sp = fmaxf(x, 0.f) + l;
              rsum += ok ? sp : 0.f;
              if (KEEP_S) {
                // sigma(z) = 1 / (1 + 2^-x) from the same exponential: 1 / (1 + e) for x >= 0, e / (1 + e) below
                const float sg = ok ? __fdividef(x >= 0.f ? 1.f : e, 1.f + e) : 0.f;
                rsum2 += sg;
                v[k] = sg - ((diag_tile && p.grow0 + i == j0 + cc * 32 + k) ? 1.f : 0.f);
              }
            }
            if (KEEP_S)
              keep_stage_store<false>(v, cc, r, h, stage_s, s.staging + h * 16384, store_issuer, &mapW, j0 + (cc >> 1) * 64, ib * BM, 0);
            continue;
          }
          if (SUMS) {
            if (!edge) {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const float e = ex2(fmaf(v[k], c, negG));
                rsum += e;
                colacc[cc * 32 + k] += e;
                if (KEEP_E) v[k] = e;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const bool ok = rowok && (j0 + cc * 32 + k < p.N);
                const float e = ok ? ex2(fmaf(v[k], c, negG)) : 0.f;
                rsum += e;
                colacc[cc * 32 + k] += e;
                if (KEEP_E) v[k] = e;
              }
            }
            if (KEEP_E)   // the exponentials of this 32-column slice join the kept panel
              keep_stage_store<SCfg<EPI>::L2_HINTS>(v, cc, r, h, stage_s, s.staging + h * 16384, store_issuer, &mapW,
                                                    j0 + (cc >> 1) * 64, ib * BM, pol_stream);
          } else {
            if ((cc & 1) == 0) {
              // the previous TMA store must have finished reading the staging box
              if (store_issuer) bulk_wait_read<0>();
              named_bar_sync(2 + h, 128);
            }
            uint32_t packed[16];
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              float dz[4];
              if (EPI == EPI_SDZ) {
                // sigma(z) = 1 / (1 + 2^-x): 0 for x -> -inf (2^-x = inf), 1 for x -> +inf
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float sg = __fdividef(1.f, 1.f + ex2(-fmaf(v[k4 * 4 + u], c, negG)));
                  dz[u] = wr_i * sg;
                  // optional row sums of sigma (d logit_bias = (g / n) (sum sigma - n)): only valid entries count
                  if (p.rowpart) rsum += (rowok && (j0 + cc * 32 + k4 * 4 + u < p.N)) ? sg : 0.f;
                }
              } else {
                const float4 w4 = ld_shared_f4(colvec_s + (h * 128 + cc * 32 + k4 * 4) * 4);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int k = k4 * 4 + u;
                  const float e = ex2(fmaf(v[k], c, negG));
                  dz[u] = e * (wr_i + wv[u]);
                }
              }
              if (diag_tile) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (p.grow0 + i == j0 + cc * 32 + k4 * 4 + u) dz[u] -= dg_i;
              }
              packed[k4 * 2] = pack_bf16x2(dz[0], dz[1]);
              packed[k4 * 2 + 1] = pack_bf16x2(dz[2], dz[3]);
            }
            // row r of the box: 128-byte line, 16-byte slots XOR-swizzled by (r % 8)
            const uint32_t line = stage_s + r * 128;
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              const uint32_t slot = static_cast<uint32_t>(((cc & 1) * 4 + v4) ^ (r & 7));
              st_shared_v4(line + slot * 16, packed[v4 * 4], packed[v4 * 4 + 1], packed[v4 * 4 + 2], packed[v4 * 4 + 3]);
            }
            if (cc & 1) {
              fence_proxy_async();            // generic-proxy smem writes -> visible to the TMA engine
              named_bar_sync(2 + h, 128);
              if (store_issuer) {
                if (SCfg<EPI>::L2_HINTS) tma_store_2d_hint(&mapW, s.staging + h * 16384, j0 + (cc >> 1) * 64, ib * BM, pol_stream);
                else tma_store_2d(&mapW, s.staging + h * 16384, j0 + (cc >> 1) * 64, ib * BM);
                bulk_commit();
              }
            }
          }
        }
        if (SUMS || EPI == EPI_RCMAX || EPI == EPI_RANK || EPI == EPI_SFWD || KEEP_S || (EPI == EPI_SDZ && p.rowpart)) {
          p.rowpart[static_cast<size_t>(jb * 2 + h) * p.ldr + i] = rsum;   // ldr covers nI*128 rows
        }
        if (KEEP_S && p.colpart) p.colpart[static_cast<size_t>(jb * 2 + h) * p.ldr + i] = rsum2;   // second row-partial buffer
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (SUMS || EPI == EPI_RCMAX || EPI == EPI_RANK) {
        // flush column sums (maxima, counts) of this item: reduce over the 32 rows of the warp, one slot per (chunk, quadrant)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = colacc[cc * 32 + k];
          if (EPI == EPI_RCMAX) warp_transpose_max32(v); else warp_transpose_sum32(v);
          p.colpart[static_cast<size_t>(ch * 4 + q) * p.ldc + j0 + cc * 32 + lane] = v[0];  // ldc covers nJ*256
        }
      }
    }
    if ((PANEL || KEEP_E || KEEP_S) && store_issuer) bulk_wait<0>();   // panel fully written before the CTA retires
    if (EPI == EPI_MAX) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(p.stats_out + 2), __float_as_uint(xmax));
      if (blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32) p.stats_out[3] = 1.f;
    }
  }
  kernel_epilogue_dealloc(tmem_base);
}

// ------------------------------------------------------------------------------------------
// Generic GEMM: C[M x Nc] = op(A) op(B), tile 128 x 256, full K per tile.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]);

constexpr int GEMM_NS = 4;
constexpr int GEMM_SMEM = smem_bytes(GEMM_NS, 0);

struct GParams {
  int M, Nc, nK;
  int nMb, nNb;
  const float* acc_in;
  float* acc_out;
  __nv_bfloat16* out;
  int ldc;
  const float* row_scale;        // optional: value *= row_scale[m] (applied after acc_in is added)
  const __nv_bfloat16* dot_mat;  // optional: rowdot_part[(nb*2+h)*ldd + m] = <unscaled value row, dot_mat row>
  float* rowdot_part;
  int ld_dot, ldd;
};

// (A variant whose epilogue TMA-stored every finished dB tile into its owner GPU - the reduce-scatter fused into
// the GEMM - and a CTA-pair variant (tcgen05.mma cta_group::2, 256 x 256 tiles) were built and measured: 1.08 vs
// 0.98 ms per step at 8 GPUs against the side-stream pull-reduce, resp. 7.13 vs 7.07 ms per step on one GPU in
// the sustained regime; both were removed in round 2, see DESIGN.md.)
template <int A_MN, int B_MN>
__global__ void __maxnreg__(128)   // 384 x 128 registers: leaves 16 K registers per SM for a co-resident exchange kernel
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GParams p) {
  OP_DYNAMIC_SMEM(smem_raw);
  const Smem s = carve_smem<GEMM_NS, 0>(smem_raw);
  const uint32_t tmem_base = kernel_prologue(s, &mapA, &mapB);
  const int warp = threadIdx.x >> 5;
  const int tiles = p.nMb * p.nNb;

  if (warp == 0) {
    reg_dealloc<40>();
    if (lane_id() == 0) {
      PipeState<GEMM_NS> ps;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int mb = t / p.nNb, nb = t % p.nNb;
        for (int kb = 0; kb < p.nK; ++kb) {
          mbar_wait(&s.tail->empty[ps.stage], ps.phase ^ 1);
          uint8_t* sa = s.stages + ps.stage * STAGE_BYTES;
          load_stage<A_MN, B_MN>(sa, sa + A_STAGE_BYTES, &mapA, &mapB, &s.tail->full[ps.stage], mb * BM, nb * BN,
                                 kb * BK);
          ps.advance();
        }
      }
    }
  } else if (warp == 1) {
    reg_dealloc<40>();
    {   // all lanes in lock step, the elected lane issues (see clip_s_kernel)
      PipeState<GEMM_NS> ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t stages_s = smem_u32(s.stages);
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        mbar_wait_warp(&s.tail->tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.nK; ++kb) {
          mbar_wait_warp(&s.tail->full[ps.stage], ps.phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = stages_s + ps.stage * STAGE_BYTES;
            mma_stage<A_MN, B_MN>(sa, sa + A_STAGE_BYTES, tmem_d, kb == 0);
            umma_commit(&s.tail->empty[ps.stage]);
          }
          __syncwarp();
          ps.advance();
        }
        if (elect_one()) umma_commit(&s.tail->tfull[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < EPI_WARP0) {
    reg_dealloc<40>();
  } else {
    reg_alloc<168>();
    const int ew = warp - EPI_WARP0;
    const int q = ew & 3, h = ew >> 2;
    const int lane = lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int mb = t / p.nNb, nb = t % p.nNb;
      const int m = mb * BM + q * 32 + lane;
      const int n0 = nb * BN + h * 128;
      const float rs = (p.row_scale && m < p.M) ? __ldg(p.row_scale + m) : 1.f;
      float rdot = 0.f;
      mbar_wait(&s.tail->tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + h * 128;
      // read this thread's whole 128-column slice, then hand the TMEM buffer back at once: the
      // global loads / stores of the epilogue no longer hold up the next tile's MMAs
      float vall[4][32];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32(taddr + cc * 32, vall[cc]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tail->tempty[acc]);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float (&v)[32] = vall[cc];
        if (m < p.M) {
          const size_t base = static_cast<size_t>(m) * p.ldc + n0 + cc * 32;
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            const int n = n0 + cc * 32 + v8 * 8;
            if (n < p.Nc) {       // Nc % 8 == 0 is required by the host wrapper
              float x[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) x[u] = v[v8 * 8 + u];
              if (p.acc_in) {
                const float4 a0 = *reinterpret_cast<const float4*>(p.acc_in + base + v8 * 8);
                const float4 a1 = *reinterpret_cast<const float4*>(p.acc_in + base + v8 * 8 + 4);
                x[0] += a0.x; x[1] += a0.y; x[2] += a0.z; x[3] += a0.w;
                x[4] += a1.x; x[5] += a1.y; x[6] += a1.z; x[7] += a1.w;
              }
              if (p.dot_mat) {
                float fd[8];
                bf16x8_to_float(*reinterpret_cast<const uint4*>(p.dot_mat + static_cast<size_t>(m) * p.ld_dot + n), fd);
#pragma unroll
                for (int u = 0; u < 8; ++u) rdot = fmaf(x[u], fd[u], rdot);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) x[u] *= rs;
              if (p.acc_out) {
                *reinterpret_cast<float4*>(p.acc_out + base + v8 * 8) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4*>(p.acc_out + base + v8 * 8 + 4) = make_float4(x[4], x[5], x[6], x[7]);
              }
              if (p.out) {
                *reinterpret_cast<uint4*>(p.out + base + v8 * 8) =
                    make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                               pack_bf16x2(x[6], x[7]));
              }
            }
          }
        }
      }
      if (p.rowdot_part && m < p.M) p.rowdot_part[static_cast<size_t>(nb * 2 + h) * p.ldd + m] = rdot;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  kernel_epilogue_dealloc(tmem_base);
}

#include "vector_kernels.cuh"   // the non-tensor-core kernels (also compiled for the CPU emulation of the tests)

#ifndef ONEPROT_KERNEL_EMULATION
// ---- NVLink SHARP (multimem) exchanges over a symmetric-memory multicast mapping ---------------
// dst_mc is the multicast alias of one buffer that exists on every GPU of the node: a
// multimem.st lands in all copies (all-gather by push), a multimem.ld_reduce returns the
// reduction over all copies computed in the switch (all-reduce / reduce-scatter by pull).
__global__ void mc_store_kernel(const uint4* __restrict__ src, uint4* dst_mc, size_t n16) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 v = src[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_mc + i), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  }
}

// op 0: fp32 add (4 per thread), op 1: u32 max (bit patterns of non-negative floats)
__global__ void mc_allreduce_kernel(const float* src_mc, float* __restrict__ dst, int count, int op) {
  const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= count) return;
  if (op == 0) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(src_mc + i4)
                 : "memory");
    *reinterpret_cast<float4*>(dst + i4) = v;          // count is padded to a multiple of 4 by the host
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t r;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.max.u32 %0, [%1];" : "=r"(r) : "l"(src_mc + i4 + u) : "memory");
      dst[i4 + u] = __uint_as_float(r);
    }
  }
}

// dst[i] = sum over GPUs of the bf16 buffer (fp32 accumulation in the switch), 8 elements per thread
__global__ void mc_reduce_bf16_kernel(const uint4* src_mc, uint4* __restrict__ dst, size_t n16) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(src_mc + i)
                 : "memory");
    dst[i] = v;
  }
}

#endif  // ONEPROT_KERNEL_EMULATION
}  // namespace op

#if !defined(ONEPROT_KERNEL_EMULATION) || defined(ONEPROT_HOST_EMULATION)
// ==========================================================================================
// Host side: C ABI
// ==========================================================================================
namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};   // process-wide: backward runs on autograd's thread

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define OP_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(ONEPROT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements,
// box {64, box_outer}, 128-byte swizzle, out-of-bounds elements read as zero.
int make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(ONEPROT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(ONEPROT_ERR_ARG, "tensor base must be 16-byte aligned and the row pitch a multiple of 8 elements");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ONEPROT_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return ONEPROT_OK;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

template <typename K>
int prep_kernel(K kernel, int smem_bytes) {
  static thread_local const void* done[24] = {nullptr};
  for (auto& p : done) {
    if (p == reinterpret_cast<const void*>(kernel)) return ONEPROT_OK;
  }
  OP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  for (auto& p : done) {
    if (!p) { p = reinterpret_cast<const void*>(kernel); break; }
  }
  return ONEPROT_OK;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// The fused all-gather forward spins on flags that are raised only after EVERY CTA of the grid has pushed its slice:
// the whole grid must be resident at once.  A cooperative launch makes the driver guarantee that (it fails instead of
// starting a grid it cannot place, e.g. beside another kernel that holds an SM) - a plain launch only happens to work.
template <typename Kern>
int launch_coresident(Kern kern, int grid, int smem, cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& w,
                      const op::SParams& p) {
#ifdef ONEPROT_HOST_EMULATION
  (void)kern; (void)grid; (void)smem; (void)st; (void)a; (void)b; (void)w; (void)p;
  return fail(ONEPROT_ERR_ARG, "the fused all-gather forward is not emulated");
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(op::NUM_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  OP_CUDA(cudaLaunchKernelEx(&cfg, kern, a, b, w, p));
  return ONEPROT_OK;
#endif
}

// Row-chunk length CI of an S-kernel work item (item = one 256-column block x CI row blocks).
// Items are dealt round-robin to one persistent CTA per SM, so the makespan is about
// ceil(items / SMs) * CI tiles; pick the CI in [ci_min, 16] that minimises it (ties: longer
// chunks, fewer column flushes; beyond 16 the cold-L2 start gets measurably worse).  Shorter chunks with the same
// makespan (fewer CTAs sharing a streamed row tile) were measured in round 2 at N = 32768 and are slower: FWD_E
// 1.806 / 1.867 / 1.956 ms and 5.89 / 5.96 / 6.14 ms per step for CI = 16 / 8 / 4 (profiles/r2f_ci_ab.txt).
void s_schedule(int rows, int N, int ci_min, op::SParams& p) {
  p.nI = cdiv(rows, op::BM);
  p.nJ = cdiv(N, op::BN);
  const int sms = num_sms();
  long long best_cost = -1;
  int best_ci = 1;
  for (int ci = std::min(ci_min, p.nI); ci <= std::min(16, p.nI); ++ci) {
    const int chunks = cdiv(p.nI, ci);
    const long long items = static_cast<long long>(p.nJ) * chunks;
    const long long cost = cdiv(static_cast<int>(std::min<long long>(items, 1 << 30)), sms) * static_cast<long long>(ci);
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && ci > best_ci)) {
      best_cost = cost;
      best_ci = ci;
    }
  }
  p.CI = std::max(1, best_ci);
  p.nChunks = cdiv(p.nI, p.CI);
}

}  // namespace

namespace opint {
// error reporting for the other translation units of the library (clip_sequence.cu)
int fail(int code, const std::string& msg) { return ::fail(code, msg); }
void count_launch(int n) { g_launches += n; }
}  // namespace opint

extern "C" {

int oneprot_abi_version(void) { return ONEPROT_ABI_VERSION; }
const char* oneprot_last_error(void) { return g_err.c_str(); }
long long oneprot_launch_count(void) { return g_launches.load(); }
void oneprot_launch_count_reset(void) { g_launches.store(0); }

int oneprot_num_sms(void) { return num_sms(); }

int oneprot_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(ONEPROT_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
  if (prop.major != 10) return fail(ONEPROT_ERR_DEVICE, "device is not compute capability 10.x (needs tcgen05/TMEM)");
  return ONEPROT_OK;
}

int oneprot_clip_rowstats(const void* A, const void* B_all, int n, int N, int d, int row_offset, float* diag,
                          float* stats, void* stream) {
  if (!A || !B_all || !diag || !stats || n <= 0 || N <= 0 || d <= 0) return fail(ONEPROT_ERR_ARG, "rowstats: bad argument");
  if (d % 8) return fail(ONEPROT_ERR_ARG, "rowstats: d must be a multiple of 8");
  if (row_offset < 0 || row_offset + n > N) return fail(ONEPROT_ERR_ARG, "rowstats: row_offset out of range");
  if (optrace::recording()) optrace::add("rowstats A=%p B=%p n=%d N=%d d=%d off=%d diag=%p stats=%p st=%p", A, B_all, n, N, d, row_offset, (void*)diag, (void*)stats, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  const int total = n > N ? n : N;
  const int blocks = std::min(cdiv(total, 8), num_sms() * 8);
  op::rowstats_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(A), static_cast<const __nv_bfloat16*>(B_all), n, N, d, row_offset, diag, stats);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

size_t oneprot_clip_fwd_scratch_bytes(int n, int N) {
  op::SParams p{};
  s_schedule(n, N, 2, p);
  const size_t ldr = static_cast<size_t>(p.nI) * op::BM, ldc = static_cast<size_t>(p.nJ) * op::BN;
  return sizeof(float) * (2 * static_cast<size_t>(p.nJ) * ldr + 4 * static_cast<size_t>(p.nChunks) * ldc);
}

int oneprot_clip_fwd_sums(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev,
                          float* stats, float* rowsum, float* colsum, void* scratch, size_t scratch_bytes,
                          void* stream) {
  return oneprot_clip_fwd_sums_ag(A, B_all, n, N, d, scale_dev, stats, nullptr, rowsum, colsum, scratch, scratch_bytes, stream);
}

int oneprot_clip_fwd_sums_ag(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev,
                             float* stats, const oneprot_ag_t* ag, float* rowsum, float* colsum, void* scratch,
                             size_t scratch_bytes, void* stream) {
  return oneprot_clip_fwd_sums_keep(A, B_all, n, N, d, scale_dev, stats, ag, rowsum, colsum, scratch, scratch_bytes, nullptr, 0, stream);
}

int oneprot_clip_fwd_sums_keep(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev,
                               float* stats, const oneprot_ag_t* ag, float* rowsum, float* colsum, void* scratch,
                               size_t scratch_bytes, void* E, int lde, void* stream) {
  if (E && (lde < N || lde % 8 || (reinterpret_cast<uintptr_t>(E) & 15)))
    return fail(ONEPROT_ERR_ARG, "fwd_sums_keep: E must be 16-byte aligned with a row pitch lde >= N that is a multiple of 8");
  if (!A || !B_all || !scale_dev || !stats || !rowsum || !colsum || !scratch) return fail(ONEPROT_ERR_ARG, "fwd_sums: null pointer");
  if (n <= 0 || N <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "fwd_sums: need n, N > 0 and d a positive multiple of 8");
  if (scratch_bytes < oneprot_clip_fwd_scratch_bytes(n, N)) return fail(ONEPROT_ERR_ARG, "fwd_sums: scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  op::SParams p{};
  s_schedule(n, N, 2, p);
  p.rows = n; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = 0;
  p.scale = scale_dev; p.stats = stats;
  p.ldr = p.nI * op::BM; p.ldc = p.nJ * op::BN;
  p.rowpart = static_cast<float*>(scratch);
  p.colpart = p.rowpart + 2 * static_cast<size_t>(p.nJ) * p.ldr;
  if (ag) {
    const int rows = ag->rows_per_rank, W = ag->world, CH = ag->chunks;
    if (!ag->src || !ag->dst_mc || !ag->counters || !ag->flags_mc || !ag->flags || !ag->stats_mc || !ag->stats_all || !ag->stats_out)
      return fail(ONEPROT_ERR_ARG, "fwd_sums_ag: null pointer in oneprot_ag_t");
    if (W <= 0 || W > 8 || ag->rank < 0 || ag->rank >= W || rows * W != N || CH <= 0 || rows % (CH * op::BN))
      return fail(ONEPROT_ERR_ARG, "fwd_sums_ag: need N = world * rows_per_rank and rows_per_rank a multiple of chunks * 256");
    p.ag_src = static_cast<const uint4*>(ag->src);
    p.ag_dst_mc = static_cast<uint4*>(ag->dst_mc);
    p.ag_counters = ag->counters;
    p.ag_flags_mc = ag->flags_mc;
    p.ag_flags = ag->flags;
    p.ag_stats_mc = ag->stats_mc;
    p.ag_stats = ag->stats_all;
    p.stats_out = ag->stats_out;
    p.ag_epoch = ag->epoch;
    p.ag_rank = ag->rank; p.ag_world = W; p.ag_chunks = CH; p.ag_rows = rows;
    p.ag_bpc = rows / CH / op::BN;
    p.ag_chunk16 = static_cast<int>(static_cast<size_t>(rows / CH) * d * 2 / 16);
  }
  if (optrace::recording()) {
    optrace::add("fwd_sums A=%p B=%p n=%d N=%d d=%d scale=%p stats=%p rowsum=%p colsum=%p scratch=%p maxpass=%d st=%p", A, B_all, n, N,
                 d, (const void*)scale_dev, (void*)stats, (void*)rowsum, (void*)colsum, scratch, (!ag && n == N) ? 1 : 0, stream);
    if (ag)
      optrace::add("  ag src=%p dst_mc=%p counters=%p flags_mc=%p flags=%p stats_mc=%p stats_all=%p stats_out=%p epoch=%u rank=%d world=%d chunks=%d rows=%d",
                   ag->src, ag->dst_mc, (void*)ag->counters, (void*)ag->flags_mc, (const void*)ag->flags, (void*)ag->stats_mc,
                   (const void*)ag->stats_all, (void*)ag->stats_out, ag->epoch, ag->rank, ag->world, ag->chunks, ag->rows_per_rank);
  }
  if (E && optrace::recording()) optrace::add("  keep E=%p lde=%d", E, lde);
  if (optrace::dry()) { g_launches += (!ag && n == N) ? 4 : 3; return ONEPROT_OK; }
  CUtensorMap mapA, mapB, mapE;
  int rc;
  if ((rc = make_map(&mapA, A, d, n, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, B_all, d, N, d, op::BN))) return rc;
  if (E && (rc = make_map(&mapE, E, N, n, lde, op::BM))) return rc;    // store side, like the dL/dZ panel: boxes {64 cols, 128 rows}
  constexpr int smem = op::SCfg<op::EPI_FWD>::SMEM;
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_FWD>, smem))) return rc;
  // fused gather: every CTA of the grid pushes a slice, so the grid must be fully co-resident (it is: <= #SMs)
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  if (!ag && n == N) {
    // Robust tier (whole matrix on this GPU): when the Cauchy-Schwarz bound exceeds the fp32 window
    // the max pass finds the exact maximum logit (stats[2], stats[3] = 1) and the sums below use it
    // as reference.  It returns immediately otherwise (device-side decision, no host sync).
    p.stats_out = stats;
    if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_MAX>, smem))) return rc;
    op::clip_s_kernel<op::EPI_MAX><<<grid, op::NUM_THREADS, smem, st>>>(mapA, mapB, mapA /*unused*/, p);
    ++g_launches;
    OP_CUDA(cudaGetLastError());
  }
  if (E) {
    constexpr int smem_e = op::SCfg<op::EPI_FWD_E>::SMEM;
    if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_FWD_E>, smem_e))) return rc;
    if (ag) { if ((rc = launch_coresident(op::clip_s_kernel<op::EPI_FWD_E>, grid, smem_e, st, mapA, mapB, mapE, p))) return rc; }
    else op::clip_s_kernel<op::EPI_FWD_E><<<grid, op::NUM_THREADS, smem_e, st>>>(mapA, mapB, mapE, p);
  } else {
    if (ag) { if ((rc = launch_coresident(op::clip_s_kernel<op::EPI_FWD>, grid, smem, st, mapA, mapB, mapA, p))) return rc; }
    else op::clip_s_kernel<op::EPI_FWD><<<grid, op::NUM_THREADS, smem, st>>>(mapA, mapB, mapA /*unused*/, p);
  }
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  op::reduce_slots_kernel<<<cdiv(n, 32), 256, 0, st>>>(p.rowpart, 2 * p.nJ, p.ldr, n, rowsum);
  op::reduce_slots_kernel<<<cdiv(N, 32), 256, 0, st>>>(p.colpart, 4 * p.nChunks, p.ldc, N, colsum);
  g_launches += 2;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_clip_loss_finalize(const float* rowsum_all, const float* colsum_all, const float* diag_all, int N, int n,
                               int row_offset, int mode, const float* scale_dev, const float* stats, float* loss_out,
                               float* inv_rowsum, float* inv_colsum, int* flag, void* scratch, void* stream) {
  return oneprot_clip_loss_finalize_ex(rowsum_all, colsum_all, diag_all, N, n, row_offset, mode, scale_dev, stats, loss_out,
                                       inv_rowsum, inv_colsum, flag, scratch, nullptr, nullptr, stream);
}

int oneprot_clip_loss_finalize_ex(const float* rowsum_all, const float* colsum_all, const float* diag_all, int N, int n,
                                  int row_offset, int mode, const float* scale_dev, const float* stats, float* loss_out,
                                  float* inv_rowsum, float* inv_colsum, int* flag, void* scratch, const float* row_ref,
                                  const float* col_ref, void* stream) {
  if (!rowsum_all || !colsum_all || !diag_all || !scale_dev || !stats || !loss_out || !inv_rowsum || !inv_colsum || !flag || !scratch)
    return fail(ONEPROT_ERR_ARG, "loss_finalize: null pointer");
  if (N <= 0 || n <= 0 || row_offset < 0 || row_offset + n > N) return fail(ONEPROT_ERR_ARG, "loss_finalize: bad sizes");
  if (reinterpret_cast<uintptr_t>(scratch) & 7) return fail(ONEPROT_ERR_ARG, "loss_finalize: scratch must be 8-byte aligned");
  if ((row_ref != nullptr) != (col_ref != nullptr)) return fail(ONEPROT_ERR_ARG, "loss_finalize: row_ref and col_ref go together");
  if (optrace::recording())
    optrace::add("loss_finalize rowsum=%p colsum=%p diag=%p N=%d n=%d off=%d mode=%d scale=%p stats=%p loss=%p inv_rs=%p inv_cs=%p flag=%p scratch=%p row_ref=%p col_ref=%p st=%p",
                 (const void*)rowsum_all, (const void*)colsum_all, (const void*)diag_all, N, n, row_offset, mode, (const void*)scale_dev,
                 (const void*)stats, (void*)loss_out, (void*)inv_rowsum, (void*)inv_colsum, (void*)flag, scratch, (const void*)row_ref, (const void*)col_ref, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  // scratch: FIN_BLOCKS doubles + one zero-initialised counter (the kernel resets it)
  double* partial = static_cast<double*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(partial + op::FIN_BLOCKS);
  const int blocks = std::min(op::FIN_BLOCKS, cdiv(N, 256));
  op::loss_finalize_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rowsum_all, colsum_all, diag_all, N, n, row_offset, mode, scale_dev, stats, loss_out, inv_rowsum, inv_colsum, flag,
      partial, counter, row_ref, col_ref);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_clip_rowcol_max(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev, float* rowmax,
                            float* colmax, void* scratch, size_t scratch_bytes, void* stream) {
  if (!A || !B_all || !scale_dev || !rowmax || !colmax || !scratch) return fail(ONEPROT_ERR_ARG, "rowcol_max: null pointer");
  if (n <= 0 || N <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "rowcol_max: need n, N > 0 and d a positive multiple of 8");
  if (scratch_bytes < oneprot_clip_fwd_scratch_bytes(n, N)) return fail(ONEPROT_ERR_ARG, "rowcol_max: scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  op::SParams p{};
  s_schedule(n, N, 2, p);
  p.rows = n; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = 0;
  p.scale = scale_dev; p.stats = nullptr;
  p.ldr = p.nI * op::BM; p.ldc = p.nJ * op::BN;
  p.rowpart = static_cast<float*>(scratch);
  p.colpart = p.rowpart + 2 * static_cast<size_t>(p.nJ) * p.ldr;
  CUtensorMap mapA, mapB;
  int rc;
  if ((rc = make_map(&mapA, A, d, n, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, B_all, d, N, d, op::BN))) return rc;
  constexpr int smem = op::SCfg<op::EPI_RCMAX>::SMEM;
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_RCMAX>, smem))) return rc;
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  op::clip_s_kernel<op::EPI_RCMAX><<<grid, op::NUM_THREADS, smem, st>>>(mapA, mapB, mapA /*unused*/, p);
  op::reduce_slots_max_kernel<<<cdiv(n, 32), 256, 0, st>>>(p.rowpart, 2 * p.nJ, p.ldr, n, rowmax);
  op::reduce_slots_max_kernel<<<cdiv(N, 32), 256, 0, st>>>(p.colpart, 4 * p.nChunks, p.ldc, N, colmax);
  g_launches += 3;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_retrieval_ranks(const void* S, const void* M, int N, int d, const float* label_dot, float* rank_s2m,
                            float* rank_m2s, void* scratch, size_t scratch_bytes, void* stream) {
  if (!S || !M || !label_dot || !rank_s2m || !rank_m2s || !scratch) return fail(ONEPROT_ERR_ARG, "retrieval_ranks: null pointer");
  if (N <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "retrieval_ranks: need N > 0 and d a positive multiple of 8");
  if (scratch_bytes < oneprot_clip_fwd_scratch_bytes(N, N)) return fail(ONEPROT_ERR_ARG, "retrieval_ranks: scratch too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  op::SParams p{};
  s_schedule(N, N, 2, p);
  p.rows = N; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = 0;
  p.wr = label_dot; p.wc = label_dot;
  p.ldr = p.nI * op::BM; p.ldc = p.nJ * op::BN;
  p.rowpart = static_cast<float*>(scratch);
  p.colpart = p.rowpart + 2 * static_cast<size_t>(p.nJ) * p.ldr;
  CUtensorMap mapA, mapB;
  int rc;
  if ((rc = make_map(&mapA, S, d, N, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, M, d, N, d, op::BN))) return rc;
  constexpr int smem = op::SCfg<op::EPI_RANK>::SMEM;
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_RANK>, smem))) return rc;
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  op::clip_s_kernel<op::EPI_RANK><<<grid, op::NUM_THREADS, smem, st>>>(mapA, mapB, mapA /*unused*/, p);
  op::reduce_slots_kernel<<<cdiv(N, 32), 256, 0, st>>>(p.rowpart, 2 * p.nJ, p.ldr, N, rank_s2m);
  op::reduce_slots_kernel<<<cdiv(N, 32), 256, 0, st>>>(p.colpart, 4 * p.nChunks, p.ldc, N, rank_m2s);
  g_launches += 3;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_augment_bf16(const void* in, int rows, int d, const float* ref, const float* scale_dev, void* out, float* ref_q,
                         void* stream) {
  if (!in || !out || !scale_dev || rows <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "augment: need d a positive multiple of 8");
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(ONEPROT_ERR_ARG, "augment: pointers must be 16-byte aligned");
  const int blocks = std::min(cdiv(rows, 8), num_sms() * 16);
  op::augment_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), rows, d, ref, scale_dev, static_cast<__nv_bfloat16*>(out), ref_q);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_clip_bwd_weights(const float* inv_rowsum, const float* inv_colsum, int N, int n, int row_offset, int mode,
                             int use_gsum, int part, int world, int rank, const float* gvec_dev,
                             const float* scale_dev, float* wr, float* wc, float* dg, float* out_scale_a,
                             float* out_scale_b, int what, void* stream) {
  if (!inv_rowsum || !inv_colsum || !gvec_dev || !scale_dev || !wr || !wc || !dg || !out_scale_a || !out_scale_b)
    return fail(ONEPROT_ERR_ARG, "bwd_weights: null pointer");
  if (world <= 0 || rank < 0 || rank >= world || N % world || n <= 0 || row_offset + n > N || what < 0 || what > 2)
    return fail(ONEPROT_ERR_ARG, "bwd_weights: bad sizes");
  if (optrace::recording())
    optrace::add("bwd_weights inv_rs=%p inv_cs=%p N=%d n=%d off=%d mode=%d use_gsum=%d part=%d world=%d rank=%d gvec=%p scale=%p wr=%p wc=%p dg=%p sA=%p sB=%p what=%d st=%p",
                 (const void*)inv_rowsum, (const void*)inv_colsum, N, n, row_offset, mode, use_gsum, part, world, rank, (const void*)gvec_dev,
                 (const void*)scale_dev, (void*)wr, (void*)wc, (void*)dg, (void*)out_scale_a, (void*)out_scale_b, what, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  op::bwd_weights_kernel<<<cdiv(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      inv_rowsum, inv_colsum, N, n, row_offset, mode, use_gsum, part, world, rank, gvec_dev, scale_dev, wr, wc, dg,
      out_scale_a, out_scale_b, what);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_clip_dz_panel(const void* A_rows, const void* B_all, int rows, int N, int d, int grow0,
                          const float* scale_dev, const float* stats, const float* wr, const float* wc, const float* dg,
                          void* Wz, int ldw, void* stream) {
  if (!A_rows || !B_all || !scale_dev || !stats || !wr || !wc || !dg || !Wz) return fail(ONEPROT_ERR_ARG, "dz_panel: null pointer");
  if (rows <= 0 || N <= 0 || d <= 0 || d % 8 || ldw < N || ldw % 8) return fail(ONEPROT_ERR_ARG, "dz_panel: bad sizes");
  if (reinterpret_cast<uintptr_t>(Wz) & 15) return fail(ONEPROT_ERR_ARG, "dz_panel: Wz must be 16-byte aligned");
  if (optrace::recording())
    optrace::add("dz_panel A=%p B=%p rows=%d N=%d d=%d grow0=%d scale=%p stats=%p wr=%p wc=%p dg=%p Wz=%p ldw=%d st=%p", A_rows, B_all, rows,
                 N, d, grow0, (const void*)scale_dev, (const void*)stats, (const void*)wr, (const void*)wc, (const void*)dg, Wz, ldw, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  op::SParams p{};
  s_schedule(rows, N, 1, p);
  p.rows = rows; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = grow0;
  p.scale = scale_dev; p.stats = stats;
  p.wr = wr; p.wc = wc; p.dg = dg;
  p.Wz = static_cast<__nv_bfloat16*>(Wz); p.ldw = ldw;
  CUtensorMap mapA, mapB;
  int rc;
  if ((rc = make_map(&mapA, A_rows, d, rows, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, B_all, d, N, d, op::BN))) return rc;
  CUtensorMap mapW;   // store side: boxes {64 cols, 128 rows}; columns >= N and rows >= `rows` are clipped
  if ((rc = make_map(&mapW, Wz, N, rows, ldw, op::BM))) return rc;
  constexpr int smem = op::SCfg<op::EPI_DZ>::SMEM;
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_DZ>, smem))) return rc;
  op::clip_s_kernel<op::EPI_DZ><<<grid, op::NUM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(mapA, mapB, mapW, p);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_clip_dz_from_exp(void* E, int rows, int N, int lde, int grow0, const float* wr, const float* wc, const float* dg,
                             void* stream) {
  if (!E || !wr || !wc || !dg) return fail(ONEPROT_ERR_ARG, "dz_from_exp: null pointer");
  if (rows <= 0 || N <= 0 || lde < N || lde % 8 || (reinterpret_cast<uintptr_t>(E) & 15)) return fail(ONEPROT_ERR_ARG, "dz_from_exp: bad sizes");
  if (optrace::recording())
    optrace::add("dz_from_exp E=%p rows=%d N=%d lde=%d grow0=%d wr=%p wc=%p dg=%p st=%p", E, rows, N, lde, grow0, (const void*)wr,
                 (const void*)wc, (const void*)dg, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  const dim3 grid(cdiv(N, 256 * 8), cdiv(rows, op::DZE_ROWS));
  op::dz_from_exp_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__nv_bfloat16*>(E), rows, N, lde, grow0, wr, wc, dg);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_siglip_fwd(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev, const float* bias_dev,
                       float* rowsum, void* scratch, size_t scratch_bytes, void* stream) {
  if (!A || !B_all || !scale_dev || !rowsum || !scratch) return fail(ONEPROT_ERR_ARG, "siglip_fwd: null pointer");
  if (n <= 0 || N <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "siglip_fwd: need n, N > 0 and d a positive multiple of 8");
  if (scratch_bytes < oneprot_clip_fwd_scratch_bytes(n, N)) return fail(ONEPROT_ERR_ARG, "siglip_fwd: scratch too small");
  if (optrace::recording())
    optrace::add("siglip_fwd A=%p B=%p n=%d N=%d d=%d scale=%p bias=%p rowsum=%p scratch=%p st=%p", A, B_all, n, N, d, (const void*)scale_dev,
                 (const void*)bias_dev, (void*)rowsum, scratch, stream);
  if (optrace::dry()) { g_launches += 2; return ONEPROT_OK; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  op::SParams p{};
  s_schedule(n, N, 2, p);
  p.rows = n; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = 0;
  p.scale = scale_dev; p.wc = bias_dev;
  p.ldr = p.nI * op::BM; p.ldc = p.nJ * op::BN;
  p.rowpart = static_cast<float*>(scratch);
  CUtensorMap mapA, mapB;
  int rc;
  if ((rc = make_map(&mapA, A, d, n, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, B_all, d, N, d, op::BN))) return rc;
  constexpr int smem = op::SCfg<op::EPI_SFWD>::SMEM;
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_SFWD>, smem))) return rc;
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  op::clip_s_kernel<op::EPI_SFWD><<<grid, op::NUM_THREADS, smem, st>>>(mapA, mapB, mapA /*unused*/, p);
  op::reduce_slots_kernel<<<cdiv(n, 32), 256, 0, st>>>(p.rowpart, 2 * p.nJ, p.ldr, n, rowsum);
  g_launches += 2;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

size_t oneprot_siglip_fwd_keep_scratch_bytes(int n, int N) { return 2 * oneprot_clip_fwd_scratch_bytes(n, N); }

int oneprot_siglip_fwd_keep(const void* A, const void* B_all, int n, int N, int d, int grow0, const float* scale_dev,
                            const float* bias_dev, float* rowsum, float* sig_rowsum, void* scratch, size_t scratch_bytes, void* S,
                            int lds, void* stream) {
  if (!A || !B_all || !scale_dev || !rowsum || !scratch || !S) return fail(ONEPROT_ERR_ARG, "siglip_fwd_keep: null pointer");
  if (n <= 0 || N <= 0 || d <= 0 || d % 8 || grow0 < 0 || grow0 + n > N) return fail(ONEPROT_ERR_ARG, "siglip_fwd_keep: bad sizes");
  if (lds < N || lds % 8 || (reinterpret_cast<uintptr_t>(S) & 15))
    return fail(ONEPROT_ERR_ARG, "siglip_fwd_keep: S must be 16-byte aligned with a row pitch lds >= N that is a multiple of 8");
  if (scratch_bytes < oneprot_siglip_fwd_keep_scratch_bytes(n, N)) return fail(ONEPROT_ERR_ARG, "siglip_fwd_keep: scratch too small");
  if (optrace::recording())
    optrace::add("siglip_fwd_keep A=%p B=%p n=%d N=%d d=%d grow0=%d scale=%p bias=%p rowsum=%p sig_rowsum=%p scratch=%p S=%p lds=%d st=%p", A, B_all,
                 n, N, d, grow0, (const void*)scale_dev, (const void*)bias_dev, (void*)rowsum, (void*)sig_rowsum, scratch, S, lds, stream);
  const int launches = sig_rowsum ? 3 : 2;
  if (optrace::dry()) { g_launches += launches; return ONEPROT_OK; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  op::SParams p{};
  s_schedule(n, N, 2, p);
  p.rows = n; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = grow0;
  p.scale = scale_dev; p.wc = bias_dev;
  p.ldr = p.nI * op::BM; p.ldc = p.nJ * op::BN;
  p.rowpart = static_cast<float*>(scratch);
  // second row-partial buffer (row sums of sigma), same [2 nJ][ldr] layout, in the second half of the scratch
  p.colpart = sig_rowsum ? reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + oneprot_clip_fwd_scratch_bytes(n, N)) : nullptr;
  CUtensorMap mapA, mapB, mapS;
  int rc;
  if ((rc = make_map(&mapA, A, d, n, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, B_all, d, N, d, op::BN))) return rc;
  if ((rc = make_map(&mapS, S, N, n, lds, op::BM))) return rc;
  constexpr int smem = op::SCfg<op::EPI_SFWD_K>::SMEM;
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_SFWD_K>, smem))) return rc;
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  op::clip_s_kernel<op::EPI_SFWD_K><<<grid, op::NUM_THREADS, smem, st>>>(mapA, mapB, mapS, p);
  op::reduce_slots_kernel<<<cdiv(n, 32), 256, 0, st>>>(p.rowpart, 2 * p.nJ, p.ldr, n, rowsum);
  if (sig_rowsum) op::reduce_slots_kernel<<<cdiv(n, 32), 256, 0, st>>>(p.colpart, 2 * p.nJ, p.ldr, n, sig_rowsum);
  g_launches += launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_siglip_finalize(const float* rowsum, const float* diag, int n, const float* scale_dev, const float* bias_dev,
                            float* loss_out, void* stream) {
  if (!rowsum || !diag || !scale_dev || !loss_out || n <= 0) return fail(ONEPROT_ERR_ARG, "siglip_finalize: bad argument");
  if (optrace::recording())
    optrace::add("siglip_finalize rowsum=%p diag=%p n=%d scale=%p bias=%p loss=%p st=%p", (const void*)rowsum, (const void*)diag, n,
                 (const void*)scale_dev, (const void*)bias_dev, (void*)loss_out, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  op::siglip_finalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(rowsum, diag, n, scale_dev, bias_dev, loss_out);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

size_t oneprot_siglip_dz_scratch_bytes(int rows, int N) {
  op::SParams p{};
  s_schedule(rows, N, 1, p);
  return sizeof(float) * 2 * static_cast<size_t>(p.nJ) * p.nI * op::BM;
}

int oneprot_siglip_dz_panel(const void* A_rows, const void* B_all, int rows, int N, int d, int grow0, const float* scale_dev,
                            const float* bias_dev, const float* wr, const float* dg, void* Wz, int ldw, float* sig_rowsum,
                            void* scratch, size_t scratch_bytes, void* stream) {
  if (!A_rows || !B_all || !scale_dev || !wr || !dg || !Wz) return fail(ONEPROT_ERR_ARG, "siglip_dz_panel: null pointer");
  if (sig_rowsum && (!scratch || scratch_bytes < oneprot_siglip_dz_scratch_bytes(rows, N)))
    return fail(ONEPROT_ERR_ARG, "siglip_dz_panel: sig_rowsum needs scratch of oneprot_siglip_dz_scratch_bytes(rows, N)");
  if (rows <= 0 || N <= 0 || d <= 0 || d % 8 || ldw < N || ldw % 8) return fail(ONEPROT_ERR_ARG, "siglip_dz_panel: bad sizes");
  if (reinterpret_cast<uintptr_t>(Wz) & 15) return fail(ONEPROT_ERR_ARG, "siglip_dz_panel: Wz must be 16-byte aligned");
  if (optrace::recording())
    optrace::add("siglip_dz_panel A=%p B=%p rows=%d N=%d d=%d grow0=%d scale=%p bias=%p wr=%p dg=%p Wz=%p ldw=%d sig_rowsum=%p scratch=%p st=%p",
                 A_rows, B_all, rows, N, d, grow0, (const void*)scale_dev, (const void*)bias_dev, (const void*)wr, (const void*)dg, Wz, ldw,
                 (void*)sig_rowsum, scratch, stream);
  if (optrace::dry()) { g_launches += sig_rowsum ? 2 : 1; return ONEPROT_OK; }
  op::SParams p{};
  s_schedule(rows, N, 1, p);
  p.rows = rows; p.N = N; p.nK = cdiv(d, op::BK); p.grow0 = grow0;
  p.scale = scale_dev; p.wc = bias_dev;
  p.wr = wr; p.dg = dg;
  p.Wz = static_cast<__nv_bfloat16*>(Wz); p.ldw = ldw;
  if (sig_rowsum) { p.rowpart = static_cast<float*>(scratch); p.ldr = p.nI * op::BM; }
  CUtensorMap mapA, mapB, mapW;
  int rc;
  if ((rc = make_map(&mapA, A_rows, d, rows, d, op::BM))) return rc;
  if ((rc = make_map(&mapB, B_all, d, N, d, op::BN))) return rc;
  if ((rc = make_map(&mapW, Wz, N, rows, ldw, op::BM))) return rc;
  constexpr int smem = op::SCfg<op::EPI_SDZ>::SMEM;
  if ((rc = prep_kernel(op::clip_s_kernel<op::EPI_SDZ>, smem))) return rc;
  const int grid = std::min(num_sms(), p.nJ * p.nChunks);
  op::clip_s_kernel<op::EPI_SDZ><<<grid, op::NUM_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(mapA, mapB, mapW, p);
  ++g_launches;
  if (sig_rowsum) {
    op::reduce_slots_kernel<<<cdiv(rows, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(p.rowpart, 2 * p.nJ, p.ldr, rows, sig_rowsum);
    ++g_launches;
  }
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int Nc, int K,
                      const float* acc_in, float* acc_out, void* out_bf16, int ldc, void* stream) {
  return oneprot_gemm_bf16_ex(A, lda, a_mn, B, ldb, b_mn, M, Nc, K, acc_in, acc_out, out_bf16, ldc, nullptr, nullptr, 0,
                              nullptr, stream);
}

size_t oneprot_gemm_rowdot_scratch_bytes(int M, int Nc) {
  return sizeof(float) * 2 * static_cast<size_t>(cdiv(Nc, op::BN)) * static_cast<size_t>(cdiv(M, op::BM) * op::BM);
}

int oneprot_gemm_bf16_ex(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int Nc, int K,
                         const float* acc_in, float* acc_out, void* out_bf16, int ldc, const float* row_scale,
                         const void* dot_mat, int ld_dot, float* rowdot_part, void* stream) {
  if (!A || !B || (!acc_out && !out_bf16)) return fail(ONEPROT_ERR_ARG, "gemm: null pointer");
  if ((dot_mat != nullptr) != (rowdot_part != nullptr) || (dot_mat && (ld_dot < Nc || ld_dot % 8)))
    return fail(ONEPROT_ERR_ARG, "gemm: dot_mat and rowdot_part go together, ld_dot >= Nc and a multiple of 8");
  if (M <= 0 || Nc <= 0 || K <= 0 || Nc % 8 || ldc % 8 || ldc < Nc) return fail(ONEPROT_ERR_ARG, "gemm: need Nc, ldc multiples of 8, ldc >= Nc");
  if (lda < (a_mn ? M : K) || ldb < (b_mn ? Nc : K)) return fail(ONEPROT_ERR_ARG, "gemm: leading dimension too small");
  if (optrace::recording())
    optrace::add("gemm A=%p lda=%d a_mn=%d B=%p ldb=%d b_mn=%d M=%d Nc=%d K=%d acc_in=%p acc_out=%p out=%p ldc=%d row_scale=%p dot_mat=%p ld_dot=%d rowdot=%p st=%p",
                 A, lda, a_mn, B, ldb, b_mn, M, Nc, K, (const void*)acc_in, (void*)acc_out, out_bf16, ldc, (const void*)row_scale, dot_mat,
                 ld_dot, (void*)rowdot_part, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  op::GParams p{};
  p.M = M; p.Nc = Nc; p.nK = cdiv(K, op::BK);
  p.nMb = cdiv(M, op::BM); p.nNb = cdiv(Nc, op::BN);
  p.acc_in = acc_in; p.acc_out = acc_out; p.out = static_cast<__nv_bfloat16*>(out_bf16); p.ldc = ldc;
  p.row_scale = row_scale;
  p.dot_mat = static_cast<const __nv_bfloat16*>(dot_mat); p.ld_dot = ld_dot;
  p.rowdot_part = rowdot_part; p.ldd = p.nMb * op::BM;
  CUtensorMap mapA, mapB;
  int rc;
  if (a_mn) rc = make_map(&mapA, A, M, K, lda, 64); else rc = make_map(&mapA, A, K, M, lda, op::BM);
  if (rc) return rc;
  if (b_mn) rc = make_map(&mapB, B, Nc, K, ldb, 64); else rc = make_map(&mapB, B, K, Nc, ldb, op::BN);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = std::min(num_sms(), p.nMb * p.nNb);
#define LAUNCH_GEMM(AM, BMJ)                                                                       \
  do {                                                                                             \
    if ((rc = prep_kernel(op::gemm_kernel<AM, BMJ>, op::GEMM_SMEM))) return rc;                    \
    op::gemm_kernel<AM, BMJ><<<grid, op::NUM_THREADS, op::GEMM_SMEM, st>>>(mapA, mapB, p);          \
  } while (0)
  if (!a_mn && !b_mn) LAUNCH_GEMM(0, 0);
  else if (!a_mn && b_mn) LAUNCH_GEMM(0, 1);
  else if (a_mn && !b_mn) LAUNCH_GEMM(1, 0);
  else LAUNCH_GEMM(1, 1);
#undef LAUNCH_GEMM
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_rowdot_bf16(const void* x, int ldx, const void* y, int ldy, int rows, int d, float* out, void* stream) {
  if (!x || !y || !out || rows <= 0 || d <= 0 || d % 8 || ldx % 8 || ldy % 8) return fail(ONEPROT_ERR_ARG, "rowdot: bad argument");
  if (optrace::recording()) optrace::add("rowdot_bf16 x=%p ldx=%d y=%p ldy=%d rows=%d d=%d out=%p st=%p", x, ldx, y, ldy, rows, d, (void*)out, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  const int blocks = std::min(cdiv(rows, 8), num_sms() * 16);
  op::rowdot_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(y), ldy, rows, d, out);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_sum_f32(const float* v, int count, float* out, void* stream) {
  if (!v || !out || count <= 0) return fail(ONEPROT_ERR_ARG, "sum: bad argument");
  if (optrace::recording()) optrace::add("sum_f32 v=%p count=%d out=%p st=%p", (const void*)v, count, (void*)out, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  op::sum_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(v, count, out);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

}  // extern "C"

namespace {
// one wave of persistent 256-thread blocks: resident blocks per SM from the occupancy API, cached per instantiation
template <typename Kern>
int row_grid(Kern kern, int& cache, int rows) {
  if (!cache) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, 0) != cudaSuccess || nb < 1) nb = 1;
    cache = nb;
  }
  return std::max(1, std::min(cdiv(rows, 8), num_sms() * cache));
}
template <bool FP32, int C>
void launch_l2norm_fwd(const void* x, void* y, float* inv_norm, int rows, int d, const float* scale_dev, float eps, cudaStream_t st) {
  static int nb = 0;
  const int blocks = row_grid(op::l2norm_fwd_rows_kernel<FP32, C>, nb, rows);
  op::l2norm_fwd_rows_kernel<FP32, C><<<blocks, 256, 0, st>>>(x, y, inv_norm, rows, d, scale_dev, eps);
}
template <bool FP32, int C>
void launch_l2norm_bwd(const void* x, const void* gy, const float* inv_norm, void* gx, float* dscale_partial, int rows, int d,
                       const float* scale_dev, float eps, cudaStream_t st) {
  static int nb = 0;
  const int blocks = row_grid(op::l2norm_bwd_rows_kernel<FP32, C>, nb, rows);
  op::l2norm_bwd_rows_kernel<FP32, C><<<blocks, 256, 0, st>>>(x, gy, inv_norm, gx, dscale_partial, rows, d, scale_dev, eps);
}
#define ROW_SWITCH_C(cneed, CALL)                                           \
  switch (cneed) {                                                          \
    case 1: CALL(1); break;                                                 \
    case 2: CALL(2); break;                                                 \
    case 3: CALL(3); break;                                                 \
    case 4: CALL(4); break;                                                 \
    case 5: CALL(5); break;                                                 \
    case 6: CALL(6); break;                                                 \
    default: CALL(8); break;                                                \
  }
}  // namespace

extern "C" {

int oneprot_l2norm_scale_fwd(const void* x, void* y, float* inv_norm, int rows, int d, int is_fp32,
                             const float* scale_dev, float eps, void* stream) {
  if (!x || !y || rows <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "l2norm_fwd: need d a positive multiple of 8");
  const int blocks = std::min(cdiv(rows, 8), num_sms() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d <= 256 * oprow::MAXC) {        // the row fits the register-resident kernel
    if (is_fp32) {
#define ROW_CALL(CC) launch_l2norm_fwd<true, CC>(x, y, inv_norm, rows, d, scale_dev, eps, st)
      ROW_SWITCH_C(cdiv(d, 256), ROW_CALL)
#undef ROW_CALL
    } else {
#define ROW_CALL(CC) launch_l2norm_fwd<false, CC>(x, y, inv_norm, rows, d, scale_dev, eps, st)
      ROW_SWITCH_C(cdiv(d, 256), ROW_CALL)
#undef ROW_CALL
    }
  } else if (is_fp32) op::l2norm_fwd_kernel<true><<<blocks, 256, 0, st>>>(x, y, inv_norm, rows, d, scale_dev, eps);
  else op::l2norm_fwd_kernel<false><<<blocks, 256, 0, st>>>(x, y, inv_norm, rows, d, scale_dev, eps);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_l2norm_scale_bwd(const void* x, const void* gy, const float* inv_norm, void* gx, float* dscale_partial,
                             int rows, int d, int is_fp32, const float* scale_dev, float eps, void* stream) {
  if (!x || !gy || !inv_norm || !gx || rows <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "l2norm_bwd: bad argument");
  const int blocks = std::min(cdiv(rows, 8), num_sms() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d <= 256 * oprow::MAXC) {
    if (is_fp32) {
#define ROW_CALL(CC) launch_l2norm_bwd<true, CC>(x, gy, inv_norm, gx, dscale_partial, rows, d, scale_dev, eps, st)
      ROW_SWITCH_C(cdiv(d, 256), ROW_CALL)
#undef ROW_CALL
    } else {
#define ROW_CALL(CC) launch_l2norm_bwd<false, CC>(x, gy, inv_norm, gx, dscale_partial, rows, d, scale_dev, eps, st)
      ROW_SWITCH_C(cdiv(d, 256), ROW_CALL)
#undef ROW_CALL
    }
  } else if (is_fp32) op::l2norm_bwd_kernel<true><<<blocks, 256, 0, st>>>(x, gy, inv_norm, gx, dscale_partial, rows, d, scale_dev, eps);
  else op::l2norm_bwd_kernel<false><<<blocks, 256, 0, st>>>(x, gy, inv_norm, gx, dscale_partial, rows, d, scale_dev, eps);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_scale_rows(const void* x, void* y, int rows, int d, int is_fp32, const float* scale_dev, void* stream) {
  if (!x || !y || !scale_dev || rows <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "scale_rows: need d a positive multiple of 8");
  const size_t total8 = static_cast<size_t>(rows) * d / 8;
  const int blocks = static_cast<int>(std::min<size_t>((total8 + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) op::scale_kernel<true><<<blocks, 256, 0, st>>>(x, y, total8, scale_dev);
  else op::scale_kernel<false><<<blocks, 256, 0, st>>>(x, y, total8, scale_dev);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_rowdot(const void* x, const void* y, int rows, int d, int is_fp32, float* out, void* stream) {
  if (!x || !y || !out || rows <= 0 || d <= 0 || d % 8) return fail(ONEPROT_ERR_ARG, "rowdot: need d a positive multiple of 8");
  const int blocks = std::min(cdiv(rows, 8), num_sms() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_fp32) op::rowdot_dense_kernel<true><<<blocks, 256, 0, st>>>(x, y, rows, d, out);
  else op::rowdot_dense_kernel<false><<<blocks, 256, 0, st>>>(x, y, rows, d, out);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_mc_store(const void* src, void* dst_mc, size_t bytes, void* stream) {
  if (!src || !dst_mc || bytes == 0 || bytes % 16 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst_mc) & 15))
    return fail(ONEPROT_ERR_ARG, "mc_store: need 16-byte aligned pointers and size");
  if (optrace::recording()) optrace::add("mc_store src=%p dst_mc=%p bytes=%zu st=%p", src, dst_mc, bytes, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  const size_t n16 = bytes / 16;
  const int blocks = static_cast<int>(std::min<size_t>((n16 + 255) / 256, static_cast<size_t>(num_sms()) * 8));
#ifndef ONEPROT_KERNEL_EMULATION   // multimem exchanges are not emulated
  op::mc_store_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst_mc), n16);
#else
  return fail(ONEPROT_ERR_DEVICE, "multimem exchanges are not available under the CPU emulation");
#endif
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_mc_allreduce_f32(const float* src_mc, float* dst, int count, int op, void* stream) {
  if (!src_mc || !dst || count <= 0 || count % 4 || (op != 0 && op != 1) || (reinterpret_cast<uintptr_t>(src_mc) & 15) ||
      (reinterpret_cast<uintptr_t>(dst) & 15))
    return fail(ONEPROT_ERR_ARG, "mc_allreduce: count must be a multiple of 4, pointers 16-byte aligned");
  if (optrace::recording()) optrace::add("mc_allreduce_f32 src_mc=%p dst=%p count=%d op=%d st=%p", (const void*)src_mc, (void*)dst, count, op, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
#ifndef ONEPROT_KERNEL_EMULATION   // multimem exchanges are not emulated
  op::mc_allreduce_kernel<<<cdiv(count / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src_mc, dst, count, op);
#else
  return fail(ONEPROT_ERR_DEVICE, "multimem exchanges are not available under the CPU emulation");
#endif
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_mc_reduce_bf16(const void* src_mc, void* dst, size_t bytes, void* stream) {
  if (!src_mc || !dst || bytes == 0 || bytes % 16 || (reinterpret_cast<uintptr_t>(src_mc) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return fail(ONEPROT_ERR_ARG, "mc_reduce_bf16: need 16-byte aligned pointers and size");
  if (optrace::recording()) optrace::add("mc_reduce_bf16 src_mc=%p dst=%p bytes=%zu st=%p", src_mc, dst, bytes, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  const size_t n16 = bytes / 16;
  const int blocks = static_cast<int>(std::min<size_t>((n16 + 255) / 256, static_cast<size_t>(num_sms()) * 8));
#ifndef ONEPROT_KERNEL_EMULATION   // multimem exchanges are not emulated
  op::mc_reduce_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src_mc), static_cast<uint4*>(dst), n16);
#else
  return fail(ONEPROT_ERR_DEVICE, "multimem exchanges are not available under the CPU emulation");
#endif
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

int oneprot_split_fp32(const float* x, void* out, int rows, int d, int side, int terms, void* stream) {
  if (!x || !out || rows <= 0 || d <= 0 || (terms != 3 && terms != 6) || (side != 0 && side != 1))
    return fail(ONEPROT_ERR_ARG, "split_fp32: bad argument");
  if (optrace::recording()) optrace::add("split_fp32 x=%p out=%p rows=%d d=%d side=%d terms=%d st=%p", (const void*)x, out, rows, d, side, terms, stream);
  if (optrace::dry()) { ++g_launches; return ONEPROT_OK; }
  const size_t total = static_cast<size_t>(rows) * d;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  op::split_fp32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(out), rows, d, side, terms);
  ++g_launches;
  OP_CUDA(cudaGetLastError());
  return ONEPROT_OK;
}

}  // extern "C"
#endif  // ONEPROT_KERNEL_EMULATION (host side)
