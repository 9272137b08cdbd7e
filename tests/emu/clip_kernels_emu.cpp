// Compiles the tensor-core kernels of oneprot_b200/csrc/clip_kernels.cu for the CPU (ptx_emu.h supplies
// TMA / mbarrier / tcgen05 / TMEM stand-ins) and exposes C entry points that set the kernel parameters up
// exactly like the CUDA host functions of that file do.  Test infrastructure only.
#include <cmath>
#define ONEPROT_KERNEL_EMULATION 1
#include "../../oneprot_b200/csrc/clip_kernels.cu"

namespace {
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
int g_sms = 3;      // a small "GPU": persistent CTAs loop over several work items each

void make_map(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  std::memset(m, 0, sizeof(*m));
  op::EmuMap e{static_cast<const uint16_t*>(ptr), inner, outer, ld, box_outer};
  std::memcpy(m, &e, sizeof(e));
}

// s_schedule of clip_kernels.cu with the emulated SM count
void s_schedule(int rows, int N, int ci_min, op::SParams& p) {
  p.nI = cdiv(rows, op::BM);
  p.nJ = cdiv(N, op::BN);
  long long best_cost = -1;
  int best_ci = 1;
  for (int ci = std::min(ci_min, p.nI); ci <= std::min(16, p.nI); ++ci) {
    const int chunks = cdiv(p.nI, ci);
    const long long items = static_cast<long long>(p.nJ) * chunks;
    const long long cost = cdiv(static_cast<int>(items), g_sms) * static_cast<long long>(ci);
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && ci > best_ci)) { best_cost = cost; best_ci = ci; }
  }
  p.CI = std::max(1, best_ci);
  p.nChunks = cdiv(p.nI, p.CI);
}

template <int EPI>
void run_s(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const op::SParams& p) {
  const int grid = std::min(g_sms, p.nJ * p.nChunks);
  emu::launch(dim3(grid), dim3(op::NUM_THREADS), [&] { op::clip_s_kernel<EPI>(mA, mB, mW, p); });
}
void reduce(const float* part, int slots, int ld, int count, float* out, bool is_max) {
  emu::launch(dim3(cdiv(count, 32)), dim3(256), [&] {
    if (is_max) op::reduce_slots_max_kernel(part, slots, ld, count, out); else op::reduce_slots_kernel(part, slots, ld, count, out);
  });
}
void s_common(op::SParams& p, int rows, int N, int d, int ci_min, float* scratch) {
  s_schedule(rows, N, ci_min, p);
  p.rows = rows; p.N = N; p.nK = cdiv(d, op::BK);
  p.ldr = p.nI * op::BM; p.ldc = p.nJ * op::BN;
  p.rowpart = scratch;
  p.colpart = scratch ? scratch + 2 * static_cast<size_t>(p.nJ) * p.ldr : nullptr;
}
}  // namespace

extern "C" {

void emu_set_sms(int sms) { g_sms = sms; }
size_t emu_s_scratch_floats(int n, int N) {
  op::SParams p{};
  s_schedule(n, N, 2, p);
  return 2 * static_cast<size_t>(p.nJ) * p.nI * op::BM + 4 * static_cast<size_t>(p.nChunks) * p.nJ * op::BN;
}

// oneprot_clip_fwd_sums: [max pass when the whole matrix is here] + forward + the two slot reductions
void emu_fwd_sums(const void* A, const void* B, int n, int N, int d, const float* scale, float* stats, float* rowsum, float* colsum,
                  float* scratch, void* E, int lde) {
  op::SParams p{};
  s_common(p, n, N, d, 2, scratch);
  p.scale = scale; p.stats = stats;
  CUtensorMap mA, mB, mE;
  make_map(&mA, A, d, n, d, op::BM);
  make_map(&mB, B, d, N, d, op::BN);
  if (n == N) { p.stats_out = stats; run_s<op::EPI_MAX>(mA, mB, mA, p); }
  if (E) {                       // stored-exponentials forward
    make_map(&mE, E, N, n, lde, op::BM);
    run_s<op::EPI_FWD_E>(mA, mB, mE, p);
  } else {
    run_s<op::EPI_FWD>(mA, mB, mA, p);
  }
  reduce(p.rowpart, 2 * p.nJ, p.ldr, n, rowsum, false);
  reduce(p.colpart, 4 * p.nChunks, p.ldc, N, colsum, false);
}

void emu_dz_panel(const void* A_rows, const void* B, int rows, int N, int d, int grow0, const float* scale, const float* stats,
                  const float* wr, const float* wc, const float* dg, void* Wz, int ldw, int l2_hints) {
  op::SParams p{};
  s_common(p, rows, N, d, 1, nullptr);
  p.grow0 = grow0; p.scale = scale; p.stats = stats; p.wr = wr; p.wc = wc; p.dg = dg;
  p.Wz = static_cast<__nv_bfloat16*>(Wz); p.ldw = ldw;
  CUtensorMap mA, mB, mW;
  make_map(&mA, A_rows, d, rows, d, op::BM);
  make_map(&mB, B, d, N, d, op::BN);
  make_map(&mW, Wz, N, rows, ldw, op::BM);
  (void)l2_hints;
  run_s<op::EPI_DZ>(mA, mB, mW, p);
}

void emu_rowcol_max(const void* A, const void* B, int n, int N, int d, const float* scale, float* rowmax, float* colmax, float* scratch) {
  op::SParams p{};
  s_common(p, n, N, d, 2, scratch);
  p.scale = scale;
  CUtensorMap mA, mB;
  make_map(&mA, A, d, n, d, op::BM);
  make_map(&mB, B, d, N, d, op::BN);
  run_s<op::EPI_RCMAX>(mA, mB, mA, p);
  reduce(p.rowpart, 2 * p.nJ, p.ldr, n, rowmax, true);
  reduce(p.colpart, 4 * p.nChunks, p.ldc, N, colmax, true);
}

void emu_retrieval_ranks(const void* S, const void* M, int N, int d, const float* label_dot, float* r_s2m, float* r_m2s, float* scratch) {
  op::SParams p{};
  s_common(p, N, N, d, 2, scratch);
  p.wr = label_dot; p.wc = label_dot;
  CUtensorMap mA, mB;
  make_map(&mA, S, d, N, d, op::BM);
  make_map(&mB, M, d, N, d, op::BN);
  run_s<op::EPI_RANK>(mA, mB, mA, p);
  reduce(p.rowpart, 2 * p.nJ, p.ldr, N, r_s2m, false);
  reduce(p.colpart, 4 * p.nChunks, p.ldc, N, r_m2s, false);
}

void emu_siglip_fwd(const void* A, const void* B, int n, int N, int d, const float* scale, const float* bias, float* rowsum, float* scratch) {
  op::SParams p{};
  s_common(p, n, N, d, 2, scratch);
  p.scale = scale; p.wc = bias;
  CUtensorMap mA, mB;
  make_map(&mA, A, d, n, d, op::BM);
  make_map(&mB, B, d, N, d, op::BN);
  run_s<op::EPI_SFWD>(mA, mB, mA, p);
  reduce(p.rowpart, 2 * p.nJ, p.ldr, n, rowsum, false);
}

void emu_siglip_fwd_keep(const void* A, const void* B, int n, int N, int d, int grow0, const float* scale, const float* bias,
                         float* rowsum, float* sig_rowsum, float* scratch, float* scratch2, void* S, int lds) {
  op::SParams p{};
  s_common(p, n, N, d, 2, scratch);
  p.grow0 = grow0; p.scale = scale; p.wc = bias;
  p.colpart = sig_rowsum ? scratch2 : nullptr;
  CUtensorMap mA, mB, mS;
  make_map(&mA, A, d, n, d, op::BM);
  make_map(&mB, B, d, N, d, op::BN);
  make_map(&mS, S, N, n, lds, op::BM);
  run_s<op::EPI_SFWD_K>(mA, mB, mS, p);
  reduce(p.rowpart, 2 * p.nJ, p.ldr, n, rowsum, false);
  if (sig_rowsum) reduce(p.colpart, 2 * p.nJ, p.ldr, n, sig_rowsum, false);
}

void emu_siglip_dz(const void* A_rows, const void* B, int rows, int N, int d, int grow0, const float* scale, const float* bias,
                   const float* wr, const float* dg, void* Wz, int ldw, float* sig_rowsum, float* scratch) {
  op::SParams p{};
  s_common(p, rows, N, d, 1, sig_rowsum ? scratch : nullptr);
  if (!sig_rowsum) p.rowpart = nullptr;
  p.grow0 = grow0; p.scale = scale; p.wc = bias; p.wr = wr; p.dg = dg;
  p.Wz = static_cast<__nv_bfloat16*>(Wz); p.ldw = ldw;
  CUtensorMap mA, mB, mW;
  make_map(&mA, A_rows, d, rows, d, op::BM);
  make_map(&mB, B, d, N, d, op::BN);
  make_map(&mW, Wz, N, rows, ldw, op::BM);
  run_s<op::EPI_SDZ>(mA, mB, mW, p);
  if (sig_rowsum) reduce(p.rowpart, 2 * p.nJ, p.ldr, rows, sig_rowsum, false);
}

// oneprot_gemm_bf16_ex
void emu_gemm(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int Nc, int K, const float* acc_in, float* acc_out,
              void* out, int ldc, const float* row_scale, const void* dot_mat, int ld_dot, float* rowdot_part) {
  op::GParams p{};
  p.M = M; p.Nc = Nc; p.nK = cdiv(K, op::BK);
  p.nMb = cdiv(M, op::BM); p.nNb = cdiv(Nc, op::BN);
  p.acc_in = acc_in; p.acc_out = acc_out; p.out = static_cast<__nv_bfloat16*>(out); p.ldc = ldc;
  p.row_scale = row_scale;
  p.dot_mat = static_cast<const __nv_bfloat16*>(dot_mat); p.ld_dot = ld_dot;
  p.rowdot_part = rowdot_part; p.ldd = p.nMb * op::BM;
  CUtensorMap mA, mB;
  if (a_mn) make_map(&mA, A, M, K, lda, 64); else make_map(&mA, A, K, M, lda, op::BM);
  if (b_mn) make_map(&mB, B, Nc, K, ldb, 64); else make_map(&mB, B, K, Nc, ldb, op::BN);
  const int grid = std::min(g_sms, p.nMb * p.nNb);
  emu::launch(dim3(grid), dim3(op::NUM_THREADS), [&] {
    if (!a_mn && !b_mn) op::gemm_kernel<0, 0>(mA, mB, p);
    else if (!a_mn && b_mn) op::gemm_kernel<0, 1>(mA, mB, p);
    else if (a_mn && !b_mn) op::gemm_kernel<1, 0>(mA, mB, p);
    else op::gemm_kernel<1, 1>(mA, mB, p);
  });
}

}  // extern "C"
