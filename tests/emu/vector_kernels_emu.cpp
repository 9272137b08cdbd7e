// Compiles oneprot_b200/csrc/vector_kernels.cuh (the non-tensor-core kernels of the ClipLoss path) for the
// CPU through tests/emu/cuda_emu.h, with the launch shapes of the CUDA host code.  Test infrastructure only.
#include "cuda_emu.h"
#include "../../include/oneprot_clip.h"
#include "../../oneprot_b200/csrc/row_regs.cuh"

namespace op {
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float G_MARGIN = 100.0f;
inline uint32_t pack_bf16x2(float lo, float hi) {        // ptx.cuh: cvt.rn.bf16x2.f32 (round to nearest even)
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  uint32_t r;
  std::memcpy(&r, &t, 4);
  return r;
}
#include "../../oneprot_b200/csrc/vector_kernels.cuh"
}  // namespace op

namespace {
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int SMS = 4;
}

extern "C" {

void emu_rowstats(const void* A, const void* B, int n, int N, int d, int off, float* diag, float* stats) {
  const int total = std::max(n, N);
  emu::launch(dim3(std::min(cdiv(total, 8), SMS * 8)), dim3(256), [&] {
    op::rowstats_kernel(static_cast<const __nv_bfloat16*>(A), static_cast<const __nv_bfloat16*>(B), n, N, d, off, diag, stats);
  });
}
void emu_reduce_slots(const float* part, int slots, int ld, int count, float* out, int is_max) {
  emu::launch(dim3(cdiv(count, 32)), dim3(256), [&] {
    if (is_max) op::reduce_slots_max_kernel(part, slots, ld, count, out); else op::reduce_slots_kernel(part, slots, ld, count, out);
  });
}
void emu_augment(const void* in, int rows, int d, const float* ref, const float* scale, void* out, float* ref_q) {
  emu::launch(dim3(std::min(cdiv(rows, 8), SMS * 16)), dim3(256), [&] {
    op::augment_kernel(static_cast<const __nv_bfloat16*>(in), rows, d, ref, scale, static_cast<__nv_bfloat16*>(out), ref_q);
  });
}
void emu_dz_from_exp(void* E, int rows, int N, int lde, int grow0, const float* wr, const float* wc, const float* dg) {
  emu::launch(dim3(cdiv(N, 256 * 8), cdiv(rows, op::DZE_ROWS)), dim3(256), [&] {
    op::dz_from_exp_kernel(static_cast<__nv_bfloat16*>(E), rows, N, lde, grow0, wr, wc, dg);
  });
}
void emu_loss_finalize(const float* rowsum, const float* colsum, const float* diag, int N, int n, int off, int mode, const float* scale,
                       const float* stats, float* loss, float* inv_rs, float* inv_cs, int* flag, double* partial, unsigned int* counter,
                       const float* row_ref, const float* col_ref) {
  emu::launch(dim3(std::min(op::FIN_BLOCKS, cdiv(N, 256))), dim3(256), [&] {
    op::loss_finalize_kernel(rowsum, colsum, diag, N, n, off, mode, scale, stats, loss, inv_rs, inv_cs, flag, partial, counter, row_ref, col_ref);
  });
}
void emu_siglip_finalize(const float* rowsum, const float* diag, int n, const float* scale, const float* bias, float* loss) {
  emu::launch(dim3(1), dim3(1024), [&] { op::siglip_finalize_kernel(rowsum, diag, n, scale, bias, loss); });
}
void emu_bwd_weights(const float* inv_rs, const float* inv_cs, int N, int n, int off, int mode, int use_gsum, int part, int world, int rank,
                     const float* gvec, const float* scale, float* wr, float* wc, float* dg, float* sa, float* sb, int what) {
  emu::launch(dim3(cdiv(N, 256)), dim3(256), [&] {
    op::bwd_weights_kernel(inv_rs, inv_cs, N, n, off, mode, use_gsum, part, world, rank, gvec, scale, wr, wc, dg, sa, sb, what);
  });
}
void emu_rowdot_bf16(const void* x, int ldx, const void* y, int ldy, int rows, int d, float* out) {
  emu::launch(dim3(std::min(cdiv(rows, 8), SMS * 16)), dim3(256), [&] {
    op::rowdot_kernel(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(y), ldy, rows, d, out);
  });
}
void emu_sum_f32(const float* v, int count, float* out) {
  emu::launch(dim3(1), dim3(1024), [&] { op::sum_kernel(v, count, out); });
}
#define EMU_ROW_SWITCH_C(cneed, CALL)                                       \
  switch (cneed) {                                                          \
    case 1: CALL(1); break;                                                 \
    case 2: CALL(2); break;                                                 \
    case 3: CALL(3); break;                                                 \
    case 4: CALL(4); break;                                                 \
    case 5: CALL(5); break;                                                 \
    case 6: CALL(6); break;                                                 \
    default: CALL(8); break;                                                \
  }
void emu_l2norm_fwd(const void* x, void* y, float* inv, int rows, int d, int fp32, const float* scale, float eps) {
  const bool lng = d > 256 * oprow::MAXC;          // register-resident rows up to 2048 elements, as the library dispatches
  emu::launch(dim3(std::min(cdiv(rows, 8), SMS * (lng ? 16 : 2))), dim3(256), [&] {
    if (lng) { if (fp32) op::l2norm_fwd_kernel<true>(x, y, inv, rows, d, scale, eps); else op::l2norm_fwd_kernel<false>(x, y, inv, rows, d, scale, eps); }
    else if (fp32) {
#define EMU_CALL(CC) op::l2norm_fwd_rows_kernel<true, CC>(x, y, inv, rows, d, scale, eps)
      EMU_ROW_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
    } else {
#define EMU_CALL(CC) op::l2norm_fwd_rows_kernel<false, CC>(x, y, inv, rows, d, scale, eps)
      EMU_ROW_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
    }
  });
}
void emu_l2norm_bwd(const void* x, const void* gy, const float* inv, void* gx, float* dsp, int rows, int d, int fp32, const float* scale, float eps) {
  const bool lng = d > 256 * oprow::MAXC;
  emu::launch(dim3(std::min(cdiv(rows, 8), SMS * (lng ? 16 : 2))), dim3(256), [&] {
    if (lng) { if (fp32) op::l2norm_bwd_kernel<true>(x, gy, inv, gx, dsp, rows, d, scale, eps); else op::l2norm_bwd_kernel<false>(x, gy, inv, gx, dsp, rows, d, scale, eps); }
    else if (fp32) {
#define EMU_CALL(CC) op::l2norm_bwd_rows_kernel<true, CC>(x, gy, inv, gx, dsp, rows, d, scale, eps)
      EMU_ROW_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
    } else {
#define EMU_CALL(CC) op::l2norm_bwd_rows_kernel<false, CC>(x, gy, inv, gx, dsp, rows, d, scale, eps)
      EMU_ROW_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
    }
  });
}
void emu_scale_rows(const void* x, void* y, int rows, int d, int fp32, const float* scale) {
  const size_t total8 = static_cast<size_t>(rows) * d / 8;
  emu::launch(dim3(static_cast<unsigned>(std::min<size_t>((total8 + 255) / 256, SMS * 16))), dim3(256), [&] {
    if (fp32) op::scale_kernel<true>(x, y, total8, scale); else op::scale_kernel<false>(x, y, total8, scale);
  });
}
void emu_rowdot(const void* x, const void* y, int rows, int d, int fp32, float* out) {
  emu::launch(dim3(std::min(cdiv(rows, 8), SMS * 16)), dim3(256), [&] {
    if (fp32) op::rowdot_dense_kernel<true>(x, y, rows, d, out); else op::rowdot_dense_kernel<false>(x, y, rows, d, out);
  });
}
void emu_split_fp32(const float* x, void* out, int rows, int d, int side, int terms) {
  const size_t total = static_cast<size_t>(rows) * d;
  emu::launch(dim3(static_cast<unsigned>(std::min<size_t>((total + 255) / 256, SMS * 16))), dim3(256), [&] {
    op::split_fp32_kernel(x, static_cast<__nv_bfloat16*>(out), rows, d, side, terms);
  });
}

}  // extern "C"
