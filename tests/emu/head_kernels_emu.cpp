// Compiles the kernel SOURCE of oneprot_b200/csrc/head_kernels.cu for the CPU (tests/emu/cuda_emu.h) and
// exposes one C entry point per kernel with the same grid / block shapes the CUDA host code uses.
// Built by tests/test_kernel_emulation_cpu.py with g++; test infrastructure only.
#include "cuda_emu.h"

#define ONEPROT_KERNEL_EMULATION 1
#include "../../oneprot_b200/csrc/head_kernels.cu"

namespace {
inline int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr int SMS = 4;          // a small "GPU": grids are sized from the SM count like on the device
int ln_row_chunks(int rows, int d) {
  const int tiles = cdiv(d, 256);
  const int want = std::max(1, 2 * SMS / tiles);
  return std::max(1, std::min(cdiv(rows, 64), want));
}
}  // namespace

extern "C" {

#define EMU_SWITCH_C(cneed, CALL)                                           \
  switch (cneed) {                                                          \
    case 1: CALL(1); break;                                                 \
    case 2: CALL(2); break;                                                 \
    case 3: CALL(3); break;                                                 \
    case 4: CALL(4); break;                                                 \
    case 5: CALL(5); break;                                                 \
    case 6: CALL(6); break;                                                 \
    default: CALL(8); break;                                                \
  }

void emu_layernorm_fwd(const void* x, const void* g, const void* b, void* y, float* mean, float* rstd, int rows, int d, int fp32, float eps) {
  const int blocks = std::min(cdiv(rows, 8), SMS * 2);
  const bool lng = d > 256 * oph::LN_MAXC;
  emu::launch(dim3(blocks), dim3(256), [&] {
    if (lng) { if (fp32) oph::layernorm_fwd_long_kernel<true>(x, g, b, y, mean, rstd, rows, d, eps); else oph::layernorm_fwd_long_kernel<false>(x, g, b, y, mean, rstd, rows, d, eps); }
    else if (fp32) {
#define EMU_CALL(CC) oph::layernorm_fwd_kernel<true, CC>(x, g, b, y, mean, rstd, rows, d, eps)
      EMU_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
    } else {
#define EMU_CALL(CC) oph::layernorm_fwd_kernel<false, CC>(x, g, b, y, mean, rstd, rows, d, eps)
      EMU_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
    }
  });
}

int emu_ln_slots(int rows, int d) { return d > 256 * oph::LN_MAXC ? ln_row_chunks(rows, d) : std::max(1, std::min(cdiv(rows, 8), SMS * 2)); }

void emu_layernorm_bwd(const void* x, const void* gy, const void* g, const float* mean, const float* rstd, void* gx, float* dgamma,
                       float* dbeta, float* scratch, int rows, int d, int fp32) {
  const bool lng = d > 256 * oph::LN_MAXC;
  const int ld = cdiv(d, 256) * 256;
  const int slots = emu_ln_slots(rows, d);
  float* pg = scratch;
  float* pb = pg + static_cast<size_t>(slots) * ld;
  if (!lng) {        // one fused pass: row part + warp-private column sums, one partial slot per block
    emu::launch(dim3(slots), dim3(256), [&] {
      if (fp32) {
#define EMU_CALL(CC) oph::layernorm_bwd_kernel<true, CC>(x, gy, g, mean, rstd, gx, pg, pb, ld, rows, d)
        EMU_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
      } else {
#define EMU_CALL(CC) oph::layernorm_bwd_kernel<false, CC>(x, gy, g, mean, rstd, gx, pg, pb, ld, rows, d)
        EMU_SWITCH_C(cdiv(d, 256), EMU_CALL)
#undef EMU_CALL
      }
    });
  } else {
    const int blocks = std::min(cdiv(rows, 8), SMS * 8);
    emu::launch(dim3(blocks), dim3(256), [&] {
      if (fp32) oph::layernorm_bwd_rows_long_kernel<true>(x, gy, g, mean, rstd, gx, rows, d); else oph::layernorm_bwd_rows_long_kernel<false>(x, gy, g, mean, rstd, gx, rows, d);
    });
    const int rpc = cdiv(rows, slots);
    emu::launch(dim3(cdiv(d, 256), slots), dim3(256), [&] {
      if (fp32) oph::layernorm_bwd_cols_kernel<true>(x, gy, mean, rstd, rows, d, rpc, pg, pb, ld);
      else oph::layernorm_bwd_cols_kernel<false>(x, gy, mean, rstd, rows, d, rpc, pg, pb, ld);
    });
  }
  emu::launch(dim3(cdiv(d, 256)), dim3(256), [&] { oph::sum_slots_f32_kernel(pg, slots, ld, d, dgamma); });
  emu::launch(dim3(cdiv(d, 256)), dim3(256), [&] { oph::sum_slots_f32_kernel(pb, slots, ld, d, dbeta); });
}
int emu_ln_scratch_floats(int rows, int d) { return 2 * emu_ln_slots(rows, d) * cdiv(d, 256) * 256; }

void emu_gelu(const void* x, const void* gy, void* out, size_t count, int fp32) {
  const size_t total8 = count / 8;
  const int blocks = static_cast<int>(std::min<size_t>((total8 + 255) / 256, SMS * 16));
  emu::launch(dim3(blocks), dim3(256), [&] {
    if (gy) { if (fp32) oph::gelu_kernel<true, true>(x, gy, out, total8); else oph::gelu_kernel<false, true>(x, gy, out, total8); }
    else { if (fp32) oph::gelu_kernel<true, false>(x, nullptr, out, total8); else oph::gelu_kernel<false, false>(x, nullptr, out, total8); }
  });
}

void emu_abs_mean_fwd(const void* x, size_t count, size_t true_count, int fp32, float* out, float* part) {
  const size_t total8 = count / 8;
  const int blocks = static_cast<int>(std::min<size_t>(std::min<size_t>((total8 + 255) / 256, SMS * 16), oph::ABS_MAX_BLOCKS));
  emu::launch(dim3(blocks), dim3(256), [&] {
    if (fp32) oph::abs_sum_kernel<true>(x, total8, part); else oph::abs_sum_kernel<false>(x, total8, part);
  });
  emu::launch(dim3(1), dim3(256), [&] { oph::abs_mean_final_kernel(part, blocks, 1.0f / static_cast<float>(true_count), out); });
}
void emu_abs_mean_bwd(const void* x, const float* g, size_t count, size_t true_count, int fp32, void* gx) {
  const size_t total8 = count / 8;
  const int blocks = static_cast<int>(std::min<size_t>((total8 + 255) / 256, SMS * 16));
  emu::launch(dim3(blocks), dim3(256), [&] {
    if (fp32) oph::abs_mean_bwd_kernel<true>(x, g, 1.0f / static_cast<float>(true_count), gx, total8);
    else oph::abs_mean_bwd_kernel<false>(x, g, 1.0f / static_cast<float>(true_count), gx, total8);
  });
}

void emu_meanpool_fwd(const void* x, const float* mask, void* y, float* inv, int B, int L, int D, int fp32, int normalize) {
  emu::launch(dim3(B, cdiv(D, 256)), dim3(256), [&] {
    if (fp32) oph::meanpool_fwd_kernel<true>(x, mask, y, inv, L, D, normalize);
    else oph::meanpool_fwd_kernel<false>(x, mask, y, inv, L, D, normalize);
  });
}

void emu_meanpool_bwd(const void* gy, const float* mask, const float* inv, void* gx, int B, int L, int D, int fp32) {
  emu::launch(dim3(B, cdiv(D, 256)), dim3(256), [&] {
    if (fp32) oph::meanpool_bwd_kernel<true>(gy, mask, inv, gx, L, D);
    else oph::meanpool_bwd_kernel<false>(gy, mask, inv, gx, L, D);
  });
}

void emu_token_dot(const void* x, const void* vec, int per_batch, const float* bias, const float* mask, float* out, int B, int L, int D, int fp32) {
  const int blocks = std::min(cdiv(B * L, 8), SMS * 16);
  const size_t stride = per_batch ? D : 0;
  emu::launch(dim3(blocks), dim3(256), [&] {
    if (fp32) oph::token_dot_kernel<true>(x, vec, stride, bias, mask, out, B, L, D);
    else oph::token_dot_kernel<false>(x, vec, stride, bias, mask, out, B, L, D);
  });
}

void emu_softmax_rows(const float* s, float* p, int B, int L) {
  emu::launch(dim3(B), dim3(256), [&] { oph::softmax_rows_kernel(s, p, L); });
}
void emu_softmax_rows_bwd(const float* p, const float* dp, float* ds, int B, int L) {
  emu::launch(dim3(B), dim3(256), [&] { oph::softmax_rows_bwd_kernel(p, dp, ds, L); });
}
void emu_attnpool_bwd_x(const void* g, const float* p, const float* ds, const void* w, void* gx, int B, int L, int D, int fp32) {
  emu::launch(dim3(B, cdiv(D, 256)), dim3(256), [&] {
    if (fp32) oph::attnpool_bwd_x_kernel<true>(g, p, ds, w, gx, L, D);
    else oph::attnpool_bwd_x_kernel<false>(g, p, ds, w, gx, L, D);
  });
}

}  // extern "C"
