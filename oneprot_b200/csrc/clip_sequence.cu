// Host-side step sequencer of the ClipLoss hot path + the launch trace (host_trace.h).
//
// One ClipLoss fwd+bwd through the Python host is ~35 host operations (ctypes calls, torch
// allocations, memsets, event records): ~0.4 ms of enqueue time per rank.  That is hidden at
// N = 32768 on one GPU (6.4 ms of kernels) but not at 8 GPUs (0.9 ms per step, where one late host
// stalls every GPU at the next in-kernel flag wait) and it dominates OneProt's shipped batch sizes
// (N = 2048-8192: 50-400 us of kernels).  The entry points below enqueue a whole phase of the step
// - memsets, kernels of clip_kernels.cu, event records / waits between the compute stream and the
// exchange stream - from ONE call, out of ONE caller-provided workspace.  They issue exactly the
// launches the Python host issues (oneprot_b200/clip_loss.py::_forward_impl/_backward_impl and
// comm.py::NvlsComm), in the same order on the same streams; tests/test_sequencer_cpu.py holds the
// two launch traces against each other.  The symmetric-memory barriers between the phases stay with
// the caller (torch.distributed._symmetric_memory), which is why a phase ends where a barrier sits.
#include "host_trace.h"
#include "../../include/oneprot_clip.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>

// ------------------------------------------------------------------------------------------
// launch trace
// ------------------------------------------------------------------------------------------
namespace optrace {
namespace {
std::mutex g_mu;
std::string g_buf;
std::atomic<int> g_state{0};   // 0 = off, 1 = recording, 2 = recording + dry run
}  // namespace

bool recording() { return g_state.load(std::memory_order_relaxed) != 0; }
bool dry() { return g_state.load(std::memory_order_relaxed) == 2; }

void add(const char* fmt, ...) {
  char line[1536];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(line, sizeof(line), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_mu);
  g_buf += line;
  g_buf += '\n';
}
}  // namespace optrace

namespace opint {
int fail(int code, const std::string& msg);   // clip_kernels.cu: sets the thread-local error string
}

namespace {

inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct Seq {
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
enum { EV_FORK = 0, EV_G = 1, EV_DB = 2, EV_RS = 3 };

#define SQ_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return opint::fail(ONEPROT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define SQ(expr)                    \
  do {                              \
    int _rc = (expr);               \
    if (_rc != ONEPROT_OK) return _rc; \
  } while (0)

int sq_memset(void* p, size_t bytes, void* st) {
  if (optrace::recording()) optrace::add("memset p=%p bytes=%zu st=%p", p, bytes, st);
  if (optrace::dry()) return ONEPROT_OK;
  SQ_CUDA(cudaMemsetAsync(p, 0, bytes, static_cast<cudaStream_t>(st)));
  return ONEPROT_OK;
}

int sq_copy(void* dst, const void* src, size_t bytes, void* st) {
  if (optrace::recording()) optrace::add("copy dst=%p src=%p bytes=%zu st=%p", dst, src, bytes, st);
  if (optrace::dry()) return ONEPROT_OK;
  SQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(st)));
  return ONEPROT_OK;
}

int sq_record(Seq* s, int e, void* st) {
  if (optrace::recording()) optrace::add("record ev=%d st=%p", e, st);
  if (optrace::dry()) return ONEPROT_OK;
  if (!s->ev[e]) SQ_CUDA(cudaEventCreateWithFlags(&s->ev[e], cudaEventDisableTiming));
  SQ_CUDA(cudaEventRecord(s->ev[e], static_cast<cudaStream_t>(st)));
  return ONEPROT_OK;
}

int sq_wait(Seq* s, int e, void* st) {
  if (optrace::recording()) optrace::add("wait ev=%d st=%p", e, st);
  if (optrace::dry()) return ONEPROT_OK;
  if (!s->ev[e]) return opint::fail(ONEPROT_ERR_ARG, "sequencer: wait on an event that was never recorded");
  SQ_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(st), s->ev[e], 0));
  return ONEPROT_OK;
}

// ---- forward workspace: [finalize scratch 512 B | complete sums (3N floats, padded) | forward scratch]
struct FwdWs {
  void* fin;
  float* sums_full;
  int cnt;            // 3N rounded up to a multiple of 4 (the multimem reduction moves float4)
  void* fscratch;
  size_t fscratch_bytes;
  size_t total;
};

FwdWs fwd_ws(void* base, int n, int N) {
  FwdWs w;
  uint8_t* p = static_cast<uint8_t*>(base);
  w.cnt = (3 * N + 3) / 4 * 4;
  w.fin = p;
  w.sums_full = reinterpret_cast<float*>(p + ONEPROT_FINALIZE_SCRATCH_BYTES);
  const size_t sums_bytes = al256(static_cast<size_t>(w.cnt) * 4);
  w.fscratch = p + ONEPROT_FINALIZE_SCRATCH_BYTES + sums_bytes;
  w.fscratch_bytes = oneprot_clip_fwd_scratch_bytes(n, N);
  w.total = ONEPROT_FINALIZE_SCRATCH_BYTES + sums_bytes + al256(w.fscratch_bytes);
  return w;
}

int check_fwd(const oneprot_fwd_seq_t* f, const char* who) {
  if (!f || !f->A || !f->B_all || !f->stats_rows || !f->scale || !f->stats || !f->saved || !f->ws)
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": null pointer");
  if (f->n <= 0 || f->N < f->n || f->d <= 0 || f->d % 8 || f->row_offset < 0 || f->row_offset + f->n > f->N)
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": bad sizes");
  if ((reinterpret_cast<uintptr_t>(f->ws) & 15) || (reinterpret_cast<uintptr_t>(f->saved) & 15))
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": ws and saved must be 16-byte aligned");
  if (f->ws_bytes < oneprot_seq_fwd_ws_bytes(f->n, f->N)) return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": workspace too small");
  if (f->E && f->lde != cdiv(f->N, 64) * 64) return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": lde must be ceil(N / 64) * 64");
  if ((f->sums == nullptr) != (f->sums_mc == nullptr) || (f->sums_mc != nullptr) != (f->ag != nullptr) ||
      (f->zero_ptr != nullptr) != (f->ag != nullptr))
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": sums, sums_mc, zero_ptr and ag go together (all NULL when the sums are complete locally)");
  // the maxima every later kernel reads live in the saved header
  const float* stats_final = f->saved + ONEPROT_SAVED_STATS_AT;
  if (f->ag ? (f->ag->stats_out != stats_final) : (f->stats != stats_final))
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": the final maxima must be written to saved + ONEPROT_SAVED_STATS_AT");
  return ONEPROT_OK;
}

// ---- backward workspace: [g placeholder | g gathered | wr dg sA wc sB | fp32 dB accumulator | Wz panel]
struct BwdPlan {
  int ldw, rows_cap, n_panels, wz_rows, W4;
  float *g_zero, *g_all, *wr, *dg, *sA, *wc, *sB, *acc;
  void* Wz;
  size_t total;
};

int panel_row_unit(int d) {
  const int sms = oneprot_num_sms();
  const int ncb = cdiv(d, 256);
  return 128 * (sms / std::gcd(sms, ncb));
}

// the panel split of clip_loss.py::_backward_impl, line for line
// kept != nullptr: the forward kept the exponentials - they ARE the panel (one panel of n rows, rescaled in place)
BwdPlan bwd_plan(void* base, int n, int N, int d, int world, int want_b, size_t panel_bytes, bool kept_panel = false,
                 void* kept = nullptr) {
  BwdPlan b{};
  b.ldw = cdiv(N, 64) * 64;
  long long cap = static_cast<long long>(panel_bytes / (2 * static_cast<size_t>(b.ldw))) / 128 * 128;
  if (cap < 128) cap = 128;
  if (kept_panel) cap = (static_cast<long long>(n) + 127) / 128 * 128;
  if (cap < n) {
    // several panels: balance them and make the dA GEMM of every full panel a whole number of waves
    const long long unit = panel_row_unit(d);
    const long long n_panels = (n + cap - 1) / cap;
    const long long target = (n + n_panels - 1) / n_panels;
    if (unit <= cap) {
      const long long up = (target + unit - 1) / unit * unit;
      cap = (up <= cap) ? up : cap / unit * unit;
    } else {
      cap = std::min(cap, (target + 127) / 128 * 128);
    }
  }
  // loop stride = allocated panel height: a cap beyond n means one panel of ceil128(n) rows
  b.rows_cap = static_cast<int>(std::min<long long>(cap, (static_cast<long long>(n) + 127) / 128 * 128));
  b.wz_rows = b.rows_cap;
  b.n_panels = cdiv(n, b.rows_cap);
  b.W4 = (world + 3) / 4 * 4;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  b.g_zero = reinterpret_cast<float*>(p + off); off += 256;
  b.g_all = reinterpret_cast<float*>(p + off); off += 256;
  float* vec = reinterpret_cast<float*>(p + off);
  off += al256((3 * static_cast<size_t>(n) + 2 * static_cast<size_t>(N)) * 4);
  b.wr = vec; b.dg = vec + n; b.sA = vec + 2 * static_cast<size_t>(n);
  b.wc = vec + 3 * static_cast<size_t>(n); b.sB = b.wc + N;
  b.acc = nullptr;
  if (want_b && b.n_panels > 1) {       // fp32 accumulator of the dB terms of all but the last panel
    b.acc = reinterpret_cast<float*>(p + off);
    off += al256(static_cast<size_t>(N) * d * 4);
  }
  if (kept_panel) {
    b.Wz = kept;
  } else {
    b.Wz = p + off;
    off += al256(static_cast<size_t>(b.wz_rows) * b.ldw * 2);
  }
  b.total = off;
  return b;
}
BwdPlan bwd_plan(const oneprot_bwd_seq_t* q) {
  return bwd_plan(q->ws, q->n, q->N, q->d, q->world, q->want_b, q->panel_bytes, q->E != nullptr, q->E);
}

int check_bwd(const oneprot_bwd_seq_t* q, const char* who) {
  if (!q || !q->A || !q->B_all || !q->scale || !q->stats || !q->inv_rowsum || !q->inv_colsum || !q->g || !q->ws)
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": null pointer");
  if (q->n <= 0 || q->world <= 0 || q->N != q->world * q->n || q->d <= 0 || q->d % 8 || q->rank < 0 || q->rank >= q->world ||
      q->row_offset != q->rank * q->n)
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": bad sizes");
  if ((q->want_a && !q->dA) || (q->want_b && !q->dB)) return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": missing gradient buffer");
  if (reinterpret_cast<uintptr_t>(q->ws) & 15) return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": ws must be 16-byte aligned");
  if (q->ws_bytes < oneprot_seq_bwd_ws_bytes_ex(q->n, q->N, q->d, q->world, q->want_b, q->panel_bytes, q->E != nullptr))
    return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": workspace too small");
  if (q->E && q->lde != cdiv(q->N, 64) * 64) return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": lde must be ceil(N / 64) * 64");
  if (q->world > 1) {
    if (!q->side_stream || !q->g_slot || !q->g_slot_mc || !q->dB_mc_mine || !q->dB_out || !q->seq || !q->want_b)
      return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": world > 1 needs the exchange fields (side_stream, g_slot, g_slot_mc, dB_mc_mine, dB_out, seq) and want_b");
    if (q->world > 8) return opint::fail(ONEPROT_ERR_ARG, std::string(who) + ": at most 8 ranks (one NVSwitch node)");
  }
  return ONEPROT_OK;
}

}  // namespace

extern "C" {

void oneprot_trace_begin(int dry_run) {
  std::lock_guard<std::mutex> lk(optrace::g_mu);
  optrace::g_buf.clear();
  optrace::g_state.store(dry_run ? 2 : 1);
}

size_t oneprot_trace_end(char* out, size_t cap) {
  std::lock_guard<std::mutex> lk(optrace::g_mu);
  optrace::g_state.store(0);
  const size_t len = optrace::g_buf.size();
  if (out && cap) {
    const size_t k = len < cap - 1 ? len : cap - 1;
    memcpy(out, optrace::g_buf.data(), k);
    out[k] = 0;
  }
  return len;
}

void oneprot_trace_note(const char* text) {
  if (optrace::recording() && text) optrace::add("%s", text);
}

int oneprot_seq_create(void** out) {
  if (!out) return opint::fail(ONEPROT_ERR_ARG, "seq_create: null pointer");
  *out = new Seq();
  return ONEPROT_OK;
}

void oneprot_seq_destroy(void* seq) {
  Seq* s = static_cast<Seq*>(seq);
  if (!s) return;
  for (auto& e : s->ev)
    if (e) cudaEventDestroy(e);
  delete s;
}

// ---- forward -------------------------------------------------------------------------------
size_t oneprot_seq_fwd_ws_bytes(int n, int N) {
  if (n <= 0 || N <= 0) return 0;
  return fwd_ws(nullptr, n, N).total;
}

int oneprot_seq_fwd_begin(const oneprot_fwd_seq_t* f) {
  SQ(check_fwd(f, "seq_fwd_begin"));
  const FwdWs w = fwd_ws(f->ws, f->n, f->N);
  const int N = f->N;
  float* sums = f->sums ? f->sums : w.sums_full;
  SQ(sq_memset(f->saved, ONEPROT_SAVED_HEADER_FLOATS * 4, f->stream));         // loss, maxima, hazard flag
  SQ(sq_memset(w.fin, ONEPROT_FINALIZE_SCRATCH_BYTES, f->stream));
  if (f->zero_ptr) SQ(sq_memset(f->zero_ptr, f->zero_bytes, f->stream));        // symmetric [g | maxima | sums] buffer
  else SQ(sq_memset(sums, static_cast<size_t>(3) * N * 4, f->stream));
  SQ(oneprot_clip_rowstats(f->A, f->stats_rows, f->n, f->stats_rows_n, f->d, f->stats_off, sums + 2 * static_cast<size_t>(N) + f->row_offset,
                           f->stats, f->stream));
  SQ(oneprot_clip_fwd_sums_keep(f->A, f->B_all, f->n, N, f->d, f->scale, f->stats, f->ag, sums + N + f->row_offset, sums, w.fscratch,
                                w.fscratch_bytes, f->E, f->lde, f->stream));
  return ONEPROT_OK;
}

int oneprot_seq_fwd_end(const oneprot_fwd_seq_t* f) {
  SQ(check_fwd(f, "seq_fwd_end"));
  const FwdWs w = fwd_ws(f->ws, f->n, f->N);
  const int N = f->N;
  const float* full = f->sums ? f->sums : w.sums_full;
  if (f->sums_mc) {   // the switch adds the partial sums of all ranks
    SQ(oneprot_mc_allreduce_f32(f->sums_mc, w.sums_full, w.cnt, 0, f->stream));
    full = w.sums_full;
  }
  float* saved = f->saved;
  SQ(oneprot_clip_loss_finalize(full + N, full, full + 2 * static_cast<size_t>(N), N, f->n, f->row_offset, f->mode, f->scale,
                                saved + ONEPROT_SAVED_STATS_AT, saved + ONEPROT_SAVED_LOSS_AT, saved + ONEPROT_SAVED_HEADER_FLOATS,
                                saved + ONEPROT_SAVED_HEADER_FLOATS + N, reinterpret_cast<int*>(saved + ONEPROT_SAVED_FLAG_AT), w.fin,
                                f->stream));
  return ONEPROT_OK;
}

int oneprot_seq_fwd(const oneprot_fwd_seq_t* f) {
  SQ(oneprot_seq_fwd_begin(f));
  return oneprot_seq_fwd_end(f);
}

// ---- backward ------------------------------------------------------------------------------
size_t oneprot_seq_bwd_ws_bytes(int n, int N, int d, int world, int want_b, size_t panel_bytes) {
  return oneprot_seq_bwd_ws_bytes_ex(n, N, d, world, want_b, panel_bytes, 0);
}

size_t oneprot_seq_bwd_ws_bytes_ex(int n, int N, int d, int world, int want_b, size_t panel_bytes, int kept_panel) {
  if (n <= 0 || N <= 0 || d <= 0 || world <= 0) return 0;
  return bwd_plan(nullptr, n, N, d, world, want_b, panel_bytes, kept_panel != 0).total;
}

int oneprot_seq_bwd_panels(int n, int N, int d, size_t panel_bytes, int* rows_per_panel, int* wz_rows) {
  const BwdPlan b = bwd_plan(nullptr, n, N, d, 1, 0, panel_bytes);
  if (rows_per_panel) *rows_per_panel = b.rows_cap;
  if (wz_rows) *wz_rows = b.wz_rows;
  return b.n_panels;
}

int oneprot_seq_bwd_begin(const oneprot_bwd_seq_t* q) {
  SQ(check_bwd(q, "seq_bwd_begin"));
  if (q->world == 1) return ONEPROT_OK;    // nothing to exchange
  const BwdPlan b = bwd_plan(q);
  Seq* s = static_cast<Seq*>(q->seq);
  void* xs = q->g_on_side ? q->side_stream : q->stream;
  if (q->g_on_side) {
    SQ(sq_memset(b.g_zero, static_cast<size_t>(b.W4) * 4, q->stream));   // placeholder read by bwd_weights(what = 1)
    SQ(sq_record(s, EV_FORK, q->stream));
    SQ(sq_wait(s, EV_FORK, q->side_stream));
  }
  // one-hot contribution of this rank; after the caller's barrier the switch-side sum is the gather
  SQ(sq_memset(q->g_slot, static_cast<size_t>(b.W4) * 4, xs));
  SQ(sq_copy(q->g_slot + q->rank, q->g, 4, xs));
  return ONEPROT_OK;
}

static int dA_gemm(const oneprot_bwd_seq_t* q, const BwdPlan& b, int r0, int rows) {
  uint8_t* dA = static_cast<uint8_t*>(q->dA);
  return oneprot_gemm_bf16_ex(b.Wz, b.ldw, 0, q->B_all, q->d, 1, rows, q->d, q->N, nullptr, nullptr,
                              dA + static_cast<size_t>(r0) * q->d * 2, q->d, b.sA + r0, nullptr, 0, nullptr, q->stream);
}

int oneprot_seq_bwd_main(const oneprot_bwd_seq_t* q) {
  SQ(check_bwd(q, "seq_bwd_main"));
  const BwdPlan b = bwd_plan(q);
  Seq* s = static_cast<Seq*>(q->seq);
  const int n = q->n, N = q->N, d = q->d;
  void* mainst = q->stream;
  bool wait_g = false;
  if (q->world == 1) {
    SQ(oneprot_clip_bwd_weights(q->inv_rowsum, q->inv_colsum, N, n, q->row_offset, q->mode, q->use_gsum, 0, 1, 0, q->g, q->scale, b.wr,
                                b.wc, b.dg, b.sA, b.sB, 0, mainst));
  } else if (q->g_on_side) {
    SQ(oneprot_mc_allreduce_f32(q->g_slot_mc, b.g_all, b.W4, 0, q->side_stream));
    SQ(oneprot_clip_bwd_weights(q->inv_rowsum, q->inv_colsum, N, n, q->row_offset, q->mode, q->use_gsum, 0, q->world, q->rank, b.g_zero,
                                q->scale, b.wr, b.wc, b.dg, b.sA, b.sB, 1, mainst));          // panel weights need no g
    SQ(oneprot_clip_bwd_weights(q->inv_rowsum, q->inv_colsum, N, n, q->row_offset, q->mode, q->use_gsum, 0, q->world, q->rank, b.g_all,
                                q->scale, b.wr, b.wc, b.dg, b.sA, b.sB, 2, q->side_stream));  // output scales from the gathered g
    SQ(sq_record(s, EV_G, q->side_stream));
    wait_g = true;
  } else {
    SQ(oneprot_mc_allreduce_f32(q->g_slot_mc, b.g_all, b.W4, 0, mainst));
    SQ(oneprot_clip_bwd_weights(q->inv_rowsum, q->inv_colsum, N, n, q->row_offset, q->mode, q->use_gsum, 0, q->world, q->rank, b.g_all,
                                q->scale, b.wr, b.wc, b.dg, b.sA, b.sB, 0, mainst));
  }
  const uint8_t* A = static_cast<const uint8_t*>(q->A);
  for (int pi = 0, r0 = 0; r0 < n; ++pi, r0 += b.rows_cap) {
    const int rows = std::min(b.rows_cap, n - r0);
    const bool first = pi == 0, last = r0 + b.rows_cap >= n;
    const void* A_rows = A + static_cast<size_t>(r0) * d * 2;
    if (q->E)    // kept exponentials: one in-place rescale of the whole panel instead of the recompute
      SQ(oneprot_clip_dz_from_exp(q->E, n, N, q->lde, q->row_offset, b.wr, b.wc, b.dg, mainst));
    else
      SQ(oneprot_clip_dz_panel(A_rows, q->B_all, rows, N, d, q->row_offset + r0, q->scale, q->stats, b.wr + r0, b.wc, b.dg + r0, b.Wz, b.ldw,
                               mainst));
    if (wait_g) {             // the GEMM epilogues read the output scales
      SQ(sq_wait(s, EV_G, mainst));
      wait_g = false;
    }
    if (q->want_b) {          // dB first: its exchange then hides under the dA GEMM
      const float* acc_in = first ? nullptr : b.acc;
      if (last)
        SQ(oneprot_gemm_bf16_ex(b.Wz, b.ldw, 1, A_rows, d, 1, N, d, rows, acc_in, nullptr, q->dB, d, b.sB, nullptr, 0, nullptr, mainst));
      else
        SQ(oneprot_gemm_bf16_ex(b.Wz, b.ldw, 1, A_rows, d, 1, N, d, rows, acc_in, b.acc, nullptr, d, nullptr, nullptr, 0, nullptr, mainst));
      if (last && q->world > 1) {
        SQ(sq_record(s, EV_DB, mainst));
        SQ(sq_wait(s, EV_DB, q->side_stream));
      }
    }
    // world > 1: the dA GEMM of the last panel is enqueued by oneprot_seq_bwd_end, AFTER the side
    // stream's barrier + pull-reduce, so that those small kernels are resident before the GEMM's
    // persistent CTAs fill the SMs (same submission order as the Python host)
    if (q->want_a && !(last && q->world > 1)) SQ(dA_gemm(q, b, r0, rows));
  }
  return ONEPROT_OK;
}

int oneprot_seq_bwd_end(const oneprot_bwd_seq_t* q) {
  SQ(check_bwd(q, "seq_bwd_end"));
  if (q->world == 1) return ONEPROT_OK;
  Seq* s = static_cast<Seq*>(q->seq);
  // every rank's partial dB is in place (caller's barrier on the side stream): pull-reduce this rank's rows
  SQ(oneprot_mc_reduce_bf16(q->dB_mc_mine, q->dB_out, static_cast<size_t>(q->n) * q->d * 2, q->side_stream));
  SQ(sq_record(s, EV_RS, q->side_stream));
  if (q->want_a) {     // last panel's dA GEMM: runs over the exchange above
    const BwdPlan b = bwd_plan(q);
    const int r0 = (b.n_panels - 1) * b.rows_cap;
    SQ(dA_gemm(q, b, r0, q->n - r0));
  }
  SQ(sq_wait(s, EV_RS, q->stream));
  return ONEPROT_OK;
}

}  // extern "C"
