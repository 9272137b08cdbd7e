// Host-side launch trace shared by the C ABI entry points (clip_kernels.cu) and the step sequencer
// (clip_sequence.cu).  While a trace is open every entry point appends one text line with its
// arguments; in DRY mode it then returns without touching CUDA (no tensor maps, no launches), which
// lets the CPU tests compare the launch sequence of the C sequencer with the one the Python host
// issues - same kernels, same arguments, same streams - on a machine without a GPU.
#pragma once
#include <cstddef>

namespace optrace {
bool recording();                 // a trace is open: entry points append their line
bool dry();                       // ... and skip all CUDA work
void add(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
}  // namespace optrace
