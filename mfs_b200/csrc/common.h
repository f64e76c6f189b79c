// Helpers shared by the translation units of libmfs_b200.so (error text, launch counter).  Defined in api.cu.
#pragma once
#include <cstdint>

namespace mfs {

// Stores printf-style text as the calling thread's last error (mfs_last_error) and returns -1.
int fail(const char* fmt, ...);

// Adds n to the process-wide kernel-launch counter (mfs_launch_count).
void count_launches(int64_t n);

// Tangent selection of the gradient kernel (filter1d_grad.cuh), filled in by mfs_filter_1d_grad.
struct GradInfo {
  int32_t tangent_ids[4];  // per tangent slot: 0..3 -> trans_params[i], 4..7 -> meas_params[i - 4], -1 -> unused
  double* grad_out;        // [B][n_slots]
  int32_t n_slots;
};

}  // namespace mfs
