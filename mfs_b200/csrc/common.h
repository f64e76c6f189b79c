// Helpers shared by the translation units of libmfs_b200.so (error text, launch counter).  Defined in api.cu.
#pragma once
#include <cstdint>

namespace mfs {

// Stores printf-style text as the calling thread's last error (mfs_last_error) and returns -1.
int fail(const char* fmt, ...);

// Adds n to the process-wide kernel-launch counter (mfs_launch_count).
void count_launches(int64_t n);

}  // namespace mfs
