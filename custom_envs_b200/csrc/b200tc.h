// Host-side entry points of the tcgen05 eval kernel (b200tc.cu), called from b200env.cu.
// Dev / StepArgs travel as opaque pointers: both translation units include b200env_shared.cuh.
#pragma once
#include <stddef.h>

bool b2e_tc2_supported(const void *dev);
size_t b2e_tc2_smem_bytes();
const char *b2e_tc2_prepare();                       // nullptr on success, else the reason
// Launches one eval over args.e_begin .. e_begin + e_count on `grid` persistent CTAs.
// dbg: 4 ints of device memory, zero unless the kernel's watchdog fired (protocol error).
int b2e_tc2_launch(const void *dev, const void *args, int second, int grid, int *dbg, void *stream);
