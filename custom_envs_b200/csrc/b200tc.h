// Host-side entry points of the tcgen05 eval kernel (b200tc.cu), called from b200env.cu.
// Dev / StepArgs travel as opaque pointers: both translation units include b200env_shared.cuh.
#pragma once
#include <stddef.h>

#include <string>

struct b2e_tc2_ctx;

bool b2e_tc2_supported(const void *dev);
// Per-handle context: shared-memory opt-in, the tensor map over the envs' first-layer matrices (d.w must
// be final), the watchdog word.  nullptr + *error on failure.
b2e_tc2_ctx *b2e_tc2_create(const void *dev, int num_sms, std::string *error);
void b2e_tc2_destroy(b2e_tc2_ctx *ctx);
int b2e_tc2_grid(const b2e_tc2_ctx *ctx);
// out[0] != 0: the kernel's watchdog fired (protocol error); out[1..3] say where.
int b2e_tc2_watchdog(const b2e_tc2_ctx *ctx, int out[4]);
// One eval over args.e_begin .. e_begin + e_count on the context's persistent grid.
int b2e_tc2_launch(const b2e_tc2_ctx *ctx, const void *dev, const void *args, int second, void *stream);
