// Internal (C++ only) view of an env handle for the other translation units of libb200env.so.
#pragma once
#include "b200env.h"

// The handle's device view (a `Dev` of b200env_shared.cuh, as an opaque pointer) for kernels outside
// b200env.cu that read env state in place (the policy kernel's ring front end, b200policy.cu).
// *ring_ok = 1 when the env keeps MultiOptLRs adjusted-history rings that such a kernel can read
// (large-problem pipeline); *device = the env's CUDA ordinal.  The view follows the handle's gradient
// ping-pong, so fetch it per call, not once.
const void *b2e_dev_view(b2e_handle h, int *ring_ok, int *device);
