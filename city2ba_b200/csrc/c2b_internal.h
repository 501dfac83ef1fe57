// c2b_internal.h — entry points shared between the library's translation units, not part of the ABI.
#pragma once
#include <cstddef>

#include "../../include/city2ba_cuda.h"

extern "C" {
// host array -> device on the ctx's stream; a pageable source is staged through the ctx's pinned ring by
// `stage_threads` host threads (c2b_api.cu: copy_in)
int c2b_internal_copy_in(c2b_ctx *ctx, void *d, const void *h, size_t bytes);
}
