// The pair kernel of the alignment pass without its statistic counters, short-read (10 CTAs / SM, NW scratch descriptor by value) build: the same source as
// kmagpu_align.cu compiled again inside a namespace (nothing but the kernel and its C-linkage launcher).
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include <string.h>
#include <algorithm>
#define KG_NO_STATS
#define KG_PAIR_VARIANT_ONLY
#define NW_SCRATCH_BYVAL
#define KG_VARIANT_LAUNCHER kg_launch_pair_fast_short
#define KG_VARIANT_MINB AL_MINB_SHORT
namespace kg_fast_short {
#include "kmagpu_nw.cuh"
#include "kmagpu_align.cu"
}
