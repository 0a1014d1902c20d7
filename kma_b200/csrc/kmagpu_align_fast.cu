// The pair kernel of the alignment pass without its statistic counters: the same source as kmagpu_align.cu compiled a
// second time inside a namespace (nothing but the kernel and its C-linkage launcher). See kg_launch_pair_nostats there.
#include "kmagpu_internal.h"
#include "kmagpu_dev.cuh"
#include <string.h>
#include <algorithm>
#define KG_NO_STATS
#define KG_PAIR_VARIANT_ONLY
namespace kg_fast {
#include "kmagpu_nw.cuh"
#include "kmagpu_align.cu"
}
