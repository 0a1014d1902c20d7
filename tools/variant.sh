#!/bin/bash
# build libkmagpu with extra -D flags on the GPU box and time the PE pipeline: tools/variant.sh "<flags>" [pairs] [steps]
cd "$(dirname "$0")/../kma_b200/csrc" && touch kmagpu_align.cu kmagpu_seed.cu && make EXTRA="$1" > /dev/null 2>&1 && cd ../.. && echo "== $1" && python tools/pe_perf.py ${2:-2000000} ${3:-3} 2>&1 | tail -1 | cut -c1-260
