#!/bin/bash
# Like phase_timing.sh, but only the three per-level stamps (-DSZB_LEVEL_ONLY): the fine-grained stamps cost
# registers and perturb what they measure.  Run on the GPU box:  bash tools/level_timing.sh
set -e
SZB_EXTRA="$SZB_EXTRA -DSZB_LEVEL_ONLY" bash "$(dirname "$0")/phase_timing.sh" "$@" | head -6
