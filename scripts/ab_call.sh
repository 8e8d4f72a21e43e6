#!/bin/bash
# A/B-only gpurun call: scripts/ab_call.sh <tag> <variant specs...>  (first = baseline); results in gpurun_out/<tag>/
T=$1; shift
O=gpurun_out/$T
mkdir -p $O
timeout 900 python scripts/ab_variants.py --out $O/ab_variants.json "$@" > $O/ab.log 2>&1
tail -2 $O/ab.log
