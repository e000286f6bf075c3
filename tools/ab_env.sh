#!/bin/bash
# A/B of run-time switches on the GPU box: tools/ab_env.sh "VAR=val VAR2=val" "..." ...   ("" = defaults)
cd "$(dirname "$0")/.."
for e in "$@"; do
  printf "%-28s " "[$e]"; env $e timeout 120 python tools/profile_frame.py --frames 4 2>&1 | tail -1
done
