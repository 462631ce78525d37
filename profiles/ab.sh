#!/bin/bash
# Same-box A/B of library builds / kernel-variant options through bench.py (value is reproducible to ~0.1 % on one box).
# usage: bash profiles/ab.sh "label1:ENV=..;ENV2=.." "label2:..." ...   e.g.  "row0fwd:DGVIT_OPTS=attention_row0=3"
for rep in $(seq 1 ${REPS:-2}); do
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  ( IFS=';'; for e in $envs; do [ -n "$e" ] && export "$e"; done
    timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', round(d['value']), round(d['ms_per_step'],4), d['act_latency']['evaluate']['p50_ms'])" )
done; done
