#!/bin/bash
# dev tool: A/B the default library against builds under nimrud_b200/lib/variants/
for v in "" nimrud_b200/lib/variants/*.so; do
  NIMRUD_B200_LIB=$v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,3), round(d['ms_per_step'],3), d['roofline']['phase_ms_per_step'])"
  NIMRUD_B200_LIB=$v python scripts/per_scale_timing.py 2>&1 | awk '{printf "%s ", $4} END{print ""}'
done
