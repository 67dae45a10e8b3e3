#!/bin/bash
# dev tool: A/B the default library against builds under nimrud_b200/lib/variants/ (feature phase of configs[1])
for v in "" nimrud_b200/lib/variants/*.so; do
  NIMRUD_B200_LIB=$v python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-config4 --no-extras 2>&1 | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['phase_ms_per_step']; print('%-44s %.3f G  %.3f ms/step  features %.3f index %.3f order %.3f' % ('${v:-default}', d['value']/1e9, d['ms_per_step'], p['features'], p['index'], p['order']))"
  if [ -n "$AB_PER_SCALE" ]; then NIMRUD_B200_LIB=$v python scripts/per_scale_timing.py 2>&1 | awk '{printf "%s ", $4} END{print ""}'; fi
done
