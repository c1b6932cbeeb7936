#!/bin/bash
# A/B of the lock-step iteration latency on one box: GF_EAGER_COUNT=1 restores the eager Newton-step counter.
for i in 1 2 3; do
  for e in 0 1; do
    GF_EAGER_COUNT=$e timeout 200 python tools/run_config.py --cfg 4 --B 128 --check 0 2>&1 | tail -1 > /tmp/ab.json
    python -c "import json; d=json.load(open('/tmp/ab.json')); print('eager_count=$e', round(d['wall_s'],4), round(d['ms_per_outer'],4))"
  done
done
