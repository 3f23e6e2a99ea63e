#!/bin/bash
# usage: tools/bench_short.sh [extra bench args]  -- prints value / ms_per_step / loss of one short bench run
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c '
import sys, json
j = json.loads(sys.stdin.read())
print({k: round(j[k], 3) for k in ("value", "ms_per_step")}, "launches", j["gpu_launches"], "loss", j["config"]["final_loss"])'
