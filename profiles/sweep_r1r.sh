#!/bin/bash
b() { python bench.py --preset $1 --envs $2 --steps 100 --warmup 5 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g env-steps/s  %.4f ms' % (d['value'], d['ms_per_step']))"; }
for p in "exp02_vFinal 65536" "exp03_vFinal 65536" "level5_c1 65536" "swarm 8192" "stage02_10lm 4096" "exp02_vFinal 8192"; do set -- $p; echo "$1 $2 default: $(b $1 $2)"; for k in 1 2 4; do echo "$1 $2 DC_SUB_BATCHES=$k: $(DC_SUB_BATCHES=$k b $1 $2)"; done; done
