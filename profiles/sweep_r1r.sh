#!/bin/bash
b() { python bench.py "$@" --steps 200 --warmup 5 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g env-steps/s  %.4f ms  launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']))"; }
echo "default: $(b)"
for k in 1 2 3 4 8; do echo "graph K=$k: $(b --graph --sub-batches $k)"; done
echo "8192 default: $(b --envs 8192)"; echo "8192 graph K=1: $(b --envs 8192 --graph --sub-batches 1)"; echo "8192 graph K=2: $(b --envs 8192 --graph --sub-batches 2)"
echo "level5 graph K=4: $(b --preset level5_c1 --graph --sub-batches 4)"; echo "exp03 graph K=4: $(b --preset exp03_vFinal --graph --sub-batches 4)"
