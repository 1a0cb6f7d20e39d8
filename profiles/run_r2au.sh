#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/r2_host_dma_bound.py > gpurun_out/r2au_host_dma_bound.txt 2> gpurun_out/r2au.err
cat gpurun_out/r2au_host_dma_bound.txt; tail -3 gpurun_out/r2au.err; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" ; nvidia-smi topo -m 2>/dev/null | head -12
