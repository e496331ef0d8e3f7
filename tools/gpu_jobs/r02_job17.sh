#!/bin/bash
# round 2, call 17: granularity of the pipelined host-buffer apply (bench `e2e`)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python tools/e2e_host_ab.py > gpurun_out/r02_e2e_host_ab.txt 2>&1; cat gpurun_out/r02_e2e_host_ab.txt
