#!/bin/bash
# Round-1 profiling recipe (run on the GPU box): the launch list of a short bench run and one `ncu --set full` capture of
# each dominant kernel.  Outputs land in gpurun_out/ (scratch); the summaries are copied to profiles/ by hand.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
if [ "$1" != "full-only" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1b_launches.csv $B > gpurun_out/r1b_launches_bench.log 2>&1
fi
NC="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NC -k "regex:graph_layer_fwd_sparse_kernel<.bool.0, .bool.0>" -s 4 -c 1 -o gpurun_out/r1b_sparse68 $B > /dev/null 2>&1
timeout 300 $NC -k "regex:graph_layer_fwd_sparse_kernel<.bool.0, .bool.1>" -s 4 -c 1 -o gpurun_out/r1b_sparse10 $B > /dev/null 2>&1
timeout 300 $NC -k "regex:gemm_tf32x3_persistent_kernel<.int.240" -s 13 -c 1 -o gpurun_out/r1b_gemm $B > /dev/null 2>&1
if [ "$1" != "full-only" ]; then
timeout 300 $NC -k regex:topic_segment_fwd -s 6 -c 1 -o gpurun_out/r1b_topic $B > /dev/null 2>&1
fi
ls -la gpurun_out/r1b_*
