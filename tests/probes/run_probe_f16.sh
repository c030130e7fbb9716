#!/bin/bash
P=tests/probes/probe_mnmajor_f16
out=gpurun_out/probe_mnmajor_f16.txt
mkdir -p gpurun_out; : > $out
timeout 20 $P 3 2 16 1024 0 >> $out 2>&1
for which in 1 2 3; do
 for cfg in "3 2 8192 1024" "3 2 1024 8192" "4 1 8192 512" "4 1 8192 1024" "3 2 8192 2048"; do
   timeout 20 $P $cfg $which 2048 >> $out 2>&1 || echo "rc=$? for $cfg $which" >> $out
 done
done
cat $out
