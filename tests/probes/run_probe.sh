#!/bin/bash
# runs the MN-major probe over a grid of configurations, one process each (a bad config may trap)
P=tests/probes/probe_mnmajor
out=gpurun_out/probe_mnmajor.txt
mkdir -p gpurun_out; : > $out
timeout 20 $P 3 2 16 1024 0 >> $out 2>&1
for which in 1 2 3; do
 for tswz in 3 4; do
  for type in 2 1; do
   for ls in "4096 1024" "1024 4096" "4096 512" "512 4096" "4096 256" "128 4096" "4096 128"; do
    for kstep in 1024; do
      timeout 20 $P $tswz $type $ls $which $kstep >> $out 2>&1 || echo "rc=$? for $tswz $type $ls $which $kstep" >> $out
    done
   done
  done
 done
done
grep -c MATCH $out
grep MATCH $out
