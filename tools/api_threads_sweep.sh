#!/bin/bash
# tools/api_threads_sweep.sh -- the nine-symbol API's end-to-end time (tools/api_bench.c) against the number of
# threads of the staging copy (HYDRIUM_B200_THREADS; 1 = the calling thread alone).  Run on the GPU box.
B=hydrium_b200/bin/api_bench
for t in 1 2 3 4 6 8; do
  echo "threads=$t tile-mode:  $(HYDRIUM_B200_THREADS=$t $B --reps 20 --warmup 3)"
done
for t in 1 4 8; do
  echo "threads=$t one-frame:  $(HYDRIUM_B200_THREADS=$t $B --one-frame --reps 10 --warmup 3)"
  echo "threads=$t shift 3:    $(HYDRIUM_B200_THREADS=$t $B --shift 3 --reps 10 --warmup 3)"
done
echo "default: $($B --reps 20 --warmup 3)"
nproc
