#!/bin/bash
# one full-section capture of the PCG kernels per library build (development tool):  tools/ncu_full.sh <tag>=<lib> ...
for kv in "$@"; do
  tag=${kv%%=*}; lib=${kv#*=}
  MOFEM_B200_LIB=$lib timeout 600 ncu --set full --import-source on --clock-control none \
    -k regex:"k_spmv_ring|k_pcg_vec|k_pcg_fused" -s 20 -c 2 -f -o gpurun_out/r2_full_$tag \
    python tools/profile_run.py 664 60 reps=1 > gpurun_out/r2_full_$tag.log 2>&1
  echo "== $tag ncu rc=$?"
done
