#!/bin/bash
# A/B of library builds (development tool): in-sequence ncu (no cache flush, no replay) of the PCG kernels + plain timing.
#   tools/ncu_ab.sh <tag>=<lib path or ""> ...
for kv in "$@"; do
  tag=${kv%%=*}; lib=${kv#*=}
  MOFEM_B200_LIB=$lib timeout 300 ncu --cache-control none --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    -k regex:"k_spmv_ring|k_pcg_vec|k_pcg_fused" -s 60 -c 8 --csv --log-file gpurun_out/r2_ncu_ab_$tag.csv \
    python tools/profile_run.py 664 200 reps=1 $EXTRA > gpurun_out/r2_ncu_ab_$tag.log 2>&1
  echo "== $tag ncu rc=$?"
  MOFEM_B200_LIB=$lib timeout 200 python tools/profile_run.py 664 3000 ${CONFIGS:-configs=ring:pdl,ring} 2>&1 | grep us_per_iteration | sed 's/.*us_per_iteration/us_per_iteration/'
done
