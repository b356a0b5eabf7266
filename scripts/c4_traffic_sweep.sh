#!/bin/bash
# DRAM traffic vs throughput of the C4 global-work-matrix kernel (profiles/r2_c4_traffic.json); run under gpurun.
out=gpurun_out/r2_c4_traffic
mkdir -p $out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
run() {  # name, env assignments...
  name=$1; shift
  env "$@" python scripts/c4_traffic.py 40 > $out/$name.plain.json 2> $out/$name.plain.err || return
  env "$@" ncu --metrics $M --clock-control none -k regex:lsi_kernel -c 1 --csv --log-file $out/$name.ncu.csv python scripts/c4_traffic.py 40 > $out/$name.ncu.log 2>&1
}
run w8_dmma     LEXLS_B200_DMMA=1
run w8_dfma     LEXLS_B200_DMMA=0
run w8_persist  LEXLS_B200_DMMA=1 LEXLS_B200_LG_PERSIST=1
run w16         LEXLS_B200_LG_WARPS=16
run w16_persist LEXLS_B200_LG_WARPS=16 LEXLS_B200_LG_PERSIST=1
run w16_grid96  LEXLS_B200_LG_WARPS=16 LEXLS_B200_LG_GRID=96
run w16_grid74  LEXLS_B200_LG_WARPS=16 LEXLS_B200_LG_GRID=74
cat $out/*.plain.json
