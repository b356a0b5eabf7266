#!/bin/bash
# Round profiles (run under gpurun): ncu --set full captures of the three kernel families, summarised ON THE BOX (the reports
# exceed what gpurun brings back), text summaries under gpurun_out/profiles_<tag>/ ; two reports are kept for source views.
tag=${1:-r2}
out=gpurun_out/profiles_$tag
mkdir -p $out
R=$(python scripts/ncu_ranges.py)
summ() {  # report kernel-substring nprob outfile title
  { echo "# $5"; python scripts/ncu_summary.py $1; echo; echo "# instructions executed / pc samples by source function (scripts/ncu_by_range.py; per problem = / $3)";
    NPROB=$3 python scripts/ncu_by_range.py $1 $2 $R; echo; echo "# executed code footprint by source function (scripts/ncu_footprint.py)";
    python scripts/ncu_footprint.py $1 $2 $3; } > $4 2>&1
}
python scripts/quick_perf.py c2 c3 > $out/quick_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:lse_kernel -s 2 -c 1 -o $out/lse_c2 python scripts/quick_perf.py c2 > $out/ncu_c2.log 2>&1
summ $out/lse_c2.ncu-rep lse_kernelILi1ELb0 65536 $out/${tag}_lse_c2_w1_full.txt "C2 LexLSE 65,536 problems: lse_kernel<1,0> (ncu --set full --clock-control none -k regex:lse_kernel -s 2 -c 1 python scripts/quick_perf.py c2)"
ncu --set full --clock-control none --import-source on -k regex:lsi_kernel -c 1 -o $out/lsi_c3_cold python scripts/quick_perf.py c3 > $out/ncu_c3c.log 2>&1
summ $out/lsi_c3_cold.ncu-rep lsi_kernelILi2ELb0ELb0ELb0 16384 $out/${tag}_lsi_c3_w2_cold_full.txt "C3 LexLSI 16,384 problems, COLD start: lsi_kernel<2,0,0> (first launch of scripts/quick_perf.py c3)"
rm -f $out/lsi_c3_cold.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:lsi_kernel -s 5 -c 1 -o $out/lsi_c3_warm python scripts/quick_perf.py c3 > $out/ncu_c3w.log 2>&1
summ $out/lsi_c3_warm.ncu-rep lsi_kernelILi2ELb0ELb0ELb0 16384 $out/${tag}_lsi_c3_w2_warm_full.txt "C3 LexLSI 16,384 problems, WARM start (the C3 configuration proper): lsi_kernel<2,0,0> (sixth launch of scripts/quick_perf.py c3)"
for v in 1 0; do
  LEXLS_B200_DMMA=$v python scripts/c4_traffic.py 40 > $out/c4_dmma$v.plain.json 2>&1
  LEXLS_B200_DMMA=$v ncu --set full --clock-control none --import-source on -k regex:lsi_kernel -c 1 -o $out/lsi_c4_dmma$v python scripts/c4_traffic.py 40 > $out/ncu_c4_$v.log 2>&1
  summ $out/lsi_c4_dmma$v.ncu-rep lsi_kernelILi8ELb0ELb1ELb0 1024 $out/${tag}_lsi_c4_w8_lodglobal_dmma${v}_full.txt "C4 LexLSI n=200, 1,024 problems, cold start, 40 factorizations each: lsi_kernel<8,0,1> with LEXLS_B200_DMMA=$v (Schur update: 1 = mma.sync m8n8k4 f64 tiles, 0 = DFMA strips)"
  rm -f $out/lsi_c4_dmma$v.ncu-rep
done
python bench.py --steps 5 --warmup 3 --no-extra > $out/bench_short.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_c2.csv python bench.py --steps 5 --warmup 3 --no-extra > $out/ncu_launch.log 2>&1
rm -f $out/lsi_c3_warm.ncu-rep
ls -la $out
