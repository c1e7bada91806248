#!/bin/bash
# experiment: how many sectors does one 16-byte gather cost at each level (L1->L2 request, L2->DRAM fill)?
OUT=gpurun_out/exp6; mkdir -p $OUT; rm -f $OUT/bench.txt
G=tools/bin/gather_bench
M="--metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct"
run() {
  echo "== $*" | tee -a $OUT/bench.txt
  $G "$@" | tee -a $OUT/bench.txt
  ncu $M --clock-control none -k regex:chase -s 1 -c 1 --csv $G "$@" 2>/dev/null | grep -E 'dram__|lts__|l1tex__' | awk -F'","' '{gsub(/"/,"",$NF); printf "   %s = %s\n", $(NF-2), $NF}' | tee -a $OUT/bench.txt
}
# <table MiB> <gran> <flavour> <threads/SM> <steps>   (steps*threads = 189440*500 = 94.7M gathers per launch)
for mib in 16 151 2048; do for fl in 0 1 4; do run $mib 0 $fl 1280 500; done; done
run 2048 32 0 1280 500
run 2048 32 1 1280 500
run 2048 128 1 1280 500
run 151 32 1 1280 500
for fl in 2 5 6 7 9; do run 2048 0 $fl 1280 500; done
