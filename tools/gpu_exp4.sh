#!/bin/bash
# experiment: which stream costs the L2 its hit rate (gather_bench2), timing + ncu DRAM bytes
OUT=gpurun_out/exp4; mkdir -p $OUT; rm -f $OUT/bench.txt
G=tools/bin/gather_bench2
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct"
run() {  # args of gather_bench2
  $G "$@" | tee -a $OUT/bench.txt
  if [ "${NCU:-1}" = "1" ]; then
    ncu $M --clock-control none -k regex:walk -s 1 -c 1 --csv $G "$@" 2>/dev/null | grep -E 'dram__|lts__' | awk -F'","' '{gsub(/"/,"",$NF); printf "   %s = %s %s\n", $(NF-2), $NF, $(NF-1)}' | tee -a $OUT/bench.txt
  fi
}
#     MiB tpsm steps nbr rd wr ob row
run 151 1280 1024 1 0 0 1 0
run 151 1280 1024 1 1 0 1 0
run 151 1280 1024 1 3 0 1 0
run 151 1280 1024 1 32 0 1 0
run 151 1280 1024 1 64 0 1 0
run 151 1280 1024 1 128 0 1 0
run 151 1280 1024 1 0 1 1 0
run 151 1280 1024 1 0 2 1 0
run 151 1280 1024 1 0 32 1 0
run 151 1280 1024 1 0 64 1 0
run 151 1280 1024 1 0 128 1 0
run 151 1280 1024 1 0 2 4 0
run 151 1280 1024 1 0 32 4 0
run 151 1280 1024 1 0 64 4 0
run 151 1280 1024 1 0 128 4 0
NCU=0
run 151 1280 1024 1 64 64 1 0
run 151 1280 1024 1 64 64 4 0
run 151 1280 1024 1 128 32 1 0
run 151 1024 1024 1 128 64 1 0
run 151 1024 1024 1 128 64 4 0
run 151 768 1024 1 128 128 1 0
run 151 768 1024 1 128 128 4 0
run 151 1280 1024 1 64 64 1 1
run 151 1280 1024 1 64 64 1 2
run 76 1280 1024 1 64 64 1 0
run 76 1280 1024 1 1 1 1 0
run 76 1280 1024 1 0 0 1 0
run 100 1280 1024 1 64 64 1 0
run 120 1280 1024 1 64 64 1 0
run 300 1280 1024 1 64 64 1 0
run 600 1280 1024 1 64 64 1 0
