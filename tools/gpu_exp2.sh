#!/bin/bash
OUT=gpurun_out/exp2; mkdir -p $OUT
G=tools/bin/gather_bench
{
for fl in 0 1 2 3 4 5 6 7 8 9; do $G 640 0 $fl 1280 1500; done
for fl in 0 1 4 7; do $G 256 0 $fl 1280 1500; $G 160 0 $fl 1280 1500; $G 64 0 $fl 1280 1500; done
for fl in 0 1 4; do for t in 128 256 512 1024 2048; do $G 64 0 $fl $t 1500; done; done
} > $OUT/gather.txt 2>&1
cat $OUT/gather.txt
for fl in 0 1 4 5 7 9; do
  ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sector_op_read_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_evict_first_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_evict_normal_lookup_miss.sum --clock-control none -k regex:chase -s 1 -c 1 --csv $G 640 0 $fl 1280 500 2>&1 | grep -E "chase|table" | awk -F'","' '{print $5, $(NF-2), $NF}' > $OUT/ncu_fl$fl.txt
  echo "== flavour $fl"; cat $OUT/ncu_fl$fl.txt
done
for fl in 0 1; do
  ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sector_op_read_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:chase -s 1 -c 1 --csv $G 200 0 $fl 1280 500 2>&1 | grep -E "chase|table" | awk -F'","' '{print $5, $(NF-2), $NF}' > $OUT/ncu200_fl$fl.txt
  echo "== 200MiB flavour $fl"; cat $OUT/ncu200_fl$fl.txt
done
