#!/bin/bash
mkdir -p gpurun_out
./microbench/random_access 160000000 64000000 0 | tee gpurun_out/micro_g0.txt
./microbench/random_access 160000000 64000000 32 | tee gpurun_out/micro_g32.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,lts__t_sectors_lookup_hit.sum,lts__t_sectors_lookup_miss.sum --clock-control none --csv --log-file gpurun_out/micro_ncu_g0.csv ./microbench/random_access 160000000 64000000 0 > /dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/micro_ncu_g32.csv ./microbench/random_access 160000000 64000000 32 > /dev/null
