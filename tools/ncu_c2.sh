#!/bin/bash
# `ncu --set full` of the fixed-base commit kernels at the C2 shape (2^18 variables to keep the replay short; same
# tables and per-variable work as 2^20).  usage: tools/ncu_c2.sh <tag>
tag=$1
out=gpurun_out/${tag}
ncu --set full --clock-control none --import-source on -k regex:"k_fixed_commit" -s 4 -c 2 -f -o ${out} \
    python tools/bench_configs.py c2 --log2n 18 --no-cpu > ${out}.log 2>&1
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null
ncu -i ${out}.ncu-rep --page details > ${out}_details.txt 2>/dev/null
ncu -i ${out}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > ${out}_source.csv.gz
sz=$(stat -c %s ${out}.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f ${out}.ncu-rep; fi
ls -la gpurun_out/ | grep ${tag}
