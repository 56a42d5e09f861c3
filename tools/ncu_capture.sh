#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> [bench args...]
# One `ncu --set full` capture of the named kernels inside a bench.py run (C5 leg only); exports the raw / details /
# source pages as text next to the report and drops the report itself when it is too big to travel back (64 MiB cap).
# NCU_EXTRA adds ncu options (e.g. NCU_EXTRA="--replay-mode application" for kernels whose kernel-replay passes fail).
tag=$1; regex=$2; skip=$3; count=$4; shift 4
out=gpurun_out/${tag}
GS_PASS_STREAMS=1 ncu --set full --clock-control none --import-source on ${NCU_EXTRA} -k regex:"${regex}" -s ${skip} -c ${count} -f -o ${out} \
    python bench.py --no-cpu-baseline --skip c3,c4,weak "$@" > ${out}.log 2>&1
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null
ncu -i ${out}.ncu-rep --page details > ${out}_details.txt 2>/dev/null
ncu -i ${out}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > ${out}_source.csv.gz
sz=$(stat -c %s ${out}.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f ${out}.ncu-rep; fi
du -sh gpurun_out; ls -la gpurun_out/ | tail -5
