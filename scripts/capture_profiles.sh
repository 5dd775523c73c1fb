#!/bin/bash
# Evidence capture for profiles/: bench (not under a profiler), ncu launch list of the same command, and one
# `ncu --set full` capture per hot kernel.  Usage (GPU box): bash scripts/capture_profiles.sh r02p
tag=${1:-cap}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
tail -c 600 $out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $out/${tag}_bench_under_ncu.log 2>&1
for spec in "k1:loglik_batch_kernel:3" "pt:pt_kernel:1" "k4:multi_loglik_kernel:2" "scan:scan_filter_kernel:1" "mle:lbfgs_kernel:1"; do
    IFS=: read name pat skip <<< "$spec"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -f -o $out/${tag}_${name} \
        python bench.py --steps 2 --warmup 3 --no-cpu --pt-iters 20 > $out/${tag}_${name}_ncu.log 2>&1
    ncu -i $out/${tag}_${name}.ncu-rep --page raw --csv > $out/${tag}_${name}_ncu_raw.csv 2>/dev/null
    ncu -i $out/${tag}_${name}.ncu-rep --page details --csv > $out/${tag}_${name}_ncu_details.csv 2>/dev/null
    ncu -i $out/${tag}_${name}.ncu-rep --page source --csv 2>/dev/null | gzip > $out/${tag}_${name}_ncu_source.csv.gz
    rm -f $out/${tag}_${name}.ncu-rep      # the reports are 20-40 MB each; gpurun_out/ is capped at 64 MiB
done
ls -la $out | tail -20
