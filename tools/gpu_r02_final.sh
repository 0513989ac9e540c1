#!/bin/bash
# Evidence run of round 2, last session (one gpurun call, one GPU): full GPU test suite, the default bench line, the reference arm,
# ncu launch list + one --set full capture of the per-epoch deconvolution kernel (each only after the same command exited 0 without
# ncu), and the cfg5 shapes at 1332 frames (nine 148-frame waves: two chunks of the 1184-frame workspace).
set -x
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests_r02c.log 2>&1; echo "tests rc=$?"; tail -4 $out/tests_r02c.log
timeout 900 python bench.py > $out/bench_r02c_1gpu.json 2> $out/bench_r02c_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/benchref_r02c_1gpu.json 2> $out/benchref_r02c_1gpu.err; echo "benchref rc=$?"
timeout 600 python bench.py --workload deconv --steps 3 --warmup 3 --iters-per-step 400 > $out/bench_r02c_deconv_1gpu.json 2> $out/bench_r02c_deconv_1gpu.err; echo "deconv rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --frames 296 --no-cpu-baseline --iters-per-step 20"
timeout 600 $CMD > $out/plain_r02c.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_r02c.csv $CMD > $out/ncu_launches_r02c.log 2>&1; echo "launch list rc=$?"
DC="python bench.py --workload deconv --steps 1 --warmup 1 --iters-per-step 20 --no-cpu-baseline"
timeout 600 $DC > $out/plain_deconv_r02c.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_deconv_epoch -s 25 -c 1 -o $out/prof_deconv_epoch_r02c $DC > $out/ncu_deconv_r02c.log 2>&1; echo "ncu deconv rc=$?"
ncu -i $out/prof_deconv_epoch_r02c.ncu-rep --page raw --csv > $out/prof_deconv_epoch_r02c_raw.csv 2>/dev/null
timeout 900 python bench.py --workload cfg5 --frames 1332 --steps 1 --warmup 1 --no-cpu-baseline > $out/bench_r02c_cfg5_1332.json 2> $out/bench_r02c_cfg5_1332.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_r02c_cfg3.json 2> $out/bench_r02c_cfg3.err; echo "cfg3 rc=$?"
