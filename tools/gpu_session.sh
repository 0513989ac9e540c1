#!/bin/bash
# One gpurun call of the build -> measure loop: GPU tests, the bench line, sweeps, sanitizer logs.  Everything lands in gpurun_out/.
# usage: tools/gpu_session.sh <tag> [steps...]   steps: tests bench sweep sanitize phase
tag=$1; shift
out=gpurun_out
mkdir -p $out
for step in "$@"; do
  case $step in
    tests)    timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $out/tests_$tag.log 2>&1; echo "tests rc=$?" ;;
    bench)    timeout 600 python bench.py --steps 3 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -c 600 $out/bench_$tag.json ;;
    benchref) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/benchref_$tag.json 2> $out/benchref_$tag.err; echo "benchref rc=$?" ;;
    sweep)    timeout 600 python tools/deconv_sweep.py --iters 100 --epochs 25 --cs 4,5,6,7,8 > $out/sweep_$tag.log 2>&1
              timeout 600 python tools/deconv_sweep.py --iters 100 --epochs 50 --cs 2,3,4,5,6,8 >> $out/sweep_$tag.log 2>&1
              timeout 600 python tools/deconv_sweep.py --iters 50 --epochs 100,200 --cs 2,3,4 >> $out/sweep_$tag.log 2>&1; echo "sweep rc=$?"; cat $out/sweep_$tag.log ;;
    sanitize) for tool in memcheck racecheck; do
                timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $out/sanitize_${tool}_$tag.log 2>&1; echo "$tool rc=$?"; tail -5 $out/sanitize_${tool}_$tag.log
              done ;;
    phase)    LCB_LIBRARY=lightcurver_b200/liblcb_timers.so timeout 300 python tools/phase_time.py cfg2 > $out/phase_$tag.log 2>&1; echo "phase rc=$?"; cat $out/phase_$tag.log ;;
  esac
done
