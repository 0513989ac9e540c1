set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/tests_r02i.log 2>&1; echo "tests rc=$?"; grep -v "^\[parity\]" gpurun_out/tests_r02i.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02i.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02i.log
timeout 600 python bench.py > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err; echo "bench rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02i.json') if l.startswith('{')][-1]; print(d['value'], d['ms_per_step'], d['e2e'], d['deconv']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/benchref_r02i.json 2> gpurun_out/benchref_r02i.err; echo "benchref rc=$?"; tail -c 400 gpurun_out/benchref_r02i.json
timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 3 > gpurun_out/bench_r02i_cfg3.json 2> gpurun_out/bench_r02i_cfg3.err; echo "cfg3 rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_r02i_cfg3.json') if l.startswith('{')][-1]; print('cfg3', d['value'], d['e2e']['value'], d['roofline']['frac'])"
CMD="python bench.py --steps 1 --warmup 1 --frames 296 --no-cpu-baseline --iters-per-step 20"
$CMD > gpurun_out/plain_r02i.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches_r02i.log 2>&1; echo "ncu launches rc=$?"
DC="python bench.py --workload deconv --steps 1 --warmup 1 --iters-per-step 20 --no-cpu-baseline"
$DC > gpurun_out/plain_deconv_r02i.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_deconv_epoch -s 25 -c 1 -o gpurun_out/prof_deconv_epoch_r02 $DC > gpurun_out/ncu_deconv_r02i.log 2>&1; echo "ncu deconv rc=$?"
ncu --set full --clock-control none -k regex:k_deconv_starlet_sm -s 25 -c 1 -o gpurun_out/prof_deconv_starlet_r02 $DC > gpurun_out/ncu_starlet_r02i.log 2>&1; echo "ncu starlet rc=$?"
