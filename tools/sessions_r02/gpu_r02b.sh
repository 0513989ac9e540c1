set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/tests_r02b.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_r02b.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r02b.json')); print(d['value'], d['e2e']['value'], d['deconv']['value'], d['deconv']['kernels'])"
CMD="python bench.py --steps 1 --warmup 1 --frames 296 --no-cpu-baseline --no-deconv"
$CMD > gpurun_out/plain_r02b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_psf_fit -s 1 -c 1 -o gpurun_out/prof_psf_fit_r02b $CMD > gpurun_out/ncu_r02b.log 2>&1; echo "ncu rc=$?"
bash tools/gpu_session.sh r02b sanitize
