set -x
mkdir -p gpurun_out
python tools/ffma2_probe.py > gpurun_out/ffma2_probe_r02.txt 2>&1; cat gpurun_out/ffma2_probe_r02.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/tests_r02e.log 2>&1; echo "tests rc=$?"; grep -v "^\[parity\]" gpurun_out/tests_r02e.log | tail -15; grep "parity\] ROI\|parity\] distortion" gpurun_out/tests_r02e.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r02e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['deconv']['value'])"
