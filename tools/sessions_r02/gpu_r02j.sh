set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_api_gpu.py tests/test_psf_gpu.py -m gpu -q -s -p no:cacheprovider -k "strided or chunked or fan_out" > gpurun_out/tests_r02j.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_r02j.log
timeout 900 python tools/inprocess_scaling.py --devices 1,2 > gpurun_out/inprocess_r02b_2gpu.jsonl 2> gpurun_out/inprocess_r02b_2gpu.err; echo "inprocess rc=$?"; cat gpurun_out/inprocess_r02b_2gpu.jsonl
