set -x
mkdir -p gpurun_out
LCB_DECONV_FUSED_MAX=0 timeout 900 python -m pytest tests/test_deconv_gpu.py -m gpu -q -p no:cacheprovider -k "two_ranks" > gpurun_out/ac_tests.log 2>&1; echo "two-rank tests (separate reduce) rc=$?"; tail -3 gpurun_out/ac_tests.log
bash tools/gpu_deconv_routes_multi.sh 2 r02x
