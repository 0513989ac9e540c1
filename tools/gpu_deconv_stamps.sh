# row-wise translation paths of the f build / shift gradient / transposed warp: parity, then timing with the phase stamps
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deconv_gpu.py tests/test_api_gpu.py tests/test_starred_api_gpu.py -m gpu -q > gpurun_out/${TAG:-w}_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/${TAG:-w}_tests.log
timeout 300 python tools/deconv_sweep.py --epochs 25,50,100,200 --cs 0 --iters 300 2>&1 | grep -v "^+" | tee gpurun_out/${TAG:-w}_sweep.log
LCB_LIBRARY=lightcurver_b200/liblcb_dctim.so timeout 300 python tools/deconv_sweep.py --epochs 25,200 --cs 0 --iters 40 2>&1 | grep -v "^+" | tee gpurun_out/${TAG:-w}_dctim.log
