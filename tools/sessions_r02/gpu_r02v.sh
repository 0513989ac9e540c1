# one-level vs two-level fused reduction at small local epoch counts (LCB_DC_ONELEVEL_MAX), tail stamps
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deconv_gpu.py -m gpu -q > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/v_tests.log
for rep in 1 2; do
for mx in 0 32 64; do echo "== one level up to $mx epochs"; LCB_DC_ONELEVEL_MAX=$mx timeout 300 python tools/deconv_sweep.py --epochs 25,50 --cs 0 --iters 300; done
done 2>&1 | grep -v "^+" | tee gpurun_out/v_ab.log
for mx in 0 32; do echo "== one level up to $mx epochs"; LCB_DC_ONELEVEL_MAX=$mx LCB_LIBRARY=lightcurver_b200/liblcb_dctim.so timeout 300 python tools/deconv_sweep.py --epochs 25 --cs 0 --iters 40; done 2>&1 | grep -v "^+" | tee gpurun_out/v_dctim.log
