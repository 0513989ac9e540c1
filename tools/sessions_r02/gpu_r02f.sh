set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_psf_gpu.py tests/test_golden_gpu.py tests/test_deconv_gpu.py -m gpu -q -s -p no:cacheprovider -k "not configured_length" > gpurun_out/tests_r02f.log 2>&1; echo "tests rc=$?"; grep -v "^\[parity\]" gpurun_out/tests_r02f.log | tail -8; grep "parity\] ROI" gpurun_out/tests_r02f.log
for lib in liblcb_nopk.so liblcb.so liblcb_nopk.so liblcb.so; do
  echo "== $lib"; LCB_LIBRARY=lightcurver_b200/$lib timeout 300 python tools/quick_time.py 592 300 2>&1 | grep "T2=300 W=False\|phot B"
done
