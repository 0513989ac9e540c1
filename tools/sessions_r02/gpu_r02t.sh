# A/B of the 16-output forward / adjoint passes of the per-epoch kernel (development switches LCB_DC_FWD8 / LCB_DC_ADJ8)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deconv_gpu.py -m gpu -q > gpurun_out/t_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_tests.log
for rep in 1 2; do
echo "== fwd16 adj16"; timeout 300 python tools/deconv_sweep.py --epochs 25,100,200 --cs 0 --iters 300
echo "== fwd8 adj16"; LCB_DC_FWD8=1 timeout 300 python tools/deconv_sweep.py --epochs 25,100,200 --cs 0 --iters 300
echo "== fwd8 adj8"; LCB_DC_FWD8=1 LCB_DC_ADJ8=1 timeout 300 python tools/deconv_sweep.py --epochs 25,100,200 --cs 0 --iters 300
echo "== fwd16 adj8"; LCB_DC_ADJ8=1 timeout 300 python tools/deconv_sweep.py --epochs 25,100,200 --cs 0 --iters 300
done 2>&1 | grep -v "^+" | tee gpurun_out/t_ab.log
