"""FP32 issue probes on the box: scalar FFMA chains, FFMA with register operands, packed FFMA2 (sm_100), and both forms mixed
with shared-memory loads.  python tools/ffma2_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lightcurver_b200 import _lib
print('FFMA chains (immediates)     %.1f TFLOP/s' % _lib.fp32_peak()[0])
print('FFMA 8x8 register operands   %.1f TFLOP/s' % _lib.fp32_peak_rrr()[0])
print('FFMA2 chains                 %.1f TFLOP/s' % _lib.fp32x2_peak(0)[0])
print('FFMA  + 1 LDS per 2 FMAs     %.1f TFLOP/s' % _lib.fp32x2_peak(1)[0])
print('FFMA2 + 1 LDS per FFMA2      %.1f TFLOP/s' % _lib.fp32x2_peak(2)[0])
