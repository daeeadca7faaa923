"""FP64 pipes of the GPU: DMMA (tensor) peak next to the CUDA-core DFMA rate and latency."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ap_vast_unofficial_b200 import _capi
tf = C.c_double(0)
_capi.check(_capi.lib().apv_bench_dmma_peak(4000, C.byref(tf)))
o = (C.c_double * 3)()
_capi.check(_capi.lib().apv_bench_dfma(4000, o))
print("DMMA peak %.1f TFLOP/s | DFMA %.2f TFLOP/s, %.1f cycles per dependent DFMA, one warp-DFMA per %.1f cycles per sub-partition"
      % (tf.value, o[0], o[1], o[2]))
