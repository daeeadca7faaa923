"""FP64 DMMA GEMM building block on the shapes of the joint diagonalisation (cfg-3 sizes): TFLOP/s per shape.
Usage: python scripts/bench_gemm_shapes.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ap_vast_unofficial_b200 import _capi

lib = _capi.lib()
tf = C.c_double(0)
_capi.check(lib.apv_bench_dmma_peak(4000, C.byref(tf)))
print("DMMA peak %.2f TFLOP/s" % tf.value)
shapes = [
    # name, M, N, K, batch, transB, tri, mirror, bn, beta
    ("band update  C -= Z1 Z2^T   (K=64, lower tiles + mirror)", 3072, 3072, 64, 2, 1, 1, 1, 64, 1.0),
    ("band update, first panel    (K=64)", 4032, 4032, 64, 2, 1, 1, 1, 64, 1.0),
    ("cholesky trailing update    (K=256, lower tiles)", 3072, 3072, 256, 2, 1, 1, 0, 0, 1.0),
    ("reduction update            (K=256, full)", 3072, 4096, 256, 2, 1, 0, 0, 0, 1.0),
    ("reduction update, NN        (K=256, full)", 3072, 4096, 256, 2, 0, 0, 0, 0, 1.0),
    ("super-block inverse product (M=256, K=256)", 256, 4096, 256, 2, 0, 0, 0, 0, 0.0),
    ("cholesky panel              (N=64, K=64)", 3072, 64, 64, 2, 1, 0, 0, 0, 0.0),
    ("skinny Y = C22 V            (N=32, K=3072)", 3072, 32, 3072, 2, 0, 0, 0, 0, 0.0),
    ("square 4096^3 NT", 4096, 4096, 4096, 1, 1, 0, 0, 0, 0.0),
    ("square 4096^3 NN, 128x64 tile", 4096, 4096, 4096, 1, 0, 0, 0, 64, 0.0),
    ("square 4096^3 NT, 128x64 tile", 4096, 4096, 4096, 1, 1, 0, 0, 64, 0.0),
    ("K = 2048 NT (auto tile)", 4096, 4096, 2048, 2, 1, 0, 0, 0, 0.0),
    ("K = 2048 NT, 128x64 tile", 4096, 4096, 2048, 2, 1, 0, 0, 64, 0.0),
]
for name, M, N, K, batch, tb, tri, mir, bn, beta in shapes:
    ms = C.c_float(0)
    _capi.check(lib.apv_bench_gemm_shape(M, N, K, batch, tb, tri, mir, bn, beta, 20, C.byref(ms)))
    flops = 2.0 * M * N * K * batch * (0.5 * (1 + 128.0 / M) if tri else 1.0)
    print("%-62s %8.3f ms  %6.2f TFLOP/s  (%4.1f %% of peak)" % (name, ms.value, flops / ms.value / 1e9,
                                                                100 * flops / ms.value / 1e9 / tf.value))
