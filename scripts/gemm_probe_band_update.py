import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
from ap_vast_unofficial_b200 import _capi
lib = _capi.lib()
def run(name, M, N, K, batch, tb, tri, mir, bn, beta):
    ms = C.c_float(0)
    _capi.check(lib.apv_bench_gemm_shape(M, N, K, batch, tb, tri, mir, bn, beta, 20, C.byref(ms)))
    flops = 2.0 * M * N * K * batch * (0.5 * (1 + 128.0 / M) if tri else 1.0)
    print("%-50s %8.3f ms %6.2f TFLOP/s" % (name, ms.value, flops / ms.value / 1e9))
for M in (4032, 2048):
    run("K=64 tri+mirror beta=1 M=%d" % M, M, M, 64, 2, 1, 1, 1, 64, 1.0)
    run("K=64 tri        beta=1 M=%d" % M, M, M, 64, 2, 1, 1, 0, 64, 1.0)
    run("K=64 tri        beta=0 M=%d" % M, M, M, 64, 2, 1, 1, 0, 64, 0.0)
    run("K=64 full       beta=1 M=%d" % M, M, M, 64, 2, 1, 0, 0, 64, 1.0)
    run("K=64 full       beta=0 M=%d" % M, M, M, 64, 2, 1, 0, 0, 64, 0.0)
    run("K=128 tri+mirror beta=1 M=%d" % M, M, M, 128, 2, 1, 1, 1, 64, 1.0)
    run("K=128 tri        beta=1 M=%d" % M, M, M, 128, 2, 1, 1, 0, 64, 1.0)
