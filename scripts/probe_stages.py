"""Quick per-stage device timing of the engine on synthetic workloads (not the bench)."""
import ctypes as C
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ap_vast_unofficial_b200 import apvast, _capi
from ap_vast_unofficial_b200.workloads import make_workload

def main():
    names = sys.argv[1:] or ["cfg2", "cfg3"]
    tf = C.c_double(0)
    _capi.check(_capi.lib().apv_bench_dmma_peak(4000, C.byref(tf)))
    print("DMMA peak TFLOP/s", tf.value)
    for nn in (1024, 4096):
        ms = C.c_float(0)
        _capi.check(_capi.lib().apv_bench_gemm(nn, 5, C.byref(ms)))
        print("gemm n=%d: %.3f ms -> %.2f TFLOP/s" % (nn, ms.value, 2.0 * nn**3 / ms.value / 1e9))
    for name in names:
        wl = make_workload(name)
        np.random.seed(0)
        t0 = time.time()
        eng = apvast(rir_A=wl["rir_A"], rir_B=wl["rir_B"], perceptual=False, stats_mode=int(os.environ.get("APV_STATS_MODE", "0")), **wl["cfg"])
        print(name, "create %.2fs" % (time.time() - t0))
        H = eng.hop_size
        for t in range(6):
            a = wl["signal_A"][t * H:(t + 1) * H]; b = wl["signal_B"][t * H:(t + 1) * H]
            t0 = time.time()
            eng.process_input_buffers(a, b)
            dt = time.time() - t0
            print(name, t, "wall %.1f ms" % (dt * 1e3), {k: round(v, 2) for k, v in eng.stage_times().items()})
        eng.close()

if __name__ == "__main__":
    main()
