"""Role timing of the tcgen05 weight-gradient kernel (DSW_OPT_DEBUG bit 512), paired and unpaired.

    python tools/diag_wgrad.py
"""
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import _lib  # noqa: E402

OPT_DEBUG, OPT_PAIR = 2, 12


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, V = 32, 12288
    for Fout, Fin in [(256, 512), (512, 256), (256, 128), (256, 64)]:
        x = torch.randn(B, V, Fin, device=dev)
        dy = torch.randn(B, V, Fout, device=dev)
        w = torch.randn(Fout, Fin, device=dev) * 0.05
        dw = torch.empty_like(w)
        ws = torch.empty(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), dtype=torch.uint8, device=dev)

        def f():
            _lib.check(lib.dsw_linear_bwd(x.data_ptr(), V * Fin, Fin, dy.data_ptr(), w.data_ptr(), None, dw.data_ptr(), None,
                                          B, V, Fin, Fout, ws.data_ptr(), ws.numel(), st), "linear_bwd")
        for no_pair in (0, 1, 2):
            lib.dsw_set_option(OPT_PAIR, 1 if no_pair == 0 else 0)
            lib.dsw_set_option(3, 1 if no_pair == 2 else 0)  # DSW_OPT_NO_TMA: register-path kernels
            ts = []
            for i in range(7):
                flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); e1.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1) * 1e3)
            buf = (ctypes.c_uint64 * 8)()
            lib.dsw_set_option(OPT_DEBUG, 512)
            f(); torch.cuda.synchronize()
            lib.dsw_debug_dense_counters(buf, 1)
            f(); torch.cuda.synchronize()
            lib.dsw_debug_dense_counters(buf, 1)
            lib.dsw_set_option(OPT_DEBUG, 0)
            n = max(buf[2], 1)
            print(f"M {Fout} x N {Fin} {('paired  ', 'unpaired', 'reg-path')[no_pair]}: {statistics.median(ts):7.1f} us | per stage and CTA (cycles): "
                  f"conv-wait-TMA {buf[0] / n:6.0f}  convert {buf[1] / n:6.0f}  producer-wait-empty {buf[3] / n:6.0f}  "
                  f"mma-wait-operands {buf[4] / n:6.0f}  stages {n}", flush=True)
        lib.dsw_set_option(OPT_PAIR, 0)
        lib.dsw_set_option(3, 0)


if __name__ == "__main__":
    main()
