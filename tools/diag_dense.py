"""Where do the tcgen05 dense kernels spend their time?  A/B timing of the channel mix (through
dsw_linear_fwd, i.e. P = 1) and the weight gradient (dsw_linear_bwd, dW only) with parts of the
pipeline switched off by DSW_OPT_DEBUG bits (results are wrong in those runs; timing only).

    python tools/diag_dense.py [B] [V]
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import _lib  # noqa: E402

OPT_DEBUG, OPT_MIX_BN, OPT_CONV = 2, 6, 7


def timed(fn, flush, iters=5, warm=2):
    ts = []
    for i in range(warm + iters):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 12288
    dev = torch.device("cuda:0")
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    N = B * V
    variants = [("base", 0, 0), ("no-store", 16, 0), ("no-A", 32, 0), ("no-B", 64, 0), ("no-AB", 96, 0), ("no-conv", 128, 0),
                ("no-mma", 256, 0), ("mma-only", 16 | 32 | 64 | 128, 0), ("none", 16 | 32 | 64 | 128 | 256, 0),
                ("BN128", 0, 128), ("cvt-F2F", 0, -1), ("cvt-int", 0, -2)]
    print("== channel mix (dsw_linear_fwd): us per call")
    print(f"{'Fin->Fout':>12s} " + " ".join(f"{n:>9s}" for n, _, _ in variants) + "   floorHBM floorMMA")
    shapes = [(256, 512), (256, 128), (512, 256), (128, 256), (64, 256), (64, 64), (256, 64)]
    if os.environ.get("DIAG_SHAPES"):  # e.g. DIAG_SHAPES=24x128,64x256
        shapes = [tuple(int(v) for v in s_.split("x")) for s_ in os.environ["DIAG_SHAPES"].split(",")]
    for Fin, Fout in shapes:
        x = torch.randn(B, V, Fin, device=dev)
        w = torch.randn(Fout, Fin, device=dev) * 0.05
        b = torch.randn(Fout, device=dev)
        y = torch.empty(B, V, Fout, device=dev)
        row = []
        for name, dbg, bn in variants:
            lib.dsw_set_option(OPT_DEBUG, dbg)
            lib.dsw_set_option(OPT_MIX_BN, max(bn, 0))
            lib.dsw_set_option(OPT_CONV, max(-bn, 0))
            ws = torch.empty(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), dtype=torch.uint8, device=dev)

            def f():
                _lib.check(lib.dsw_linear_fwd(x.data_ptr(), V * Fin, Fin, w.data_ptr(), b.data_ptr(), y.data_ptr(), B, V, Fin, Fout,
                                              ws.data_ptr(), ws.numel(), st), "linear_fwd")
            row.append(timed(f, flush))
        lib.dsw_set_option(OPT_DEBUG, 0)
        lib.dsw_set_option(OPT_MIX_BN, 0)
        lib.dsw_set_option(OPT_CONV, 0)
        hbm = 4 * N * (Fin + Fout) / 6.4e12 * 1e6
        mma = 3 * 2 * N * Fin * Fout / 1.6e15 * 1e6
        print(f"{Fin:5d}->{Fout:<5d} " + " ".join(f"{t:9.1f}" for t in row) + f"   {hbm:8.1f} {mma:8.1f}", flush=True)
        del x, y

    if os.environ.get("DIAG_MIX_ONLY"):
        return
    wvariants = [("base", 0), ("no-A", 32), ("no-B", 64), ("no-AB", 96), ("no-conv", 128), ("no-mma", 256),
                 ("mma-only", 32 | 64 | 128), ("none", 32 | 64 | 128 | 256), ("cvt-F2F", -1), ("cvt-int", -2)]
    print("== weight gradient (dsw_linear_bwd, dW only): us per call;  M side = dy channels, N side = x channels")
    print(f"{'M x N':>12s} " + " ".join(f"{n:>9s}" for n, _ in wvariants) + "   floorHBM floorMMA")
    for Fout, Fin in [(256, 512), (128, 256), (128, 512), (64, 64), (64, 256), (512, 256)]:
        x = torch.randn(B, V, Fin, device=dev)
        dy = torch.randn(B, V, Fout, device=dev)
        w = torch.randn(Fout, Fin, device=dev) * 0.05
        dw = torch.empty_like(w)
        row = []
        for name, dbg in wvariants:
            lib.dsw_set_option(OPT_DEBUG, max(dbg, 0))
            lib.dsw_set_option(OPT_CONV, max(-dbg, 0))
            ws = torch.empty(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), dtype=torch.uint8, device=dev)

            def f():
                _lib.check(lib.dsw_linear_bwd(x.data_ptr(), V * Fin, Fin, dy.data_ptr(), w.data_ptr(), None, dw.data_ptr(), None,
                                              B, V, Fin, Fout, ws.data_ptr(), ws.numel(), st), "linear_bwd")
            row.append(timed(f, flush))
        lib.dsw_set_option(OPT_DEBUG, 0)
        lib.dsw_set_option(OPT_CONV, 0)
        hbm = 4 * N * (Fin + Fout) / 6.4e12 * 1e6
        mma = 3 * 2 * N * Fin * Fout / 1.6e15 * 1e6
        print(f"{Fout:5d}x{Fin:<5d}  " + " ".join(f"{t:9.1f}" for t in row) + f"   {hbm:8.1f} {mma:8.1f}", flush=True)
        del x, dy


if __name__ == "__main__":
    main()
