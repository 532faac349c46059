"""Role cycle counters of the TMA-fed channel mix (DSW_OPT_DEBUG bit 2048; dsw_debug_mix_counters) per tile, for short-K
shapes where the per-tile hand-over chain — not the tensor pipe — bounds the kernel.

    python tools/diag_mix_roles.py [FinxFout,...]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepsphere_weather_b200 import _lib  # noqa: E402

NAMES = ["prod wait stage", "conv wait TMA", "converting", "mma wait acc", "mma wait operands", "epi wait acc", "epi tmem->smem",
         "epi stores", "tiles", "kernel cycles"]


def main():
    shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[1] if len(sys.argv) > 1 else "24x128,64x256,256x64,256x512").split(",")]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    B, V = 32, 12288
    st = torch.cuda.current_stream().cuda_stream
    for Fin, Fout in shapes:
        x = torch.randn(B, V, Fin, device=dev)
        w = torch.randn(Fout, Fin, device=dev) * 0.05
        b = torch.randn(Fout, device=dev)
        y = torch.empty(B, V, Fout, device=dev)
        ws = torch.empty(lib.dsw_linear_workspace_bytes(B, V, Fin, Fout), dtype=torch.uint8, device=dev)
        for name, dbg in [("base", 0), ("none", 16 | 32 | 64 | 128 | 256)]:
            lib.dsw_set_option(2, dbg | 2048)
            out = (C.c_uint64 * 16)()
            for it in range(3):
                if it == 2:
                    torch.cuda.synchronize()
                    lib.dsw_debug_mix_counters(out, 1)
                _lib.check(lib.dsw_linear_fwd(x.data_ptr(), V * Fin, Fin, w.data_ptr(), b.data_ptr(), y.data_ptr(), B, V, Fin, Fout,
                                              ws.data_ptr(), ws.numel(), st), "linear_fwd")
            torch.cuda.synchronize()
            lib.dsw_debug_mix_counters(out, 1)
            lib.dsw_set_option(2, 0)
            tiles = max(int(out[8]), 1)
            print(f"{Fin}->{Fout} {name}: tiles/CTA {tiles / 148:.1f}; kernel {int(out[9])} cycles = {int(out[9]) / (tiles / 148):.0f} per tile; per tile: "
                  + ", ".join(f"{n} {int(out[i]) / tiles:.0f}" for i, n in enumerate(NAMES[:8])))


if __name__ == "__main__":
    main()
