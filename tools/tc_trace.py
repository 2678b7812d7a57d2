"""Timeline of the tcgen05 prefill kernel's hand-overs (development build -DB200_TC_TRACE=1 -> libb200_trace.so).

Runs one 4096 x 4096 Q4_0 mat-mul over 256 columns through b200_q4_0_matmul(path=1) and prints, for CTA 0's first block
steps, when each role reached each hand-over point (cycles relative to the first TMA issue)."""
import ctypes, os, sys
import numpy as np

os.environ.setdefault("B200_LIB", os.path.join(os.path.dirname(__file__), "..", "llama.swift_b200", "libb200_trace.so"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import llama_swift_b200 as lsb  # noqa: E402
from llama_swift_b200 import ggml_format as gf  # noqa: E402


def main():
    M, K, N = 4096, 4096, 256
    rng = np.random.default_rng(0)
    w = gf.quantize_q4_0((rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32))
    x = rng.standard_normal((N, K), dtype=np.float32)
    out, ms = lsb.q4_0_matmul(w, x, path=1, timed=True)
    print(f"kernel {ms:.3f} ms")
    lib = ctypes.CDLL(os.environ["B200_LIB"])
    buf = (ctypes.c_longlong * (12 * 512))()
    assert lib.b200_debug_tc_trace(buf, 12 * 512) == 0
    t = np.frombuffer(buf, dtype=np.int64).reshape(12, 512).copy()
    t0 = t[0, 0]
    t = t - t0
    print("quad: tma_issue | unpack: raw_full seen, stores done, fence done | epi: quad done")
    for g in range(40, 64):
        print(f"g={g:3d} tma {t[0,g]:8d} | raw_full {t[1,g]:8d} stored {t[2,g]:8d} fenced {t[3,g]:8d} | epi quad done {t[9,g]:8d}")
    print("step: mma: ab_full seen, tm_empty seen, committed | epi: wait start, tm_full seen, ld done")
    for k in range(160, 256):
        print(f"k={k:3d} mma ab_full {t[4,k]:8d} tm_empty {t[5,k]:8d} commit {t[6,k]:8d} | epi start {t[10,k]:8d} tm_full {t[7,k]:8d} ld done {t[8,k]:8d}")
    d = np.diff(t[6, 64:448])
    print(f"steady state: {d.mean():.1f} cycles per block step (steps 64..448)")
    for name, a, b in (("tma issue -> raw_full seen by unpack", t[0, 16:112], t[1, 16:112]), ("unpack raw_full -> stored", t[1, 16:112], t[2, 16:112]),
                       ("unpack stored -> fenced", t[2, 16:112], t[3, 16:112]), ("unpack fenced -> mma saw ab_full (first block of quad)", t[3, 16:112], t[4, 64:448:4]),
                       ("mma ab_full -> tm_empty", t[4, 64:448], t[5, 64:448]), ("mma tm_empty -> commit issued", t[5, 64:448], t[6, 64:448]),
                       ("commit -> epi saw tm_full", t[6, 64:448], t[7, 64:448]), ("epi tm_full -> ld done", t[7, 64:448], t[8, 64:448]),
                       ("epi ld done -> next wait start", t[8, 64:447], t[10, 65:448]), ("epi wait start -> tm_full", t[10, 64:448], t[7, 64:448]),
                       ("epi quad done -> tma issue of quad+3", t[9, 16:109], t[0, 19:112])):
        dd = b - a
        print(f"{name:55s} mean {dd.mean():8.1f}  min {dd.min():6d}  max {dd.max():6d}")


if __name__ == "__main__":
    main()
