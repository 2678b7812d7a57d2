"""Parity at the BENCHMARKED configurations: the full 32-layer LLaMA-7B file that bench.py times and a
full 40-layer two-part LLaMA-13B file (BASELINE.json configs[1] / configs[3]), 16 greedy tokens after the 8-token prompt,
against the UNMODIFIED reference compiled under oracle/_ref.  Bar: logits within 1e-3 relative and identical arg-max
(BASELINE.json north_star); the run also reports how many steps are bit-identical."""
import os
import sys

import numpy as np
import pytest

import llama_swift_b200 as lsb
from conftest import CACHE, ROOT, CpuModel, bits, rel_l2
from llama_swift_b200 import ggml_format as gf

pytestmark = pytest.mark.gpu

PROMPT = np.array([1, 15043, 3186, 29892, 590, 1024, 338, 29871], np.int32)     # bench.py's prompt
N_GEN = 16


def _bench_model_7b():
    sys.path.insert(0, ROOT)
    import bench
    return bench.ensure_model(32)


def _model_13b():
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, "ggml-model-13b-q4_0-direct.bin")
    if not (os.path.exists(path) and os.path.exists(path + ".1")):
        hp = gf.HParams(n_vocab=32000, n_embd=5120, n_head=40, n_layer=40)
        gf.write_synthetic_model(path + ".tmp", hp, seed=0, mode="direct")
        os.replace(path + ".tmp", path)
        os.replace(path + ".tmp.1", path + ".1")
    return path


def _greedy_vs_reference(ref_lib, path, n_threads=8):
    ref = CpuModel(ref_lib, "ref_llama", path, 64)
    gpu = lsb.llama_model_load(path, n_ctx=64)
    try:
        want = ref.eval(n_threads, 0, PROMPT)
        got = lsb.llama_eval(gpu, n_threads, 0, PROMPT)
        worst, exact, n_past = rel_l2(got, want), int(np.array_equal(bits(got), bits(want))), len(PROMPT)
        assert int(got.argmax()) == int(want.argmax())
        cur = int(want.argmax())
        for _ in range(N_GEN):
            t = np.array([cur], np.int32)
            want = ref.eval(n_threads, n_past, t)
            got = lsb.llama_eval(gpu, n_threads, n_past, t)
            r = rel_l2(got, want)
            worst = max(worst, r)
            exact += int(np.array_equal(bits(got), bits(want)))
            assert r <= 1e-3, f"n_past {n_past}: rel-L2 {r:.3e} exceeds 1e-3"
            assert int(got.argmax()) == int(want.argmax()), f"n_past {n_past}: arg-max differs"
            cur = int(want.argmax())
            n_past += 1
        print(f"[parity-full] {os.path.basename(path)}: {N_GEN + 1} evals, worst rel-L2 {worst:.3e}, {exact} bit-identical")
        # the device-resident loop bench.py times must produce the same greedy stream as the C-ABI loop above
        first = int(lsb.llama_eval(gpu, n_threads, 0, PROMPT).argmax())
        toks, _, _ = gpu.decode_device(len(PROMPT), first, N_GEN, n_threads=n_threads)
        ref2 = CpuModel(ref_lib, "ref_llama", path, 64)
        try:
            c = int(ref2.eval(n_threads, 0, PROMPT).argmax())
            for i in range(N_GEN):
                c2 = int(ref2.eval(n_threads, len(PROMPT) + i, np.array([c], np.int32)).argmax())
                assert int(toks[i]) == c2, f"greedy token {i}: GPU {int(toks[i])} vs reference {c2}"
                c = c2
        finally:
            ref2.free()
        return worst, exact
    finally:
        ref.free()
        gpu.free()


def test_full_7b_bench_model_vs_reference(ref_lib):
    worst, exact = _greedy_vs_reference(ref_lib, _bench_model_7b())
    assert worst <= 1e-3


def test_full_13b_two_part_vs_reference(ref_lib):
    worst, exact = _greedy_vs_reference(ref_lib, _model_13b())
    assert worst <= 1e-3
