"""Soak test of the barrier-free whole-token kernel: > 10^4 teacher-forced decode steps through
the device-resident loop, every step compared with the oracle.  Answers two questions with a measurement instead of
prose: (1) does the flagged-word / arrival-counter exchange ever deliver a stale or torn value over many launches
(epoch counter, never-reset hint counters, parity double buffering), (2) how often does the one non-mirrored detail -- the
tree-ordered double sums of the LayerNorm -- flip a result bit.  Bar per step: rel-L2 <= 1e-3 and identical arg-max; the
number of bit-identical steps is reported (expected: all)."""
import numpy as np
import pytest

import llama_swift_b200 as lsb
from conftest import CpuModel, bits, rel_l2

pytestmark = pytest.mark.gpu

SEGMENTS, STEPS = 20, 512          # 10 240 steps; every segment restarts at position 0 with a new token stream


def test_decode_soak_10k_steps_vs_oracle(oracle_lib, small_model):
    ora = CpuModel(oracle_lib, "ora", small_model, STEPS + 8)
    gpu = lsb.llama_model_load(small_model, n_ctx=STEPS + 8)
    try:
        exact = total = 0
        worst = 0.0
        for seg in range(SEGMENTS):
            rng = np.random.default_rng(1000 + seg)
            stream = rng.integers(3, 512, size=STEPS + 1).astype(np.int32)
            toks, logits, _ = gpu.decode_device(0, int(stream[0]), STEPS, n_threads=8, forced_tokens=stream[1:], want_logits=True)
            for i in range(STEPS):
                want = ora.eval(8, i, stream[i:i + 1])
                same = np.array_equal(bits(logits[i]), bits(want))
                if not same:
                    r = rel_l2(logits[i], want)
                    worst = max(worst, r)
                    assert r <= 1e-3 and int(toks[i]) == int(want.argmax()), f"segment {seg} step {i}: rel-L2 {r:.3e}"
                exact += int(same)
                total += 1
        print(f"[soak] {total} steps: {exact} bit-identical, {total - exact} within tolerance (worst rel-L2 {worst:.3e}); "
              f"{(total - exact) / total:.2e} non-identical steps per step (LayerNorm tree-order flips)")
        assert exact >= total - max(2, total // 1000)
    finally:
        ora.free()
        gpu.free()
