"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/b200_llama.h declares, and refuses to run without a GPU (no CPU fallback) with the reference's error codes."""
import ctypes as C
import os
import re

import pytest

import llama_swift_b200 as lsb
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200_llama.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", text)))


def test_header_symbols_exported():
    L = lsb.lib()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/b200_llama.h but not exported"


def test_no_oracle_in_product():
    """The product never links or mentions oracle/ (a CPU fallback would void every parity claim)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "llama.swift_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".swift", ".mm")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in src and "llama_oracle" not in src and "libllama_ref" not in src, f


def test_fails_loudly_without_gpu_or_file(tmp_path):
    import torch
    if torch.cuda.is_available():
        with pytest.raises(lsb.LlamaError) as ei:
            lsb.llama_model_load(str(tmp_path / "missing.bin"))
        assert ei.value.code == lsb.ERR_LOAD and "failed to open" in ei.value.message   # PO.mm:100-104
    else:
        with pytest.raises(lsb.LlamaError) as ei:
            lsb.llama_model_load(str(tmp_path / "missing.bin"))
        assert ei.value.code == lsb.ERR_LOAD and "no CUDA device" in ei.value.message


def test_runner_fails_loudly_without_gpu(tmp_path):
    """b200_llama_run (the token loop of -main) posts startedLoadingModel, then failedWithError(-1000) when the model cannot
    be loaded -- on a box without a GPU that is every run: no CPU path exists."""
    import torch
    events = []
    runner = lsb.LlamaRunner(str(tmp_path / "missing.bin"))
    with pytest.raises(lsb.LlamaError) as ei:
        runner.run("hello", on_event=lambda kind, piece, code: events.append((kind, code, piece)))
    assert ei.value.code == lsb.ERR_LOAD
    assert events[0][0] == lsb.EVENT_STARTED_LOADING_MODEL
    assert events[-1][0] == lsb.EVENT_FAILED and events[-1][1] == lsb.ERR_LOAD
    if not torch.cuda.is_available():
        assert b"no CUDA device" in events[-1][2]
