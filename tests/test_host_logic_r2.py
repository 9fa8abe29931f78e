"""Host-side logic added in round 2 that runs without a GPU: argument validation of the new wrappers (no CPU fallback: CPU
tensors raise before any kernel is touched), chunk schedules of the host pipeline and of the sharded gather, the gather
table of bench.py."""
import numpy as np
import pytest
import torch

import bench
from ddsp_svc_b200 import frontend, mel, pipeline, sharding
from ddsp_svc_b200.unit2control import Unit2Control, split_to_dict


def test_wrappers_refuse_cpu_tensors_and_bad_shapes():
    with pytest.raises(ValueError, match="CUDA"):
        frontend.volume_extract(torch.zeros(1, 4096), 512)
    with pytest.raises(ValueError, match="CUDA"):
        frontend.mask_apply_(torch.zeros(1, 1024), torch.ones(1, 2), 512)
    with pytest.raises(ValueError, match="CUDA"):
        frontend.cross_fade(torch.zeros(8), torch.zeros(8), 4)
    with pytest.raises(ValueError, match="CUDA"):
        mel.STFT(44100, 128, 2048, 2048, 512, 40, 16000).get_mel(torch.zeros(1, 8192))
    with pytest.raises(NotImplementedError):
        mel.STFT(44100, 128, 2048, 2048, 512, 40, 16000).get_mel(torch.zeros(1, 8192), keyshift=1)
    u = Unit2Control(16, 1, {"a": 3, "b": 5})
    with pytest.raises(ValueError, match="CUDA"):
        u(torch.zeros(1, 4, 16), torch.zeros(1, 4, 1), torch.zeros(1, 4, 1), torch.zeros(1, 4, 1))
    with pytest.raises(ValueError):
        Unit2Control.gemm_precision = "fp16"
        try:
            from ddsp_svc_b200.unit2control import _Gemm
            _Gemm(Unit2Control.gemm_precision)
        finally:
            Unit2Control.gemm_precision = "3xtf32"


def test_split_to_dict_returns_views_of_one_tensor():
    e = torch.arange(2 * 3 * 8, dtype=torch.float32).reshape(2, 3, 8)
    d = split_to_dict(e, {"a": 3, "b": 5})
    assert list(d) == ["a", "b"] and d["a"].shape == (2, 3, 3) and d["b"].shape == (2, 3, 5)
    assert d["b"].stride(1) == 8 and d["b"].data_ptr() == e.data_ptr() + 3 * 4          # strided views, no copy


def test_mel_filter_support_and_volume_extractor_contract():
    fb = mel.mel_filterbank(44100, 2048, 128, 40, 16000)
    lohi = mel._support(fb)
    assert lohi.dtype == np.int32 and lohi.shape == (128, 2) and (lohi[:, 1] > lohi[:, 0]).all()
    assert int((lohi[:, 1] - lohi[:, 0]).sum()) < 0.03 * fb.size                         # sparse: ~2 % of the dense matrix
    ve = frontend.Volume_Extractor(441)
    assert ve.hop_size == 441


def test_chunk_schedules_cover_the_batch_exactly_once():
    for chunks in (1, 4, 7, (6, 10, 10, 6), (4, 8, 12, 6, 2), (1, 0, 3)):
        for B in (1, 5, 32):
            spans = pipeline.chunk_bounds(B, chunks)
            assert spans[0][0] == 0 and spans[-1][1] == B and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(hi > lo for lo, hi in spans)
    with pytest.raises(ValueError):
        pipeline.chunk_bounds(8, (0, 0))
    for n in (1, 6, 32):
        for c in (1, 3, 4, 40):
            spans = sharding._chunk_bounds(n, c)
            assert spans[0][0] == 0 and spans[-1][1] == n and sum(hi - lo for lo, hi in spans) == n


def test_bench_gather_table_and_workload_registry():
    for n, (mode, chunks, streams) in bench.AUTO_GATHER.items():
        assert n in (2, 4, 8) and mode in ("peer", "peer-chunks", "peer-copy", "nccl") and chunks >= 1 and streams >= 1
    assert set(bench.OTHER_WORKLOADS) <= set(bench.WORKLOADS) and "sins" not in bench.OTHER_WORKLOADS
    for name, w in bench.WORKLOADS.items():
        assert bench.algorithmic_bytes(w, 861) > 0 and w["label"]
