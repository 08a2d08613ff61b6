"""CPU-side checks of the product libraries: they build, load, export every
symbol the headers declare, validate arguments, and refuse to run without a GPU
(no CPU fallback). No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__
    __graft_entry__.build()
    import kafka_lag_based_assignor_b200 as p
    return p


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lagassign_[a-z_0-9]+)\s*\(", src)))


def test_headers_declare_expected_entry_points():
    names = _declared("lagassign.h")
    for must in ("lagassign_create", "lagassign_assign_batch", "lagassign_upload", "lagassign_execute",
                 "lagassign_download", "lagassign_destroy", "lagassign_last_error"):
        assert must in names


def test_cuda_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.library_path())
    for name in _declared("lagassign.h"):
        assert hasattr(lib, name), name
    assert lib.lagassign_abi_version() == 3


def test_host_library_exports_every_declared_symbol(pkg):
    pkg.load_library()
    lib = ctypes.CDLL(os.path.join(os.path.dirname(pkg.library_path()), "liblagassign_host.so"))
    for name in _declared("lagassign_host.h"):
        assert hasattr(lib, name), name


def test_library_is_sm100a_only(pkg):
    out = os.popen(f"cuobjdump --list-elf {pkg.library_path()} 2>/dev/null").read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu(pkg):
    assert pkg.device_count() == 0
    with pytest.raises(pkg.LagAssignError) as e:
        pkg.FlatAssignor(0)
    assert e.value.code == -5
    with pytest.raises(pkg.LagAssignError) as e:
        pkg.FlatAssignor([0, 1])                       # lagassign_create_multi: no device, no fallback either
    assert e.value.code == -5
    with pytest.raises(pkg.LagAssignError) as e:
        pkg.FlatAssignor([])
    assert e.value.code == -1
    with pytest.raises(pkg.LagAssignError):
        a = pkg.LagBasedPartitionAssignor()
        a.assign({"t": [pkg.TopicPartitionLag("t", 0, 1)]}, {"c": ["t"]})


def test_host_mirror_configure_and_name_without_gpu(pkg):
    """configure / name / computePartitionLag are host-only (Main:97-135, :376-404)."""
    a = pkg.LagBasedPartitionAssignor()
    assert a.name() == "lag"                                   # Main:134
    with pytest.raises(ValueError) as e:                       # Main:108-113
        a.configure({"bootstrap.servers": "x:9092"})
    assert "group.id cannot be null" in str(e.value)
    a.configure({"group.id": "g1", "auto.offset.reset": "earliest"})
    assert a.metadata_client_id == "g1.assignor"               # Main:119-120
    # the four reference vectors, Test:21-80
    assert pkg.compute_partition_lag(5555, 1111, 9999, "none") == 4444
    assert pkg.compute_partition_lag(5555, 0, 0, "none") == 0
    assert pkg.compute_partition_lag(None, 1111, 9999, "latest") == 0
    assert pkg.compute_partition_lag(None, 1111, 9999, "earliest") == 8888
    assert pkg.compute_partition_lag(None, 1111, 9999, "LaTeSt") == 0


def test_host_mirror_empty_inputs_need_no_gpu(pkg):
    """Members with no topics / topics with no lags never reach the device (Main:171-174, :182)."""
    a = pkg.LagBasedPartitionAssignor()
    assert a.assign({}, {"idle": [], "busy": ["ghost"]}) == {"idle": [], "busy": []}
    assert a.topic_order == ["ghost"]
