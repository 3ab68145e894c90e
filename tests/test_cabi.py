"""CPU suite: the C-ABI boundary.  The CUDA library must load without a GPU, export every symbol include/zkb.h
declares, agree with the ctypes / numpy mirrors on every struct layout, and refuse to run without a device
(there is no CPU fallback behind it)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from era_zk_evm_b200 import _binding, records

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zkb.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("zkb_create", "zkb_destroy", "zkb_load_bytecode", "zkb_populate_storage", "zkb_push_bootloader_context",
                 "zkb_run", "zkb_sync", "zkb_vm_status", "zkb_read_local_state", "zkb_read_stream", "zkb_fetch_stream_packed"):
        assert must in names
    assert len(names) >= 28


@pytest.fixture(scope="module")
def cuda_lib():
    from era_zk_evm_b200 import load_library
    return load_library()       # raises if the extension is not built: the product has no fallback


def test_cuda_library_exports_every_declared_symbol(cuda_lib):
    missing = [n for n in declared_functions() if not hasattr(cuda_lib, n)]
    assert not missing, missing


def test_oracle_exports_the_same_surface(oracle_mod):
    """the checker mirrors the boundary (prefix orc_) so parity tests drive both through one code path"""
    lib = oracle_mod.lib()
    shared = [n for n in declared_functions() if n not in (
        "zkb_stream_device_view", "zkb_fetch_stream_packed", "zkb_fetch_stream_packed_async", "zkb_pack_stream_device", "zkb_pack_stream_device_async", "zkb_snapshot", "zkb_restore",
        "zkb_transfer_stats", "zkb_encode_streams_device", "zkb_fetch_encoded_kinds_async", "zkb_gather_streams", "zkb_exchange_logs", "zkb_comm_unique_id", "zkb_comm_create", "zkb_comm_destroy", "zkb_comm_wait_packed", "zkb_exchange_step", "zkb_push_step", "zkb_push_result", "zkb_consume", "zkb_snapshot_counts", "zkb_read_snapshots", "zkb_read_queue_digests", "zkb_fetch_consumed_async", "zkb_sort_log_queries", "zkb_alu_microbench", "zkb_peer_push_async", "zkb_peer_sink_create", "zkb_peer_sink_open", "zkb_peer_sink_close")]
    missing = [n for n in shared if not hasattr(lib, n.replace("zkb_", "orc_", 1))]
    assert not missing, missing


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof/offsetof from a C compiler over include/*.h == the ctypes structures and numpy record dtypes"""
    probe = tmp_path / "probe.c"
    fields = {
        "ZkbConfig": ["n_vms", "device", "witness_mode", "cap_records", "stack_words", "heap_bytes", "n_heap_slabs",
                      "max_far_depth", "max_depth", "storage_slots", "journal_entries", "host_mirror", "schedule"],
        "ZkbFrame": [n for n, _ in _binding.ZkbFrame._fields_],
        "ZkbLocalState": [n for n, _ in _binding.ZkbLocalState._fields_],
        "ZkbVmStatus": ["code", "cycles"],
        "ZkbCycleRow": list(records.ROW_DTYPE.names),
        "ZkbMemoryQueryRec": list(records.MEM_DTYPE.names),
        "ZkbLogQueryRec": list(records.LOG_DTYPE.names),
        "ZkbDecommitRec": list(records.DECOMMIT_DTYPE.names),
        "ZkbFrameRec": list(records.FRAME_DTYPE.names),
        "ZkbRefundRec": list(records.REFUND_DTYPE.names),
        "ZkbStorageInit": ["shard_id", "reserved", "address", "key_be", "value_be"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for s, fs in fields.items():
        lines.append(f'printf("{s} size %zu\\n", sizeof({s}));')
        for f in fs:
            lines.append(f'printf("{s} {f} %zu\\n", offsetof({s}, {f}));')
    lines.append("return 0;}")
    probe.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c11", "-o", str(exe), str(probe)])
    out = subprocess.check_output([str(exe)], text=True)
    c_layout = {}
    for line in out.splitlines():
        s, f, v = line.split()
        c_layout[(s, f)] = int(v)
    ct = {"ZkbConfig": _binding.ZkbConfig, "ZkbFrame": _binding.ZkbFrame, "ZkbLocalState": _binding.ZkbLocalState,
          "ZkbVmStatus": _binding.ZkbVmStatus}
    for s, cls in ct.items():
        assert C.sizeof(cls) == c_layout[(s, "size")], s
        for f in fields[s]:
            assert getattr(cls, f).offset == c_layout[(s, f)], (s, f)
    nd = {"ZkbCycleRow": records.ROW_DTYPE, "ZkbMemoryQueryRec": records.MEM_DTYPE, "ZkbLogQueryRec": records.LOG_DTYPE,
          "ZkbDecommitRec": records.DECOMMIT_DTYPE, "ZkbFrameRec": records.FRAME_DTYPE, "ZkbRefundRec": records.REFUND_DTYPE,
          "ZkbStorageInit": _binding.STORAGE_INIT_DTYPE}
    for s, dt in nd.items():
        assert dt.itemsize == c_layout[(s, "size")], s
        for f in fields[s]:
            assert dt.fields[f][1] == c_layout[(s, f)], (s, f)
    for kind, nbytes in enumerate(records.RECORD_BYTES):
        assert records.DTYPES[kind].itemsize == nbytes


def test_no_cpu_fallback_without_a_device(cuda_lib):
    """zkb_create must fail loudly (ZKB_ERR_NO_DEVICE / ZKB_ERR_CUDA) when no GPU is visible"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is exercised in the CPU container")
    from era_zk_evm_b200 import GpuVmBatch, ZkbError, default_config
    with pytest.raises(ZkbError) as e:
        GpuVmBatch(default_config(4))
    assert "status 5" in str(e.value) or "status 2" in str(e.value)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under era_zk_evm_b200/ may import, load or link it"""
    pkg = os.path.join(ROOT, "era_zk_evm_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M) or re.search(r"(CDLL|dlopen)\([^)]*orc", text):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
