"""SURVEY §8 row f-4 — bytecode ingestion: versioned code hashes (ContractCodeSha256 layout parsed at
/root/reference/src/opcodes/execution/far_call.rs:169-252; the key of SimpleDecommitter::populate,
/root/reference/src/reference_impls/decommitter.rs:23-28).

CPU suite: the oracle's restatement against hashlib.sha256 (independent of the oracle) on empty / one-word / odd /
even / block-boundary / maximum lengths.  -m gpu: the CUDA kernel behind zkb_hash_bytecodes against both, and a VM run
whose contracts were populated through zkb_ingest_bytecodes (hash on the GPU, then decommit by that hash)."""
import hashlib

import numpy as np
import pytest

from era_zk_evm_b200 import isa
from era_zk_evm_b200.asm import Imm, Program, R, bytecode_hash
from era_zk_evm_b200.isa import C

VM_ENDED = 1   # ZKB_VM_ENDED (include/zkb.h)
LENGTHS = [0, 1, 2, 3, 4, 5, 7, 8, 63, 64, 65, 127, 1023, 4096, 65535]


def _codes(seed=7):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, size=32 * n, dtype=np.uint8).tobytes() for n in LENGTHS]


def _expected(code: bytes, marker: int) -> int:
    raw = bytes([C.CODE_HASH_VERSION_BYTE, marker]) + (len(code) // 32).to_bytes(2, "big") + hashlib.sha256(code).digest()[4:]
    return int.from_bytes(raw, "big")


@pytest.mark.parametrize("marker", [C.CODE_AT_REST_MARKER, C.YET_CONSTRUCTED_MARKER])
def test_oracle_versioned_hashes_match_hashlib(oracle_mod, marker):
    codes = _codes()
    got = oracle_mod.hash_bytecodes(codes, marker)
    assert got == [_expected(c, marker) for c in codes]
    assert got == [bytecode_hash(c, marker) for c in codes]            # the assembler's host-side helper agrees


def test_oracle_rejects_overlong_bytecode(oracle_mod):
    with pytest.raises(RuntimeError):
        oracle_mod.hash_bytecodes([bytes(32 * 65536)])


def test_oracle_ingest_registers_the_code(oracle_mod):
    from vm_harness import small_config
    b = oracle_mod.OracleBatch(small_config(1))
    codes = _codes()[1:5]
    hashes = b.ingest_bytecodes(codes)
    assert hashes == [_expected(c, C.CODE_AT_REST_MARKER) for c in codes]
    with pytest.raises(Exception):                                       # decommitter.rs:25 asserts the hash is new
        b.ingest_bytecodes(codes[:1])


@pytest.mark.gpu
@pytest.mark.parametrize("marker", [C.CODE_AT_REST_MARKER, C.YET_CONSTRUCTED_MARKER])
def test_gpu_versioned_hashes_match_oracle_and_hashlib(oracle_mod, marker):
    from era_zk_evm_b200 import hash_bytecodes
    codes = _codes()
    got = hash_bytecodes(codes, marker)
    assert got == oracle_mod.hash_bytecodes(codes, marker)
    assert got == [_expected(c, marker) for c in codes]


@pytest.mark.gpu
def test_gpu_many_contracts_one_launch(oracle_mod):
    """an ingest batch of 5 000 contracts of ragged lengths (one thread per contract)"""
    from era_zk_evm_b200 import hash_bytecodes
    rng = np.random.default_rng(11)
    codes = [rng.integers(0, 256, size=32 * int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 200, size=5000)]
    got = hash_bytecodes(codes)
    assert got == oracle_mod.hash_bytecodes(codes)
    assert got[:50] == [_expected(c, 0) for c in codes[:50]]


@pytest.mark.gpu
def test_gpu_rejects_overlong_bytecode():
    from era_zk_evm_b200 import hash_bytecodes
    with pytest.raises(RuntimeError):
        hash_bytecodes([bytes(32 * 65536)])


def _ingest_and_call(batch_cls):
    """bootloader far-calls a contract that was populated through ingest_bytecodes (hash computed by the backend)"""
    from era_zk_evm_b200._binding import FIELD_MEMORY_PAGE_COUNTER, make_frame, storage_entries
    from era_zk_evm_b200.asm import far_call_abi
    from vm_harness import BOOT_ADDRESS, BOOT_PAGE, small_config
    callee = Program()
    callee.add(Imm(0x1234), 0, 5)
    callee.ret(isa.RET_OK, R(0))
    boot = Program()
    boot.const("abi", far_call_abi(1 << 20))
    from era_zk_evm_b200.asm import Code
    boot.add(Code("abi"), 0, 1)
    boot.add(Imm(0x9999), 0, 2)
    boot.far_call(R(1), 2, "fail")
    boot.ret(isa.RET_OK, R(0))
    boot.label("fail")
    boot.ret(isa.RET_PANIC, R(0))
    b = batch_cls(small_config(4))
    h_boot, h_callee = b.ingest_bytecodes([boot.bytecode(), callee.bytecode()])
    assert h_boot == bytecode_hash(boot.bytecode()) and h_callee == bytecode_hash(callee.bytecode())
    b.set_block_properties(h_boot, False)
    b.populate_code(BOOT_PAGE, h_boot)
    b.set_local_field(FIELD_MEMORY_PAGE_COUNTER, 1024)
    b.push_bootloader_context(make_frame(this_address=BOOT_ADDRESS, msg_sender=0, code_address=BOOT_ADDRESS,
                                         base_memory_page=BOOT_PAGE, code_page=BOOT_PAGE, ergs_remaining=1 << 30))
    b.populate_storage(storage_entries([(0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, 0x9999, h_callee)]))
    b.run()
    return b


def test_oracle_runs_code_populated_by_ingest(oracle_mod):
    from era_zk_evm_b200 import records
    b = _ingest_and_call(oracle_mod.OracleBatch)
    assert b.execution_has_ended() and (b.vm_status()[:, 0] == VM_ENDED).all()
    dec = b.read_stream(0, records.STREAM_DECOMMIT)
    assert len(dec) == 1 and dec[0]["is_fresh"] == 1       # the callee was decommitted by the ingested hash


@pytest.mark.gpu
def test_gpu_runs_code_populated_by_ingest(oracle_mod):
    from era_zk_evm_b200 import GpuVmBatch
    from parity_util import compare_batches
    gpu, orc = _ingest_and_call(GpuVmBatch), _ingest_and_call(oracle_mod.OracleBatch)
    assert (gpu.vm_status()[:, 0] == VM_ENDED).all()
    assert compare_batches(gpu, orc, range(4)) == []
