"""Runs small hand-written EraVM programs on any backend with the `_binding.Batch` surface
(the CPU oracle in the CPU suite, the CUDA batch in the `-m gpu` suite)."""
import numpy as np

from era_zk_evm_b200 import isa, records
from era_zk_evm_b200._binding import (FIELD_ERGS_PER_PUBDATA, FIELD_MEMORY_PAGE_COUNTER, default_config, make_frame,
                                      storage_entries)
from era_zk_evm_b200.asm import Program, bytecode_hash
from era_zk_evm_b200.isa import C

BOOT_ADDRESS = 0x8001
BOOT_PAGE = 8


def be32_rows(values):
    """list of ints -> [n, 32] uint8 big-endian"""
    return np.frombuffer(b"".join(int(v).to_bytes(32, "big") for v in values), dtype=np.uint8).reshape(len(values), 32).copy()


def small_config(n_vms, max_cycles=256, **over):
    cfg = default_config(n_vms, max_cycles=max_cycles)
    cfg.stack_words = 64
    cfg.heap_bytes = 4096
    cfg.n_heap_slabs = 10
    cfg.max_far_depth = 5
    cfg.max_depth = 12
    cfg.storage_slots = 64
    cfg.journal_entries = 64
    for k, v in over.items():
        if k == "warm_write_refund_bytes":      # ZkbConfig.reserved[1]: the refund-aware storage oracle (row f-3)
            cfg.reserved[1] = v
        else:
            setattr(cfg, k, v)
    return cfg


def launch(batch_cls, prog: Program, n_vms=1, *, regs=None, ptr_regs=(), contracts=None, storage=(), heap=None,
           ergs=1 << 30, this_address=BOOT_ADDRESS, ergs_per_pubdata=0, is_static=False, heap_bound=0, run=True,
           max_cycles=256, default_aa=None, cfg_over=None):
    """regs: {reg index 1..15: int | list of per-VM ints}; contracts: {address: Program} deployed through the
    deployer's storage (far_call.rs:131-145); storage: [(shard, address, key, value)]."""
    cfg = small_config(n_vms, max_cycles=max_cycles, **(cfg_over or {}))
    b = batch_cls(cfg)
    code = prog.bytecode()
    h = bytecode_hash(code)
    b.load_bytecode(h, code)
    entries = list(storage)
    for addr, cprog in (contracts or {}).items():
        ccode = cprog.bytecode()
        ch = bytecode_hash(ccode)
        b.load_bytecode(ch, ccode)
        entries.append((0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, addr, ch))
    b.set_block_properties(default_aa if default_aa is not None else h, False)
    b.populate_code(BOOT_PAGE, h)
    b.set_local_field(FIELD_MEMORY_PAGE_COUNTER, 1024)
    if ergs_per_pubdata:
        b.set_local_field(FIELD_ERGS_PER_PUBDATA, ergs_per_pubdata)
    b.push_bootloader_context(make_frame(this_address=this_address, msg_sender=0, code_address=this_address,
                                         base_memory_page=BOOT_PAGE, code_page=BOOT_PAGE, ergs_remaining=ergs,
                                         is_static=is_static, heap_bound=heap_bound, aux_heap_bound=heap_bound))
    if entries:
        b.populate_storage(storage_entries(entries))
    if heap is not None:
        b.populate_heap(heap)
    for r, v in (regs or {}).items():
        if isinstance(v, (list, tuple, np.ndarray)):
            b.set_register(r - 1, be32_rows(v), is_pointer=r in ptr_regs, per_vm=True)
        else:
            b.set_register(r - 1, int(v), is_pointer=r in ptr_regs)
    if run:
        b.run()
    return b


def rows(b, vm=0):
    return b.read_stream(vm, records.STREAM_ROWS)


def val(limbs) -> int:
    return records.limbs_to_int(limbs)


def family_of(row) -> str:
    return isa.FAMILY_NAMES[isa.VARIANTS[int(row["masked_variant"])].family]
