"""Synthetic workloads of BASELINE.json `configs`, written in EraVM assembly via asm.py.

Every workload exposes ``setup(batch, vm_ids)``: it populates ANY object with the `_binding.Batch`
surface (the CUDA batch or, in tests, the CPU oracle) for the given GLOBAL VM ids, so a subset of a
large batch can be re-created elsewhere for differential checks.  All per-VM inputs come from the
counter-based RNG splitmix64(seed ^ vm_index) (default seed 0x5EED0001).
"""
from __future__ import annotations

import numpy as np

from . import isa
from ._binding import (FIELD_ERGS_PER_PUBDATA, FIELD_MEMORY_PAGE_COUNTER, default_config, make_frame,
                       storage_entries)
from .asm import (Code, DStackAbs, Imm, Program, R, StackAbs, bytecode_hash, far_call_abi, ret_abi)
from .isa import C

DEFAULT_SEED = 0x5EED0001
BOOTLOADER_ADDRESS = 0x8001
BOOT_BASE_PAGE = 8
INITIAL_MEMORY_PAGE_COUNTER = 1024
TOKEN_ADDRESS = 0x00C0FFEE00000000000000000000000000ABCDEF
MASK64 = (1 << 64) - 1


# ---------------------------------------------------------------------------------------------
# RNG + batched keccak (numpy; workload generation only)
# ---------------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def vm_random_u64(seed: int, vm_ids: np.ndarray, n: int) -> np.ndarray:
    """[len(vm_ids), n] uint64 stream: splitmix64 chain started from seed ^ vm_index."""
    state = (np.uint64(seed) ^ vm_ids.astype(np.uint64))
    out = np.empty((len(vm_ids), n), dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(n):
            state = (state + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
            z = state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            out[:, i] = z ^ (z >> np.uint64(31))
    return out


def u64_to_be_bytes(words: np.ndarray) -> np.ndarray:
    """[..., k] uint64 (most significant first) -> [..., 8k] big-endian bytes."""
    return np.ascontiguousarray(words.astype(">u8")).view(np.uint8).reshape(*words.shape[:-1], words.shape[-1] * 8)


_KECCAK_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808a, 0x8000000080008000, 0x000000000000808b,
              0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008a, 0x0000000000000088,
              0x0000000080008009, 0x000000008000000a, 0x000000008000808b, 0x800000000000008b, 0x8000000000008089,
              0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800a, 0x800000008000000a,
              0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_KECCAK_ROT = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]


def _rotl(x, n):
    n = n % 64
    if n == 0:
        return x
    return (x << np.uint64(n)) | (x >> np.uint64(64 - n))


def keccak256_batch(msgs: np.ndarray) -> np.ndarray:
    """keccak256 of each row of a [n, L] uint8 array (all rows same length) -> [n, 32] uint8."""
    msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
    n, length = msgs.shape
    rate = 136
    n_blocks = length // rate + 1
    padded = np.zeros((n, n_blocks * rate), dtype=np.uint8)
    padded[:, :length] = msgs
    padded[:, length] ^= 0x01
    padded[:, -1] ^= 0x80
    st = [np.zeros(n, dtype=np.uint64) for _ in range(25)]
    for b in range(n_blocks):
        blk = padded[:, b * rate:(b + 1) * rate].copy().view("<u8")
        for i in range(17):
            st[i] = st[i] ^ blk[:, i]
        for rnd in range(24):
            c = [st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20] for x in range(5)]
            d = [c[(x + 4) % 5] ^ _rotl(c[(x + 1) % 5], 1) for x in range(5)]
            st = [st[i] ^ d[i % 5] for i in range(25)]
            bb = [None] * 25
            for x in range(5):
                for y in range(5):
                    bb[y + 5 * ((2 * x + 3 * y) % 5)] = _rotl(st[x + 5 * y], _KECCAK_ROT[x + 5 * y])
            st = [bb[x + 5 * y] ^ (~bb[(x + 1) % 5 + 5 * y] & bb[(x + 2) % 5 + 5 * y]) for y in range(5) for x in range(5)]
            st[0] = st[0] ^ np.uint64(_KECCAK_RC[rnd])
    out = np.stack(st[:4], axis=1).astype("<u8")
    return np.ascontiguousarray(out).view(np.uint8).reshape(n, 32)


# ---------------------------------------------------------------------------------------------
# system contracts shared by the workloads
# ---------------------------------------------------------------------------------------------
def keccak_system_contract() -> Program:
    """Contract at 0x8010: hashes its calldata slice with the keccak256 precompile, returns the 32-byte digest.
    r1 = calldata fat pointer (kernel mode: metadata not erased, cycle.rs:374-396)."""
    p = Program()
    p.const("mask32", 0xFFFFFFFF)
    p.const("ret32", ret_abi(start=0, length=32))
    p.shift(isa.SHR, Imm(64), 1, 2, swap=True)            # r2 = r1 >> 64        = start | len << 32
    p.binop(isa.AND, Code("mask32"), 1, 3)                # r3 = offset
    p.add(R(2), 3, 4)                                     # r4 = (start + offset) | len << 32
    p.shift(isa.SHR, Imm(32), 1, 5, swap=True)            # r5 = r1 >> 32
    p.binop(isa.AND, Code("mask32"), 5, 5)                # r5 = memory page of the calldata
    p.shift(isa.SHL, Imm(128), 5, 5, swap=True)           # r5 = page << 128     = memory_page_to_read
    p.binop(isa.OR, R(4), 5, 4)                           # r4 = PrecompileCallABI (out word 0, own heap)
    p.precompile(4, 0, 6)                                 # log.precompile
    p.add(Code("ret32"), 0, 7)
    p.ret(isa.RET_OK, R(7))
    return p


def event_writer_contract() -> Program:
    """Contract at 0x800d: calldata = (key, value) pairs; emits them as events (first flag on the first)."""
    p = Program()
    p.ld_ptr(R(1), 2, 1, inc=True)         # r2 = key0 ; r1 += 32
    p.ld_ptr(R(1), 3, 1, inc=True)         # r3 = value0
    p.event(2, 3, first=True)
    p.ld_ptr(R(1), 2, 1, inc=True)
    p.ld_ptr(R(1), 3, 1, inc=True)
    p.event(2, 3, first=False)
    p.ret(isa.RET_OK, R(0))
    return p


class Workload:
    name = "base"
    max_cycles_hint = 2048

    def __init__(self, seed: int = DEFAULT_SEED):
        self.seed = seed
        self.codes = {}      # name -> (hash:int, code:bytes)

    def _add_code(self, name: str, prog: Program):
        code = prog.bytecode()
        self.codes[name] = (bytecode_hash(code), code)

    def config(self, n_vms: int, device: int = 0, witness: bool = True):
        return default_config(n_vms, device=device, max_cycles=self.max_cycles_hint, witness=witness)

    def _common(self, batch, boot_code: str, ergs: int = 1 << 31, heap_bound: int = 0):
        for h, code in self.codes.values():
            batch.load_bytecode(h, code)
        batch.set_block_properties(self.codes[boot_code][0], False)
        batch.populate_code(BOOT_BASE_PAGE, self.codes[boot_code][0])
        batch.set_local_field(FIELD_MEMORY_PAGE_COUNTER, INITIAL_MEMORY_PAGE_COUNTER)
        frame = make_frame(this_address=BOOTLOADER_ADDRESS, msg_sender=0, code_address=BOOTLOADER_ADDRESS,
                           base_memory_page=BOOT_BASE_PAGE, code_page=BOOT_BASE_PAGE, ergs_remaining=ergs,
                           heap_bound=heap_bound, aux_heap_bound=heap_bound)
        batch.push_bootloader_context(frame)

    def setup(self, batch, vm_ids):
        raise NotImplementedError


# ---------------------------------------------------------------------------------------------
# config 1: register-only ADD/SUB/MUL/jump loop, exactly `cycles` cycles (default 1000)
# ---------------------------------------------------------------------------------------------
class AluLoop(Workload):
    name = "alu_loop"

    def __init__(self, cycles: int = 1000, seed: int = DEFAULT_SEED):
        super().__init__(seed)
        assert cycles >= 10 and (cycles - 4) % 6 == 0, "cycles must be 4 + 6k"
        self.cycles = cycles
        self.max_cycles_hint = cycles + 8
        k = (cycles - 4) // 6
        p = Program()
        p.add(Imm(0), 0, 7)                                 # r7 = 0
        p.add(R(1), 0, 8)                                   # prologue filler
        p.add(R(2), 0, 9)
        p.label("loop")
        p.add(R(1), 2, 3)                                   # r3 = r1 + r2
        p.sub(R(3), 1, 4, set_flags=True)                   # r4 = r3 - r1
        p.mul(R(3), 4, 5, 6)                                # r5:r6 = r3 * r4
        p.add(Imm(1), 7, 7)                                 # r7 += 1
        p.sub(Imm(k), 7, 0, set_flags=True, swap=True)      # flags(r7 - k)
        p.jump("loop", cond="lt")
        p.ret(isa.RET_OK, R(0))
        self._add_code("boot", p)

    def config(self, n_vms, device=0, witness=True):
        cfg = super().config(n_vms, device, witness)
        cfg.cap_records[1] = self.max_cycles_hint // 2 + 16
        cfg.cap_records[2] = 16
        cfg.cap_records[4] = 16
        cfg.cap_records[5] = 16
        return cfg

    def setup(self, batch, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        self._common(batch, "boot")
        rnd = vm_random_u64(self.seed, vm_ids, 8)
        batch.set_register(0, u64_to_be_bytes(rnd[:, 0:4]), per_vm=True)
        batch.set_register(1, u64_to_be_bytes(rnd[:, 4:8]), per_vm=True)


# ---------------------------------------------------------------------------------------------
# config 2: ERC-20-shaped transfers (mimic far call -> token -> keccak system contract x2 -> SLOAD/SSTORE
#           -> event writer), T transfers per VM, 1/64 of the VMs start with balance 0 (revert + rollback)
# ---------------------------------------------------------------------------------------------
TRANSFER_SELECTOR = 0xA9059CBB
TRANSFER_TOPIC = int.from_bytes(bytes.fromhex("ddf252ad1be2c89b69c2b068fc378daa952ba7f163c4a11628f55a4df523b3ef"), "big")
CALLDATA_OFF = 1024


def token_contract() -> Program:
    p = Program()
    p.const("keccak_abi", far_call_abi(0xFFFFFFFF, start=0, length=64))
    p.const("event_abi", far_call_abi(0xFFFFFFFF, start=64, length=128))
    p.const("topic", TRANSFER_TOPIC)
    p.const("ret32", ret_abi(start=0, length=32))
    p.const("selector", TRANSFER_SELECTOR)
    # --- decode calldata (r1 = fat pointer; 4-byte selector makes the argument reads unaligned) ---
    p.ld_ptr(R(1), 2)                                       # r2 = first word
    p.shift(isa.SHR, Imm(224), 2, 2, swap=True)             # r2 = selector
    p.sub(Code("selector"), 2, 0, set_flags=True)
    p.jump("revert", cond="ne")
    p.ptr(isa.PTR_ADD, Imm(4), 1, 3, swap=True)             # r3 = ptr + 4
    p.ld_ptr(R(3), 4)                                       # r4 = to
    p.ptr(isa.PTR_ADD, Imm(36), 1, 3, swap=True)
    p.ld_ptr(R(3), 5)                                       # r5 = amount
    p.context(isa.CTX_CALLER, 6)                            # r6 = from
    p.add(R(4), 0, DStackAbs(0))                            # spill: stack[0] = to
    p.add(R(5), 0, DStackAbs(1))                            #        stack[1] = amount
    p.add(R(6), 0, DStackAbs(2))                            #        stack[2] = from
    # --- slot(from) = keccak(from ++ 0) through the keccak system contract ---
    p.st(Imm(0), 6)
    p.st(Imm(32), 0)
    p.add(Imm(C.KECCAK256_PRECOMPILE_ADDRESS), 0, 7)
    p.far_call(Src_reg_code("keccak_abi", p, 8), 7, "revert")
    p.ld_ptr(R(1), 9)                                       # r9 = slot(from)
    p.add(R(9), 0, DStackAbs(3))
    p.sload(9, 10)                                          # r10 = balance[from]
    p.add(StackAbs(1), 0, 5)                                # r5 = amount
    p.sub(R(10), 5, 11, set_flags=True)                     # r11 = balance - amount
    p.jump("revert", cond="lt")
    p.sstore(9, 11)
    # --- slot(to) ---
    p.add(StackAbs(0), 0, 4)
    p.st(Imm(0), 4)
    p.add(Imm(C.KECCAK256_PRECOMPILE_ADDRESS), 0, 7)
    p.far_call(Src_reg_code("keccak_abi", p, 8), 7, "revert")
    p.ld_ptr(R(1), 9)                                       # r9 = slot(to)
    p.sload(9, 10)
    p.add(StackAbs(1), 10, 10)                              # balance[to] += amount
    p.sstore(9, 10)
    # --- Transfer(from, to, amount) through the event writer ---
    p.add(Code("topic"), 0, 2)
    p.st(Imm(64), 2)
    p.add(StackAbs(2), 0, 2)
    p.st(Imm(96), 2)
    p.add(StackAbs(0), 0, 2)
    p.st(Imm(128), 2)
    p.add(StackAbs(1), 0, 2)
    p.st(Imm(160), 2)
    p.add(Imm(C.ADDRESS_EVENT_WRITER), 0, 7)
    p.far_call(Src_reg_code("event_abi", p, 8), 7, "revert")
    # --- return true ---
    p.add(Imm(1), 0, 2)
    p.st(Imm(0), 2)
    p.add(Code("ret32"), 0, 3)
    p.ret(isa.RET_OK, R(3))
    p.label("revert")
    p.ret(isa.RET_REVERT, R(0))
    return p


def Src_reg_code(const_name: str, p: Program, tmp_reg: int):
    """far_call takes its ABI from a register: load the code constant into tmp_reg first."""
    p.add(Code(const_name), 0, tmp_reg)
    return R(tmp_reg)


def erc20_bootloader(n_transfers: int) -> Program:
    """heap image per VM: [0,32) from | [64 + 64 i, +32) to_i | [96 + 64 i, +32) amount_i."""
    p = Program()
    p.const("call_abi", far_call_abi(0xFFFFFFFF, start=CALLDATA_OFF, length=68))
    p.const("token", TOKEN_ADDRESS)
    p.const("selector_word", TRANSFER_SELECTOR << 224)
    p.add(Imm(0), 0, DStackAbs(0))                          # stack[0] = i
    p.add(Imm(0), 0, DStackAbs(1))                          # stack[1] = failures
    p.label("loop")
    p.add(StackAbs(0), 0, 2)                                # r2 = i
    p.shift(isa.SHL, Imm(6), 2, 3, swap=True)               # r3 = 64 i
    p.add(Imm(64), 3, 3)
    p.ld(R(3), 4, 3, inc=True)                              # r4 = to ; r3 += 32
    p.ld(R(3), 5)                                           # r5 = amount
    p.add(Code("selector_word"), 0, 6)
    p.st(Imm(CALLDATA_OFF), 6)
    p.st(Imm(CALLDATA_OFF + 4), 4)                          # unaligned stores
    p.st(Imm(CALLDATA_OFF + 36), 5)
    p.ld(Imm(0), 15)                                        # r15 = from  (mimic-call implicit parameter)
    p.add(Code("token"), 0, 7)
    p.add(Code("call_abi"), 0, 8)
    p.far_call(R(8), 7, "failed", sub=isa.FC_MIMIC)
    p.label("next")
    p.add(StackAbs(0), 0, 2)
    p.add(Imm(1), 2, 2)
    p.add(R(2), 0, DStackAbs(0))
    p.sub(Imm(n_transfers), 2, 0, set_flags=True, swap=True)
    p.jump("loop", cond="lt")
    p.ret(isa.RET_OK, R(0))
    p.label("failed")
    p.add(StackAbs(1), 0, 3)
    p.add(Imm(1), 3, 3)
    p.add(R(3), 0, DStackAbs(1))
    p.jump("next")
    return p


class Erc20(Workload):
    name = "erc20"

    def __init__(self, n_transfers: int = 8, seed: int = DEFAULT_SEED):
        super().__init__(seed)
        self.n_transfers = n_transfers
        self.max_cycles_hint = 160 * n_transfers + 64
        self._add_code("boot", erc20_bootloader(n_transfers))
        self._add_code("token", token_contract())
        self._add_code("keccak", keccak_system_contract())
        self._add_code("event_writer", event_writer_contract())

    def config(self, n_vms, device=0, witness=True):
        cfg = super().config(n_vms, device, witness)
        c = self.max_cycles_hint
        cfg.cap_records[0] = c
        cfg.cap_records[1] = c + c // 2
        cfg.cap_records[2] = 16 * self.n_transfers + 16
        cfg.cap_records[3] = 8 * self.n_transfers + 8
        cfg.cap_records[4] = 12 * self.n_transfers + 8
        cfg.cap_records[5] = 4 * self.n_transfers + 8
        cfg.stack_words = 16
        cfg.heap_bytes = 2048
        # every returned token heap stays reachable from the bootloader frame until it ends (memory.rs:702-712)
        cfg.n_heap_slabs = min(32, self.n_transfers + 6)
        cfg.max_far_depth = 5
        cfg.max_depth = 8
        cfg.storage_slots = 64
        cfg.journal_entries = 4 * self.n_transfers + 8
        return cfg

    def inputs(self, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        t = self.n_transfers
        rnd = vm_random_u64(self.seed, vm_ids, 3 + 5 * t)
        n = len(vm_ids)
        heap = np.zeros((n, 64 + 64 * t), dtype=np.uint8)
        frm = rnd[:, 0:3].copy()
        frm[:, 0] &= np.uint64(0xFFFFFFFF)                 # 160-bit address
        frm[:, 0] |= np.uint64(0x10000000)                 # never a kernel address
        heap[:, 8:32] = u64_to_be_bytes(frm)
        for i in range(t):
            to = rnd[:, 3 + 5 * i: 6 + 5 * i].copy()
            to[:, 0] &= np.uint64(0xFFFFFFFF)
            to[:, 0] |= np.uint64(0x20000000)
            heap[:, 64 + 64 * i + 8: 64 + 64 * i + 32] = u64_to_be_bytes(to)
            amount = rnd[:, 6 + 5 * i: 8 + 5 * i]            # uniform in [0, 2^128)
            heap[:, 96 + 64 * i + 16: 96 + 64 * i + 32] = u64_to_be_bytes(amount)
        # balance slot of `from`: keccak(from ++ slot 0)
        pre = np.zeros((n, 64), dtype=np.uint8)
        pre[:, 0:32] = heap[:, 0:32]
        slot = keccak256_batch(pre)
        broke = (vm_ids % np.uint64(64)) == np.uint64(63)
        return heap, slot, broke

    def setup(self, batch, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        self._common(batch, "boot", heap_bound=4096)
        batch.set_local_field(FIELD_ERGS_PER_PUBDATA, 1)
        code_entries = storage_entries([
            (0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, TOKEN_ADDRESS, self.codes["token"][0]),
            (0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, C.KECCAK256_PRECOMPILE_ADDRESS, self.codes["keccak"][0]),
            (0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, C.ADDRESS_EVENT_WRITER, self.codes["event_writer"][0]),
        ])
        batch.populate_storage(code_entries)
        heap, slot, broke = self.inputs(vm_ids)
        batch.populate_heap(heap, per_vm=True)
        from ._binding import STORAGE_INIT_DTYPE
        ent = np.zeros(len(vm_ids), dtype=STORAGE_INIT_DTYPE)
        ent["shard_id"] = 0
        ent["address"] = np.frombuffer(TOKEN_ADDRESS.to_bytes(20, "big"), dtype=np.uint8)
        ent["key_be"] = slot
        bal = np.zeros((len(vm_ids), 32), dtype=np.uint8)
        bal[~broke, 15] = 2                                  # 2^129
        ent["value_be"] = bal
        batch.populate_storage(ent, per_vm=True)


# ---------------------------------------------------------------------------------------------
# config 3: keccak-heavy — K precompile calls over a 4 KiB preimage (half of the VMs unaligned by 31 bytes)
# ---------------------------------------------------------------------------------------------
class KeccakHeavy(Workload):
    name = "keccak"

    def __init__(self, n_calls: int = 8, preimage_bytes: int = 4096, seed: int = DEFAULT_SEED):
        super().__init__(seed)
        self.n_calls, self.preimage_bytes = n_calls, preimage_bytes
        self.max_cycles_hint = 24 * n_calls + 32
        p = Program()
        p.const("abi_aligned", far_call_abi(0xFFFFFFFF, start=0, length=preimage_bytes))
        p.const("abi_unaligned", far_call_abi(0xFFFFFFFF, start=31, length=preimage_bytes))
        p.add(Imm(0), 0, DStackAbs(0))
        p.add(R(1), 0, DStackAbs(1))                       # r1 (set by host) = 1 for unaligned VMs
        p.label("loop")
        p.add(Code("abi_aligned"), 0, 8)
        p.sub(StackAbs(1), 0, 0, set_flags=True)
        p.add(Code("abi_unaligned"), 0, 8, cond="ne")
        p.add(Imm(C.KECCAK256_PRECOMPILE_ADDRESS), 0, 7)
        p.far_call(R(8), 7, "fail")
        p.ld_ptr(R(1), 2)                                  # digest
        p.st(Imm(preimage_bytes + 64), 2)                  # keep it (outside the preimage)
        p.add(StackAbs(0), 0, 3)
        p.add(Imm(1), 3, 3)
        p.add(R(3), 0, DStackAbs(0))
        p.sub(Imm(n_calls), 3, 0, set_flags=True, swap=True)
        p.jump("loop", cond="lt")
        p.ret(isa.RET_OK, R(0))
        p.label("fail")
        p.ret(isa.RET_PANIC, R(0))
        self._add_code("boot", p)
        self._add_code("keccak", keccak_system_contract())

    def config(self, n_vms, device=0, witness=True):
        cfg = super().config(n_vms, device, witness)
        words = self.preimage_bytes // 32 + 3
        cfg.cap_records[0] = self.max_cycles_hint
        cfg.cap_records[1] = self.n_calls * (words + 16) + 64
        cfg.cap_records[2] = 3 * self.n_calls + 8
        cfg.cap_records[3] = self.n_calls + 4
        cfg.cap_records[4] = 2 * self.n_calls + 8
        cfg.cap_records[5] = 8
        cfg.stack_words = 8
        cfg.heap_bytes = ((self.preimage_bytes + 31 + 96 + 31) // 32 + 1) * 32
        cfg.n_heap_slabs = 6
        cfg.max_far_depth = 3
        cfg.max_depth = 4
        cfg.storage_slots = 32
        cfg.journal_entries = 8
        return cfg

    def setup(self, batch, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        self._common(batch, "boot", heap_bound=self.preimage_bytes + 128)
        batch.populate_storage(storage_entries([
            (0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, C.KECCAK256_PRECOMPILE_ADDRESS, self.codes["keccak"][0])]))
        n_words64 = (self.preimage_bytes + 32) // 8
        rnd = vm_random_u64(self.seed, vm_ids, n_words64)
        heap = u64_to_be_bytes(rnd)[:, : self.preimage_bytes + 31]
        batch.populate_heap(np.ascontiguousarray(heap), per_vm=True)
        unaligned = np.zeros((len(vm_ids), 32), dtype=np.uint8)
        unaligned[:, 31] = (vm_ids & np.uint64(1)).astype(np.uint8)
        batch.set_register(0, unaligned, per_vm=True)


# ---------------------------------------------------------------------------------------------
# config 4: SLOAD/SSTORE-heavy with panicking near calls (rollback path, storage.rs:156-180)
# ---------------------------------------------------------------------------------------------
class StorageHeavy(Workload):
    name = "storage"

    def __init__(self, n_iters: int = 64, n_keys: int = 16, seed: int = DEFAULT_SEED):
        super().__init__(seed)
        self.n_iters, self.n_keys = n_iters, n_keys
        self.max_cycles_hint = 14 * n_iters + 32
        p = Program()
        p.add(Imm(0), 0, 1)                                 # r1 = i ; r2 = key base (host)
        p.label("loop")
        p.binop(isa.AND, Imm(n_keys - 1), 1, 3)
        p.add(R(3), 2, 4)                                   # r4 = key
        p.binop(isa.AND, Imm(7), 1, 5)
        p.sub(Imm(7), 5, 0, set_flags=True, swap=True)      # (i & 7) == 7 ?
        p.jump("wrapped", cond="eq")
        p.sload(4, 6)
        p.add(R(6), 1, 6)
        p.sstore(4, 6)
        p.jump("next")
        p.label("wrapped")
        p.near_call(0, "body", "next")
        p.label("next")
        p.add(Imm(1), 1, 1)
        p.sub(Imm(n_iters), 1, 0, set_flags=True, swap=True)
        p.jump("loop", cond="lt")
        p.ret(isa.RET_OK, R(0))
        p.label("body")
        p.sload(4, 6)
        p.add(R(6), 1, 6)
        p.sstore(4, 6)
        p.ret(isa.RET_PANIC, R(0))
        self._add_code("boot", p)

    def config(self, n_vms, device=0, witness=True):
        cfg = super().config(n_vms, device, witness)
        cfg.cap_records[0] = self.max_cycles_hint
        cfg.cap_records[1] = self.max_cycles_hint // 2 + 16
        cfg.cap_records[2] = 2 * self.n_iters + 8
        cfg.cap_records[3] = 4
        cfg.cap_records[4] = self.n_iters // 2 + 8
        cfg.cap_records[5] = self.n_iters + 8
        cfg.stack_words = 8
        cfg.heap_bytes = 256
        cfg.n_heap_slabs = 4
        cfg.max_far_depth = 2
        cfg.max_depth = 4
        cfg.storage_slots = 64
        cfg.journal_entries = 2 * self.n_iters + 8
        return cfg

    def setup(self, batch, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        self._common(batch, "boot")
        batch.set_local_field(FIELD_ERGS_PER_PUBDATA, 2)
        from ._binding import STORAGE_INIT_DTYPE
        n, k = len(vm_ids), self.n_keys
        rnd = vm_random_u64(self.seed, vm_ids, 4 + 4 * k)
        base = rnd[:, 0:4].copy()
        base[:, 3] &= np.uint64(0xFFFFFFFFFFFFFF00)         # low byte free so base + j never carries
        batch.set_register(1, u64_to_be_bytes(base), per_vm=True)
        ent = np.zeros((n, k), dtype=STORAGE_INIT_DTYPE)
        ent["address"] = np.frombuffer(BOOTLOADER_ADDRESS.to_bytes(20, "big"), dtype=np.uint8)
        for j in range(k):
            key = base.copy()
            key[:, 3] += np.uint64(j)
            ent["key_be"][:, j, :] = u64_to_be_bytes(key)
            ent["value_be"][:, j, :] = u64_to_be_bytes(rnd[:, 4 + 4 * j: 8 + 4 * j])
        batch.populate_storage(ent.reshape(-1), per_vm=True)


WORKLOADS = {"alu_loop": AluLoop, "erc20": Erc20, "keccak": KeccakHeavy, "storage": StorageHeavy}
