"""Synthetic workloads of BASELINE.json `configs`, written in EraVM assembly via asm.py.

Every workload exposes ``setup(batch, vm_ids)``: it populates ANY object with the `_binding.Batch`
surface (the CUDA batch or, in tests, the CPU oracle) for the given GLOBAL VM ids, so a subset of a
large batch can be re-created elsewhere for differential checks.  All per-VM inputs come from the
counter-based RNG splitmix64(seed ^ vm_index) (default seed 0x5EED0001).
"""
from __future__ import annotations

import os

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
        self._prepared = {}

    def prepared(self, vm_ids):
        """host-side input arrays of a VM range, generated once (workload GENERATION is not part of any timed region)"""
        key = (int(vm_ids[0]) if len(vm_ids) else 0, len(vm_ids))
        if key not in self._prepared:
            self._prepared[key] = self.inputs(np.asarray(vm_ids, dtype=np.uint64))
        return self._prepared[key]

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


class DivLoop(AluLoop):
    """DIV-heavy variant of the register loop (U256 division): every iteration divides a 256-bit by a
    ~128-bit value (quotient and remainder ~128 bits: four limb steps of the Knuth-D divider) and feeds both back."""
    name = "div_loop"

    def __init__(self, cycles: int = 1000, seed: int = DEFAULT_SEED):
        Workload.__init__(self, seed)
        self.cycles = cycles
        self.max_cycles_hint = cycles + 16
        k = max(1, (cycles - 4) // 7)
        p = Program()
        p.add(Imm(0), 0, 7)
        p.add(R(1), 0, 8)
        p.add(R(2), 0, 9)
        p.label("loop")
        p.div(R(1), 2, 3, 4)                                # r3 = r1 / r2, r4 = r1 % r2
        p.mul(R(3), 2, 5, 6)                                # r5:r6 = q * b
        p.add(R(5), 4, 1)                                   # r1 = q * b + rem (= the old r1)  ...
        p.binop(isa.XOR, R(4), 1, 1)                        # ... stirred with the remainder
        p.sub(Imm(k), 7, 0, set_flags=True, swap=True)
        p.add(Imm(1), 7, 7)
        p.jump("loop", cond="lt")
        p.ret(isa.RET_OK, R(0))
        self._add_code("boot", p)

    def setup(self, batch, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        self._common(batch, "boot")
        rnd = vm_random_u64(self.seed, vm_ids, 8)
        rnd[:, 4:6] = 0                                     # divisor: 128 bits (the two most significant words cleared)
        rnd[:, 6] |= np.uint64(1) << np.uint64(63)
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
        broke_mod = int(os.environ.get("ZKB_ERC20_BROKE_MOD", "64"))       # experiment knob; 0 = every VM funded
        broke = (vm_ids % np.uint64(broke_mod)) == np.uint64(broke_mod - 1) if broke_mod else np.zeros(len(vm_ids), dtype=bool)
        from ._binding import STORAGE_INIT_DTYPE
        ent = np.zeros(len(vm_ids), dtype=STORAGE_INIT_DTYPE)
        ent["shard_id"] = 0
        ent["address"] = np.frombuffer(TOKEN_ADDRESS.to_bytes(20, "big"), dtype=np.uint8)
        ent["key_be"] = slot
        bal = np.zeros((len(vm_ids), 32), dtype=np.uint8)
        bal[~broke, 15] = 16                                 # 2^132 > T * 2^128: every transfer of a funded VM succeeds
        ent["value_be"] = bal
        return heap, ent

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
        heap, ent = self.prepared(vm_ids)
        batch.populate_heap(heap, per_vm=True)
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
        cfg.n_heap_slabs = min(32, self.n_calls + 4)      # every returned digest page stays reachable until the end (memory.rs:702-712)
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


WORKLOADS = {"alu_loop": AluLoop, "div_loop": DivLoop, "erc20": Erc20, "keccak": KeccakHeavy, "storage": StorageHeavy}


# ---------------------------------------------------------------------------------------------
# config 5: mixed-opcode synthetic block — a pool of RNG-generated programs over all 15 opcode families
# (ALU 55 %, stack/UMA 25 %, jump/near-call/ret 10 %, log 7 %, far-call 2 %, precompile 1 %), ~512 cycles per VM.
# Also the fuzzing corpus of the parity tests: every exception path the generator can reach (panicking near calls,
# static violations, out-of-ergs frames, pointer-op panics, invalid opcodes, reverting callees) is exercised.
# ---------------------------------------------------------------------------------------------
class _Rng:
    """splitmix64 stream (scalar) so program generation is reproducible and independent of numpy's generators"""

    def __init__(self, seed: int):
        self.s = seed & MASK64

    def u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.u64() % n

    def chance(self, p: float) -> bool:
        return (self.u64() >> 11) / float(1 << 53) < p

    def pick(self, seq):
        return seq[self.below(len(seq))]

    def u256(self) -> int:
        bits = self.pick([8, 16, 32, 64, 128, 200, 256])
        v = 0
        for _ in range(4):
            v = (v << 64) | self.u64()
        return v & ((1 << bits) - 1)


MIXED_CALLEE_BASE = 0x00A11CE000000000000000000000000000000100
_CONDS = ["always", "always", "always", "gt", "lt", "eq", "ge", "le", "ne", "gtlt"]
_GP = list(range(1, 11))        # general-purpose registers r1..r10; r11..r15 are reserved for control / pointers


class _Gen:
    """statement generator shared by the bootloader programs, their near-call subroutines and the callee contracts"""

    def __init__(self, rng: _Rng, p: Program, kernel: bool):
        self.r, self.p, self.kernel = rng, p, kernel
        self.n_const = 0
        self.n_label = 0

    def const(self, v: int) -> str:
        name = f"k{self.n_const}"
        self.n_const += 1
        self.p.const(name, v)
        return name

    def label(self) -> str:
        self.n_label += 1
        return f"L{self.n_label}"

    def src(self):
        k = self.r.below(10)
        if k < 5:
            return R(self.r.pick(_GP))
        if k < 7:
            return Imm(self.r.below(1 << 16) if self.r.chance(0.5) else self.r.below(40))
        if k < 9:
            return StackAbs(self.r.below(8))
        return Code(self.const(self.r.u256()))

    def dst(self):
        return DStackAbs(self.r.below(8)) if self.r.chance(0.2) else self.r.pick(_GP)

    def alu(self, cond=None):
        r, p = self.r, self.p
        cond = cond or r.pick(_CONDS)
        k = r.below(20)
        sf = r.chance(0.4)
        if k < 5:
            p.add(self.src(), r.pick(_GP), self.dst(), set_flags=sf, cond=cond)
        elif k < 9:
            p.sub(self.src(), r.pick(_GP), self.dst(), set_flags=sf, swap=r.chance(0.5), cond=cond)
        elif k < 12:
            p.mul(self.src(), r.pick(_GP), self.dst(), r.pick(_GP), set_flags=sf, cond=cond)
        elif k < 14:
            p.div(self.src(), r.pick(_GP), self.dst(), r.pick(_GP), set_flags=sf, swap=r.chance(0.5), cond=cond)
        elif k < 17:
            p.shift(r.below(4), self.src(), r.pick(_GP), self.dst(), set_flags=sf, swap=r.chance(0.5), cond=cond)
        else:
            p.binop(r.below(3), self.src(), r.pick(_GP), self.dst(), set_flags=sf, cond=cond)

    def memory(self):
        r, p = self.r, self.p
        k = r.below(10)
        if k < 2:       # balanced push / pop above the absolute window
            p.add(R(r.pick(_GP)), 0, DStackPush(1))
            p.add(StackPop(1), r.pick(_GP), r.pick(_GP))
        elif k < 3:     # sp-relative read of the slot just below sp, context.sp
            p.add(StackRel(1), 0, r.pick(_GP))
            p.context(isa.CTX_SP, r.pick(_GP))
        else:
            off = r.below(32) * 32 + (r.below(32) if r.chance(0.3) else 0)
            aux = r.chance(0.25)
            use_reg = r.chance(0.4)
            if use_reg:   # r11 = bounded offset derived from data
                p.binop(isa.AND, Imm(0x3FF), r.pick(_GP), 11)
            addr = R(11) if use_reg else Imm(off)
            if r.chance(0.5):
                (p.st_aux if aux else p.st)(addr, r.pick(_GP), 12 if r.chance(0.3) else 0, inc=r.chance(0.3))
            else:
                (p.ld_aux if aux else p.ld)(addr, r.pick(_GP), 12, inc=r.chance(0.3))

    def storage(self, allow_write=True):
        r, p = self.r, self.p
        p.binop(isa.AND, Imm(0xF), r.pick(_GP), 11)        # 16 hot keys per contract
        k = r.below(10)
        if k < 5 or not allow_write:
            p.sload(11, r.pick(_GP))
        elif k < 9:
            p.sstore(11, r.pick(_GP))
        elif self.kernel:
            (p.to_l1 if r.chance(0.3) else p.event)(11, r.pick(_GP), first=r.chance(0.5))
        else:
            p.sload(11, r.pick(_GP))

    def context(self):
        r, p = self.r, self.p
        sub = r.pick([isa.CTX_THIS, isa.CTX_CALLER, isa.CTX_CODE_ADDRESS, isa.CTX_META, isa.CTX_ERGS_LEFT, isa.CTX_SP,
                      isa.CTX_GET_U128] + ([isa.CTX_SET_U128, isa.CTX_INC_TX] if self.kernel else []))
        if sub in (isa.CTX_SET_U128,):
            p.context(sub, 0, r.pick(_GP))
        elif sub == isa.CTX_INC_TX:
            p.context(sub)
        else:
            p.context(sub, r.pick(_GP))

    def safe_statement(self):
        k = self.r.below(100)
        if k < 62:
            self.alu()
        elif k < 90:
            self.memory()
        elif k < 97:
            self.storage()
        else:
            self.context()

    def risky_statement(self):
        """may raise a pending exception / panic: only used inside near-call subroutines and callees"""
        r, p = self.r, self.p
        k = r.below(10)
        if k < 3:       # pointer arithmetic on a register that may or may not hold a pointer
            p.ptr(r.below(4), Imm(r.below(64)) if r.chance(0.6) else R(r.pick(_GP)), r.pick([1, 12, r.pick(_GP)]),
                  r.pick(_GP), swap=True)
        elif k < 5:     # fat-pointer read through r1 / r12 (non-pointer => panic)
            p.ld_ptr(R(r.pick([1, 1, 12])), r.pick(_GP), 12, inc=r.chance(0.5))
        elif k < 6:     # far heap offset: growth cost exceeds the frame's ergs
            p.add(Code(self.const(r.pick([1 << 20, (1 << 32) - 33, (1 << 32) - 32, 1 << 40]))), 0, 11)
            p.ld(R(11), r.pick(_GP))
        elif k < 7:     # undecodable instruction (variant index beyond the valid range)
            p.ins.append((isa.NOP, 0, R(0), 0, DR_ZERO, 0, 0, "always", None, None))
            p.raw_patch = getattr(p, "raw_patch", {})
            p.raw_patch[len(p.ins) - 1] = isa.N_VALID_VARIANTS + r.below(2048 - isa.N_VALID_VARIANTS)
        elif k < 8 and not self.kernel:   # kernel-only opcode in user mode
            p.event(r.pick(_GP), r.pick(_GP))
        else:
            self.storage()


from .asm import DR as _DR, DStackPush, StackPop, StackRel  # noqa: E402
DR_ZERO = _DR(0)


def _patched_words(p: Program) -> bytes:
    """Program.bytecode() with the undecodable-instruction patches of _Gen.risky_statement applied"""
    patch = getattr(p, "raw_patch", {})
    if not patch:
        return p.bytecode()
    ins = p.encode()
    for i, variant in patch.items():
        ins[i] = (ins[i] & ~((1 << isa.VARIANT_BITS) - 1)) | variant
    ins += [0] * (-len(ins) % 4)
    words = [(ins[i] << 192) | (ins[i + 1] << 128) | (ins[i + 2] << 64) | ins[i + 3] for i in range(0, len(ins), 4)]
    words += [v for _, v in p.consts]
    if len(words) % 2 == 0:
        words.append(0)
    return b"".join(w.to_bytes(32, "big") for w in words)


def mixed_callee(rng: _Rng, kernel: bool) -> Program:
    """a callee contract: reads calldata through r1, random statements, returns / reverts / panics"""
    p = Program()
    g = _Gen(rng, p, kernel)
    p.ld_ptr(R(1), 2, 12, inc=True)              # r2 = calldata word 0 ; r12 = advanced pointer
    p.ld_ptr(R(12), 3)
    for _ in range(6 + rng.below(10)):
        g.safe_statement() if rng.chance(0.8) else g.risky_statement()
    p.st(Imm(0), rng.pick(_GP))
    p.st(Imm(32), rng.pick(_GP))
    k = rng.below(10)
    abi = g.const(ret_abi(start=rng.below(2) * 16, length=rng.pick([0, 32, 64, 40])))
    p.add(Code(abi), 0, 13)
    if k < 7:
        p.ret(isa.RET_OK, R(13))
    elif k < 9:
        p.ret(isa.RET_REVERT, R(13))
    else:
        p.ret(isa.RET_PANIC, R(0))
    return p


def mixed_bootloader(rng: _Rng, n_body: int, n_iters: int, callee_addresses) -> Program:
    p = Program()
    g = _Gen(rng, p, kernel=True)
    subs = []                                     # (label, generator thunk) emitted after the main loop
    p.nop(R(0), DStackPush(16))                   # sp = 16: slots 0..7 absolute window, 8..15 control
    for i in range(8):
        p.add(R(1 + i % 6), 0, DStackAbs(i))
    p.add(Imm(0), 0, DStackAbs(8))                # loop counter
    p.label("loop")
    # every far-call return keeps its returndata page alive until the bootloader frame ends (memory.rs:702-712): the
    # bounded slab pool of the device build (<= 32 per VM) caps the far calls of one VM
    far_budget = max(0, 24 // n_iters)
    for _ in range(n_body):
        k = rng.below(100)
        if k >= 96 and far_budget == 0:
            k = 0
        if k >= 96:
            far_budget -= 1
        if k < 88:
            g.safe_statement()
        elif k < 91:                              # forward conditional jump over 1..3 statements
            lab = g.label()
            p.jump(lab, cond=rng.pick(_CONDS[3:]))
            for _ in range(1 + rng.below(3)):
                g.alu(cond="always")
            p.label(lab)
        elif k < 96:                              # near call into a (possibly panicking) subroutine
            sub, cont = g.label(), g.label()
            ergs_limited = rng.chance(0.3)
            if ergs_limited:
                p.add(Imm(rng.pick([30, 60, 200, 1000])), 0, 13)
            p.near_call(13 if ergs_limited else 0, sub, cont)
            p.label(cont)
            subs.append((sub, rng.below(1 << 30)))
        elif k < 98:                              # far call (normal / delegate / mimic, sometimes static)
            addr = rng.pick(callee_addresses + [0xDEAD0000 + rng.below(4)])       # unknown address => default AA
            length = rng.pick([0, 32, 64, 100])
            fwd_aux = rng.chance(0.2)
            abi = g.const(far_call_abi(rng.pick([0xFFFFFFFF, 100000, 3000]), start=rng.below(3) * 32, length=length,
                                       fwd=C.FWD_USE_AUX_HEAP if fwd_aux else C.FWD_USE_HEAP))
            p.add(Code(abi), 0, 13)
            p.add(Code(g.const(addr)), 0, 14)
            cont = g.label()
            p.far_call(R(13), 14, cont, sub=rng.pick([isa.FC_NORMAL, isa.FC_NORMAL, isa.FC_DELEGATE, isa.FC_MIMIC]),
                       static=rng.chance(0.25))
            p.label(cont)
            p.ld_ptr(R(1), rng.pick(_GP), 12, inc=True, cond="always")   # returndata (r1 is a pointer after far ret)
        else:                                     # keccak through the system contract
            abi = g.const(far_call_abi(0xFFFFFFFF, start=rng.below(64), length=rng.pick([0, 1, 64, 135, 136, 200])))
            p.add(Code(abi), 0, 13)
            p.add(Imm(C.KECCAK256_PRECOMPILE_ADDRESS), 0, 14)
            cont = g.label()
            p.far_call(R(13), 14, cont)
            p.label(cont)
            p.ld_ptr(R(1), rng.pick(_GP))
    p.add(StackAbs(8), 0, 13)
    p.add(Imm(1), 13, 13)
    p.add(R(13), 0, DStackAbs(8))
    p.sub(Imm(n_iters), 13, 0, set_flags=True, swap=True)
    p.jump("loop", cond="lt")
    p.ret(isa.RET_OK, R(0))
    for label, sub_seed in subs:
        sr = _Rng(sub_seed)
        sg = _Gen(sr, p, kernel=True)
        sg.n_const, sg.n_label = g.n_const + 1000 * (1 + len(p.labels)), g.n_label + 1000 * (1 + len(p.labels))
        p.label(label)
        for _ in range(2 + sr.below(6)):
            sg.safe_statement() if sr.chance(0.7) else sg.risky_statement()
        k = sr.below(10)
        if k < 6:
            p.ret(isa.RET_OK, R(0))
        elif k < 8:
            p.ret(isa.RET_REVERT, R(0))
        else:
            p.ret(isa.RET_PANIC, R(0))
    return p


class Mixed(Workload):
    name = "mixed"

    def __init__(self, n_programs: int = 64, target_cycles: int = 512, vms_per_program: int = 32, seed: int = DEFAULT_SEED):
        super().__init__(seed)
        self.n_programs, self.vms_per_program = n_programs, vms_per_program
        self.max_cycles_hint = 2 * target_cycles + 256
        rng = _Rng(seed ^ 0xC0FFEE)
        self.callees = []
        for i in range(4):
            addr = (MIXED_CALLEE_BASE + i) if i < 3 else 0x8123       # one callee lives in kernel space
            prog = mixed_callee(_Rng(rng.u64()), kernel=addr < (1 << 16))
            code = _patched_words(prog)
            self.codes[f"callee{i}"] = (bytecode_hash(code), code)
            self.callees.append(addr)
        self._add_code("keccak", keccak_system_contract())
        n_body = 24
        n_iters = max(1, target_cycles // (n_body * 2))
        self.boot_names = []
        for i in range(n_programs):
            prog = mixed_bootloader(_Rng(rng.u64()), n_body, n_iters, self.callees)
            code = _patched_words(prog)
            self.codes[f"boot{i}"] = (bytecode_hash(code), code)
            self.boot_names.append(f"boot{i}")

    def config(self, n_vms, device=0, witness=True):
        cfg = super().config(n_vms, device, witness)
        c = self.max_cycles_hint
        cfg.cap_records[0] = c
        cfg.cap_records[1] = 2 * c
        cfg.cap_records[2] = c // 2
        cfg.cap_records[3] = c // 8 + 8
        cfg.cap_records[4] = c // 2
        cfg.cap_records[5] = c // 4
        cfg.stack_words = 64
        cfg.heap_bytes = 2048
        cfg.n_heap_slabs = 32
        cfg.max_far_depth = 4
        cfg.max_depth = 12
        cfg.storage_slots = 128
        cfg.journal_entries = c // 2
        return cfg

    def setup(self, batch, vm_ids):
        vm_ids = np.asarray(vm_ids, dtype=np.uint64)
        n = len(vm_ids)
        for h, code in self.codes.values():
            batch.load_bytecode(h, code)
        batch.set_block_properties(self.codes["callee0"][0], False)      # default AA = callee0's code
        prog_of = ((vm_ids // np.uint64(self.vms_per_program)) % np.uint64(self.n_programs)).astype(np.int64)
        start = 0
        while start < n:                                                 # one populate_code per run of equal programs
            end = start + 1
            while end < n and prog_of[end] == prog_of[start]:
                end += 1
            batch.populate_code(BOOT_BASE_PAGE, self.codes[self.boot_names[prog_of[start]]][0], vm_lo=start, vm_hi=end)
            start = end
        batch.set_local_field(FIELD_MEMORY_PAGE_COUNTER, INITIAL_MEMORY_PAGE_COUNTER)
        batch.set_local_field(FIELD_ERGS_PER_PUBDATA, 1)
        frame = make_frame(this_address=BOOTLOADER_ADDRESS, msg_sender=0, code_address=BOOTLOADER_ADDRESS,
                           base_memory_page=BOOT_BASE_PAGE, code_page=BOOT_BASE_PAGE, ergs_remaining=1 << 31,
                           heap_bound=2048, aux_heap_bound=2048)
        batch.push_bootloader_context(frame)
        entries = [(0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, addr, self.codes[f"callee{i}"][0]) for i, addr in enumerate(self.callees)]
        entries.append((0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, C.KECCAK256_PRECOMPILE_ADDRESS, self.codes["keccak"][0]))
        batch.populate_storage(storage_entries(entries))
        rnd = vm_random_u64(self.seed, vm_ids, 6 * 4 + 32)
        for r in range(6):
            words = rnd[:, 4 * r: 4 * r + 4].copy()
            if r % 2 == 1:
                words[:, 0:3] = 0                                           # small values too (shift amounts, divisors)
            batch.set_register(r, u64_to_be_bytes(words), per_vm=True)
        batch.populate_heap(np.ascontiguousarray(u64_to_be_bytes(rnd[:, 24:56])), per_vm=True)


WORKLOADS["mixed"] = Mixed


class MixedShuffled(Mixed):
    """the same corpus with every VM of a warp on a DIFFERENT program (a block of unrelated
    transactions): consecutive VMs cycle through the program pool, so the four octets of a warp never converge"""
    name = "mixed_shuffled"

    def __init__(self, **kw):
        kw.setdefault("vms_per_program", 1)
        super().__init__(**kw)


WORKLOADS["mixed_shuffled"] = MixedShuffled
