"""Tiny EraVM assembler over the ISA table in ``isa.py`` (test / workload infrastructure).

Encoding (EncodingModeProduction; reference decodes with ``E::integer_representaiton_from_u256`` at
``src/vm_state/cycle.rs:94`` — sub_pc 0 is the most significant 8 bytes of the big-endian code word):

  bits [0,11) variant index | [13,16) condition | [16,20) src0 reg | [20,24) src1 reg |
  [24,28) dst0 reg | [28,32) dst1 reg | [32,48) imm0 | [48,64) imm1
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass

from . import isa
from .isa import C


@dataclass(frozen=True)
class Src:
    mode: int
    reg: int = 0
    imm: object = 0     # int or label name (str)


@dataclass(frozen=True)
class Dst:
    mode: int
    reg: int = 0
    imm: object = 0


def R(i: int) -> Src:
    return Src(isa.SRC_REG, i, 0)


def Imm(v) -> Src:
    return Src(isa.SRC_IMM, 0, v)


def Code(idx, reg: int = 0) -> Src:
    """constant from the code page at word index reg+idx (``idx`` may be a constant label)."""
    return Src(isa.SRC_CODE, reg, idx)


def StackPop(imm: int = 1, reg: int = 0) -> Src:
    return Src(isa.SRC_STACK_POP, reg, imm)


def StackRel(imm: int, reg: int = 0) -> Src:
    return Src(isa.SRC_STACK_REL, reg, imm)


def StackAbs(imm: int, reg: int = 0) -> Src:
    return Src(isa.SRC_STACK_ABS, reg, imm)


def DR(i: int) -> Dst:
    return Dst(isa.DST_REG, i, 0)


def DStackPush(imm: int = 1, reg: int = 0) -> Dst:
    return Dst(isa.DST_STACK_PUSH, reg, imm)


def DStackRel(imm: int, reg: int = 0) -> Dst:
    return Dst(isa.DST_STACK_REL, reg, imm)


def DStackAbs(imm: int, reg: int = 0) -> Dst:
    return Dst(isa.DST_STACK_ABS, reg, imm)


def _as_src(x) -> Src:
    if isinstance(x, Src):
        return x
    if isinstance(x, int):
        return R(x)
    raise TypeError(x)


def _as_dst(x) -> Dst:
    if isinstance(x, Dst):
        return x
    if isinstance(x, int):
        return DR(x)
    raise TypeError(x)


class Program:
    """Collects instructions + 256-bit constants; ``words()`` renders the big-endian code words."""

    def __init__(self):
        self.ins = []          # (family, sub, Src, src1_reg, Dst, dst1_reg, flags, cond, imm0_override, imm1_override)
        self.labels = {}
        self.consts = []       # (label, int value)

    # -- plumbing -------------------------------------------------------------------------
    @property
    def pc(self) -> int:
        return len(self.ins)

    def label(self, name: str):
        assert name not in self.labels, name
        self.labels[name] = len(self.ins)
        return self

    def const(self, name: str, value: int):
        self.consts.append((name, value & ((1 << 256) - 1)))
        return self

    def raw(self, family, sub=0, src0=0, src1=0, dst0=0, dst1=0, flags=0, cond="always", imm0=None, imm1=None):
        self.ins.append((family, sub, _as_src(src0), src1, _as_dst(dst0), dst1, flags, cond, imm0, imm1))
        return self

    # -- mnemonics ------------------------------------------------------------------------
    def nop(self, src0=0, dst0=0, cond="always"):
        return self.raw(isa.NOP, 0, src0, 0, dst0, cond=cond)

    def add(self, src0, src1, dst0, set_flags=False, cond="always"):
        return self.raw(isa.ADD, 0, src0, src1, dst0, flags=int(set_flags), cond=cond)

    def sub(self, src0, src1, dst0, set_flags=False, swap=False, cond="always"):
        return self.raw(isa.SUB, 0, src0, src1, dst0, flags=int(set_flags) | (int(swap) << 1), cond=cond)

    def mul(self, src0, src1, dst0, dst1=0, set_flags=False, cond="always"):
        return self.raw(isa.MUL, 0, src0, src1, dst0, dst1, flags=int(set_flags), cond=cond)

    def div(self, src0, src1, dst0, dst1=0, set_flags=False, swap=False, cond="always"):
        return self.raw(isa.DIV, 0, src0, src1, dst0, dst1, flags=int(set_flags) | (int(swap) << 1), cond=cond)

    def jump(self, target, cond="always"):
        src = Imm(target) if isinstance(target, (str, int)) else target
        return self.raw(isa.JUMP, 0, src, cond=cond)

    def context(self, sub, dst0=0, src0=0, cond="always"):
        return self.raw(isa.CONTEXT, sub, src0, 0, dst0, cond=cond)

    def shift(self, sub, src0, src1, dst0, set_flags=False, swap=False, cond="always"):
        return self.raw(isa.SHIFT, sub, src0, src1, dst0, flags=int(set_flags) | (int(swap) << 1), cond=cond)

    def binop(self, sub, src0, src1, dst0, set_flags=False, cond="always"):
        return self.raw(isa.BINOP, sub, src0, src1, dst0, flags=int(set_flags), cond=cond)

    def ptr(self, sub, src0, src1, dst0, swap=False, cond="always"):
        return self.raw(isa.PTR, sub, src0, src1, dst0, flags=int(swap), cond=cond)

    def near_call(self, abi_reg, target, handler, cond="always"):
        return self.raw(isa.NEAR_CALL, 0, abi_reg, cond=cond, imm0=target, imm1=handler)

    def log(self, sub, src0=0, src1=0, dst0=0, first=False, cond="always"):
        return self.raw(isa.LOG, sub, src0, src1, dst0, flags=int(first), cond=cond)

    def sload(self, key, dst0, cond="always"):
        return self.log(isa.LOG_SLOAD, key, 0, dst0, cond=cond)

    def sstore(self, key, value, cond="always"):
        return self.log(isa.LOG_SSTORE, key, value, cond=cond)

    def event(self, key, value, first=False, cond="always"):
        return self.log(isa.LOG_EVENT, key, value, first=first, cond=cond)

    def to_l1(self, key, value, first=False, cond="always"):
        return self.log(isa.LOG_TO_L1, key, value, first=first, cond=cond)

    def precompile(self, abi_reg, extra_ergs_reg, dst0, cond="always"):
        return self.log(isa.LOG_PRECOMPILE, abi_reg, extra_ergs_reg, dst0, cond=cond)

    def far_call(self, abi_reg, dest_reg, handler, sub=isa.FC_NORMAL, static=False, shard=False, cond="always"):
        return self.raw(isa.FAR_CALL, sub, abi_reg, dest_reg, flags=int(shard) | (int(static) << 1),
                        cond=cond, imm0=handler)

    def ret(self, sub=isa.RET_OK, abi_reg=0, label=None, cond="always"):
        return self.raw(isa.RET, sub, abi_reg, flags=int(label is not None), cond=cond,
                        imm0=label if label is not None else None)

    def uma(self, sub, src0, src1=0, dst0=0, dst1=0, inc=False, cond="always"):
        return self.raw(isa.UMA, sub, src0, src1, dst0, dst1, flags=int(inc), cond=cond)

    def ld(self, addr, dst0, dst1=0, inc=False, cond="always"):
        return self.uma(isa.UMA_HEAP_READ, addr, 0, dst0, dst1, inc, cond)

    def st(self, addr, value_reg, dst0=0, inc=False, cond="always"):
        return self.uma(isa.UMA_HEAP_WRITE, addr, value_reg, dst0, 0, inc, cond)

    def ld_aux(self, addr, dst0, dst1=0, inc=False, cond="always"):
        return self.uma(isa.UMA_AUX_READ, addr, 0, dst0, dst1, inc, cond)

    def st_aux(self, addr, value_reg, dst0=0, inc=False, cond="always"):
        return self.uma(isa.UMA_AUX_WRITE, addr, value_reg, dst0, 0, inc, cond)

    def ld_ptr(self, ptr_reg, dst0, dst1=0, inc=False, cond="always"):
        return self.uma(isa.UMA_PTR_READ, ptr_reg, 0, dst0, dst1, inc, cond)

    # -- rendering ------------------------------------------------------------------------
    def _resolve(self, v, const_base):
        if v is None:
            return 0
        if isinstance(v, str):
            if v in self.labels:
                return self.labels[v]
            for i, (name, _) in enumerate(self.consts):
                if name == v:
                    return const_base + i
            raise KeyError(v)
        return int(v) & 0xFFFF

    def encode(self):
        n_code_words = (len(self.ins) + 3) // 4
        const_base = n_code_words
        out = []
        for (family, sub, src, src1, dst, dst1, flags, cond, imm0o, imm1o) in self.ins:
            key = (family, sub, src.mode, dst.mode, flags)
            if key not in isa.VARIANT_INDEX:
                raise ValueError(f"no such variant {isa.FAMILY_NAMES[family]} {key}")
            imm0 = self._resolve(imm0o if imm0o is not None else src.imm, const_base)
            imm1 = self._resolve(imm1o if imm1o is not None else dst.imm, const_base)
            w = isa.VARIANT_INDEX[key] | (isa.COND[cond] << isa.COND_SHIFT)
            w |= (src.reg & 15) << isa.SRC0_REG_SHIFT | (src1 & 15) << isa.SRC1_REG_SHIFT
            w |= (dst.reg & 15) << isa.DST0_REG_SHIFT | (dst1 & 15) << isa.DST1_REG_SHIFT
            w |= (imm0 & 0xFFFF) << isa.IMM0_SHIFT | (imm1 & 0xFFFF) << isa.IMM1_SHIFT
            out.append(w)
        return out

    def words(self) -> list[int]:
        """code page as 256-bit integers (instructions, then constants); odd length as Era requires."""
        ins = self.encode()
        ins += [0] * (-len(ins) % 4)
        words = []
        for i in range(0, len(ins), 4):
            words.append((ins[i] << 192) | (ins[i + 1] << 128) | (ins[i + 2] << 64) | ins[i + 3])
        words += [v for _, v in self.consts]
        if len(words) % 2 == 0:
            words.append(0)
        return words

    def bytecode(self) -> bytes:
        return b"".join(w.to_bytes(32, "big") for w in self.words())


def bytecode_hash(code: bytes, marker: int = C.CODE_AT_REST_MARKER) -> int:
    """versioned code hash (ContractCodeSha256 layout: version, marker, BE u16 length in words, sha256 tail;
    consumed at /root/reference/src/opcodes/execution/far_call.rs:169-252)."""
    assert len(code) % 32 == 0
    n_words = len(code) // 32
    digest = hashlib.sha256(code).digest()
    raw = bytes([C.CODE_HASH_VERSION_BYTE, marker]) + n_words.to_bytes(2, "big") + digest[4:]
    return int.from_bytes(raw, "big")


def far_call_abi(ergs: int, *, offset=0, page=0, start=0, length=0, fwd=C.FWD_USE_HEAP, shard=0,
                 constructor=False, to_system=False) -> int:
    """FarCallABI as a U256 (low 128 bits = fat pointer; far_call.rs:82)."""
    v = (offset & 0xFFFFFFFF) | (page & 0xFFFFFFFF) << 32 | (start & 0xFFFFFFFF) << 64 | (length & 0xFFFFFFFF) << 96
    v |= (ergs & 0xFFFFFFFF) << 192 | (fwd & 0xFF) << 224 | (shard & 0xFF) << 232
    v |= (int(constructor) & 0xFF) << 240 | (int(to_system) & 0xFF) << 248
    return v


def ret_abi(*, offset=0, page=0, start=0, length=0, fwd=C.FWD_USE_HEAP) -> int:
    v = (offset & 0xFFFFFFFF) | (page & 0xFFFFFFFF) << 32 | (start & 0xFFFFFFFF) << 64 | (length & 0xFFFFFFFF) << 96
    return v | (fwd & 0xFF) << 224


def precompile_abi(*, in_off=0, in_len=0, out_off=0, out_len=0, page_read=0, page_write=0, data=0) -> int:
    """PrecompileCallABI (field order cited at src/testing/tests/precompiles/keccak256.rs:103-111)."""
    return (in_off & 0xFFFFFFFF) | (in_len & 0xFFFFFFFF) << 32 | (out_off & 0xFFFFFFFF) << 64 | \
        (out_len & 0xFFFFFFFF) << 96 | (page_read & 0xFFFFFFFF) << 128 | (page_write & 0xFFFFFFFF) << 160 | \
        (data & 0xFFFFFFFFFFFFFFFF) << 192
