"""EraVM ISA data table (the single source of ISA data for oracle, kernels and assembler).

The reference consumes this data from the external crate ``zkevm_opcode_defs`` (git branch
``v1.4.1``; ``/root/reference/Cargo.toml:15``) whose source is NOT in ``/root/reference``.  Everything
here is therefore a reconstruction ("parity unpinned") that is kept
in ONE place so it can be replaced wholesale by the crate's real values:

* ``OPCODE_TABLE``  — the 2048-entry decode table indexed by the low 11 bits of an instruction
  (reference use: ``E::parse_preliminary_variant_and_absolute_number``, ``src/vm_state/cycle.rs:135``)
* ``OPCODE_PRICES`` — ergs per raw variant (reference use: ``OPCODES_PRICES[idx]``, ``cycle.rs:147``)
* ``C``             — system parameters (``STARTING_TIMESTAMP`` etc.; uses listed in SURVEY §8c)

``gen_header()`` renders the same data as ``csrc/isa_tables.inc`` (C arrays + #defines) which both
the CPU oracle (``oracle/``) and the CUDA kernels include.  Only DATA is shared between oracle and
product; the interpreter logic is written twice, independently.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

# ---------------------------------------------------------------------------------------------
# opcode families (order = reference's `Opcode` enum order as dispatched in
# src/opcodes/parsing.rs:61-78, with Invalid first as table entry 0)
# ---------------------------------------------------------------------------------------------
INVALID, NOP, ADD, SUB, MUL, DIV, JUMP, CONTEXT, SHIFT, BINOP, PTR, NEAR_CALL, LOG, FAR_CALL, RET, UMA = range(16)
FAMILY_NAMES = ["invalid", "nop", "add", "sub", "mul", "div", "jump", "context", "shift", "binop",
                "ptr", "near_call", "log", "far_call", "ret", "uma"]

# sub-variants (names cited by use in the reference handlers)
CTX_THIS, CTX_CALLER, CTX_CODE_ADDRESS, CTX_META, CTX_ERGS_LEFT, CTX_SP, CTX_GET_U128, CTX_SET_U128, \
    CTX_SET_ERGS_PER_PUBDATA, CTX_INC_TX = range(10)          # context.rs:36-98
SHL, SHR, ROL, ROR = range(4)                                  # shift.rs:45-46
XOR, AND, OR = range(3)                                        # binop.rs:42-46
PTR_ADD, PTR_SUB, PTR_PACK, PTR_SHRINK = range(4)              # ptr.rs:33,96,140
LOG_SLOAD, LOG_SSTORE, LOG_TO_L1, LOG_EVENT, LOG_PRECOMPILE = range(5)   # log.rs:70-252
FC_NORMAL, FC_DELEGATE, FC_MIMIC = range(3)                    # far_call.rs:510-523
RET_OK, RET_REVERT, RET_PANIC = range(3)                       # ret.rs:35,105,196
UMA_HEAP_READ, UMA_HEAP_WRITE, UMA_AUX_READ, UMA_AUX_WRITE, UMA_PTR_READ = range(5)  # uma.rs:80-108

# operand addressing modes (mem_ops.rs:37-122)
SRC_REG, SRC_STACK_POP, SRC_STACK_REL, SRC_STACK_ABS, SRC_IMM, SRC_CODE = range(6)
DST_REG, DST_STACK_PUSH, DST_STACK_REL, DST_STACK_ABS = range(4)

# conditions (cycle.rs:193-210)
COND = {"always": 0, "gt": 1, "lt": 2, "eq": 3, "ge": 4, "le": 5, "ne": 6, "gtlt": 7}

# decode-table entry bit layout (u32)
E_FAMILY_SHIFT, E_SUB_SHIFT, E_SRC_SHIFT, E_DST_SHIFT = 0, 4, 8, 11
E_FLAG0, E_FLAG1 = 1 << 13, 1 << 14
E_KERNEL_ONLY, E_STATIC_FORBIDDEN = 1 << 15, 1 << 16
E_SRC0_PTR_OK, E_SRC1_PTR_OK = 1 << 17, 1 << 18
E_SWAP, E_INVALID = 1 << 19, 1 << 20

# instruction word bit layout (EncodingModeProduction, N = 8)
VARIANT_BITS = 11
COND_SHIFT = 13
SRC0_REG_SHIFT, SRC1_REG_SHIFT, DST0_REG_SHIFT, DST1_REG_SHIFT = 16, 20, 24, 28
IMM0_SHIFT, IMM1_SHIFT = 32, 48


class C:
    """system parameters (zkevm_opcode_defs::system_params and friends) — reconstructed."""
    REGISTERS_COUNT = 15
    STARTING_TIMESTAMP = 1024
    STARTING_BASE_PAGE = 8
    TIME_DELTA_PER_CYCLE = 4
    NEW_MEMORY_PAGES_PER_FAR_CALL = 8
    UNMAPPED_PAGE = 0
    BOOTLOADER_CALLDATA_PAGE = 3
    INITIAL_SP_ON_FAR_CALL = 0
    VM_INITIAL_FRAME_ERGS = 0xFFFFFFFF
    CALL_LIKE_ERGS_COST = 20
    VM_MAX_STACK_DEPTH = 0xFFFFFFFF // 20 + 80
    MEMORY_GROWTH_ERGS_PER_BYTE = 1
    NEW_FRAME_MEMORY_STIPEND = 4096
    ERGS_PER_CODE_WORD_DECOMMITTMENT = 4
    STORAGE_AUX_BYTE, EVENT_AUX_BYTE, L1_MESSAGE_AUX_BYTE, PRECOMPILE_AUX_BYTE = 0, 1, 2, 3
    INITIAL_STORAGE_WRITE_PUBDATA_BYTES = 64
    L1_MESSAGE_PUBDATA_BYTES = 88
    DEPLOYER_SYSTEM_CONTRACT_ADDRESS = 0x8002
    ADDRESS_MSG_VALUE = 0x8009
    ADDRESS_EVENT_WRITER = 0x800D
    KECCAK256_PRECOMPILE_ADDRESS = 0x8010
    SHA256_PRECOMPILE_ADDRESS = 0x02
    ECRECOVER_PRECOMPILE_ADDRESS = 0x01
    MAX_OFFSET_FOR_ADD_SUB = 1 << 32
    MAX_OFFSET_TO_DEREF = (1 << 32) - 33
    MAX_STACK_PAGE_SIZE_IN_WORDS = 1 << 16
    MAX_CODE_PAGE_SIZE_IN_WORDS = 1 << 16
    # far-call / ret register ABI (0-based index into registers[]; far_call.rs:573-610, ret.rs:213-233)
    CALL_IMPLICIT_CALLDATA_FAT_PTR_REGISTER = 0
    CALL_IMPLICIT_CONSTRUCTOR_MARKER_REGISTER = 1
    CALL_SYSTEM_ABI_REGISTERS = (2, 12)
    CALL_RESERVED_RANGE = (12, 14)
    CALL_IMPLICIT_PARAMETER_REG_IDX = 14
    RET_IMPLICIT_RETURNDATA_PARAMS_REGISTER = 0
    # forwarding modes
    FWD_USE_HEAP, FWD_FORWARD_FAT_POINTER, FWD_USE_AUX_HEAP = 0, 1, 2
    # versioned hash
    CODE_HASH_VERSION_BYTE = 1
    CODE_AT_REST_MARKER, YET_CONSTRUCTED_MARKER = 0, 1
    # memory types (zk_evm_abstractions::vm::MemoryType)
    MEM_STACK, MEM_HEAP, MEM_AUX_HEAP, MEM_FAT_PTR, MEM_CODE = range(5)


# ergs price building blocks
_VM_CYCLE, _RAM = 4, 1
_AVERAGE = _VM_CYCLE + 2 * _RAM          # 6
_RICH = _VM_CYCLE + 4 * _RAM             # 8


def _price(family: int, sub: int) -> int:
    if family == INVALID:
        return 0xFFFFFFFF
    if family in (NOP, ADD, SUB, JUMP, SHIFT, BINOP, PTR):
        return _RICH
    if family in (MUL, DIV):
        return 2 * _VM_CYCLE + 4 * _RAM                      # 12
    if family == CONTEXT:
        return _AVERAGE
    if family == NEAR_CALL:
        return _AVERAGE + 19                                 # 25
    if family == LOG:
        return {LOG_SLOAD: 158, LOG_SSTORE: 258, LOG_TO_L1: 156250, LOG_EVENT: 34, LOG_PRECOMPILE: 6}[sub]
    if family == FAR_CALL:
        return 182
    if family == RET:
        return _AVERAGE
    if family == UMA:
        return _VM_CYCLE + (5 if sub in (UMA_HEAP_WRITE, UMA_AUX_WRITE) else 3) * _RAM
    raise AssertionError


PREDICATES = ("kernel_only", "static_forbidden", "src0_ptr_ok", "src1_ptr_ok", "swap")


@dataclass(frozen=True)
class Variant:
    family: int
    sub: int
    src: int
    dst: int
    flags: int  # bit0 = flag0, bit1 = flag1
    # per-variant predicates and price; None = derived from (family, sub, flags) by the reconstruction below, a value =
    # taken from a dump of the real crate (load_json)
    given: tuple = field(default=None, compare=False)   # (kernel_only, static_forbidden, src0_ptr_ok, src1_ptr_ok, swap, price)

    @property
    def predicates(self) -> dict:
        if self.given is not None:
            return dict(zip(PREDICATES, (bool(x) for x in self.given[:5])))
        f, s = self.family, self.sub
        return {
            # OpcodeVariant::requires_kernel_mode (cycle.rs:174)
            "kernel_only": (f == CONTEXT and s in (CTX_SET_U128, CTX_SET_ERGS_PER_PUBDATA, CTX_INC_TX)) or
                           (f == LOG and s in (LOG_TO_L1, LOG_EVENT, LOG_PRECOMPILE)) or (f == FAR_CALL and s == FC_MIMIC),
            # !OpcodeVariant::can_be_used_in_static_context (cycle.rs:178)
            "static_forbidden": (f == CONTEXT and s in (CTX_SET_U128, CTX_SET_ERGS_PER_PUBDATA, CTX_INC_TX)) or
                                (f == LOG and s in (LOG_SSTORE, LOG_TO_L1, LOG_EVENT)),
            # Opcode::src0_can_be_pointer (cycle.rs:379)
            "src0_ptr_ok": f in (PTR, FAR_CALL, RET) or (f == UMA and s == UMA_PTR_READ),
            # Opcode::src1_can_be_pointer (cycle.rs:387).  Recollection of zkevm_opcode_defs 1.4.x: the function returns
            # `false` for every opcode.  (SURVEY Appendix A guessed "Ptr* only, so that ptr.rs:41 is reachable"; with
            # `false` that check is still reachable -- in kernel mode, where cycle.rs:374-396 erases nothing -- and in user
            # mode a pointer in src1 is demoted to an integer before ptr.rs sees it.)  Unpinned either way: a dump of the
            # real crate (load_json) overrides it per variant.
            "src1_ptr_ok": False,
            # swap flag: arithmetic families carry {set_flags=flag0, swap=flag1}; ptr carries {swap=flag0}
            "swap": (f in (SUB, DIV, SHIFT) and bool(self.flags & 2)) or (f == PTR and bool(self.flags & 1)),
        }

    @property
    def price(self) -> int:
        return int(self.given[5]) if self.given is not None else _price(self.family, self.sub)

    @property
    def entry(self) -> int:
        f, s = self.family, self.sub
        e = (f << E_FAMILY_SHIFT) | (s << E_SUB_SHIFT) | (self.src << E_SRC_SHIFT) | (self.dst << E_DST_SHIFT)
        if self.flags & 1:
            e |= E_FLAG0
        if self.flags & 2:
            e |= E_FLAG1
        pr = self.predicates
        kernel_only, static_forbidden, src0_ptr_ok, src1_ptr_ok, swap = (pr[k] for k in PREDICATES)
        if kernel_only:
            e |= E_KERNEL_ONLY
        if static_forbidden:
            e |= E_STATIC_FORBIDDEN
        if src0_ptr_ok:
            e |= E_SRC0_PTR_OK
        if src1_ptr_ok:
            e |= E_SRC1_PTR_OK
        if swap:
            e |= E_SWAP
        if f == INVALID:
            e |= E_INVALID
        return e


_FULL_SRC = (SRC_REG, SRC_STACK_POP, SRC_STACK_REL, SRC_STACK_ABS, SRC_IMM, SRC_CODE)
_FULL_DST = (DST_REG, DST_STACK_PUSH, DST_STACK_REL, DST_STACK_ABS)


def _synthesize():
    """family × sub-variant × src0 mode × dst0 mode × flag combinations, Invalid-padded to 2048."""
    out = [Variant(INVALID, 0, SRC_REG, DST_REG, 0)]

    def emit(family, subs, srcs, dsts, nflags):
        for sub in range(subs):
            for s in srcs:
                for d in dsts:
                    for fl in range(1 << nflags):
                        out.append(Variant(family, sub, s, d, fl))

    emit(NOP, 1, _FULL_SRC, _FULL_DST, 0)
    emit(ADD, 1, _FULL_SRC, _FULL_DST, 1)          # set_flags
    emit(SUB, 1, _FULL_SRC, _FULL_DST, 2)          # set_flags, swap
    emit(MUL, 1, _FULL_SRC, _FULL_DST, 1)
    emit(DIV, 1, _FULL_SRC, _FULL_DST, 2)
    emit(JUMP, 1, _FULL_SRC, (DST_REG,), 0)
    emit(CONTEXT, 10, (SRC_REG,), (DST_REG,), 0)
    emit(SHIFT, 4, _FULL_SRC, _FULL_DST, 2)
    emit(BINOP, 3, _FULL_SRC, _FULL_DST, 1)
    emit(PTR, 4, _FULL_SRC, _FULL_DST, 1)          # swap
    emit(NEAR_CALL, 1, (SRC_REG,), (DST_REG,), 0)
    emit(LOG, 5, (SRC_REG,), (DST_REG,), 1)        # is_first
    emit(FAR_CALL, 3, (SRC_REG,), (DST_REG,), 2)   # flag0 = shard, flag1 = static
    emit(RET, 3, (SRC_REG,), (DST_REG,), 1)        # to_label
    emit(UMA, 5, (SRC_REG, SRC_IMM), (DST_REG,), 1)  # increment
    assert len(out) <= (1 << VARIANT_BITS)
    n_valid = len(out)
    out.extend([Variant(INVALID, 0, SRC_REG, DST_REG, 0)] * ((1 << VARIANT_BITS) - len(out)))
    return out, n_valid


# ---------------------------------------------------------------------------------------------
# pinning hook: the whole table as JSON.  `python -m era_zk_evm_b200.isa --to-json f.json` writes the reconstruction;
# a dump of the REAL zkevm_opcode_defs tables in the same schema (INTEGRATION.md §7 has the 40-line Rust program that
# prints it from OPCODES_TABLE / OPCODES_PRICES / system_params) is installed with `--from-json f.json` (or the
# environment variable ZKB_ISA_JSON at import time): oracle, kernels, assembler and tests then run on the crate's own
# variant indices, prices, predicates and constants without a code change.
# ---------------------------------------------------------------------------------------------
JSON_SCHEMA = "zkb-isa/1"
ISA_SOURCE = "reconstruction (era_zk_evm_b200/isa.py)"


def to_json() -> dict:
    consts = {k: (list(v) if isinstance(v, tuple) else v) for k, v in sorted(vars(C).items()) if not k.startswith("_")}
    variants = []
    for v in VARIANTS[:N_VALID_VARIANTS]:
        d = {"family": FAMILY_NAMES[v.family], "sub": v.sub, "src": v.src, "dst": v.dst, "flags": v.flags, "price": v.price}
        d.update({k: bool(x) for k, x in v.predicates.items()})
        variants.append(d)
    return {"schema": JSON_SCHEMA, "source": ISA_SOURCE, "variant_bits": VARIANT_BITS, "constants": consts, "variants": variants}


def load_json(obj) -> None:
    """replaces the table, prices, predicates and constants of this module by those of `obj` (a dict or a path)"""
    global VARIANTS, N_VALID_VARIANTS, ISA_SOURCE
    if not isinstance(obj, dict):
        with open(obj) as f:
            obj = json.load(f)
    if obj.get("schema") != JSON_SCHEMA or obj.get("variant_bits") != VARIANT_BITS:
        raise ValueError(f"not a {JSON_SCHEMA} table with {VARIANT_BITS} variant bits")
    for k, v in obj["constants"].items():
        if not hasattr(C, k):
            raise ValueError(f"unknown constant {k}")
        setattr(C, k, tuple(v) if isinstance(v, list) else int(v))
    out = []
    for d in obj["variants"]:
        fam = FAMILY_NAMES.index(d["family"])
        given = tuple(bool(d[k]) for k in PREDICATES) + (int(d["price"]),)
        out.append(Variant(fam, int(d["sub"]), int(d["src"]), int(d["dst"]), int(d["flags"]), given))
    if not out or out[0].family != INVALID or len(out) > (1 << VARIANT_BITS):
        raise ValueError("variant 0 must be the invalid opcode and the table must fit the variant bits")
    n_valid = len(out)
    out.extend([Variant(INVALID, 0, SRC_REG, DST_REG, 0)] * ((1 << VARIANT_BITS) - len(out)))
    VARIANTS, N_VALID_VARIANTS = out, n_valid
    ISA_SOURCE = str(obj.get("source", "json"))
    _index()


def _index() -> None:
    global OPCODE_TABLE, OPCODE_PRICES, VARIANT_INDEX, NOP_VARIANT_IDX, PANIC_VARIANT_IDX, NOP_ENCODING, EXCEPTION_REVERT_ENCODING
    OPCODE_TABLE = [v.entry for v in VARIANTS]
    OPCODE_PRICES = [v.price for v in VARIANTS]
    VARIANT_INDEX = {}
    for i, v in enumerate(VARIANTS[:N_VALID_VARIANTS]):
        VARIANT_INDEX.setdefault((v.family, v.sub, v.src, v.dst, v.flags), i)
    NOP_VARIANT_IDX = VARIANT_INDEX[(NOP, 0, SRC_REG, DST_REG, 0)]
    PANIC_VARIANT_IDX = VARIANT_INDEX[(RET, RET_PANIC, SRC_REG, DST_REG, 0)]
    NOP_ENCODING = NOP_VARIANT_IDX              # E::nop_encoding()           (cycle.rs:126)
    EXCEPTION_REVERT_ENCODING = PANIC_VARIANT_IDX  # E::exception_revert_encoding() (cycle.rs:115)


VARIANTS, N_VALID_VARIANTS = _synthesize()
_index()
_OVERRIDE = os.path.join(os.path.dirname(__file__), "isa_pinned.json")   # written by --from-json
if os.environ.get("ZKB_ISA_JSON") or os.path.exists(_OVERRIDE):
    load_json(os.environ.get("ZKB_ISA_JSON") or _OVERRIDE)


def gen_header() -> str:
    lines = ["// GENERATED by era_zk_evm_b200/isa.py (python -m era_zk_evm_b200.isa) -- do not edit.",
             "// Reconstructed ISA data standing in for the external crate zkevm_opcode_defs@v1.4.1",
             "// (absent from /root/reference; parity unpinned, SURVEY.md Appendix A).",
             "#pragma once", "#include <stdint.h>", ""]
    if not ISA_SOURCE.startswith("reconstruction"):
        lines[1:3] = [f"// ISA data source: {ISA_SOURCE} (installed with isa.py --from-json)"]
    for k, v in sorted(vars(C).items()):
        if k.startswith("_"):
            continue
        if isinstance(v, tuple):
            lines.append(f"#define ZK_{k}_LO {v[0]}u")
            lines.append(f"#define ZK_{k}_HI {v[1]}u")
        else:
            lines.append(f"#define ZK_{k} {v}u" if v < (1 << 32) else f"#define ZK_{k} {v}ull")
    lines.append("")
    g = globals()
    for name in ("INVALID NOP ADD SUB MUL DIV JUMP CONTEXT SHIFT BINOP PTR NEAR_CALL LOG FAR_CALL RET UMA").split():
        lines.append(f"#define ZK_OP_{name} {g[name]}u")
    for name in [n for n in g if n.startswith(("CTX_", "PTR_", "LOG_", "FC_", "RET_", "UMA_", "SRC_", "DST_"))] + \
            ["SHL", "SHR", "ROL", "ROR", "XOR", "AND", "OR"]:
        if isinstance(g[name], int):
            lines.append(f"#define ZK_{name} {g[name]}u")
    for name in [n for n in g if n.startswith("E_")]:
        lines.append(f"#define ZK_{name} {g[name]}u")
    lines.append(f"#define ZK_VARIANT_BITS {VARIANT_BITS}u")
    lines.append(f"#define ZK_COND_SHIFT {COND_SHIFT}u")
    lines.append(f"#define ZK_NOP_ENCODING {NOP_ENCODING}ull")
    lines.append(f"#define ZK_EXCEPTION_REVERT_ENCODING {EXCEPTION_REVERT_ENCODING}ull")
    lines.append(f"#define ZK_PANIC_VARIANT_IDX {PANIC_VARIANT_IDX}u")
    lines.append(f"#define ZK_NOP_VARIANT_IDX {NOP_VARIANT_IDX}u")
    lines.append(f"#define ZK_N_VALID_VARIANTS {N_VALID_VARIANTS}u")
    lines.append("")
    lines.append("#ifndef ZK_TABLE_QUALIFIER\n#define ZK_TABLE_QUALIFIER static const\n#endif")
    lines.append("ZK_TABLE_QUALIFIER uint32_t ZK_OPCODE_TABLE[2048] = {")
    for i in range(0, 2048, 8):
        lines.append("  " + ", ".join(f"0x{e:06x}u" for e in OPCODE_TABLE[i:i + 8]) + ",")
    lines.append("};")
    lines.append("ZK_TABLE_QUALIFIER uint32_t ZK_OPCODE_PRICES[2048] = {")
    for i in range(0, 2048, 8):
        lines.append("  " + ", ".join(f"{p}u" for p in OPCODE_PRICES[i:i + 8]) + ",")
    lines.append("};")
    return "\n".join(lines) + "\n"


def write_header(path: str | None = None) -> str:
    path = path or os.path.join(os.path.dirname(__file__), "csrc", "isa_tables.inc")
    text = gen_header()
    old = open(path).read() if os.path.exists(path) else None
    if old != text:
        with open(path, "w") as f:
            f.write(text)
    return path


if __name__ == "__main__":
    import sys
    argv = sys.argv[1:]
    if argv[:1] == ["--to-json"]:
        with open(argv[1], "w") as f:
            json.dump(to_json(), f, indent=1)
        print(argv[1], N_VALID_VARIANTS, "variants,", ISA_SOURCE)
    elif argv[:1] == ["--from-json"]:
        load_json(argv[1])                       # validates
        with open(argv[1]) as f, open(_OVERRIDE, "w") as g:
            g.write(f.read())
        print(write_header(), N_VALID_VARIANTS, "variants from", ISA_SOURCE, "(pinned in", _OVERRIDE + "; delete it to go back)")
    else:
        print(write_header(), N_VALID_VARIANTS, "variants,", ISA_SOURCE)
