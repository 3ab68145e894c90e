"""CPU suite: the ISA pinning hook.  The decode table / prices / predicates / constants are a reconstruction of the absent
crate zkevm_opcode_defs; `isa.py --to-json / --from-json` lets a maintainer with the real crate drop
its dump in.  Here: the dump round-trips to the identical generated header, a permuted / repriced
table is honoured everywhere the module derives data from it, malformed tables are rejected."""
import copy
import importlib

import pytest


@pytest.fixture()
def isa():
    from era_zk_evm_b200 import isa as mod
    yield mod
    importlib.reload(mod)            # other tests must see the in-tree table again


def test_dump_round_trips_to_the_identical_header(isa):
    before = isa.gen_header()
    dump = isa.to_json()
    assert dump["schema"] == isa.JSON_SCHEMA and len(dump["variants"]) == isa.N_VALID_VARIANTS
    isa.load_json(copy.deepcopy(dump))
    after = isa.gen_header()
    strip = lambda t: "\n".join(l for l in t.split("\n") if not l.startswith("//"))   # the provenance comment differs
    assert strip(after) == strip(before)


def test_a_different_table_is_honoured(isa):
    dump = isa.to_json()
    v = dump["variants"]
    key = lambda d: (d["family"], d["sub"], d["src"], d["dst"], d["flags"])
    i_add = next(i for i, d in enumerate(v) if key(d) == ("add", 0, 0, 0, 0))
    i_sub = next(i for i, d in enumerate(v) if key(d) == ("sub", 0, 0, 0, 0))
    v[i_add], v[i_sub] = v[i_sub], v[i_add]                      # the crate numbers its variants differently
    for d in v:
        if d["family"] == "near_call":
            d["price"] = 26                                       # ... prices them differently
        if d["family"] == "ptr":
            d["src1_ptr_ok"] = True                               # ... and answers src1_can_be_pointer differently
    dump["constants"]["BOOTLOADER_CALLDATA_PAGE"] = 7
    dump["source"] = "unit test"
    isa.load_json(dump)
    assert isa.VARIANT_INDEX[(isa.ADD, 0, isa.SRC_REG, isa.DST_REG, 0)] == i_sub
    assert isa.VARIANT_INDEX[(isa.SUB, 0, isa.SRC_REG, isa.DST_REG, 0)] == i_add
    assert isa.OPCODE_TABLE[i_sub] & 15 == isa.ADD
    near = isa.VARIANT_INDEX[(isa.NEAR_CALL, 0, isa.SRC_REG, isa.DST_REG, 0)]
    assert isa.OPCODE_PRICES[near] == 26
    ptr = isa.VARIANT_INDEX[(isa.PTR, isa.PTR_ADD, isa.SRC_REG, isa.DST_REG, 0)]
    assert isa.OPCODE_TABLE[ptr] & isa.E_SRC1_PTR_OK
    assert isa.C.BOOTLOADER_CALLDATA_PAGE == 7
    text = isa.gen_header()
    assert "#define ZK_BOOTLOADER_CALLDATA_PAGE 7u" in text and "ISA data source: unit test" in text
    assert isa.PANIC_VARIANT_IDX == isa.VARIANT_INDEX[(isa.RET, isa.RET_PANIC, isa.SRC_REG, isa.DST_REG, 0)]


def test_malformed_tables_are_rejected(isa):
    dump = isa.to_json()
    bad = copy.deepcopy(dump)
    bad["schema"] = "something else"
    with pytest.raises(ValueError):
        isa.load_json(bad)
    bad = copy.deepcopy(dump)
    bad["variants"][0]["family"] = "add"                          # entry 0 must stay the invalid opcode
    with pytest.raises(ValueError):
        isa.load_json(bad)
    bad = copy.deepcopy(dump)
    bad["constants"]["NOT_A_CONSTANT"] = 1
    with pytest.raises(ValueError):
        isa.load_json(bad)
