"""Hand-derived per-quirk programs.  Every expectation below is derived from the reference
SOURCE (file:line cited), not from the oracle: the CPU suite runs them on the oracle (pinning the restatement), the
`-m gpu` suite runs the same functions on the CUDA batch."""
import numpy as np

from era_zk_evm_b200 import isa, records
from era_zk_evm_b200.asm import (Code, DStackAbs, DStackPush, DStackRel, Imm, Program, R, StackAbs, StackPop, StackRel,
                                  far_call_abi, ret_abi)
from era_zk_evm_b200.isa import C

import vm_harness as H

M256 = (1 << 256) - 1
USER = 0x00000000000000000000000000000000DEADBEEF      # > 2^16: not a kernel address (execution_stack.rs:83-87)
ROOT_ERGS = C.VM_INITIAL_FRAME_ERGS


def case_ergs_charged_before_masking(B):
    """cycle.rs:147-163,187-190: the raw variant's price is charged first; on shortfall ergs := 0, NOT_ENOUGH_ERGS,
    the opcode becomes ret.panic; ret.rs:243 returns the (zero) remainder to the parent frame."""
    p = Program()
    p.add(R(1), 2, 3)
    p.add(R(1), 2, 3)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, regs={1: 5, 2: 6}, ergs=10)
    r = H.rows(b)
    assert len(r) == 2
    assert int(r[0]["ergs_after"]) == 10 - isa.OPCODE_PRICES[isa.VARIANT_INDEX[(isa.ADD, 0, isa.SRC_REG, isa.DST_REG, 0)]] == 2
    assert H.val(r[0]["dst0"]) == 11
    assert int(r[1]["error_flags"]) == 2 and int(r[1]["masked_variant"]) == isa.PANIC_VARIANT_IDX      # helpers.rs:344-352
    assert int(r[1]["cond_resolved"]) == 1                                                            # panic encoding is "always"
    assert int(r[1]["callstack_depth"]) == 0 and int(r[1]["ergs_after"]) == ROOT_ERGS - 10             # helpers.rs:295-303 + ret.rs:243
    assert int(r[1]["flags_after"]) == 1                                                              # ret.rs:262-264
    assert b.vm_status()[0, 0] == 1
    b.close()


def case_condition_false_is_nop_and_real_nop_moves_sp(B):
    """cycle.rs:212-217: a false condition masks to NOP with zero operands (SP untouched, nothing written);
    cycle.rs:298-301 + noop.rs:18-19: a real NOP with stack operands moves SP (pop then push, u16 wrapping,
    mem_ops.rs:55-86) but reads and writes nothing."""
    p = Program()
    p.add(Imm(1), 0, DStackPush(1), cond="gt")              # flags are all clear -> masked
    p.nop(StackPop(2), DStackPush(5))
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1)
    r = H.rows(b)
    assert int(r[0]["masked_variant"]) == isa.NOP_VARIANT_IDX and int(r[0]["cond_resolved"]) == 0 and int(r[0]["error_flags"]) == 0
    assert int(r[0]["sp_after"]) == 0 and int(r[0]["n_mem"]) == 1        # only the instruction fetch
    assert int(r[0]["bits"]) & 0xC0 == 0                                 # no dst written
    assert int(r[1]["sp_after"]) == (0 - 2 + 5) & 0xFFFF and int(r[1]["n_mem"]) == 0
    mem = b.read_stream(0, records.STREAM_MEM)
    assert all(int(m["memory_type"]) == C.MEM_CODE for m in mem)
    b.close()


def case_pop_then_push_in_one_instruction(B):
    """cycle.rs:278-297: src0 addressing mutates SP before dst0 addressing; mem_ops.rs:71-86 pop = subtract then
    use the new SP, :55-70 push = use the old SP then add.  Reads at t+0, dst write at t+3 (mod.rs:220-231)."""
    p = Program()
    p.nop(R(0), DStackPush(3))                              # sp = 3
    p.add(Imm(7), 0, DStackAbs(2))                          # stack[2] = 7
    p.add(StackPop(1), 1, DStackPush(1))                    # reads stack[2], writes stack[2] = 7 + r1, sp back to 3
    p.add(StackAbs(2), 0, 5)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, regs={1: 100})
    r = H.rows(b)
    assert int(r[2]["sp_after"]) == 3 and H.val(r[2]["src0"]) == 7 and H.val(r[2]["dst0"]) == 107
    assert H.val(r[3]["dst0"]) == 107
    mem = [m for m in b.read_stream(0, records.STREAM_MEM) if int(m["memory_type"]) == C.MEM_STACK]
    ts = int(r[2]["timestamp"])
    rd = [m for m in mem if int(m["timestamp"]) == ts]
    wr = [m for m in mem if int(m["timestamp"]) == ts + 3]
    assert len(rd) == 1 and int(rd[0]["index"]) == 2 and int(rd[0]["rw_flag"]) == 0 and H.val(rd[0]["value"]) == 7
    assert len(wr) == 1 and int(wr[0]["index"]) == 2 and int(wr[0]["rw_flag"]) == 1 and H.val(wr[0]["value"]) == 107
    assert int(wr[0]["page"]) == H.BOOT_PAGE + 1                          # stack page = base + 1 (execution_stack.rs:71-73)
    b.close()


def case_pointer_erasure_only_in_user_mode(B):
    """cycle.rs:374-396: for opcodes that cannot take pointers the top 128 bits of a pointer operand are erased --
    outside kernel mode only.  The row's SRC0_PTR bit is the pre-erasure marker (quirk 16)."""
    ptr = (0xAAAA << 200) | (5 << 96) | (7 << 64) | (9 << 32) | 3
    for this, expect in ((USER, ptr & ((1 << 128) - 1)), (H.BOOT_ADDRESS, ptr)):
        p = Program()
        p.add(R(1), 0, 3)
        p.ret(isa.RET_OK, R(0))
        b = H.launch(B, p, 1, regs={1: ptr}, ptr_regs=(1,), this_address=this)
        r = H.rows(b)[0]
        assert H.val(r["src0"]) == expect and H.val(r["dst0"]) == expect
        assert int(r["bits"]) & records_bit("SRC0_PTR") and not int(r["bits"]) & records_bit("DST0_PTR")   # add.rs: result never a pointer
        b.close()


def records_bit(name):
    return {"SRC0_PTR": 0x01, "SRC1_PTR": 0x02, "DST0_PTR": 0x04, "DST1_PTR": 0x08, "PENDING": 0x10, "SKIP": 0x20,
            "DST0_VALID": 0x40, "DST1_VALID": 0x80}[name]


def case_imm_zero_extended_and_r0(B):
    """cycle.rs:331-335: imm16 operands are zero-extended; helpers.rs:318-334: r0 reads as zero, writes to it vanish."""
    p = Program()
    p.add(Imm(0xFFFF), 0, 3)
    p.sub(Imm(0xFFFF), 0, 0, set_flags=True)                # result discarded, flags kept
    p.add(R(0), 0, 4)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1)
    r = H.rows(b)
    assert H.val(r[0]["dst0"]) == 0xFFFF
    assert H.val(r[1]["dst0"]) == 0xFFFF and int(r[1]["flags_after"]) == 4         # 0xFFFF - 0 > 0: GT (sub.rs:40-44)
    assert H.val(r[2]["src0"]) == 0 and H.val(r[2]["dst0"]) == 0
    b.close()


def case_sload_reports_written_equals_read(B):
    """helpers.rs:145-148: for reads the tracer sees written_value = read_value; log.rs:163-194 dst0 = value; t+1."""
    p = Program()
    p.add(Imm(0x42), 0, 1)
    p.sload(1, 2)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, storage=[(0, H.BOOT_ADDRESS, 0x42, 0x1234)])
    r = H.rows(b)
    lg = b.read_stream(0, records.STREAM_LOG)
    assert len(lg) == 1 and H.val(r[1]["dst0"]) == 0x1234
    q = lg[0]
    assert H.val(q["key"]) == 0x42 and H.val(q["read_value"]) == 0x1234 == H.val(q["written_value"])
    assert int(q["rw_flag"]) == 0 and int(q["aux_byte"]) == C.STORAGE_AUX_BYTE and int(q["timestamp"]) == int(r[1]["timestamp"]) + 1
    assert bytes(q["address"]) == H.BOOT_ADDRESS.to_bytes(20, "big")
    b.close()


def case_sstore_out_of_ergs_records_refund_only(B):
    """log.rs:99-102 the refund is recorded first; :136-144 on shortfall ergs := 0 and spent_pubdata +=
    min(ergs_available, pubdata cost); :196-199 the write itself is skipped (no log query, storage untouched)."""
    price = isa.OPCODE_PRICES[isa.VARIANT_INDEX[(isa.LOG, isa.LOG_SSTORE, isa.SRC_REG, isa.DST_REG, 0)]]
    epp = 3
    cost = epp * C.INITIAL_STORAGE_WRITE_PUBDATA_BYTES
    avail = 50
    p = Program()
    p.sstore(1, 2)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, regs={1: 7, 2: 9}, ergs=price + avail, ergs_per_pubdata=epp, storage=[(0, H.BOOT_ADDRESS, 7, 1)])
    r = H.rows(b)
    assert avail < cost
    assert int(r[0]["ergs_after"]) == 0 and int(r[0]["spent_pubdata"]) == min(avail, cost)
    assert len(b.read_stream(0, records.STREAM_REFUND)) == 1 and len(b.read_stream(0, records.STREAM_LOG)) == 0
    assert b.read_storage(0, 0, H.BOOT_ADDRESS, 7) == 1
    assert int(r[0]["bits"]) & records_bit("PENDING") == 0                  # "DO NOT set any pending" (log.rs:49-53)
    b.close()
    # with enough ergs: ergs -= cost, spent += cost, one refund + one log query (old value in read_value)
    b = H.launch(B, p, 1, regs={1: 7, 2: 9}, ergs=price + cost + 100, ergs_per_pubdata=epp, storage=[(0, H.BOOT_ADDRESS, 7, 1)])
    r = H.rows(b)
    lg = b.read_stream(0, records.STREAM_LOG)
    assert int(r[0]["ergs_after"]) == 100 and int(r[0]["spent_pubdata"]) == cost
    assert len(lg) == 1 and H.val(lg[0]["read_value"]) == 1 and H.val(lg[0]["written_value"]) == 9 and int(lg[0]["rw_flag"]) == 1
    assert b.read_storage(0, 0, H.BOOT_ADDRESS, 7) == 9
    b.close()


def case_near_call_panic_rolls_storage_back(B):
    """near_call.rs:27-46 (pc = imm0, handler = imm1, ergs 0 => all); ret.rs:35-41,196-251 panic: frame finished with
    panicked = true, pc := exception handler, LT flag; storage.rs:156-176 the frame's writes are undone in reverse."""
    p = Program()
    p.add(Imm(7), 0, 1)
    p.add(Imm(99), 0, 2)
    p.near_call(0, "body", "handler")
    p.label("after")
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.sload(1, 5)
    p.jump("after")
    p.label("body")
    p.sstore(1, 2)
    p.add(Imm(100), 0, 2)
    p.sstore(1, 2)
    p.ret(isa.RET_PANIC, R(0))
    b = H.launch(B, p, 1, storage=[(0, H.BOOT_ADDRESS, 7, 5)], ergs=1 << 20)
    r = H.rows(b)
    fams = [H.family_of(x) for x in r]
    assert fams == ["add", "add", "near_call", "log", "add", "log", "ret", "log", "jump", "ret"]
    nc, panic_ret, sload = r[2], r[6], r[7]
    assert int(nc["pc_after"]) == p.labels["body"] and int(nc["exception_handler"]) == p.labels["handler"]
    assert int(nc["callstack_depth"]) == 2 and int(nc["frame_bits"]) & 0x02            # is_local_frame (near_call.rs:59-63)
    assert int(panic_ret["pc_after"]) == p.labels["handler"] and int(panic_ret["flags_after"]) == 1
    assert int(panic_ret["callstack_depth"]) == 1
    assert H.val(sload["dst0"]) == 5                                                   # both writes rolled back
    fr = b.read_stream(0, records.STREAM_FRAME)
    assert [(int(f["kind"]), int(f["panicked"])) for f in fr] == [(1, 0), (1, 0), (2, 1), (2, 0)]
    assert b.read_storage(0, 0, H.BOOT_ADDRESS, 7) == 5
    # all ergs passed on a zero ABI, remainder returned on ret (near_call.rs:32-46, ret.rs:243)
    before = int(r[1]["ergs_after"])
    spent_inside = sum(isa.OPCODE_PRICES[int(x["raw_opcode"]) & 0x7FF] for x in r[3:7])
    nc_price = isa.OPCODE_PRICES[int(nc["raw_opcode"]) & 0x7FF]
    assert int(nc["ergs_after"]) == before - nc_price
    assert int(panic_ret["ergs_after"]) == before - nc_price - spent_inside
    b.close()


def case_refund_aware_storage_oracle(B):
    """SURVEY §8 row f-3.  log.rs:99-119: estimate_refunds_for_write is asked BEFORE the write executes, its answer goes
    to the tracer (helpers.rs:130-134) and its pubdata_refund() is subtracted from INITIAL_STORAGE_WRITE_PUBDATA_BYTES on
    the rollup shard.  The oracle behind ZkbConfig.reserved[1] answers RepeatedWrite(n) for a slot whose cold/warm marker
    is set; the markers are set by every executed read and write (storage.rs:105-110,126-131) and finish_frame never
    touches them (storage.rs:144-186), so a rolled-back write leaves its slot warm."""
    refund, epp = 24, 2
    full, net = epp * C.INITIAL_STORAGE_WRITE_PUBDATA_BYTES, epp * (C.INITIAL_STORAGE_WRITE_PUBDATA_BYTES - refund)
    p = Program()
    p.add(Imm(7), 0, 1)
    p.add(Imm(11), 0, 4)
    p.add(Imm(13), 0, 6)
    p.add(Imm(99), 0, 2)
    p.sstore(1, 2)                       # W1  key 7 cold            -> None, full price
    p.sstore(1, 4)                       # W2  key 7 warm (W1)       -> RepeatedWrite
    p.sload(4, 5)                        #     key 11 absent: the read sets its marker
    p.sstore(4, 2)                       # W3  key 11 warm (read)    -> RepeatedWrite
    p.near_call(0, "body", "handler")
    p.label("after")
    p.sstore(6, 4)                       # W5  key 13 warm (W4, rolled back: the marker stays) -> RepeatedWrite
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.jump("after")
    p.label("body")
    p.sstore(6, 2)                       # W4  key 13 cold           -> None
    p.ret(isa.RET_PANIC, R(0))
    for policy in (refund, 0):
        b = H.launch(B, p, 1, ergs=1 << 20, ergs_per_pubdata=epp, cfg_over=dict(warm_write_refund_bytes=policy))
        r = H.rows(b)
        assert [H.family_of(x) for x in r] == ["add"] * 4 + ["log", "log", "log", "log", "near_call", "log", "ret", "jump", "log", "ret"]
        rf = b.read_stream(0, records.STREAM_REFUND)
        want = [(0, 0), (1, refund), (1, refund), (0, 0), (1, refund)] if policy else [(0, 0)] * 5
        assert [(int(x["refund_type"]), int(x["refund_value"])) for x in rf] == want
        # ergs: every SSTORE pays its opcode price + ergs_per_pubdata * net bytes; spent_pubdata accumulates the same
        price = isa.OPCODE_PRICES[isa.VARIANT_INDEX[(isa.LOG, isa.LOG_SSTORE, isa.SRC_REG, isa.DST_REG, 0)]]
        writes = [r[4], r[5], r[7], r[9], r[12]]
        prev = [r[3], r[4], r[6], r[8], r[11]]
        costs = [full, net, net, full, net] if policy else [full] * 5
        spent = 0
        for w, pv, c, i in zip(writes, prev, costs, range(5)):
            if i not in (3, 4):                         # W4 / W5 sit next to frame switches: checked through spent_pubdata below
                assert int(pv["ergs_after"]) - int(w["ergs_after"]) == price + c
            spent += c
            assert int(w["spent_pubdata"]) == spent    # spent_pubdata_counter is never rolled back (log.rs:136-147)
        assert b.read_storage(0, 0, H.BOOT_ADDRESS, 7) == 11 and b.read_storage(0, 0, H.BOOT_ADDRESS, 11) == 99
        assert b.read_storage(0, 0, H.BOOT_ADDRESS, 13) == 11       # W4 rolled back, W5 applied
        lg = b.read_stream(0, records.STREAM_LOG)
        assert [int(x["rw_flag"]) for x in lg] == [1, 1, 0, 1, 1, 1]
        assert H.val(lg[5]["read_value"]) == 0                      # W5 saw the rolled-back slot
        b.close()


def case_ret_label_bounds_and_unidirectional_forwarding(B):
    """ret.rs quirks.
    (a) :202 the to-label variant is honoured only when the FINISHED frame is local: pc := imm0 instead of the saved pc;
        :254-259 a near-call return propagates the (grown) heap bound to the caller's frame; :243 unspent ergs return.
    (b) :202 for a far frame the label is ignored: the caller resumes at its saved pc.
    (c) :61-75 returning a FORWARDED pointer whose page lies below the callee's own base page (its calldata) is a
        panic: the frame finishes with panicked = true, pc := the far call's exception handler, LT flag set (:262-264),
        r1 := the empty fat pointer (still marked as a pointer, :213-218), everything else zeroed."""
    price = lambda row: isa.OPCODE_PRICES[int(row["raw_opcode"]) & 0x7FF]
    # ---- (a)
    p = Program()
    p.add(Imm(77), 0, 2)
    p.near_call(0, "body", "handler")
    p.label("after")
    p.jump("bad")                              # the saved pc: a label return must not come back here
    p.label("handler")
    p.jump("bad")
    p.label("lbl")
    p.ret(isa.RET_OK, R(0))
    p.label("bad")
    p.ret(isa.RET_PANIC, R(0))
    p.label("body")
    p.st(Imm(5000), 2)                         # grows the heap bound 4096 -> 5032 inside the near frame
    p.ret(isa.RET_OK, R(0), label="lbl")
    b = H.launch(B, p, 1, ergs=1 << 20, heap_bound=4096, cfg_over=dict(heap_bytes=8192))   # the device build's heap slabs are bounded
    r = H.rows(b)
    assert [H.family_of(x) for x in r] == ["add", "near_call", "uma", "ret", "ret"]
    nc, st, nret = r[1], r[2], r[3]
    assert int(st["heap_bound"]) == 5032 and int(st["callstack_depth"]) == 2
    assert int(nret["pc_after"]) == p.labels["lbl"] and int(nret["callstack_depth"]) == 1
    assert int(nret["heap_bound"]) == 5032 and int(nret["flags_after"]) == 0          # bound propagated, no panic flag
    growth = 5032 - 4096
    assert int(nret["ergs_after"]) == int(r[0]["ergs_after"]) - price(nc) - price(st) - growth - price(nret)
    assert b.vm_status()[0, 0] == 1
    b.close()
    # ---- (b) + (c)
    ignores_label = Program()
    ignores_label.ret(isa.RET_OK, R(0), label=7)
    forwards_calldata = Program()
    forwards_calldata.const("fwd", C.FWD_FORWARD_FAT_POINTER << 224)
    forwards_calldata.add(Code("fwd"), 0, 4)
    forwards_calldata.ptr(isa.PTR_PACK, 1, 4, 3)        # low 128 of r1 (the calldata pointer) | forwarding byte; still a pointer
    forwards_calldata.ret(isa.RET_OK, R(3))
    p = Program()
    p.const("abi", far_call_abi(1 << 16, start=0, length=64))
    p.add(Code("abi"), 0, 1)
    p.add(Imm(0x1111), 0, 2)
    p.far_call(R(1), 2, "h1")
    p.add(Code("abi"), 0, 1)                   # pc 3: where a well-behaved return resumes
    p.add(Imm(0x2222), 0, 2)
    p.add(Imm(55), 0, 9)                       # r9 must be zeroed by the panicking return
    p.far_call(R(1), 2, "h2")
    p.ret(isa.RET_OK, R(0))                    # not reached: the second callee panics
    p.label("h1")
    p.ret(isa.RET_PANIC, R(0))
    p.label("h2")
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, ergs=1 << 24, heap_bound=4096, contracts={0x1111: ignores_label, 0x2222: forwards_calldata})
    r = H.rows(b)
    fams = [H.family_of(x) for x in r]
    assert fams == ["add", "add", "far_call", "ret", "add", "add", "add", "far_call", "add", "ptr", "ret", "ret"]
    ret_b = r[3]
    assert int(ret_b["pc_after"]) == 3 and int(ret_b["flags_after"]) == 0 and int(ret_b["callstack_depth"]) == 1
    ret_c = r[10]
    assert int(ret_c["bits"]) & records_bit("SRC0_PTR")                                # it WAS a pointer (ret.rs:61 passes)
    assert int(ret_c["pc_after"]) == p.labels["h2"] and int(ret_c["flags_after"]) == 1 and int(ret_c["callstack_depth"]) == 1
    assert H.val(ret_c["dst0"]) == 0 and int(ret_c["bits"]) & records_bit("DST0_PTR")   # r1 = FatPointer::empty(), is_pointer
    fr = b.read_stream(0, records.STREAM_FRAME)
    assert [(int(f["kind"]), int(f["panicked"])) for f in fr] == [(1, 0), (1, 0), (2, 0), (1, 0), (2, 1), (2, 0)]
    ls = b.read_local_state(0)
    assert all(int(x) == 0 for x in np.asarray(ls.registers).reshape(15, 8)[8])        # r9 zeroed (ret.rs:224-231)
    b.close()


def case_far_call_delegate_and_mimic(B):
    """far_call.rs:505-534: Normal -> (this, sender) = (callee, caller); Delegate -> the callee runs with the CALLER's
    this / msg_sender and the caller FRAME's context value, only code_address is the callee's; Mimic (kernel only) ->
    msg_sender comes from r15 (CALL_IMPLICIT_PARAMETER_REG_IDX).  :524-534 Normal / Mimic take the context value from
    the context_u128 REGISTER, :558 which every far call then zeroes."""
    d = Program()
    d.context(isa.CTX_THIS, 1)
    d.context(isa.CTX_CALLER, 2)
    d.context(isa.CTX_CODE_ADDRESS, 3)
    d.context(isa.CTX_GET_U128, 4)
    d.ret(isa.RET_OK, R(0))
    a = Program()
    a.const("abi", far_call_abi(1 << 18))
    a.add(Code("abi"), 0, 1)
    a.add(Imm(0x2222), 0, 2)
    a.far_call(R(1), 2, "fail", sub=isa.FC_DELEGATE)
    a.context(isa.CTX_GET_U128, 6)
    a.ret(isa.RET_OK, R(0))
    a.label("fail")
    a.ret(isa.RET_PANIC, R(0))
    m = Program()
    m.context(isa.CTX_THIS, 1)
    m.context(isa.CTX_CALLER, 2)
    m.context(isa.CTX_GET_U128, 4)
    m.ret(isa.RET_OK, R(0))
    p = Program()
    p.const("abi", far_call_abi(1 << 20))
    p.add(Imm(0xABCD), 0, 5)
    p.context(isa.CTX_SET_U128, 0, 5)
    p.add(Code("abi"), 0, 1)
    p.add(Imm(0x1111), 0, 2)
    p.far_call(R(1), 2, "fail")
    p.add(Imm(0x55), 0, 5)
    p.context(isa.CTX_SET_U128, 0, 5)
    p.add(Code("abi"), 0, 1)
    p.add(Imm(0x3333), 0, 2)
    p.add(Imm(0x7777), 0, 15)
    p.far_call(R(1), 2, "fail", sub=isa.FC_MIMIC)
    p.ret(isa.RET_OK, R(0))
    p.label("fail")
    p.ret(isa.RET_PANIC, R(0))
    b = H.launch(B, p, 1, ergs=1 << 24, contracts={0x1111: a, 0x2222: d, 0x3333: m})
    r = H.rows(b)
    fams = [H.family_of(x) for x in r]
    assert fams == ["add", "context", "add", "add", "far_call", "add", "add", "far_call"] + ["context"] * 4 + ["ret", "context", "ret"] + \
        ["add", "context", "add", "add", "add", "far_call"] + ["context"] * 3 + ["ret", "ret"]
    assert all(int(x["error_flags"]) == 0 for x in r)
    assert H.val(r[1]["context_u128"]) == 0xABCD and H.val(r[4]["context_u128"]) == 0          # set, then zeroed by the call
    # inside the delegate callee: the caller's identity, the callee's code
    assert [H.val(r[i]["dst0"]) for i in (8, 9, 10, 11)] == [0x1111, H.BOOT_ADDRESS, 0x2222, 0xABCD]
    assert H.val(r[13]["dst0"]) == 0xABCD                                                      # A's own frame value
    # inside the mimic callee: msg_sender from r15, context value from the register
    assert [H.val(r[i]["dst0"]) for i in (21, 22, 23)] == [0x3333, 0x7777, 0x55]
    assert b.vm_status()[0, 0] == 1
    b.close()


def case_far_call_to_malformed_code_hash(B):
    """far_call.rs:176-246 a stored code hash that is not a versioned ContractCodeSha256 hash (version byte != 1) is an
    exception: no decommit (:435-439 the callee's code page is UNMAPPED_PAGE), set_shorthand_panic (helpers.rs:336-338);
    :502-503 the page counter still moves; the new frame still starts and the NEXT cycle executes the exception-revert
    encoding in it without a code fetch (cycle.rs:104-115): the frame finishes as panicked, pc := the call's handler,
    LT flag (ret.rs:262-264)."""
    target = 0xDEAD4444                                # not a kernel address: a zero hash would fall back to the default AA
    p = Program()
    p.const("abi", far_call_abi(1 << 16))
    p.const("target", target)
    p.add(Code("abi"), 0, 1)
    p.add(Code("target"), 0, 2)
    p.far_call(R(1), 2, "handler")
    p.ret(isa.RET_PANIC, R(0))                         # pc 3: never reached
    p.label("handler")
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, ergs=1 << 20, storage=[(0, C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, target, 0x1234)])
    r = H.rows(b)
    assert [H.family_of(x) for x in r] == ["add", "add", "far_call", "ret", "ret"]
    fc, ex, last = r[2], r[3], r[4]
    assert int(fc["callstack_depth"]) == 2 and int(fc["code_page"]) == C.UNMAPPED_PAGE and int(fc["pc_after"]) == 0
    assert int(fc["memory_page_counter"]) == 1024 + C.NEW_MEMORY_PAGES_PER_FAR_CALL
    assert int(fc["bits"]) & records_bit("PENDING") and int(fc["n_log"]) == 1
    assert len(b.read_stream(0, records.STREAM_DECOMMIT)) == 0
    assert int(ex["raw_opcode"]) == isa.EXCEPTION_REVERT_ENCODING and int(ex["masked_variant"]) == isa.PANIC_VARIANT_IDX
    assert int(ex["n_mem"]) == 0 and int(ex["error_flags"]) == 0                       # no fetch, not an error mask
    assert int(ex["callstack_depth"]) == 1 and int(ex["pc_after"]) == p.labels["handler"] and int(ex["flags_after"]) == 1
    assert int(ex["bits"]) & records_bit("PENDING") == 0
    fr = b.read_stream(0, records.STREAM_FRAME)
    assert [(int(f["kind"]), int(f["panicked"])) for f in fr] == [(1, 0), (1, 0), (2, 1), (2, 0)]
    assert int(last["callstack_depth"]) == 0 and b.vm_status()[0, 0] == 1
    b.close()


def callee_returning(value_word: int, sub=isa.RET_OK) -> Program:
    c = Program()
    c.const("v", value_word)
    c.const("abi", ret_abi(start=0, length=32))
    c.add(Code("v"), 0, 2)
    c.st(Imm(0), 2)
    c.add(Code("abi"), 0, 3)
    c.ret(sub, R(3))
    return c


def case_far_call_ergs_and_decommit_refund(B):
    """far_call.rs:468-487 callee ergs = min(requested, floor(e / 64) * 63); :423-433 decommit cost = 4 ergs per code
    word, :450-453 refunded when the hash was already decommitted (decommitter.rs:38-47 => is_fresh = false, the tracer
    is still told, quirk 14); :502-503 page counter += 8 per call; :573-610 r1 = calldata pointer, others cleared;
    ret.rs:213-236 r1 = returndata pointer, r2.. = 0."""
    callee = callee_returning(0xABCDEF)
    n_words = len(callee.bytecode()) // 32
    p = Program()
    p.const("abi", far_call_abi(0xFFFFFFFF, start=0, length=32))
    p.const("callee", USER)
    for _ in range(2):
        p.add(Code("abi"), 0, 8)
        p.add(Code("callee"), 0, 7)
        p.add(Imm(55), 0, 9)
        p.far_call(R(8), 7, "fail")
        p.ld_ptr(R(1), 4)
    p.ret(isa.RET_OK, R(0))
    p.label("fail")
    p.ret(isa.RET_PANIC, R(0))
    b = H.launch(B, p, 1, contracts={USER: callee}, ergs=1 << 24, heap_bound=64)
    r = H.rows(b)
    calls = [i for i, x in enumerate(r) if H.family_of(x) == "far_call"]
    assert len(calls) == 2
    fc_price = isa.OPCODE_PRICES[isa.VARIANT_INDEX[(isa.FAR_CALL, isa.FC_NORMAL, isa.SRC_REG, isa.DST_REG, 0)]]
    dec = b.read_stream(0, records.STREAM_DECOMMIT)
    assert [int(d["is_fresh"]) for d in dec] == [1, 0] and int(dec[0]["memory_page"]) == int(dec[1]["memory_page"])
    assert all(int(d["decommitted_length"]) == n_words for d in dec)
    fr = b.read_stream(0, records.STREAM_FRAME)
    for k, i in enumerate(calls):
        e_before = int(r[i - 1]["ergs_after"]) - fc_price
        e_after_decommit = e_before - (C.ERGS_PER_CODE_WORD_DECOMMITTMENT * n_words if k == 0 else 0)
        passed = (e_after_decommit // 64) * 63
        new_frame = [f for f in fr if int(f["kind"]) == 1 and int(f["cycle"]) == int(r[i]["cycle"])][0]
        assert int(new_frame["ergs_remaining"]) == passed == int(r[i]["ergs_after"])
        assert int(new_frame["prev_ergs_remaining"]) == e_after_decommit - passed
        assert int(r[i]["memory_page_counter"]) == 1024 + 8 * (k + 1) and int(new_frame["base_memory_page"]) == 1024 + 8 * k
        assert int(new_frame["heap_bound"]) == C.NEW_FRAME_MEMORY_STIPEND and int(new_frame["is_local_frame"]) == 0
        assert bytes(new_frame["msg_sender"]) == H.BOOT_ADDRESS.to_bytes(20, "big")
        assert bytes(new_frame["this_address"]) == USER.to_bytes(20, "big")
        assert int(r[i]["bits"]) & records_bit("DST0_PTR") and H.val(r[i]["dst0"]) == (32 << 96) | ((H.BOOT_PAGE + 2) << 32)
        assert int(r[i]["pc_after"]) == 0 and int(r[i]["sp_after"]) == 0
    rets = [x for x in r if H.family_of(x) == "ret" and int(x["callstack_depth"]) == 1]
    assert len(rets) == 2 and all(int(x["bits"]) & records_bit("DST0_PTR") for x in rets)
    loads = [x for x in r if H.family_of(x) == "uma" and int(x["callstack_depth"]) == 1]
    assert [H.val(x["dst0"]) for x in loads] == [0xABCDEF, 0xABCDEF]
    st = b.read_local_state(0)
    b.close()


def case_static_and_kernel_violations(B):
    """cycle.rs:173-179: kernel-only opcode outside kernel mode => PRIVILAGED_ACCESS (4); writes in a static frame =>
    WRITE_IN_STATIC_CONTEXT (8); both mask to panic.  far_call.rs:500: static is sticky for the callee."""
    callee = Program()
    callee.add(Imm(1), 0, 1)
    callee.sstore(1, 1)
    callee.ret(isa.RET_OK, R(0))
    p = Program()
    p.const("abi", far_call_abi(0xFFFFFFFF))
    p.const("callee", USER)
    p.add(Code("abi"), 0, 8)
    p.add(Code("callee"), 0, 7)
    p.far_call(R(8), 7, "handler", static=True)
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.add(Imm(77), 0, 6)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, contracts={USER: callee}, ergs=1 << 22)
    r = H.rows(b)
    bad = [x for x in r if int(x["error_flags"])]
    assert len(bad) == 1 and int(bad[0]["error_flags"]) == 8 and int(bad[0]["masked_variant"]) == isa.PANIC_VARIANT_IDX
    assert int(bad[0]["pc_after"]) == p.labels["handler"] and int(bad[0]["flags_after"]) == 1
    assert len(b.read_stream(0, records.STREAM_LOG)) == 1                  # only the code-hash read (far_call.rs:131-145)
    b.close()
    q = Program()
    q.event(1, 2)
    q.ret(isa.RET_OK, R(0))
    b = H.launch(B, q, 1, this_address=USER)
    r = H.rows(b)
    assert int(r[0]["error_flags"]) == 4 and int(r[0]["masked_variant"]) == isa.PANIC_VARIANT_IDX and len(r) == 1
    b.close()


def case_uma_unaligned_and_fat_pointer_tail(B):
    """uma.rs:232-238,299-303 unaligned load = (w0 << 8u) | (w1 >> 8(32-u)); :349-400 unaligned store merges into both
    words; :152-217 growth = max(0, offset + 32 - bound) at 1 erg/byte; :110-116 fat-pointer read past the slice
    returns 0 without an exception; :305-320 bytes beyond `length` read as zero."""
    w0 = int.from_bytes(bytes(range(1, 33)), "big")
    w1 = int.from_bytes(bytes(range(33, 65)), "big")
    p = Program()
    p.ld(Imm(5), 1)                                           # unaligned heap load inside the bound
    p.add(Imm(0), 0, 2)
    p.sub(Imm(1), 2, 2, swap=True)                            # r2 = 0 - 1 = all ones
    p.st(Imm(100), 2)                                         # unaligned store: grows the heap to 132
    p.ld(Imm(96), 3)
    p.ld(Imm(128), 4)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, heap=w0.to_bytes(32, "big") + w1.to_bytes(32, "big"), heap_bound=64, ergs=1 << 20)
    r = H.rows(b)
    whole = (w0 << 256) | w1
    assert H.val(r[0]["dst0"]) == (whole >> (8 * (32 - 5))) & M256 and int(r[0]["n_mem"]) == 3        # fetch + 2 words
    st_row = r[3]
    st_price = isa.OPCODE_PRICES[int(st_row["raw_opcode"]) & 0x7FF]
    assert int(st_row["heap_bound"]) == 132 and int(r[2]["ergs_after"]) - int(st_row["ergs_after"]) == st_price + (132 - 64)
    assert int(st_row["n_mem"]) == 4                                                                # 2 reads + 2 writes
    assert H.val(r[4]["dst0"]) == (1 << (8 * 28)) - 1                                              # bytes 100..127 set
    assert H.val(r[5]["dst0"]) == ((1 << 32) - 1) << (8 * 28)                                        # bytes 128..131 set
    b.close()
    # fat pointer: slice [start 8, length 40) of the caller's heap, offset walking past the end
    callee = Program()
    callee.ld_ptr(R(1), 2, 1, inc=True)                       # bytes 8..39 ; r1.offset = 32
    callee.ld_ptr(R(1), 3, 1, inc=True)                       # 8 bytes left: tail zeroed ; offset = 64
    callee.ld_ptr(R(1), 4)                                    # offset >= length: reads 0, no exception
    callee.ret(isa.RET_OK, R(0))
    p = Program()
    p.const("abi", far_call_abi(0xFFFFFFFF, start=8, length=40))
    p.const("callee", USER)
    p.add(Code("abi"), 0, 8)
    p.add(Code("callee"), 0, 7)
    p.far_call(R(8), 7, "fail")
    p.ret(isa.RET_OK, R(0))
    p.label("fail")
    p.ret(isa.RET_PANIC, R(0))
    b = H.launch(B, p, 1, contracts={USER: callee}, heap=w0.to_bytes(32, "big") + w1.to_bytes(32, "big"), heap_bound=64,
                 ergs=1 << 22)
    r = [x for x in H.rows(b) if H.family_of(x) == "uma"]
    assert H.val(r[0]["dst0"]) == (whole >> (8 * 24)) & M256
    assert H.val(r[1]["dst0"]) == int.from_bytes(bytes(range(41, 49)) + bytes(24), "big")     # heap bytes 40..47, then zeros
    assert H.val(r[2]["dst0"]) == 0 and int(r[2]["n_mem"]) == 0 and not int(r[2]["bits"]) & records_bit("PENDING")
    assert H.val(r[0]["dst1"]) & 0xFFFFFFFF == 32 and int(r[0]["bits"]) & records_bit("DST1_PTR")      # uma.rs:335-344
    assert not any(int(x["error_flags"]) for x in H.rows(b))
    b.close()


def case_uma_exceptions_mutate_bound_and_ergs_first(B):
    """uma.rs:152-217 the heap bound is raised and the growth charged BEFORE the exceptions are looked at; :195-217 a
    shortfall zeroes the frame's ergs (NOT_ENOUGH_ERGS_TO_GROW_MEMORY), :127-132 an offset above MAX_OFFSET_TO_DEREF costs
    u32::MAX (so it always ends the same way) and leaves the bound alone; :345-347 any exception = no register write,
    no memory access, set_shorthand_panic: the NEXT cycle runs the exception-revert encoding (cycle.rs:104-115), whose
    own price can no longer be paid (cycle.rs:147-163: error flag NOT_ENOUGH_ERGS) before it panics the near frame;
    ret.rs:254-259 the raised bound survives the panicking near return."""
    def run(address_value):
        p = Program()
        p.const("addr", address_value)
        p.add(Imm(9), 0, 3)
        p.add(Imm(1000), 0, 4)
        p.add(Code("addr"), 0, 1)
        p.near_call(4, "body", "handler")          # passes exactly 1000 ergs
        p.label("after")
        p.ret(isa.RET_PANIC, R(0))                 # not reached
        p.label("handler")
        p.ret(isa.RET_OK, R(0))
        p.label("body")
        p.ld(R(1), 3)
        p.ret(isa.RET_OK, R(0))                    # not reached
        b = H.launch(B, p, 1, ergs=1 << 20, heap_bound=4096, cfg_over=dict(heap_bytes=8192))
        r = H.rows(b)
        assert [H.family_of(x) for x in r] == ["add", "add", "add", "near_call", "uma", "ret", "ret"]
        return p, b, r
    # (i) growth the frame cannot pay for: 6000 + 32 - 4096 = 1936 > 1000 - price
    p, b, r = run(6000)
    nc, uma, ex = r[3], r[4], r[5]
    assert int(nc["ergs_after"]) == 1000
    assert int(uma["heap_bound"]) == 6032 and int(uma["ergs_after"]) == 0
    assert int(uma["bits"]) & records_bit("PENDING") and not int(uma["bits"]) & records_bit("DST0_VALID")
    assert int(uma["n_mem"]) == 1                                                      # the code fetch only: access skipped
    assert int(ex["raw_opcode"]) == isa.EXCEPTION_REVERT_ENCODING and int(ex["masked_variant"]) == isa.PANIC_VARIANT_IDX
    assert int(ex["error_flags"]) == 2 and int(ex["n_mem"]) == 0
    assert int(ex["callstack_depth"]) == 1 and int(ex["pc_after"]) == p.labels["handler"] and int(ex["flags_after"]) == 1
    assert int(ex["heap_bound"]) == 6032                                               # propagated by the near return
    nc_price = isa.OPCODE_PRICES[int(nc["raw_opcode"]) & 0x7FF]
    assert int(ex["ergs_after"]) == int(r[2]["ergs_after"]) - nc_price - 1000            # nothing came back
    b.close()
    # (ii) offset above MAX_OFFSET_TO_DEREF: offset + 32 wraps, the bound stays, the cost is u32::MAX
    p, b, r = run((1 << 32) - 20)
    uma, ex = r[4], r[5]
    assert (1 << 32) - 20 > C.MAX_OFFSET_TO_DEREF
    assert int(uma["heap_bound"]) == 4096 and int(uma["ergs_after"]) == 0
    assert int(uma["bits"]) & records_bit("PENDING") and not int(uma["bits"]) & records_bit("DST0_VALID") and int(uma["n_mem"]) == 1
    assert int(ex["masked_variant"]) == isa.PANIC_VARIANT_IDX and int(ex["pc_after"]) == p.labels["handler"]
    assert int(ex["heap_bound"]) == 4096 and int(ex["flags_after"]) == 1
    b.close()


def case_near_call_ergs_clamp_and_precompile_shortfall(B):
    """near_call.rs:32-46 asking for more ergs than the frame has passes everything and leaves 0 (it is not an error);
    log.rs:120-147,255-264 a precompile call whose extra cost (src1.low_u32) exceeds the frame's ergs zeroes them and
    writes 0 to dst0 WITHOUT a log query or a precompile run; the opcode flags reset at the near call (near_call.rs:24)."""
    p = Program()
    p.const("big", 0xFFFFFFFF)
    p.add(Imm(0), 0, 5)
    p.sub(Imm(1), 5, 5, set_flags=True, swap=True)     # 0 - 1: sets LT so that the reset below is observable
    p.add(Code("big"), 0, 4)
    p.near_call(4, "body", "handler")                  # wants 2^32 - 1 ergs
    p.label("after")
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.ret(isa.RET_OK, R(0))
    p.label("body")
    p.add(Imm(7), 0, 6)
    p.precompile(0, 4, 6)                              # extra cost r4 = 2^32 - 1 > ergs: dst0 (r6) := 0
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, ergs=1 << 20)
    r = H.rows(b)
    assert [H.family_of(x) for x in r] == ["add", "sub", "add", "near_call", "add", "log", "ret", "ret"]
    assert int(r[1]["flags_after"]) == 1
    nc, pre, nret = r[3], r[5], r[6]
    nc_price = isa.OPCODE_PRICES[int(nc["raw_opcode"]) & 0x7FF]
    assert int(nc["flags_after"]) == 0
    assert int(nc["ergs_after"]) == int(r[2]["ergs_after"]) - nc_price           # the callee got everything that was left
    fr = b.read_stream(0, records.STREAM_FRAME)
    near = [f for f in fr if int(f["kind"]) == 1 and int(f["is_local_frame"]) == 1][0]
    assert int(near["prev_ergs_remaining"]) == 0 and int(near["ergs_remaining"]) == int(nc["ergs_after"])
    assert int(pre["ergs_after"]) == 0 and int(pre["n_log"]) == 0 and int(pre["n_mem"]) <= 1
    assert int(pre["bits"]) & records_bit("DST0_VALID") and H.val(pre["dst0"]) == 0 and not int(pre["bits"]) & records_bit("PENDING")
    assert len(b.read_stream(0, records.STREAM_LOG)) == 0
    # the near frame's `ret.ok` can no longer pay its own price: NOT_ENOUGH_ERGS masks it into a panic (cycle.rs:147-163)
    assert int(nret["error_flags"]) == 2 and int(nret["masked_variant"]) == isa.PANIC_VARIANT_IDX
    assert int(nret["pc_after"]) == p.labels["handler"] and int(nret["callstack_depth"]) == 1 and int(nret["ergs_after"]) == 0
    assert b.vm_status()[0, 0] == 1
    b.close()


def case_context_setters_reach_the_log_queries(B):
    """context.rs:41-50 (kernel only) set_ergs_per_pubdata_byte := src0.low_u32(), increment_tx_number := wrapping +1;
    log.rs:72-74,87 every LogQuery carries the current tx_number_in_block, log.rs:119 the pubdata cost of a storage write
    is ergs_per_pubdata * 64 and goes to spent_pubdata_counter (:146); aux bytes: storage 0, event 1 (log.rs:88,236)."""
    p = Program()
    p.add(Imm(5), 0, 1)
    p.context(isa.CTX_SET_ERGS_PER_PUBDATA, 0, 1)
    p.context(isa.CTX_INC_TX)
    p.context(isa.CTX_INC_TX)
    p.add(Imm(7), 0, 2)
    p.add(Imm(9), 0, 3)
    p.sstore(2, 3)
    p.event(2, 3, first=True)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, ergs=1 << 20)
    r = H.rows(b)
    assert [H.family_of(x) for x in r] == ["add", "context", "context", "context", "add", "add", "log", "log", "ret"]
    assert all(int(x["error_flags"]) == 0 for x in r)
    assert [int(x["ergs_per_pubdata"]) for x in r[:3]] == [0, 5, 5]
    assert [int(x["tx_number"]) for x in r[1:5]] == [0, 1, 2, 2]
    sst = r[6]
    price = isa.OPCODE_PRICES[int(sst["raw_opcode"]) & 0x7FF]
    cost = 5 * C.INITIAL_STORAGE_WRITE_PUBDATA_BYTES
    assert int(r[5]["ergs_after"]) - int(sst["ergs_after"]) == price + cost and int(sst["spent_pubdata"]) == cost
    assert int(r[7]["spent_pubdata"]) == cost                                        # events cost no pubdata
    lg = b.read_stream(0, records.STREAM_LOG)
    assert [int(x["tx_number_in_block"]) for x in lg] == [2, 2]
    assert [int(x["aux_byte"]) for x in lg] == [C.STORAGE_AUX_BYTE, C.EVENT_AUX_BYTE]
    assert [int(x["is_service"]) for x in lg] == [0, 1]                               # FIRST_MESSAGE_FLAG (log.rs:70)
    assert H.val(lg[1]["key"]) == 7 and H.val(lg[1]["written_value"]) == 9 and bytes(lg[1]["address"]) == H.BOOT_ADDRESS.to_bytes(20, "big")
    b.close()


def case_ptr_arithmetic_and_its_panics(B):
    """ptr.rs:33-94 add / sub move FatPointer.offset (u32, checked), keep the rest of the low 128 bits and src0's high
    128 bits, result is a pointer; :96-139 pack = low128(src0) | high128(src1) and needs src1.low_u128() == 0;
    :140-192 shrink takes src1 off the length.  Panics (set_shorthand_panic, no dst write): src0 not a pointer (:35),
    src1 a pointer (:41), src1 >= MAX_OFFSET_FOR_ADD_SUB (:47), offset under/overflow (:65-74), pack with dirty low half
    (:110), shrink below zero."""
    page = H.BOOT_PAGE + 2
    ptr0 = (page << 32) | (8 << 64) | (40 << 96)               # offset 0, caller's heap page, start 8, length 40

    def call(callee):
        p = Program()
        p.const("abi", far_call_abi(1 << 16, start=8, length=40))
        p.const("callee", USER)
        p.add(Code("abi"), 0, 8)
        p.add(Code("callee"), 0, 7)
        p.far_call(R(8), 7, "handler")
        p.ret(isa.RET_OK, R(0))
        p.label("handler")
        p.ret(isa.RET_OK, R(0))
        b = H.launch(B, p, 1, contracts={USER: callee}, ergs=1 << 22, heap_bound=64)
        return p, b, [x for x in H.rows(b) if int(x["callstack_depth"]) == 2 or H.family_of(x) == "ptr"], H.rows(b)

    ok = Program()
    ok.const("hi", 0xAB << 128)
    ok.add(Imm(8), 0, 2)
    ok.ptr(isa.PTR_ADD, 1, 2, 3)                               # offset 8
    ok.add(Imm(3), 0, 2)
    ok.ptr(isa.PTR_SUB, 3, 2, 4)                               # offset 5
    ok.add(Imm(10), 0, 2)
    ok.ptr(isa.PTR_SHRINK, 1, 2, 5)                            # length 30
    ok.add(Code("hi"), 0, 6)
    ok.ptr(isa.PTR_PACK, 1, 6, 7)
    ok.ret(isa.RET_OK, R(0))
    _, b, _, rows = call(ok)
    pr = [x for x in rows if H.family_of(x) == "ptr"]
    assert [H.val(x["dst0"]) for x in pr] == [ptr0 | 8, ptr0 | 5, (ptr0 & ~(0xFFFFFFFF << 96)) | (30 << 96), ptr0 | (0xAB << 128)]
    assert all(int(x["bits"]) & records_bit("DST0_PTR") and not int(x["bits"]) & records_bit("PENDING") for x in pr)
    b.close()

    def failing(build):
        c = Program()
        c.const("big", 1 << 32)
        c.const("dirty", (0xAB << 128) | 1)
        build(c)
        c.ret(isa.RET_OK, R(0))                                # not reached
        p, b, _, rows = call(c)
        bad = [x for x in rows if H.family_of(x) == "ptr"][-1]
        nxt = rows[[int(x["cycle"]) for x in rows].index(int(bad["cycle"])) + 1]
        assert int(bad["bits"]) & records_bit("PENDING") and not int(bad["bits"]) & records_bit("DST0_VALID"), build.__name__
        assert int(nxt["raw_opcode"]) == isa.EXCEPTION_REVERT_ENCODING and int(nxt["pc_after"]) == p.labels["handler"]
        assert int(nxt["flags_after"]) == 1 and int(nxt["callstack_depth"]) == 1
        b.close()

    def src0_not_a_pointer(c):
        c.add(Imm(8), 0, 2)
        c.ptr(isa.PTR_ADD, 2, 2, 3)

    def src1_is_a_pointer(c):
        c.ptr(isa.PTR_ADD, 1, 1, 3)

    def offset_too_far(c):
        c.add(Code("big"), 0, 2)
        c.ptr(isa.PTR_ADD, 1, 2, 3)

    def offset_underflow(c):
        c.add(Imm(1), 0, 2)
        c.ptr(isa.PTR_SUB, 1, 2, 3)

    def pack_dirty_low_half(c):
        c.add(Code("dirty"), 0, 2)
        c.ptr(isa.PTR_PACK, 1, 2, 3)

    def shrink_below_zero(c):
        c.add(Imm(41), 0, 2)
        c.ptr(isa.PTR_SHRINK, 1, 2, 3)

    for build in (src0_not_a_pointer, src1_is_a_pointer, offset_too_far, offset_underflow, pack_dirty_low_half, shrink_below_zero):
        failing(build)


def case_l1_message_pubdata_and_decommit_shortfall(B):
    """log.rs:121-124 an L1 message costs ergs_per_pubdata * L1_MESSAGE_PUBDATA_BYTES on top of its price and counts as
    spent pubdata (:146); aux byte 2 (log.rs:228).  far_call.rs:423-433 a frame that cannot pay 4 ergs per code word of a
    FRESH decommitment gets NOT_ENOUGH_ERGS_TO_DECOMMIT: nothing is burnt ("do not burn"), nothing is decommitted, the
    callee frame starts on the unmapped page with 63/64 of what is left (:468-487) and panics in the next cycle."""
    callee = Program()
    for _ in range(400):
        callee.add(Imm(1), 1, 1)
    callee.ret(isa.RET_OK, R(0))
    n_words = len(callee.bytecode()) // 32
    p = Program()
    p.const("abi", far_call_abi(0xFFFFFFFF))
    p.const("callee", USER)
    p.add(Imm(2), 0, 1)
    p.context(isa.CTX_SET_ERGS_PER_PUBDATA, 0, 1)
    p.add(Imm(7), 0, 2)
    p.to_l1(2, 2)
    p.add(Imm(300), 0, 4)
    p.near_call(4, "body", "handler")                   # the body runs on 300 ergs
    p.label("after")
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.ret(isa.RET_OK, R(0))
    p.label("body")
    p.add(Code("abi"), 0, 8)
    p.add(Code("callee"), 0, 7)
    p.far_call(R(8), 7, "inner")
    p.label("inner")
    p.ret(isa.RET_PANIC, R(0))                          # handler of the far call, inside the near frame
    b = H.launch(B, p, 1, contracts={USER: callee}, ergs=1 << 20)
    r = H.rows(b)
    fams = [H.family_of(x) for x in r]
    assert fams == ["add", "context", "add", "log", "add", "near_call", "add", "add", "far_call", "ret", "ret", "ret"]
    l1 = r[3]
    cost = 2 * C.L1_MESSAGE_PUBDATA_BYTES
    assert int(r[2]["ergs_after"]) - int(l1["ergs_after"]) == isa.OPCODE_PRICES[int(l1["raw_opcode"]) & 0x7FF] + cost
    assert int(l1["spent_pubdata"]) == cost
    lg = b.read_stream(0, records.STREAM_LOG)
    assert int(lg[0]["aux_byte"]) == C.L1_MESSAGE_AUX_BYTE and int(lg[0]["rw_flag"]) == 1
    fc = r[8]
    fc_price = isa.OPCODE_PRICES[int(fc["raw_opcode"]) & 0x7FF]
    left = int(r[7]["ergs_after"]) - fc_price
    assert C.ERGS_PER_CODE_WORD_DECOMMITTMENT * n_words > left                       # the premise of this case
    assert int(fc["bits"]) & records_bit("PENDING") and int(fc["code_page"]) == C.UNMAPPED_PAGE
    assert int(fc["ergs_after"]) == (left // 64) * 63                                 # nothing burnt before the 63/64 split
    assert len(b.read_stream(0, records.STREAM_DECOMMIT)) == 0
    ex = r[9]
    assert int(ex["raw_opcode"]) == isa.EXCEPTION_REVERT_ENCODING and int(ex["pc_after"]) == p.labels["inner"]
    assert int(ex["callstack_depth"]) == 2 and int(ex["flags_after"]) == 1
    assert int(ex["ergs_after"]) == left - isa.OPCODE_PRICES[isa.PANIC_VARIANT_IDX]   # callee's rest came back (ret.rs:243)
    b.close()


def case_far_call_forwarding_modes(B):
    """far_call.rs:284-312 ForwardFatPointer narrows the forwarded pointer (start += offset, length -= offset, offset = 0)
    and keeps its page; UseAuxHeap points the fresh slice at the caller's aux heap page (base + 3); :247-256 forwarding a
    register that is not a pointer is INPUT_IS_NOT_POINTER_WHEN_EXPECTED: the pointer is masked to FatPointer::empty(),
    the callee frame starts on the unmapped page and panics in its first cycle."""
    heap = bytes(range(1, 65))
    boot_heap_page = H.BOOT_PAGE + 2
    callee_b = Program()
    callee_b.ld_ptr(R(1), 2)
    callee_b.ret(isa.RET_OK, R(0))
    a = Program()
    a.const("fwd_hi", ((1 << 16) << 192) | (C.FWD_FORWARD_FAT_POINTER << 224))
    a.const("aux_abi", far_call_abi(1 << 16, start=0, length=64, fwd=C.FWD_USE_AUX_HEAP))
    a.const("bad_abi", far_call_abi(1 << 16, fwd=C.FWD_FORWARD_FAT_POINTER))
    a.add(Imm(8), 0, 2)
    a.ptr(isa.PTR_ADD, 1, 2, 3)                         # offset 8 into the calldata slice [8, 48)
    a.add(Code("fwd_hi"), 0, 4)
    a.ptr(isa.PTR_PACK, 3, 4, 5)                        # pointer in the low half, ergs + forwarding byte in the high half
    a.add(Imm(0x2222), 0, 6)
    a.far_call(R(5), 6, "fail")                         # (1) forward
    a.add(Imm(0xA5A5), 0, 9)
    a.st_aux(Imm(0), 9)
    a.add(Code("aux_abi"), 0, 7)
    a.add(Imm(0x2222), 0, 6)
    a.far_call(R(7), 6, "fail")                         # (2) fresh slice of the aux heap
    a.add(Code("bad_abi"), 0, 8)
    a.add(Imm(0x2222), 0, 6)
    a.far_call(R(8), 6, "fail2")                        # (3) "forward" a plain value
    a.ret(isa.RET_PANIC, R(0))                          # not reached
    a.label("fail2")
    a.ret(isa.RET_OK, R(0))
    a.label("fail")
    a.ret(isa.RET_PANIC, R(0))
    p = Program()
    p.const("abi", far_call_abi(1 << 20, start=8, length=40))
    p.add(Code("abi"), 0, 1)
    p.add(Imm(0x1111), 0, 2)
    p.far_call(R(1), 2, "bfail")
    p.ret(isa.RET_OK, R(0))
    p.label("bfail")
    p.ret(isa.RET_PANIC, R(0))
    b = H.launch(B, p, 1, contracts={0x1111: a, 0x2222: callee_b}, heap=heap, heap_bound=64, ergs=1 << 24)
    r = H.rows(b)
    calls = [x for x in r if H.family_of(x) == "far_call"]
    loads = [x for x in r if H.family_of(x) == "uma" and int(x["callstack_depth"]) == 3]
    assert len(calls) == 4 and len(loads) == 2
    a_base = 1024
    assert H.val(calls[0]["dst0"]) == (boot_heap_page << 32) | (8 << 64) | (40 << 96)
    assert H.val(calls[1]["dst0"]) == (boot_heap_page << 32) | (16 << 64) | (32 << 96) and int(calls[1]["bits"]) & records_bit("DST0_PTR")
    assert H.val(loads[0]["dst0"]) == int.from_bytes(heap[16:48], "big")
    assert H.val(calls[2]["dst0"]) == ((a_base + 3) << 32) | (0 << 64) | (64 << 96)
    assert H.val(loads[1]["dst0"]) == 0xA5A5
    bad = calls[3]
    assert H.val(bad["dst0"]) == 0 and int(bad["bits"]) & records_bit("PENDING") and int(bad["code_page"]) == C.UNMAPPED_PAGE
    nxt = r[[int(x["cycle"]) for x in r].index(int(bad["cycle"])) + 1]
    assert int(nxt["raw_opcode"]) == isa.EXCEPTION_REVERT_ENCODING and int(nxt["pc_after"]) == a.labels["fail2"] and int(nxt["flags_after"]) == 1
    assert b.vm_status()[0, 0] == 1 and int(r[-1]["callstack_depth"]) == 0 and int(r[-1]["flags_after"]) == 0
    b.close()


def case_far_revert_returns_data_and_rolls_storage_back(B):
    """ret.rs:232 Revert finishes the frame as panicked (storage.rs:156-176 its writes are undone) and continues at the
    call's exception handler (:248), but unlike Panic it keeps its returndata (:35-41 only Panic zeroes src0) and does
    not set the LT flag (:262-264); helpers.rs:255-261 the tracer sees finish_execution_context(panicked = true)."""
    callee = Program()
    callee.const("abi", ret_abi(start=0, length=32))
    callee.add(Imm(7), 0, 1)
    callee.add(Imm(99), 0, 2)
    callee.sstore(1, 2)                                 # 5 -> 99 under the callee's own address: rolled back
    callee.st(Imm(0), 2)                                # returndata word = 99
    callee.add(Code("abi"), 0, 3)
    callee.ret(isa.RET_REVERT, R(3))
    p = Program()
    p.const("abi", far_call_abi(1 << 20))
    p.const("callee", USER)
    p.add(Code("abi"), 0, 8)
    p.add(Code("callee"), 0, 7)
    p.far_call(R(8), 7, "handler")
    p.ret(isa.RET_PANIC, R(0))                          # a revert never resumes here
    p.label("handler")
    p.ld_ptr(R(1), 4)
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, contracts={USER: callee}, storage=[(0, USER, 7, 5)], ergs=1 << 24, ergs_per_pubdata=0)
    r = H.rows(b)
    assert [H.family_of(x) for x in r] == ["add", "add", "far_call", "add", "add", "log", "uma", "add", "ret", "uma", "ret"]
    rev = r[8]
    assert int(rev["pc_after"]) == p.labels["handler"] and int(rev["flags_after"]) == 0 and int(rev["callstack_depth"]) == 1
    assert int(rev["bits"]) & records_bit("DST0_PTR") and (H.val(rev["dst0"]) >> 96) & 0xFFFFFFFF == 32   # returndata kept
    assert H.val(r[9]["dst0"]) == 99                                                                    # and readable
    fr = b.read_stream(0, records.STREAM_FRAME)
    assert [(int(f["kind"]), int(f["panicked"])) for f in fr] == [(1, 0), (1, 0), (2, 1), (2, 0)]
    assert b.read_storage(0, 0, USER, 7) == 5                                                            # rolled back
    lg = [x for x in b.read_stream(0, records.STREAM_LOG) if int(x["rw_flag"]) == 1]
    assert len(lg) == 1 and H.val(lg[0]["read_value"]) == 5 and H.val(lg[0]["written_value"]) == 99       # the witness keeps the write
    b.close()


def case_stack_and_code_operand_addressing(B):
    """mem_ops.rs:14-125: every stack address is (reg.low_u64() as u16).wrapping_add(imm) (:34-35); absolute = that
    (:111-121), relative = sp - that (:88-98), push writes at the OLD sp and adds (:55-70), pop subtracts and reads at
    the NEW sp (:71-86), code-page operands read constant words of the current code page (:100-110).  Reads are
    witnessed at t + 0 and the destination write at t + 3 (mod.rs:220-231); stack page = base page + 1."""
    p = Program()
    p.const("k", 0x1234567890ABCDEF << 64)
    p.add(Imm(3), 0, 1)                                  # r1 = 3 (used as the register part of addresses)
    p.add(Code("k"), 0, DStackAbs(2, reg=1))             # stack[3 + 2] = k                    (absolute, reg + imm)
    p.nop(R(0), DStackPush(8))                           # sp = 8
    p.add(StackRel(0, reg=1), 0, DStackRel(1))           # reads stack[8 - 3] = k, writes stack[8 - 1]
    p.add(StackPop(1), 0, 2)                             # sp = 7, reads stack[7] = k -> r2
    p.add(Imm(1), 2, DStackPush(0, reg=1))               # writes stack[7] = k + 1, sp = 7 + 3 = 10
    p.add(StackAbs(7), 0, 4)                             # r4 = k + 1
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, ergs=1 << 20)
    r = H.rows(b)
    k = 0x1234567890ABCDEF << 64
    assert [int(x["sp_after"]) for x in r[:7]] == [0, 0, 8, 8, 7, 10, 10]
    assert H.val(r[3]["src0"]) == k and H.val(r[4]["dst0"]) == k and H.val(r[6]["dst0"]) == k + 1
    mem = [m for m in b.read_stream(0, records.STREAM_MEM) if int(m["memory_type"]) == C.MEM_STACK]
    got = [(int(m["index"]), int(m["rw_flag"]), H.val(m["value"]), int(m["timestamp"]) - C.STARTING_TIMESTAMP) for m in mem]
    dt = C.TIME_DELTA_PER_CYCLE
    assert got == [(5, 1, k, 1 * dt + 3), (5, 0, k, 3 * dt), (7, 1, k, 3 * dt + 3), (7, 0, k, 4 * dt), (7, 1, k + 1, 5 * dt + 3),
                   (7, 0, k + 1, 6 * dt)]
    assert all(int(m["page"]) == H.BOOT_PAGE + 1 for m in mem)
    code_reads = [m for m in b.read_stream(0, records.STREAM_MEM) if int(m["memory_type"]) == C.MEM_CODE and H.val(m["value"]) == k]
    assert len(code_reads) == 1 and int(code_reads[0]["page"]) == H.BOOT_PAGE
    b.close()


def case_context_and_cycle_bookkeeping(B):
    """mod.rs:232-234 timestamp += TIME_DELTA_PER_CYCLE per cycle from STARTING_TIMESTAMP; cycle.rs:59-100 one code
    fetch per code word (4 instructions); context.rs:53-64,87-88 getters; jump.rs:24-25 pc = low 16 bits of src0."""
    p = Program()
    p.context(isa.CTX_THIS, 1)
    p.context(isa.CTX_CALLER, 2)
    p.context(isa.CTX_ERGS_LEFT, 3)
    p.context(isa.CTX_SP, 4)
    p.jump("x")
    p.add(Imm(1), 0, 9)                                     # skipped
    p.label("x")
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, ergs=100000)
    r = H.rows(b)
    assert [int(x["timestamp"]) for x in r] == [C.STARTING_TIMESTAMP + C.TIME_DELTA_PER_CYCLE * i for i in range(len(r))]
    assert [int(x["cycle"]) for x in r] == list(range(len(r)))
    assert H.val(r[0]["dst0"]) == H.BOOT_ADDRESS and H.val(r[1]["dst0"]) == 0
    ctx_price = isa.OPCODE_PRICES[int(r[0]["raw_opcode"]) & 0x7FF]
    assert H.val(r[2]["dst0"]) == 100000 - 3 * ctx_price == int(r[2]["ergs_after"])       # ergs left AFTER paying for itself
    assert H.val(r[3]["dst0"]) == 0
    assert int(r[4]["pc_after"]) == p.labels["x"] == 6 and H.family_of(r[5]) == "ret" and len(r) == 6
    assert [int(x["n_mem"]) for x in r] == [1, 0, 0, 0, 1, 0]                          # fetches at pc 0 and pc 4; pc 6 shares word 1
    b.close()


ALL = [v for k, v in sorted(globals().items()) if k.startswith("case_")]


def case_invalid_opcode_burns_the_frame(B):
    """cycle.rs:142-144 an opcode whose variant is the explicit panic (index 0: what the zero padding behind the last
    instruction of a code word decodes to) sets INVALID_OPCODE; cycle.rs:146-161 its table price is charged like any other
    -- the invalid entry is priced u32::MAX (ISA datum, unpinned like the rest of the table), so the subtraction underflows:
    ergs := 0 and NOT_ENOUGH_ERGS is set as well; cycle.rs:187-190 masks to ret.panic (condition Always, :212-217);
    ret.rs:243-249,262-264 the frame returns its zero ergs, pc goes to the exception handler, LT is set."""
    p = Program()
    p.add(R(1), 2, 3)                    # pc 0; pc 1..3 of the word are zero = the invalid opcode
    b = H.launch(B, p, 1, regs={1: 5, 2: 6}, ergs=1000)
    r = H.rows(b)
    assert len(r) == 2 and H.val(r[0]["dst0"]) == 11
    assert int(r[1]["raw_opcode"]) == 0 and int(r[1]["pc_before"]) == 1
    assert int(r[1]["error_flags"]) == 1 | 2 and int(r[1]["masked_variant"]) == isa.PANIC_VARIANT_IDX and int(r[1]["cond_resolved"]) == 1
    assert int(r[1]["callstack_depth"]) == 0 and int(r[1]["ergs_after"]) == ROOT_ERGS - 1000      # nothing comes back (ret.rs:243)
    assert int(r[1]["flags_after"]) == 1
    assert b.vm_status()[0, 0] == 1                                                                # mod.rs:96-98
    b.close()


def case_jump_takes_the_low_16_bits_of_a_register(B):
    """jump.rs:23-25 "we use lowest 16 bits of src0 as a jump destination": the destination comes from a full 256-bit
    operand; everything above bit 15 is dropped.  The instruction fetch then follows the new pc (cycle.rs:59-100)."""
    p = Program()
    p.jump(R(1))                         # r1 = 2^200 + 2^16 + 5  -> pc 5
    p.add(Imm(1), 0, 2)                  # skipped
    p.add(Imm(2), 0, 2)
    p.add(Imm(3), 0, 2)
    p.add(Imm(4), 0, 2)                  # pc 4 (second code word), skipped
    p.add(Imm(77), 0, 2)                 # pc 5
    p.ret(isa.RET_OK, R(0))
    b = H.launch(B, p, 1, regs={1: (1 << 200) | (1 << 16) | 5})
    r = H.rows(b)
    assert len(r) == 3
    assert int(r[0]["pc_before"]) == 0 and int(r[0]["pc_after"]) == 5 and int(r[0]["bits"]) & 0xC0 == 0     # no dst written
    assert int(r[1]["pc_before"]) == 5 and H.val(r[1]["dst0"]) == 77
    mem = b.read_stream(0, records.STREAM_MEM)
    assert [int(m["index"]) for m in mem if int(m["memory_type"]) == C.MEM_CODE] == [0, 1]                 # word 0, then word 1 (pc 5)
    b.close()


def case_far_call_without_code_runs_the_default_aa(B):
    """far_call.rs:131-158: the callee's code hash is READ from the deployer's storage (a LOG query, read_value as stored); a
    ZERO hash at a non-kernel address is masked into block_properties.default_aa_code_hash -- the query still reports the
    zero -- and it is the default AA's hash that is decommitted (helpers.rs:164-194) and its code that runs in the new
    frame, whose this / code address are the CALLED address (far_call.rs:520-560)."""
    from era_zk_evm_b200.asm import bytecode_hash
    aa = Program()
    aa.add(Imm(42), 0, 2)
    aa.ret(isa.RET_OK, R(0))
    aa_home, target = 0xAA00AA00, 0xDEAD5555          # the AA's bytecode is loaded by deploying it somewhere else; target has no code
    p = Program()
    p.const("abi", far_call_abi(1 << 16))
    p.const("target", target)
    p.add(Code("abi"), 0, 1)
    p.add(Code("target"), 0, 2)
    p.far_call(R(1), 2, "handler")
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.ret(isa.RET_PANIC, R(0))
    aa_hash = bytecode_hash(aa.bytecode())
    b = H.launch(B, p, 1, ergs=1 << 20, contracts={aa_home: aa}, default_aa=aa_hash)
    r = H.rows(b)
    assert [H.family_of(x) for x in r] == ["add", "add", "far_call", "add", "ret", "ret"]
    lg = b.read_stream(0, records.STREAM_LOG)
    assert len(lg) == 1 and H.val(lg[0]["key"]) == target and H.val(lg[0]["read_value"]) == 0 and int(lg[0]["rw_flag"]) == 0
    assert bytes(lg[0]["address"]) == C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS.to_bytes(20, "big")
    dec = b.read_stream(0, records.STREAM_DECOMMIT)
    assert len(dec) == 1 and H.val(dec[0]["hash"]) == aa_hash and int(dec[0]["is_fresh"]) == 1
    fr = b.read_stream(0, records.STREAM_FRAME)
    call = [f for f in fr if int(f["kind"]) == 1][-1]
    assert bytes(call["this_address"]) == target.to_bytes(20, "big") == bytes(call["code_address"])
    assert H.val(r[3]["dst0"]) == 42 and int(r[3]["callstack_depth"]) == 2
    assert int(r[4]["callstack_depth"]) == 1 and int(r[4]["pc_after"]) == 3            # normal return: back behind the call
    assert int(r[5]["callstack_depth"]) == 0 and b.vm_status()[0, 0] == 1
    b.close()


def case_flattened_histories_after_a_panicking_near_call(B):
    """The backends' own post-processing.  storage.rs:98-120 every write pushes a forward and a rollback
    query; storage.rs:156-180 / event_sink.rs:166-170 finish_frame(panicked) appends the frame's forward log and then its
    rollbacks IN REVERSE to the parent's forward log; event_sink.rs:82-131 net events = forward queries whose timestamp was
    not cancelled by a rollback, in timestamp order."""
    p = Program()
    p.add(Imm(7), 0, 1)
    p.add(Imm(99), 0, 2)
    p.event(1, 2, first=True)                     # E0, in the bootloader frame: survives
    p.near_call(0, "body", "handler")
    p.label("after")
    p.sload(1, 5)
    p.ret(isa.RET_OK, R(0))
    p.label("handler")
    p.jump("after")
    p.label("body")
    p.sstore(1, 2)                                # W1: 5 -> 99
    p.event(1, 2)                                 # E1
    p.add(Imm(100), 0, 2)
    p.sstore(1, 2)                                # W2: 99 -> 100
    p.to_l1(1, 2)                                 # M1
    p.ret(isa.RET_PANIC, R(0))
    b = H.launch(B, p, 2, storage=[(0, H.BOOT_ADDRESS, 7, 5)], ergs=1 << 24)
    b.flatten_logs()
    for vm in range(2):
        sh, eh = b.read_flat(vm, 0), b.read_flat(vm, 1)
        got = [(int(q["rw_flag"]), int(q["rollback"]), H.val(q["read_value"]), H.val(q["written_value"])) for q in sh]
        assert got == [(1, 0, 5, 99), (1, 0, 99, 100), (1, 1, 99, 100), (1, 1, 5, 99), (0, 0, 5, 0)]      # W1 W2 ~W2 ~W1 R
        assert [(int(q["aux_byte"]), int(q["rollback"])) for q in eh] == [(1, 0), (1, 0), (2, 0), (2, 1), (1, 1)]   # E0 E1 M1 ~M1 ~E1
        net_e, net_l1 = b.read_flat(vm, 2), b.read_flat(vm, 3)
        assert len(net_e) == 1 and int(net_e[0]["is_service"]) == 1 and H.val(net_e[0]["written_value"]) == 99 and len(net_l1) == 0
        assert int(sh[2]["timestamp"]) == int(sh[1]["timestamp"]) and int(eh[4]["timestamp"]) == int(eh[1]["timestamp"])
    assert (b.flat_counts(0)[1] == 0).all()
    b.close()


ALL = [v for k, v in sorted(globals().items()) if k.startswith("case_")]
