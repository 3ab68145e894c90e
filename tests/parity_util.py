"""Differential check helpers: CUDA batch vs CPU oracle on the same seeded inputs (bit-exact)."""
import numpy as np

from era_zk_evm_b200 import isa, records


def describe_row(r):
    v = isa.VARIANTS[int(r["masked_variant"])]
    return (f"cycle={r['cycle']} ts={r['timestamp']} {isa.FAMILY_NAMES[v.family]}.{v.sub} pc {r['pc_before']}->{r['pc_after']} "
            f"sp={r['sp_after']} ergs={r['ergs_after']} fl={r['flags_after']} err={r['error_flags']} bits={r['bits']:#x} "
            f"depth={r['callstack_depth']} m/l/dfr={r['n_mem']}/{r['n_log']}/{r['n_dfr']:#x} "
            f"src0={records.limbs_to_int(r['src0']):#x} src1={records.limbs_to_int(r['src1']):#x} "
            f"dst0={records.limbs_to_int(r['dst0']):#x} dst1={records.limbs_to_int(r['dst1']):#x}")


def first_mismatch(kind, a, b):
    """a = gpu records, b = oracle records (structured arrays). Returns None or a message."""
    n = min(len(a), len(b))
    if n == 0:
        return None if len(a) == len(b) else f"{records.STREAM_NAMES[kind]}: length differs gpu={len(a)} oracle={len(b)}"
    ab, bb = a[:n].view(np.uint8).reshape(n, -1), b[:n].view(np.uint8).reshape(n, -1)
    bad = np.nonzero((ab != bb).any(axis=1))[0]
    if len(bad) == 0:
        if len(a) != len(b):
            return f"{records.STREAM_NAMES[kind]}: length differs gpu={len(a)} oracle={len(b)}"
        return None
    i = int(bad[0])
    fields = [f for f in a.dtype.names if not np.array_equal(a[i][f], b[i][f])]
    msg = f"{records.STREAM_NAMES[kind]}[{i}] differs in {fields}\n  gpu:    {a[i]}\n  oracle: {b[i]}"
    if kind == records.STREAM_ROWS:
        msg += f"\n  gpu:    {describe_row(a[i])}\n  oracle: {describe_row(b[i])}"
        if i > 0:
            msg += f"\n  prev:   {describe_row(b[i - 1])}"
    return msg


def compare_batches(gpu, orc, vms=None, max_report=3, allow_capacity_stops=0):
    """Compares status, every stream and the final local state of the selected VMs. Returns a list of messages.
    allow_capacity_stops: how many VMs may have stopped on a device capacity limit (ZKB_VM_CAP_*) and are then only
    checked as a prefix of the oracle's streams.  Default 0: a parity test that is not ABOUT capacity must not pass
    because a config change made VMs stop early (the reference's pages and logs are unbounded)."""
    n = gpu.n_vms
    vms = range(n) if vms is None else vms
    problems = []
    gs, os_ = gpu.vm_status(), orc.vm_status()
    capped = [int(vm) for vm in vms if int(gs[vm][0]) >= 16]
    if len(capped) > allow_capacity_stops:
        problems.append(f"{len(capped)} VMs stopped on a device capacity limit (allowed: {allow_capacity_stops}), e.g. vm {capped[0]} "
                        f"status {int(gs[capped[0]][0])} after {int(gs[capped[0]][1])} cycles")
    for vm in vms:
        if len(problems) >= max_report:
            break
        if int(gs[vm][0]) >= 16:
            # the device build stopped this VM on one of ITS capacity limits (ZKB_VM_CAP_*; the reference's pages and
            # logs are unbounded): everything it emitted before stopping must be a prefix of the oracle's streams
            for kind in range(records.N_STREAMS):
                a, b = gpu.read_stream(vm, kind), orc.read_stream(vm, kind)
                if len(a) > len(b) or a.tobytes() != b[:len(a)].tobytes():
                    m = first_mismatch(kind, a, b[:len(a)]) or "gpu stream longer than the oracle's"
                    problems.append(f"vm {vm} (capacity status {int(gs[vm][0])}): not a prefix: {m}")
                    break
            continue
        if tuple(gs[vm]) != tuple(os_[vm]):
            problems.append(f"vm {vm}: status/cycles gpu={tuple(gs[vm])} oracle={tuple(os_[vm])}")
        for kind in range(records.N_STREAMS):
            m = first_mismatch(kind, gpu.read_stream(vm, kind), orc.read_stream(vm, kind))
            if m:
                problems.append(f"vm {vm}: {m}")
                break
        a, b = gpu.read_local_state(vm), orc.read_local_state(vm)
        if bytes(a) != bytes(b):
            fields = []
            for name, _ in a._fields_:
                va, vb = getattr(a, name), getattr(b, name)
                ba = bytes(va) if hasattr(va, "_length_") or hasattr(va, "_fields_") else va
                bb = bytes(vb) if hasattr(vb, "_length_") or hasattr(vb, "_fields_") else vb
                if ba != bb:
                    fields.append(name)
            problems.append(f"vm {vm}: final local state differs in {fields}")
    return problems
