"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit-exact, on every workload."""
import numpy as np
import pytest

from era_zk_evm_b200 import records, workloads

from parity_util import compare_batches

pytestmark = pytest.mark.gpu


def _pair(w, vm_ids, oracle_mod):
    from era_zk_evm_b200 import GpuVmBatch
    cfg = w.config(len(vm_ids))
    gpu = GpuVmBatch(cfg)
    orc = oracle_mod.OracleBatch(cfg)
    w.setup(gpu, vm_ids)
    w.setup(orc, vm_ids)
    return gpu, orc


@pytest.mark.parametrize("name,kwargs,n", [
    ("alu_loop", dict(cycles=1000), 8),
    ("alu_loop", dict(cycles=100), 300),
    ("storage", dict(), 70),
    ("keccak", dict(n_calls=3), 40),
    ("keccak", dict(n_calls=2, preimage_bytes=200), 16),
    ("erc20", dict(n_transfers=4), 200),
])
def test_workload_parity(name, kwargs, n, oracle_mod):
    w = workloads.WORKLOADS[name](**kwargs)
    vm_ids = list(range(n))
    gpu, orc = _pair(w, vm_ids, oracle_mod)
    gpu.run()
    orc.run_threads(0, 0)
    st = gpu.vm_status()
    assert (st[:, 0] == 1).all(), f"not all VMs ended: {st[st[:, 0] != 1][:5]}"
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)
    gc, gb = gpu.totals()
    oc, ob = orc.totals()
    assert (gc, gb) == (oc, ob)


def test_resumable_run_matches_single_run(oracle_mod):
    w = workloads.Erc20(n_transfers=2)
    gpu, orc = _pair(w, list(range(33)), oracle_mod)
    for _ in range(64):
        gpu.run(max_cycles_per_vm=37)
        if gpu.execution_has_ended():
            break
    orc.run_threads(0, 0)
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)
