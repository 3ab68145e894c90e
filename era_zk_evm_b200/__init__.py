"""era_zk_evm_b200 — B200-native batched out-of-circuit EraVM witness generator (hot path of matter-labs/era-zk_evm)."""
from . import asm, isa, records, workloads  # noqa: F401
from ._binding import ZkbConfig, ZkbError, default_config, make_frame, storage_entries  # noqa: F401
from .batch import GpuVmBatch, hash_bytecodes, load_library  # noqa: F401
