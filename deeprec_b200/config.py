"""Typed option objects (one config layer instead of the reference's protobuf + python options +
~40 env vars).  Names and defaults follow ``python/ops/variables.py:179-302``.

Environment overrides kept for parity/debugging (read once, at object construction):
``INFERENCE_MODE``, ``TF_EV_SAVE_FILTERED_FEATURES``, ``TF_EV_RESET_VERSION``,
``COLLECTIVE_STRATEGY``, ``DEEPREC_HOST_THREADS``.
"""
from __future__ import annotations

import enum
import os
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import torch


class StorageType(enum.IntEnum):
    """storage_config.h:22-47 (tiers that make sense on a B200 box)."""
    DRAM = 0
    HBM = 1
    HBM_DRAM = 2
    DRAM_SSDHASH = 3
    HBM_DRAM_SSDHASH = 4


class CacheStrategy(enum.IntEnum):
    LFU = 0
    LRU = 1


class FilterType(enum.IntEnum):
    NONE = 0
    COUNTER = 1
    BLOOM = 2


@dataclass
class InitializerOption:
    """variables.py:179 -- new key ``k`` starts from row ``k % default_value_dim`` of a
    pre-generated matrix drawn from ``initializer``."""
    initializer: Optional[Callable[[torch.Tensor], None]] = None
    default_value_dim: int = 4096
    default_value_no_permission: float = 0.0


@dataclass
class GlobalStepEvict:
    steps_to_live: int = 0


@dataclass
class L2WeightEvict:
    l2_weight_threshold: float = -1.0


@dataclass
class CounterFilter:
    filter_freq: int = 0

    def __post_init__(self):
        if self.filter_freq >= 4096:
            # counter_filter_descriptor_impl.h:43 -- un-admitted keys keep freq in 12 bits
            raise ValueError("CounterFilter.filter_freq must be < 4096")


@dataclass
class CBFFilter:
    filter_freq: int = 0
    max_element_size: int = 0
    false_positive_probability: float = -1.0
    counter_type: torch.dtype = torch.int64

    def counter_bits(self) -> int:
        return {torch.uint8: 8, torch.int8: 8, torch.int16: 16, torch.int32: 32, torch.int64: 64}[self.counter_type]


@dataclass
class StorageOption:
    storage_type: StorageType = StorageType.DRAM
    storage_path: Optional[str] = None
    storage_size: Sequence[int] = (1 << 30,)       # bytes per tier (tier 0 first)
    cache_strategy: CacheStrategy = CacheStrategy.LFU
    layout: Optional[str] = None


@dataclass
class CheckpointOption:
    ckpt_to_load_from: Optional[str] = None
    tensor_name_in_ckpt: Optional[str] = None
    always_load_from_specific_ckpt: bool = False
    init_data_source: Optional[str] = None


@dataclass
class EmbeddingVariableOption:
    ht_type: str = ""
    ht_partition_num: int = 16
    evict_option: Optional[object] = None            # GlobalStepEvict | L2WeightEvict
    ckpt: Optional[CheckpointOption] = None
    filter_option: Optional[object] = None           # CounterFilter | CBFFilter
    storage_option: StorageOption = field(default_factory=StorageOption)
    init_option: InitializerOption = field(default_factory=InitializerOption)
    init_capacity: int = 1 << 16                     # initial key capacity (grows)
    # accepted for API parity (TF_RECORD_FREQ / TF_RECORD_VERSION): both engines keep freq / version in the key's metadata slot
    # (same cache line / DRAM sector as the key), so they are always maintained and these switches save nothing
    record_freq: bool = True
    record_version: bool = True


# The one int64 value that is never a key (the hash tables' empty-slot marker).  Host tables treat it as PADDING: a lookup returns a zero
# row, nothing is created / counted / updated for it -- so a dense ``[B, L]`` history tensor can mark its unused positions with it instead
# of looking up (and polluting the statistics of) a real id.
PAD_KEY = -(1 << 63)


def env_flag(name: str, default: bool = False) -> bool:
    v = os.environ.get(name)
    if v is None:
        return default
    return v.strip().lower() in ("1", "true", "yes", "on")


def inference_mode() -> bool:
    """``INFERENCE_MODE`` env (kv_variable_ops.cc:200-204): EV lookups never create keys."""
    return env_flag("INFERENCE_MODE")
