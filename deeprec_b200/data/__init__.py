from .kafka_dataset import KafkaDataset, KafkaGroupIODataset, merge_group_states  # noqa: F401
from .parquet_dataset import DataFrameField, DataFrameValue, ParquetDataset, parquet_fields, read_parquet  # noqa: F401
from .staged import (AsyncEmbeddingStage, PackedHostBatch, PrefetchRunner, SmartStageOptions, Staged, StagingBuffer,  # noqa: F401
                     make_prefetch_hook, smart_stage, staged)
from .synthetic import criteo_batch, taobao_batch  # noqa: F401
from .work_queue import WorkQueue  # noqa: F401
