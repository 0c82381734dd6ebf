from .feature_column import (CategoricalColumn, EmbeddingColumn, InputLayer, NumericColumn, WeightedCategoricalColumn,  # noqa: F401
                             categorical_column_with_adaptive_embedding, categorical_column_with_embedding,
                             categorical_column_with_hash_bucket, categorical_column_with_identity, categorical_column_with_multihash,
                             embedding_column, group_embedding_column_scope, input_layer, numeric_column,
                             sequence_categorical_column_with_embedding, shared_embedding_columns, sparse_column_with_embedding,
                             weighted_categorical_column)
