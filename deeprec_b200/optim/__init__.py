from .optimizers import (OPTIMIZERS, Adagrad, AdagradDecay, AdagradDecayOptimizer, AdagradOptimizer, Adam, AdamAsync,
                         AdamAsyncOptimizer, AdamOptimizer, AdamW, AdamWOptimizer, DeepRecOptimizer, Ftrl,
                         FtrlOptimizer, GlobalStep, GradientDescentOptimizer, SGD, collect_embedding_variables,
                         get_or_create_global_step, make_optimizer)
