"""Sample-aware graph compression (python/graph_optimizer/sample_awared_graph_compression.py in the reference): in ranking
requests the user-side features are identical for every candidate item, so they are sent ONCE per request (``[1, D_user]``),
run through the user-side sub-network once, and tiled to the item count as late as possible."""
from __future__ import annotations

from typing import Callable

import torch


class SampleAwareCompression:
    def __init__(self, user_net: Callable[[torch.Tensor], torch.Tensor], item_net: Callable[[torch.Tensor], torch.Tensor],
                 head: Callable[[torch.Tensor], torch.Tensor]):
        self.user_net, self.item_net, self.head = user_net, item_net, head

    @torch.no_grad()
    def __call__(self, user_features: torch.Tensor, item_features: torch.Tensor) -> torch.Tensor:
        """user_features [1, Du] (or [R, Du] with ``item_features`` [R, N, Di]); returns scores [N] (or [R, N])."""
        u = self.user_net(user_features)                       # computed once per request, not once per candidate
        if item_features.dim() == 2:
            it = self.item_net(item_features)
            return self.head(torch.cat([u.expand(it.shape[0], -1), it], -1)).squeeze(-1)
        R, N, _ = item_features.shape
        it = self.item_net(item_features.reshape(R * N, -1)).view(R, N, -1)
        return self.head(torch.cat([u.unsqueeze(1).expand(R, N, -1), it], -1)).squeeze(-1)


def enable_sample_awared_graph_compression(user_net, item_net, head) -> SampleAwareCompression:
    """``tf.graph_optimizer.enable_sample_awared_graph_compression`` analogue."""
    return SampleAwareCompression(user_net, item_net, head)
