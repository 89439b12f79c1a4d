"""Decode state of the causal SEA attention: mirror of the reference's PerlinAttentionState
(src/models/perlin_attention/attention_state.py:238-360), which carries three stateful ops -- the incremental causal Performer
(:43-98), the sliding-window causal CNN (:142-187) and the running mean of v (:205-224).  Here:
  * `performer`  fp32 running sums S | z | vsum per (n, h), advanced by sea_performer_causal_state_fwd;
  * `cnn_in_win`, `conv1_win`  the last 4 rows of the CNN input and of the first dilated conv's output: the two 3x3 / dilation-2
    causal convolutions of a new row t read rows t-4, t-2, t of their input, so a 5-row window reproduces the prefill result exactly;
  * `t`  number of tokens consumed.
The reference clones the state at every step (functional style); `clone()` does the same when a caller wants to branch."""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class PerlinAttentionState:
    t: int = 0
    performer: Optional[torch.Tensor] = None
    cnn_in_win: Optional[torch.Tensor] = None
    conv1_win: Optional[torch.Tensor] = None

    def clone(self) -> 'PerlinAttentionState':
        c = lambda x: None if x is None else x.clone()
        return PerlinAttentionState(self.t, c(self.performer), c(self.cnn_in_win), c(self.conv1_win))
