"""Plain PyTorch fp32 reference of the DQN / DDQN net and its update (test infrastructure).

The architecture is restated from /root/reference/agents/ddqn_agent.py:18-52 (Conv2d(3, 32, 3, padding 1) +
LeakyReLU + [Dropout(0.2): left out, see DESIGN.md] + MaxPool2d(2, 2); cat(conv features, state) -> 1024 ->
LeakyReLU -> 512 -> ReLU -> 4) and the update from :113-152.  Parameter names equal the reference module's, so a
state_dict moves between the two unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class RefDQN(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(3, 32, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(), nn.Identity(), nn.MaxPool2d(2, 2))
        self.fc = nn.Sequential(nn.Linear(32 * 7 * 7 + 6, 1024), nn.LeakyReLU(), nn.Linear(1024, 512), nn.ReLU(), nn.Linear(512, 4))

    def forward(self, x):
        s, w = x
        fw = self.conv(w)
        fw = fw.view(fw.shape[0], -1)
        return self.fc(torch.cat((fw, s), dim=1))


def pack_windows(win: torch.Tensor) -> torch.Tensor:
    """[n, 3, 15, 15] of {0, 1} -> [n, 24] int32 in the replay ring's format (word ch * 8 + k = window rows 2 k in
    bits 0-14 and 2 k + 1 in bits 16-30)."""
    n = win.shape[0]
    w = torch.zeros((n, 3, 16, 16), dtype=torch.int64, device=win.device)
    w[:, :, :15, :15] = win.to(torch.int64)
    bits = w.view(n, 3, 8, 32)                      # rows 2k (16 cols) then 2k + 1 (16 cols)
    weights = (1 << torch.arange(32, device=win.device, dtype=torch.int64)).view(1, 1, 1, 32)
    words = (bits * weights).sum(-1)                # [n, 3, 8] in 0 .. 2^32 - 1
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.view(n, 24).to(torch.int32).contiguous()


def ddqn_loss(source: nn.Module, target: nn.Module, state, action, reward, next_state, gamma: float):
    """ddqn_agent.py:131-143."""
    qsa = source(state).gather(1, action.long().unsqueeze(1))
    with torch.no_grad():
        best = source(next_state).max(1)[1].unsqueeze(1)
        nxt = target(next_state).gather(1, best).squeeze(1)
    expected = nxt * gamma + reward
    return F.mse_loss(qsa, expected.unsqueeze(1)), qsa.squeeze(1)
