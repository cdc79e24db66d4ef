#!/usr/bin/env python
"""AlphaZero-style self-play throughput on one B200: device-resident PUCT forest + batched GPU leaf expansion feeding
a torch policy/value net of the reference's ResNet shape (BASELINE.json configs[3]; config/alphazero_blokus_20x20.yml:
25 simulations per move).  The net is random-init (no checkpoints offline); weights do not change the cost.

  python tools/selfplay_bench.py [games] [sims] [plies]
Reference context (derived from docs/README.md:159): ~5.2 simulations/s, ~0.026 plies/s (RTX 3070 Ti, 200 sims/move).
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch import nn

from blokus_rl_b200 import BlokusEngine
from blokus_rl_b200.gpu_puct import GpuPuct
from blokus_rl_b200.mcts import TorchNetEvaluator


class PolicyValueNet(nn.Module):
    """Same tensor shapes as the reference's ResNet (blokus_rl/models/blokus_nnet.py:88-151): 8x20x20 in, 64-channel
    trunk, 2-filter policy head into a 30,433-way linear layer (24.3 M of the 24.3-24.6 M parameters), 4-way value."""

    def __init__(self, planes=8, size=20, actions=30433, players=4, channels=64, blocks=4):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(planes, channels, 3, padding=1), nn.BatchNorm2d(channels), nn.ReLU())
        self.trunk = nn.ModuleList(nn.Sequential(
            nn.Conv2d(channels, channels, 3, padding=1), nn.BatchNorm2d(channels), nn.ReLU(),
            nn.Conv2d(channels, channels, 3, padding=1), nn.BatchNorm2d(channels)) for _ in range(blocks))
        self.pi = nn.Sequential(nn.Conv2d(channels, 2, 1), nn.BatchNorm2d(2), nn.ReLU(), nn.Flatten(), nn.Linear(2 * size * size, actions))
        self.v = nn.Sequential(nn.Conv2d(channels, 1, 1), nn.BatchNorm2d(1), nn.ReLU(), nn.Flatten(), nn.Linear(size * size, 64), nn.ReLU(),
                               nn.Linear(64, players), nn.Tanh())

    def forward(self, x):
        h = self.stem(x)
        for block in self.trunk:
            h = torch.relu(h + block(h))
        return self.pi(h), self.v(h)


def main():
    games = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    sims = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    plies = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    eng = BlokusEngine(20, 4)
    net = PolicyValueNet().cuda().eval()
    print(f"net parameters: {sum(p.numel() for p in net.parameters()) / 1e6:.2f} M")
    torch.backends.cudnn.benchmark = True
    search = GpuPuct(eng, TorchNetEvaluator(net), num_trees=games, max_simulations=(sims + 1) * (plies + 2) + 2, mean_edges_per_node=400)
    search.set_roots(eng.new_states(games))

    def one_move():
        for _ in range(sims):
            search.simulate(1.0)
        search.advance(search.best_actions_device())

    one_move()                                       # warm-up (cuDNN autotune, first-move expansion)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(plies):
        one_move()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    search.check()
    print(f"{games} concurrent games x {plies} plies x {sims} sims: {games * plies * sims / dt:.3e} simulations/s, "
          f"{games * plies / dt:.3e} self-play plies/s  ({dt / (plies * sims) * 1e3:.2f} ms per lockstep simulation incl. the net)")
    # the net alone, same batch
    obs = eng.observe(search.root_states())
    with torch.inference_mode():
        for _ in range(3):
            net(obs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            net(obs)
        torch.cuda.synchronize()
    print(f"net forward alone at batch {games}: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
