"""Create random-init TorchScript nets with the REFERENCE's own network code.

Imports /root/reference/minizero/network/py (only possible in the build container) and writes
`.pt` files the compiled reference (oracle/_ref/ref_*) and this repo's worker both load.
Outputs go to oracle/_ref/nets/ (git-ignored, shipped to the GPU box by gpurun).
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import torch

REF = os.environ.get("MZ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "nets")

# name: (game_name, C, H, W, hidden, hH, hW, action_feat_ch, blocks, A, value_hidden, discrete, type)
NETS = {
    "ttt_az_2bx32": ("tictactoe", 4, 3, 3, 32, 3, 3, 1, 2, 9, 256, 1, "alphazero"),
    "go5_az_1bx16": ("go_5x5", 18, 5, 5, 16, 5, 5, 1, 1, 26, 64, 1, "alphazero"),
    "go9_az_1bx16": ("go_9x9", 18, 9, 9, 16, 9, 9, 1, 1, 82, 64, 1, "alphazero"),
    "go19_az_1bx16": ("go_19x19", 18, 19, 19, 16, 19, 19, 1, 1, 362, 64, 1, "alphazero"),
    "nogo9_az_1bx16": ("nogo_9x9", 18, 9, 9, 16, 9, 9, 1, 1, 82, 64, 1, "alphazero"),
    "gomoku15_az_1bx16": ("gomoku_15x15", 4, 15, 15, 16, 15, 15, 1, 1, 225, 64, 1, "alphazero"),
    "hex11_az_1bx16": ("hex_11x11", 4, 11, 11, 16, 11, 11, 1, 1, 121, 64, 1, "alphazero"),
    "killallgo7_az_1bx16": ("killallgo_7x7", 18, 7, 7, 16, 7, 7, 1, 1, 50, 64, 1, "alphazero"),
    "go9_az_2bx64": ("go_9x9", 18, 9, 9, 64, 9, 9, 1, 2, 82, 256, 1, "alphazero"),
    "go9_az_6bx256": ("go_9x9", 18, 9, 9, 256, 9, 9, 1, 6, 82, 256, 1, "alphazero"),
    "go19_az_2bx128": ("go_19x19", 18, 19, 19, 128, 19, 19, 1, 2, 362, 64, 1, "alphazero"),     # smallest 19x19 net the fused tower takes
    "go19_az_20bx256": ("go_19x19", 18, 19, 19, 256, 19, 19, 1, 20, 362, 256, 1, "alphazero"),  # BASELINE configs[3]
    "go5_mz_1bx16": ("go_5x5", 18, 5, 5, 16, 5, 5, 1, 1, 26, 64, 1, "muzero"),
    "ttt_mz_1bx16": ("tictactoe", 4, 3, 3, 16, 3, 3, 1, 1, 9, 32, 1, "muzero"),
    "othello_mz_1bx32": ("othello_8x8", 4, 8, 8, 32, 8, 8, 1, 1, 65, 64, 1, "muzero"),
    "othello_mz_3bx128": ("othello_8x8", 4, 8, 8, 128, 8, 8, 1, 3, 65, 256, 1, "muzero"),
    # Atari MuZero (muzero_atari_network.py): 32 x 96 x 96 planes, 6 x 6 hidden state, 18 action planes, 601-bin value / reward heads.
    # BASELINE configs[4] does not pin the size; the reference's defaults are 1 block x 256 channels (configuration.cpp:70-71)
    "atari_mz_1bx32": ("atari_ms_pacman", 32, 96, 96, 32, 6, 6, 18, 1, 18, 32, 601, "muzero"),
    "atari_mz_1bx256": ("atari_ms_pacman", 32, 96, 96, 256, 6, 6, 18, 1, 18, 256, 601, "muzero"),
}


def main(names):
    sys.path.insert(0, REF)
    from minizero.network.py.create_network import create_network  # noqa: E402

    os.makedirs(OUT, exist_ok=True)
    for name in names:
        path = os.path.join(OUT, name + ".pt")
        torch.manual_seed(0)
        net = create_network(*NETS[name])
        # perturb BatchNorm statistics so that BN folding is actually exercised by parity tests
        g = torch.Generator().manual_seed(1234)
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g))
                m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        net.eval()
        torch.jit.script(net).save(path)
        print(path, sum(p.numel() for p in net.parameters()))


if __name__ == "__main__":
    main(sys.argv[1:] or list(NETS))
