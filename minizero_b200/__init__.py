"""minizero_b200 — B200-native self-play engine for MiniZero's batched-MCTS hot path.

Python mirror of the C ABI in include/mz_b200.h (ctypes). The compute path is libmzb200.so (hand-written
sm_100a CUDA); there is no CPU path — importing works anywhere, but creating an Engine without the built
library or without a B200 raises.
"""
from .engine import Engine, EngineError, GAME_ATARI, GAME_GO, GAME_GOMOKU, GAME_HEX, GAME_KILLALLGO, GAME_NOGO, GAME_OTHELLO, GAME_TICTACTOE, build_library, library_path  # noqa: F401
