#!/usr/bin/env python
"""CPU rate of the REFERENCE's own env path: blokus_rl/colossumrl/blokus_wrapper.py, unmodified, doing what
its random player and arena do per ply -- get_sample_move -> get_next_state -> get_valid_moves -> get_game_ended -- over this
repo's colosseumrl shim on the CPU oracle.  The Python overheads the reference really pays are included (243 KB float64
mask per call, string <-> id dictionaries); the engine under the shim is the restatement, not colosseumrl.
Needs /root/reference (build container only).   python tools/ref_wrapper_cpu_rate.py [plies] [processes]
With processes > 1 the same loop runs in that many processes at once, one env each ("all host cores")."""
import os
import sys
import tempfile
import time
import types
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ref_stubs
from oracle_backend import OracleBackend
from blokus_rl_b200 import colosseum_shim

plies_wanted = int(sys.argv[1]) if len(sys.argv) > 1 else 600
procs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if procs > 1:                                   # one env per process, all at once; every child prints its own rate
    import subprocess
    t0 = time.perf_counter()
    kids = [subprocess.Popen([sys.executable, __file__, str(plies_wanted)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for _ in range(procs)]
    outs = [k.communicate()[0] for k in kids]
    for tag in ("20x20 4p", "7x7 2p"):
        rates = [float(ln.split(" = ")[1].split(" plies/s")[0]) for o in outs for ln in o.splitlines() if ln.startswith(tag)]
        print(f"{tag}: {procs} processes at once: {sum(rates):.0f} plies/s in total ({min(rates):.0f}..{max(rates):.0f} per process)")
    sys.exit(0)
for (N, P) in ((20, 4), (7, 2)):
    backend = OracleBackend(N, P)
    colosseum_shim.set_backend(backend)
    colosseum_shim.install()
    ref_stubs.install_stubs()
    work = Path(tempfile.mkdtemp())
    cwd = os.getcwd()
    os.chdir(work)                       # the reference writes debug.log / states/ into the CWD
    try:
        from blokus_rl.colossumrl.blokus_wrapper import ColosseumBlokusGameWrapper
        hp = types.SimpleNamespace(board_size=N, number_of_players=P, states_dir=work / "states")
        t0 = time.perf_counter()
        game = ColosseumBlokusGameWrapper(hp)          # builds its action table through the shim's Board API
        t_table = time.perf_counter() - t0
        plies = games = 0
        t0 = time.perf_counter()
        while plies < plies_wanted:
            s, p = game.get_init_board()
            while True:
                a = game.get_sample_move(s)                        # players/random_player.py:13
                s, p = game.get_next_state(s, p, a)                # :14
                game.get_valid_moves(s, p)                         # what every consumer asks next (mcts.py:63, trainer.py:118)
                plies += 1
                if game.get_game_ended(s) is not None:             # arena.py:85
                    games += 1
                    break
        dt = time.perf_counter() - t0
        print(f"{N}x{N} {P}p: reference wrapper path over the CPU oracle: {plies} plies / {games} games in {dt:.1f} s = "
              f"{plies / dt:.0f} plies/s on 1 core (action table built by the reference's loop in {t_table:.1f} s)")
    finally:
        os.chdir(cwd)
        colosseum_shim.set_backend(None)
    for m in [k for k in sys.modules if k.startswith("blokus_rl.") or k == "blokus_rl"]:
        del sys.modules[m]
