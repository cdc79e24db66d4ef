"""Golden vectors recovered from the reference's OWN rendered games.

The reference's env engine (colosseumrl / blokus-gym) is absent, but its docs ship renders produced by
`ColosseumBlokusGameWrapper.render` (blokus_wrapper.py:248-279: one matplotlib polygon per cell,
`board_contents[y][x]` drawn at (x, y), colours {0: lightgrey, 1: red, 2: blue, 3: yellow, 4: green})
of games played by the real engine:

  docs/images/AlphaZero/blokus_20/arena.gif            62 frames, one 20x20 4-player game, ply by ply
  docs/images/AlphaZero/blokus_7/step_1_win.gif        11 frames, 7x7 2-player, player 1 wins
  docs/images/AlphaZero/blokus_7/step_104_draw.gif     10 frames, 7x7 2-player, a draw
  docs/images/AlphaZero/blokus_20/sample_game_blokus_20.png   a 20x20 position
  docs/images/AlphaZero/blokus_7/sample_game_blokus_7.png     a 7x7 position
  docs/images/AlphaZero/blokus_20/blokus20_observation.png    the 8 planes of one `canonical_board`

This script (run in the build container, where /root/reference exists and PIL is importable) samples
the cell colours of every frame and writes the boards to tests/golden/ref_render_games.json.  It does
NOT use the oracle: the fixture is reference output only.  tests/test_ref_render_golden.py then
replays the games through the oracle (CPU) and through the CUDA engine (GPU).

    python tests/golden/make_ref_render_golden.py
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
from PIL import Image

REF = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("/root/reference")
OUT = Path(__file__).resolve().parent / "ref_render_games.json"

# matplotlib named colours used by render() (blokus_wrapper.py:259)
CELL_RGB = np.array([[211, 211, 211], [255, 0, 0], [0, 0, 255], [255, 255, 0], [0, 128, 0]], float)
VIRIDIS_RGB = np.array([[68, 1, 84], [253, 231, 37]], float)   # imshow 0 / 1


def _axes_box(img: np.ndarray):
    """Pixel box of the axes = bounding box of the pixels drawn in one of the five cell colours."""
    d = ((img[:, :, None, :] - CELL_RGB[None, None]) ** 2).sum(3).min(2)
    m = d < 300
    cols = np.where(m.sum(0) > 0.3 * img.shape[0])[0]
    rows = np.where(m.sum(1) > 0.3 * img.shape[1])[0]
    return cols.min(), cols.max() + 1, rows.min(), rows.max() + 1


def _sample_board(img: np.ndarray, n: int, box) -> tuple[np.ndarray, float]:
    """board[y][x] from the mean colour of a patch at the centre of cell (x, y); y axis points up."""
    x0, x1, y0, y1 = box
    board = np.zeros((n, n), int)
    worst = np.inf
    r = max(2, int(0.25 * (y1 - y0) / n))
    for y in range(n):
        for x in range(n):
            px = int(x0 + (x + 0.5) * (x1 - x0) / n)
            py = int(y1 - (y + 0.5) * (y1 - y0) / n)
            patch = img[py - r:py + r + 1, px - r:px + r + 1].reshape(-1, 3).mean(0)
            d = ((CELL_RGB - patch) ** 2).sum(1)
            o = np.argsort(d)
            board[y, x] = o[0]
            worst = min(worst, d[o[1]] / max(d[o[0]], 1.0))
    return board, worst


def _rows(board: np.ndarray) -> list[str]:
    """Row strings, index 0 = board_contents[0] (the bottom row of the render)."""
    return ["".join(".1234"[v] for v in row) for row in board]


def decode_gif(rel: str, n: int, p: int, result: str) -> dict:
    im = Image.open(REF / rel)
    frames, worst = [], np.inf
    # the gif frames are savefig-sized 640x480 figures with default subplot margins; the axes box is
    # found on the last (fullest) frame and reused for every frame of the file
    im.seek(im.n_frames - 1)
    box = _axes_box(np.array(im.convert("RGB")).astype(float))
    for f in range(im.n_frames):
        im.seek(f)
        b, w = _sample_board(np.array(im.convert("RGB")).astype(float), n, box)
        frames.append(_rows(b))
        worst = min(worst, w)
    return {"file": rel, "N": n, "P": p, "result": result, "frames": frames,
            "min_colour_margin": round(float(worst), 1)}


def decode_png(rel: str, n: int, p: int) -> dict:
    img = np.array(Image.open(REF / rel).convert("RGB")).astype(float)
    b, w = _sample_board(img, n, _axes_box(img))
    return {"file": rel, "N": n, "P": p, "board": _rows(b), "min_colour_margin": round(float(w), 1)}


def _ranges(v: np.ndarray):
    idx = np.where(v)[0]
    out, s, prev = [], idx[0], idx[0]
    for i in idx[1:]:
        if i != prev + 1:
            out.append((s, prev + 1))
            s = i
        prev = i
    out.append((s, prev + 1))
    return out


def decode_observation(rel: str, n: int, planes: int) -> dict:
    """8 imshow panels (row 0 at the top = array row 0), titles 'Warstwa k' = plane k, 2 rows x 4."""
    img = np.array(Image.open(REF / rel).convert("RGB")).astype(float)
    d = ((img[:, :, None, :] - VIRIDIS_RGB[None, None]) ** 2).sum(3)
    m = d.min(2) < 400
    cr, rr = _ranges(m.sum(0) > 50), _ranges(m.sum(1) > 50)
    assert len(cr) * len(rr) == planes, (cr, rr)
    out = []
    for (y0, y1) in rr:
        for (x0, x1) in cr:
            b = np.zeros((n, n), int)
            for y in range(n):
                for x in range(n):
                    b[y, x] = int(np.argmin(d[int(y0 + (y + 0.5) * (y1 - y0) / n), int(x0 + (x + 0.5) * (x1 - x0) / n)]))
            out.append(["".join(".#"[v] for v in row) for row in b])
    return {"file": rel, "N": n, "planes": out}


def main() -> None:
    doc = {
        "generated_by": "tests/golden/make_ref_render_golden.py",
        "what": "cell colours sampled from renders shipped in the reference's docs/ (real engine output); "
                "row strings are board_contents[y], '.'=0 empty, '1'..'4' = colour",
        "games": [
            decode_gif("docs/images/AlphaZero/blokus_20/arena.gif", 20, 4, "unknown"),
            decode_gif("docs/images/AlphaZero/blokus_7/step_1_win.gif", 7, 2, "player_1_wins"),
            decode_gif("docs/images/AlphaZero/blokus_7/step_104_draw.gif", 7, 2, "draw"),
        ],
        "positions": [
            decode_png("docs/images/AlphaZero/blokus_20/sample_game_blokus_20.png", 20, 4),
            decode_png("docs/images/AlphaZero/blokus_7/sample_game_blokus_7.png", 7, 2),
        ],
        "observation": decode_observation("docs/images/AlphaZero/blokus_20/blokus20_observation.png", 20, 8),
    }
    OUT.write_text(json.dumps(doc, indent=1) + "\n")
    for g in doc["games"]:
        print(g["file"], len(g["frames"]), "frames, colour margin", g["min_colour_margin"])
    print("wrote", OUT)


if __name__ == "__main__":
    main()
