"""Map loading (env.py:45-57 format) and the six stock layouts as compact row bitmasks.

The reference ships its maps as text files next to env.py (`map.txt`, `map1.txt` .. `map5.txt`).
They are input data, so the build carries them as one hex word per row (bit = obstacle, MSB = column
0) instead of as text copies; `Environment(map_file="map1.txt")` resolves a stock name when no such
file exists on disk, and `write_map_file` materialises the text form for tools that want a path.
"""
from __future__ import annotations

import os

import numpy as np

_STOCK = {
    "map.txt": (7, 7, "7f 41 41 41 41 41 7f"),
    "map1.txt": (10, 10, "3ff 201 201 201 201 201 201 201 201 3ff"),
    "map2.txt": (20, 20, "fffff 83f81 83f81 83f81 80001 83f81 83f81 83f81 83f81 80001 83f81 83f81 83f81 83f81 "
                         "80001 83f81 83f81 83f81 83f81 fffff"),
    "map3.txt": (20, 20, "fffff 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 9fff9 "
                         "9fff9 9fff9 9fff9 80001 80001 fffff"),
    "map4.txt": (20, 20, "fffff 80001 aaaa9 80001 aaaa9 80001 aaaa9 80001 aaaa9 80001 aaaa9 80001 aaaa9 80001 "
                         "aaaa9 80001 aaaa9 80001 aaaa9 fffff"),
    "map5.txt": (20, 20, "fffff 80221 b8a2d 8baa9 ba2ab 92aa9 a22ad 920a9 c47ab a80a9 abead b0089 9029f a3af5 "
                         "a3a65 97b05 ac071 a97df 82001 fffff"),
}


def stock_map(name: str) -> np.ndarray:
    rows, cols, words = _STOCK[os.path.basename(name)]
    g = np.zeros((rows, cols), np.uint8)
    for r, w in enumerate(words.split()):
        v = int(w, 16)
        for c in range(cols):
            g[r, c] = (v >> (cols - 1 - c)) & 1
    return g


def parse_map_text(text: str) -> np.ndarray:
    """Each line: cells separated by single spaces (env.py:53-56)."""
    rows = []
    for line in text.split("\n"):
        if line.strip() == "":
            continue
        rows.append([int(tok) for tok in line.strip().split(" ")])
    if not rows or any(len(r) != len(rows[0]) for r in rows):
        raise ValueError("malformed map: rows of unequal length")
    return np.array(rows, dtype=np.uint8)


def load_map(map_file) -> np.ndarray:
    """Path to a map text file, a stock name, or an already parsed 2-D array."""
    if isinstance(map_file, np.ndarray):
        return np.ascontiguousarray(map_file.astype(np.uint8))
    if isinstance(map_file, (list, tuple)) and map_file and isinstance(map_file[0], (list, tuple, np.ndarray)):
        return np.ascontiguousarray(np.asarray(map_file, dtype=np.uint8))
    path = os.fspath(map_file)
    if os.path.exists(path):
        with open(path, "r") as f:
            return parse_map_text(f.read())
    if os.path.basename(path) in _STOCK:
        return stock_map(path)
    raise FileNotFoundError(path)


def write_map_file(name: str, path: str) -> str:
    g = stock_map(name)
    with open(path, "w") as f:
        f.write("\n".join(" ".join(str(int(x)) for x in row) for row in g) + "\n")
    return path
