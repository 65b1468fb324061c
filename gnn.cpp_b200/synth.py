"""Deterministic synthetic GCN problems, numpy side.

Counter-based splitmix64: value i of a stream depends only on (seed, stream, i), so numpy here, the C
oracle (oracle/gcn_oracle.c: orc_synth_*) and the C++ driver (host/main.cpp) generate bit-identical
graphs, features, labels and weights.  Weight init bound follows nn::Linear::reset_parameters
(reference src/nn.cpp:198-204): U(-1/sqrt(in), 1/sqrt(in)).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
STREAM_EDGE_U, STREAM_EDGE_V, STREAM_X, STREAM_Y, STREAM_W0 = 1, 2, 3, 4, 16


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def hash3(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)
        return _mix64(_mix64(np.asarray(base, dtype=np.uint64)) + idx.astype(np.uint64))


def uniform(seed: int, stream: int, n: int, lo: float, hi: float, chunk: int = 1 << 24) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    lo32, span = np.float32(lo), np.float32(hi) - np.float32(lo)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        h = hash3(seed, stream, np.arange(s, e, dtype=np.uint64))
        u = (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
        out[s:e] = lo32 + span * u
    return out


def labels(seed: int, stream: int, n: int, C: int) -> np.ndarray:
    h = hash3(seed, stream, np.arange(n, dtype=np.uint64))
    return (h % np.uint64(C)).astype(np.int32)


def _gcd(a, b):
    while b:
        a, b = b, a % b
    return a


def _endpoint(h: np.ndarray, N: int, powerlaw: bool) -> np.ndarray:
    if not powerlaw:
        return (h % np.uint64(N)).astype(np.int32)
    x = (h >> np.uint64(32)).astype(np.float64) * (2.0 ** -32)
    t = x * np.sqrt(x)
    ids = np.minimum((t * float(N)).astype(np.int64), N - 1).astype(np.uint64)
    pa = 0x9E3779B1 % N or 1
    while _gcd(pa, N) != 1:
        pa += 1
    pb = 0x7F4A7C15 % N
    return ((np.uint64(pa) * ids + np.uint64(pb)) % np.uint64(N)).astype(np.int32)


def edges(seed: int, E: int, N: int, powerlaw: bool = False, chunk: int = 1 << 24, band: int = 0):
    """E directed entries from E//2 undirected pairs (symmetrised); duplicates and self loops kept.
    band > 0: a graph WITH locality (numpy generator only; not one of BASELINE.json's configs): pair k links a uniform
    node u to u + 1 + (hash mod band), so a contiguous row block only needs feature rows within `band` of its ends —
    the case the halo-only exchange and the interior/boundary overlap of the partitioned trainer are built for."""
    npairs = E // 2
    src = np.empty(E, dtype=np.int32)
    dst = np.empty(E, dtype=np.int32)
    for s in range(0, npairs + (E & 1), chunk):
        e = min(npairs + (E & 1), s + chunk)
        k = np.arange(s, e, dtype=np.uint64)
        u = _endpoint(hash3(seed, STREAM_EDGE_U, k), N, powerlaw)
        if band > 0:
            off = (hash3(seed, STREAM_EDGE_V, k) % np.uint64(band)).astype(np.int64) + 1
            v = ((u.astype(np.int64) + off) % N).astype(np.int32)
        else:
            v = _endpoint(hash3(seed, STREAM_EDGE_V, k), N, powerlaw)
        m = min(e, npairs) - s
        if m > 0:
            src[s:s + m] = u[:m]; dst[s:s + m] = v[:m]
            src[npairs + s:npairs + s + m] = v[:m]; dst[npairs + s:npairs + s + m] = u[:m]
        if e > npairs:  # odd E: last entry is a single directed edge
            src[E - 1] = u[-1]; dst[E - 1] = v[-1]
    return src, dst


@dataclass
class Config:
    name: str
    N: int
    E: int
    dims: List[int]
    powerlaw: bool = False
    config_id: int = 0
    band: int = 0            # > 0: locality graph (see edges)

    @property
    def seed(self) -> int:
        return 1234 + self.config_id


# BASELINE.json configs[0..4]; class counts where BASELINE omits them follow SURVEY.md §8.
CONFIGS = {
    "cora":     Config("cora", 2708, 10556, [1433, 16, 7], False, 1),
    "pubmed":   Config("pubmed", 19717, 88648, [500, 64, 3], False, 2),
    "arxiv":    Config("arxiv", 169343, 1170000, [128, 256, 256, 40], False, 3),
    "reddit":   Config("reddit", 232965, 114600000, [602, 128, 41], True, 4),
    "products": Config("products", 2450000, 61900000, [100, 256, 256, 47], True, 5),
    # NOT a BASELINE config: the products-shaped sizes on a graph with locality (neighbours within 32,768 ids), to show what
    # the halo-only exchange + interior/boundary overlap do when a row block does not need every remote row
    "products_local": Config("products_local", 2450000, 61900000, [100, 256, 256, 47], False, 6, 32768),
    # small shapes for tests / smoke
    "toy":      Config("toy", 5, 8, [10, 20, 4], False, 90),
    "tiny":     Config("tiny", 200, 1200, [24, 16, 5], False, 91),
    "tiny_pl":  Config("tiny_pl", 3000, 60000, [32, 48, 7], True, 92),
}


@dataclass
class Problem:
    cfg: Config
    src: np.ndarray
    dst: np.ndarray
    X: np.ndarray
    y: np.ndarray
    W: List[np.ndarray] = field(default_factory=list)
    b: List[np.ndarray] = field(default_factory=list)


def weights(cfg: Config):
    Ws, bs = [], []
    for l in range(1, len(cfg.dims)):
        fi, fo = cfg.dims[l - 1], cfg.dims[l]
        bound = np.float32(1.0) / np.sqrt(np.float32(fi))
        Ws.append(uniform(cfg.seed, STREAM_W0 + 2 * l, fo * fi, -bound, bound).reshape(fo, fi))
        bs.append(uniform(cfg.seed, STREAM_W0 + 2 * l + 1, fo, -bound, bound))
    return Ws, bs


def make_problem(cfg: Config, with_features: bool = True) -> Problem:
    src, dst = edges(cfg.seed, cfg.E, cfg.N, cfg.powerlaw, band=cfg.band)
    if cfg.name == "toy":  # the reference's own 8-edge fixture, tests/graph.test.cpp:19-20
        src = np.array([1, 2, 3, 0, 4, 1, 2, 3], dtype=np.int32)
        dst = np.array([1, 2, 0, 1, 2, 2, 1, 1], dtype=np.int32)
    X = uniform(cfg.seed, STREAM_X, cfg.N * cfg.dims[0], -1.0, 1.0).reshape(cfg.N, cfg.dims[0]) if with_features else None
    y = labels(cfg.seed, STREAM_Y, cfg.N, cfg.dims[-1])
    W, b = weights(cfg)
    return Problem(cfg, src, dst, X, y, W, b)
