"""Test helpers: an independent (numpy) LAS / LAST file writer and record-set comparison."""
from __future__ import annotations

import struct

import numpy as np

FORMAT_LEN = {0: 20, 1: 28, 2: 26, 3: 34, 6: 30, 7: 36}
COLOR_OFF = {2: 20, 3: 28, 5: 28}


def las_header(n, fmt, record_len, scale, offset, mn, mx, version=(1, 2), fmt_byte=None) -> bytes:
    h = bytearray(227 if version < (1, 3) else (235 if version < (1, 4) else 375))
    h[0:4] = b"LASF"
    h[24], h[25] = version
    struct.pack_into("<H", h, 94, len(h))
    struct.pack_into("<I", h, 96, len(h))
    h[104] = fmt if fmt_byte is None else fmt_byte
    struct.pack_into("<H", h, 105, record_len)
    struct.pack_into("<I", h, 107, n if version < (1, 4) or fmt < 6 else 0)
    struct.pack_into("<3d", h, 131, *scale)
    struct.pack_into("<3d", h, 155, *offset)
    struct.pack_into("<6d", h, 179, mx[0], mn[0], mx[1], mn[1], mx[2], mn[2])
    if version >= (1, 4):
        struct.pack_into("<Q", h, 247, n)
    return bytes(h)


def make_file(xyz, cls, rgb=None, fmt=1, scale=(0.01, 0.01, 0.01), offset=(0.0, 0.0, 0.0), layout="las",
              record_len=None, hdr_min=None, hdr_max=None, version=(1, 2), fmt_byte=None, seed=1) -> np.ndarray:
    """Builds a LAS (row-major) or LAST (transposed) file image from raw integer coordinates."""
    xyz = np.asarray(xyz, dtype=np.int32).reshape(-1, 3)
    n = xyz.shape[0]
    cls = np.asarray(cls, dtype=np.uint8).reshape(n)
    flen = FORMAT_LEN[fmt]
    R = record_len or flen
    rng = np.random.default_rng(seed)
    rec = rng.integers(0, 256, size=(n, R), dtype=np.uint8)  # every other field: noise
    rec[:, 0:12] = xyz.view(np.uint8).reshape(n, 12)
    cls_k = 15 if fmt <= 5 else 16
    rec[:, cls_k] = cls
    if fmt in COLOR_OFF:
        if rgb is None:
            rgb = rng.integers(0, 65536, size=(n, 3), dtype=np.uint16)
        rec[:, COLOR_OFF[fmt]: COLOR_OFF[fmt] + 6] = np.asarray(rgb, dtype="<u2").reshape(n, 3).view(np.uint8).reshape(n, 6)
    pos = xyz.astype(np.float64) * np.array(scale) + np.array(offset)
    mn = hdr_min if hdr_min is not None else (pos.min(axis=0) if n else np.array(offset, dtype=float))
    mx = hdr_max if hdr_max is not None else (pos.max(axis=0) if n else np.array(offset, dtype=float))
    hdr = np.frombuffer(las_header(n, fmt, R, scale, offset, mn, mx, version, fmt_byte), dtype=np.uint8)
    if layout == "las":
        body = rec.reshape(-1)
    else:
        # LAST: each record field becomes one column; column of the field at record offset k starts at k*N
        fields = [(0, 12), (12, 2), (14, 1), (15, 1), (16, 1), (17, 1), (18, 2)]
        if fmt >= 6:
            fields = [(0, 12), (12, 2), (14, 2), (16, 1), (17, 1), (18, 2), (20, 2), (22, 8)]
            if fmt == 7:
                fields.append((30, 6))
        else:
            if fmt in (1, 3):
                fields.append((20, 8))
            if fmt in (2, 3):
                fields.append((COLOR_OFF[fmt], 6))
        if R > flen:
            fields.append((flen, R - flen))
        body = np.concatenate([rec[:, o: o + s].reshape(-1) for o, s in fields]) if n else np.zeros(0, np.uint8)
    return np.concatenate([hdr, body]).astype(np.uint8)


def sort_points(p: np.ndarray) -> np.ndarray:
    """canonical order for set comparison of 31-byte records"""
    raw = np.ascontiguousarray(p).view(np.uint8).reshape(-1, 31)
    order = np.lexsort(raw.T[::-1])
    return raw[order]


def same_point_set(a: np.ndarray, b: np.ndarray) -> bool:
    if len(a) != len(b):
        return False
    return bool(np.array_equal(sort_points(a), sort_points(b)))


def same_point_seq(a: np.ndarray, b: np.ndarray) -> bool:
    if len(a) != len(b):
        return False
    return bool(np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)))
