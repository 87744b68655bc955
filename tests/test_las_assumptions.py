"""The assumptions about the un-vendored crates (`las` 0.7.4 Header::from_raw, pasture-core AABB), written
down a THIRD time and from a different source: the field table of the ASPRS LAS 1.0-1.4 public header block, not the
text of csrc/host_logic.cpp or oracle/pcq_oracle.c (the two are "the same text twice" and could share a
misreading).  A table-driven parser built from the specification's (offset, type) pairs is run against both on random
headers; what all three must agree on is listed rule by rule.  CPU only."""
import ctypes as C
import struct

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle as orc

# ASPRS LAS specification, public header block: (offset, struct format, name); R1.3 adds one field, R1.4 five more
PUBLIC_HEADER_BLOCK = [
    (0, "4s", "file_signature"), (4, "H", "file_source_id"), (6, "H", "global_encoding"), (8, "16s", "project_id"),
    (24, "B", "version_major"), (25, "B", "version_minor"), (26, "32s", "system_identifier"), (58, "32s", "generating_software"),
    (90, "H", "file_creation_day"), (92, "H", "file_creation_year"), (94, "H", "header_size"), (96, "I", "offset_to_point_data"),
    (100, "I", "number_of_vlrs"), (104, "B", "point_data_record_format"), (105, "H", "point_data_record_length"),
    (107, "I", "legacy_number_of_point_records"), (111, "5I", "legacy_number_of_points_by_return"),
    (131, "d", "x_scale_factor"), (139, "d", "y_scale_factor"), (147, "d", "z_scale_factor"),
    (155, "d", "x_offset"), (163, "d", "y_offset"), (171, "d", "z_offset"),
    (179, "d", "max_x"), (187, "d", "min_x"), (195, "d", "max_y"), (203, "d", "min_y"), (211, "d", "max_z"), (219, "d", "min_z"),
]
R13 = [(227, "Q", "start_of_waveform_data_packet_record")]
R14 = [(235, "Q", "start_of_first_evlr"), (243, "I", "number_of_evlrs"), (247, "Q", "number_of_point_records"),
       (255, "15Q", "number_of_points_by_return")]
# point data record lengths of formats 0..10 (specification tables 7-17)
RECORD_LENGTH = [20, 28, 26, 34, 57, 63, 30, 36, 38, 59, 67]


def spec_parse(buf: bytes, mask_format: bool):
    """-> dict of the fields the scan path consumes, or the name of the rule that rejects the header"""
    if len(buf) < 227:
        return "short"
    f = {}
    for off, fmt, name in PUBLIC_HEADER_BLOCK:
        v = struct.unpack_from("<" + fmt, buf, off)
        f[name] = v[0] if len(v) == 1 else v
    if f["file_signature"] != b"LASF":
        return "signature"
    version = (f["version_major"], f["version_minor"])
    fields = list(PUBLIC_HEADER_BLOCK)
    if version >= (1, 3):
        fields += R13
    if version >= (1, 4):
        fields += R14
    need = max(off + struct.calcsize("<" + fmt) for off, fmt, _ in fields)
    if len(buf) < need:
        return "short"
    for off, fmt, name in fields[len(PUBLIC_HEADER_BLOCK):]:
        v = struct.unpack_from("<" + fmt, buf, off)
        f[name] = v[0] if len(v) == 1 else v
    fmt_byte = f["point_data_record_format"]
    if mask_format:
        fmt_byte &= 0b1111  # last.rs:222 / last_reader.rs:76-79
    if fmt_byte > 10:
        return "format"           # las::point::Format::new rejects it
    if f["point_data_record_length"] < RECORD_LENGTH[fmt_byte]:
        return "record_length"    # extra bytes are allowed, missing bytes are not
    if fmt_byte >= 6 and version < (1, 4):
        return "format_version"   # formats 6-10 exist from LAS 1.4 on
    legacy = f["legacy_number_of_point_records"]
    n = legacy if legacy > 0 else (f["number_of_point_records"] if version >= (1, 4) else 0)
    return {"format": fmt_byte, "record_len": f["point_data_record_length"], "off": f["offset_to_point_data"], "n": n,
            "scale": (f["x_scale_factor"], f["y_scale_factor"], f["z_scale_factor"]),
            "offset": (f["x_offset"], f["y_offset"], f["z_offset"]),
            "min": (f["min_x"], f["min_y"], f["min_z"]), "max": (f["max_x"], f["max_y"], f["max_z"])}


finite = st.floats(allow_nan=False, allow_infinity=False, width=64)


@st.composite
def headers(draw):
    minor = draw(st.integers(0, 4))
    size = 227 + (8 if minor >= 3 else 0) + (140 if minor >= 4 else 0)
    buf = bytearray(draw(st.binary(min_size=size, max_size=size)))
    buf[0:4] = b"LASF" if draw(st.integers(0, 15)) else b"LASX"
    buf[24], buf[25] = 1, minor
    struct.pack_into("<H", buf, 94, size)
    fmt = draw(st.sampled_from([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0x81, 0x86, 0xC3, 0x1F]))
    buf[104] = fmt
    base = RECORD_LENGTH[fmt & 0xF] if (fmt & 0xF) <= 10 else 20
    struct.pack_into("<H", buf, 105, max(0, base + draw(st.sampled_from([0, 0, 0, 5, -1, -20]))))
    struct.pack_into("<I", buf, 107, draw(st.sampled_from([0, 0, 1, 77, 2**32 - 1])))
    if minor >= 4:
        struct.pack_into("<Q", buf, 247, draw(st.sampled_from([0, 5, 2**33 + 1])))
    for off in range(131, 227, 8):
        struct.pack_into("<d", buf, off, draw(finite))
    cut = draw(st.sampled_from([0, 0, 0, 1, 9, 150]))
    return bytes(buf[: len(buf) - cut]) if cut else bytes(buf)


@settings(max_examples=400, deadline=None, derandomize=True)
@given(headers(), st.booleans())
def test_three_statements_of_the_header_rules_agree(pcq, buf, mask):
    want = spec_parse(buf, mask)
    arr = np.frombuffer(buf, dtype=np.uint8)
    d = pcq.FileDesc()
    rc = pcq.lib.pcq_parse_header(C.c_void_p(arr.ctypes.data), arr.nbytes, 0, int(mask), C.byref(d))
    try:
        oh = orc.parse_header(arr, mask)
        oerr = None
    except orc.OracleError as e:
        oh, oerr = None, e
    if isinstance(want, str):
        assert rc != 0 and oerr is not None, f"spec rejects the header ({want}); product rc {rc}, oracle {oerr}"
        return
    assert rc == 0 and oerr is None, f"spec accepts the header; product rc {rc}, oracle {oerr}"
    assert (d.format, d.record_len, d.point_data_off, d.n_points) == (want["format"], want["record_len"], want["off"], want["n"])
    assert tuple(d.scale) == want["scale"] and tuple(d.offset) == want["offset"]
    assert tuple(d.hdr_min) == want["min"] and tuple(d.hdr_max) == want["max"]
    assert (oh.format, oh.record_len, oh.offset_to_point_data, oh.n_points) == (want["format"], want["record_len"], want["off"], want["n"])
    assert tuple(oh.scale) == want["scale"] and tuple(oh.offset) == want["offset"]
    assert tuple(oh.min) == want["min"] and tuple(oh.max) == want["max"]


# pasture-core 0.1.0 AABB, as the survey assumes it (closed intervals; from_min_max panics on min > max)
def spec_intersects(amin, amax, bmin, bmax):
    return all(amin[i] <= bmax[i] and amax[i] >= bmin[i] for i in range(3))


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.lists(st.sampled_from([-2.0, -1.0, 0.0, 0.5, 1.0, 2.0, 3.0, float("inf"), float("nan")]), min_size=12, max_size=12))
def test_file_box_test_is_closed_interval_overlap(pcq, v):
    hmin, hmax, qmin, qmax = v[0:3], v[3:6], v[6:9], v[9:12]
    d = pcq.FileDesc()
    for i in range(3):
        d.hdr_min[i], d.hdr_max[i] = hmin[i], hmax[i]
    out = C.c_int(-1)
    rc = pcq.lib.pcq_file_intersects(C.byref(d), pcq.binding.d3(qmin), pcq.binding.d3(qmax), C.byref(out))
    if any(hmin[i] > hmax[i] for i in range(3)):
        assert rc == pcq.binding.PCQ_ERR_PANIC  # AABB::from_min_max(header bounds), las.rs:61
        return
    assert rc == 0 and bool(out.value) == spec_intersects(hmin, hmax, qmin, qmax)
