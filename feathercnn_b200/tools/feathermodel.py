"""Single-file ``.feathermodel`` container.

The upstream flatbuffers ``.feathermodel`` schema is not part of the reference snapshot (SURVEY.md §0.1:
``grep feathermodel`` hits only README.md), so byte compatibility with upstream model files is NOT claimed.
This repo's container simply concatenates the ncnn files the snapshot does load:

    b"FTHRB200" | uint64 little-endian param_len | param text | bin bytes

``feather::Net::InitFromPath`` / ``InitFromBuffer`` (include/feather/net.h) read it.
"""
from __future__ import annotations

import struct
from pathlib import Path

MAGIC = b"FTHRB200"


def pack(param_path, bin_path, out_path) -> str:
    text = Path(param_path).read_bytes()
    blob = Path(bin_path).read_bytes()
    Path(out_path).write_bytes(MAGIC + struct.pack("<Q", len(text)) + text + blob)
    return str(out_path)


def unpack(model_path, prefix) -> tuple[str, str]:
    data = Path(model_path).read_bytes()
    assert data[:8] == MAGIC, "not a feather-b200 .feathermodel container"
    (n,) = struct.unpack_from("<Q", data, 8)
    Path(str(prefix) + ".param").write_bytes(data[16:16 + n])
    Path(str(prefix) + ".bin").write_bytes(data[16 + n:])
    return str(prefix) + ".param", str(prefix) + ".bin"


if __name__ == "__main__":
    import sys
    print(pack(*sys.argv[1:4]))
