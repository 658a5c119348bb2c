"""Reader for the oracle's ARBD1 stage dumps (oracle/dump_hooks.cpp) -- test infrastructure."""
import os, glob
import numpy as np

_DT = {b"u1": np.uint8, b"u2": np.uint16, b"u4": np.uint32, b"i4": np.int32, b"u8": np.uint64, b"f4": np.float32, b"f8": np.float64}

def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:6] == b"ARBD1\n", path
    p = 6
    while p < len(buf):
        nl = int(np.frombuffer(buf, np.uint32, 1, p)[0]); p += 4
        name = buf[p:p + nl].decode(); p += nl
        dt = _DT[buf[p:p + 2]]; p += 2
        cnt = int(np.frombuffer(buf, np.uint64, 1, p)[0]); p += 8
        out[name] = np.frombuffer(buf, dt, cnt, p).copy(); p += cnt * np.dtype(dt).itemsize   # a copy: views into the file sit at odd offsets, the C ABI expects naturally aligned arrays
    return out

def read_dir(d):
    """-> ordered list of (stage, arrays)"""
    res = []
    for path in sorted(glob.glob(os.path.join(d, "*.bin"))):
        stage = os.path.basename(path)[3:-4]
        res.append((stage, read_dump(path)))
    return res

def stage(dumps, name, occurrence=0):
    hits = [a for (s, a) in dumps if s == name]
    return hits[occurrence]
