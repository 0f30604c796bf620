"""Reader/writer of the reference's on-disk index (src/data_io.cpp:165-220 writer, :233-304 reader):

  u64 nK; u32 k[nK];
  u64 T;  T x { u64 idLen; id; u64 seqLen; seq; i32 length }
  u64 nMaps; nMaps x { u32 k; u64 nKeys; nKeys x { u32 hash; u64 deg; deg x { u64 tidLen; tid } } }

native little-endian, no magic.  Record order is unspecified upstream, so the reader is order-agnostic and
returns postings as CSR over DENSE transcript ids (position in the file's transcript section).
The C++ host (host/index_file.cpp) implements the same format for the CLI; this module serves tests/tools.
"""
import struct

import numpy as np


def write_index(path, ks, names, seqs, postings, lengths=None):
    """postings: dict k -> (keys u32[n], off u64[n+1], tids u32[...]) with dense ids into `names`.
    seqs: list of bytes (may be empty strings).  length field: 0 like the reference's libstdc++ build."""
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(ks)))
        for k in ks:
            f.write(struct.pack("<I", k))
        f.write(struct.pack("<Q", len(names)))
        for i, nm in enumerate(names):
            b = nm.encode() if isinstance(nm, str) else nm
            s = seqs[i] if seqs is not None else b""
            f.write(struct.pack("<Q", len(b)) + b + struct.pack("<Q", len(s)) + s)
            f.write(struct.pack("<i", 0 if lengths is None else int(lengths[i])))
        f.write(struct.pack("<Q", len(postings)))
        enc = [(nm.encode() if isinstance(nm, str) else nm) for nm in names]
        for k, (keys, off, tids) in postings.items():
            f.write(struct.pack("<IQ", k, len(keys)))
            for i in range(len(keys)):
                b0, b1 = int(off[i]), int(off[i + 1])
                f.write(struct.pack("<IQ", int(keys[i]), b1 - b0))
                for j in range(b0, b1):
                    nm = enc[int(tids[j])]
                    f.write(struct.pack("<Q", len(nm)) + nm)


def read_index(path):
    """-> (ks list, names list[str], seqs list[bytes], postings dict k -> (keys, off, tids))  (keys ascending)"""
    data = open(path, "rb").read()
    pos = 0

    def u64():
        nonlocal pos
        v = struct.unpack_from("<Q", data, pos)[0]
        pos += 8
        return v

    def u32():
        nonlocal pos
        v = struct.unpack_from("<I", data, pos)[0]
        pos += 4
        return v

    nk = u64()
    ks = [u32() for _ in range(nk)]
    T = u64()
    names, seqs, index = [], [], {}
    for _ in range(T):
        n = u64()
        nm = data[pos:pos + n].decode()
        pos += n
        m = u64()
        seqs.append(data[pos:pos + m])
        pos += m
        pos += 4  # length (unused by quant; 0 when written by a libstdc++ build)
        if nm not in index:
            index[nm] = len(names)
            names.append(nm)
    postings = {}
    nmaps = u64()
    for _ in range(nmaps):
        k = u32()
        nkeys = u64()
        recs = []
        for _ in range(nkeys):
            h = u32()
            deg = u64()
            ts = []
            for _ in range(deg):
                n = u64()
                ts.append(index[data[pos:pos + n].decode()])
                pos += n
            recs.append((h, sorted(set(ts))))
        recs.sort()
        keys = np.array([r[0] for r in recs], dtype=np.uint32)
        off = np.zeros(len(recs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(r[1]) for r in recs])
        tids = np.array([t for r in recs for t in r[1]], dtype=np.uint32)
        postings[k] = (keys, off, tids)
    assert pos == len(data), "trailing bytes in index file"
    return ks, names, seqs, postings
