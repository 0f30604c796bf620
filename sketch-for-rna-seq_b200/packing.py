"""2-bit packing of admitted reads (A=0 C=1 G=2 T=3, 16 bases per little-endian uint32; include/sketchquant.h)
and the reference's admission rule."""
import numpy as np

_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def is_valid_sequence(seq: bytes) -> bool:
    """reference src/data_io.cpp:17-34: upper-case A, C, G, T only"""
    if len(seq) == 0:
        return True
    return bool((_CODE[np.frombuffer(seq, dtype=np.uint8)] != 255).all())


def admit(seq: bytes, ks) -> bool:
    """reference src/main.cpp:131-138"""
    return is_valid_sequence(seq) and len(seq) >= max(ks)


def pack_reads(seqs, align=4):
    """seqs: list of bytes (ACGT only).  Each read starts at the next multiple of `align` bases.
    Returns (words uint32[n_words], base_off uint32[n], length uint32[n])."""
    n = len(seqs)
    length = np.fromiter((len(s) for s in seqs), dtype=np.uint32, count=n)
    padded = (length.astype(np.uint64) + (align - 1)) // align * align
    base_off = np.zeros(n, dtype=np.uint64)
    if n > 1:
        base_off[1:] = np.cumsum(padded[:-1])
    total = int(base_off[-1] + length[-1]) if n else 0
    n_words = ((total + 15) // 16 + 3) // 4 * 4 + 4
    codes = np.zeros(n_words * 16, dtype=np.uint8)
    if n:
        flat = np.frombuffer(b"".join(seqs), dtype=np.uint8)
        c = _CODE[flat]
        if (c == 255).any():
            raise ValueError("pack_reads: non-ACGT character (admission must happen first)")
        starts = np.repeat(base_off, length.astype(np.int64))
        within = np.arange(flat.shape[0], dtype=np.uint64) - np.repeat(
            np.concatenate(([0], np.cumsum(length.astype(np.uint64))[:-1])), length.astype(np.int64))
        codes[(starts + within).astype(np.int64)] = c
    return pack_code_array(codes), base_off.astype(np.uint32), length


def pack_code_array(codes: np.ndarray) -> np.ndarray:
    """uint8 codes (multiple of 16 long) -> uint32 words"""
    c = codes.reshape(-1, 16).astype(np.uint32)
    shifts = (2 * np.arange(16, dtype=np.uint32))[None, :]
    return np.bitwise_or.reduce(c << shifts, axis=1).astype(np.uint32)


def unpack_read(words: np.ndarray, base_off: int, length: int) -> bytes:
    idx = base_off + np.arange(length, dtype=np.int64)
    codes = (words[idx >> 4] >> ((idx & 15) * 2).astype(np.uint32)) & 3
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes()
