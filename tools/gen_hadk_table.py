"""Freeze the reference's 11 non-power-of-two Hadamard factors as a bit-packed table.

Generator-side only (runs in the build container where /root/reference is mounted).
Source of the matrices: ``get_had{12,20,28,36,40,52,60,108,140,156,172}`` in
/root/reference/fake_quant/hadamard_utils.py:201-97479 (Sloane's library, +-1 entries
written out as 97 k lines of float literals).  The online transform must use the SAME
matrices as the offline weight rotation (apply_exact_had_to_linear,
hadamard_utils.py:135-191), so they are data the new build has to carry; we carry
them as 1 bit per entry (12.4 KB total) instead of source text.

File format (little endian):
    magic  b"MQHADK1\\0"
    u32    count
    count x { u32 K ; ceil(K*K/8) bytes, row-major, bit i of byte j <-> entry 8*j+i,
              bit 1 = +1, bit 0 = -1 }
"""
import hashlib
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZES = [12, 20, 28, 36, 40, 52, 60, 108, 140, 156, 172]


def main():
    from oracle._ref_import import import_reference

    _, hu = import_reference()
    blob = bytearray(b"MQHADK1\0")
    blob += struct.pack("<I", len(SIZES))
    for k in SIZES:
        h = getattr(hu, f"get_had{k}")().numpy()
        assert h.shape == (k, k) and set(np.unique(h).tolist()) == {-1.0, 1.0}
        assert np.array_equal(h @ h.T, k * np.eye(k, dtype=h.dtype)), k
        bits = (h.reshape(-1) > 0).astype(np.uint8)
        packed = np.packbits(bits, bitorder="little")
        blob += struct.pack("<I", k) + packed.tobytes()
    for out in (os.path.join(ROOT, "mquant_b200", "data", "hadk.bin"),
                os.path.join(ROOT, "tests", "golden", "hadk.bin")):
        with open(out, "wb") as f:
            f.write(bytes(blob))
    print(len(blob), "bytes; sha256", hashlib.sha256(bytes(blob)).hexdigest())


if __name__ == "__main__":
    main()
