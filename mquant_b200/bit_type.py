"""Integer bit types (mirror of fake_quant/bit_type.py:7-50: same names, same bounds)."""


class BitType:
    def __init__(self, bits, signed, name=None):
        self.bits = bits
        self.signed = signed
        self.name = name if name is not None else ("int" if signed else "uint") + str(bits)

    @property
    def upper_bound(self):
        return (1 << (self.bits - 1)) - 1 if self.signed else (1 << self.bits) - 1

    @property
    def lower_bound(self):
        return -(1 << (self.bits - 1)) if self.signed else 0

    @property
    def range(self):
        return 1 << self.bits


BIT_TYPE_LIST = [
    BitType(4, False, "uint4"),
    BitType(8, True, "int8"),
    BitType(8, False, "uint8"),
    BitType(16, True, "int16"),
    BitType(20, True, "int20"),
    BitType(18, True, "int18"),
]
BIT_TYPE_DICT = {b.name: b for b in BIT_TYPE_LIST}
