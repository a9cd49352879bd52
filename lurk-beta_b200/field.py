"""Field / curve identifiers and byte helpers (reference: src/field.rs:40-50 LanguageField, src/field.rs:72-91)."""
import numpy as np

from . import _capi

MODULUS = {
    _capi.FIELD_BN254_FR: 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
    _capi.FIELD_BN254_FQ: 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
    _capi.FIELD_PALLAS_FQ: 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001,
    _capi.FIELD_PALLAS_FP: 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001,
}
# LanguageField names (src/field.rs:53-62)
LANGUAGE_FIELD = {"bn256": _capi.FIELD_BN254_FR, "grumpkin": _capi.FIELD_BN254_FQ, "pallas": _capi.FIELD_PALLAS_FQ,
                  "vesta": _capi.FIELD_PALLAS_FP}
CURVE_BASE_FIELD = {_capi.CURVE_BN254_G1: _capi.FIELD_BN254_FQ, _capi.CURVE_GRUMPKIN: _capi.FIELD_BN254_FR,
                    _capi.CURVE_PALLAS: _capi.FIELD_PALLAS_FP, _capi.CURVE_VESTA: _capi.FIELD_PALLAS_FQ}
CURVE_SCALAR_FIELD = {_capi.CURVE_BN254_G1: _capi.FIELD_BN254_FR, _capi.CURVE_GRUMPKIN: _capi.FIELD_BN254_FQ,
                      _capi.CURVE_PALLAS: _capi.FIELD_PALLAS_FQ, _capi.CURVE_VESTA: _capi.FIELD_PALLAS_FP}
R = 1 << 256


def to_bytes(x):
    """canonical 32-byte little-endian repr (to_repr, src/field.rs:72-75)"""
    return int(x).to_bytes(32, "little")


def from_bytes(b):
    return int.from_bytes(bytes(b), "little")


def pack(values):
    """list of ints -> uint8 array of 32-byte LE elements"""
    return np.frombuffer(b"".join(to_bytes(v) for v in values), dtype=np.uint8).copy()


def unpack(buf):
    b = np.ascontiguousarray(buf, dtype=np.uint8).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def hex_digits(x):
    """big-endian hex as LurkField::hex_digits prints it (src/field.rs:84-91)"""
    return "%064x" % int(x)


def to_montgomery(field_id, x):
    return int(x) * R % MODULUS[field_id]


def from_montgomery(field_id, x):
    p = MODULUS[field_id]
    return int(x) * pow(R, -1, p) % p
