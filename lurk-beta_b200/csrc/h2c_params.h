// Host-side construction of the hash-to-curve parameters (h2c.cuh) and SHAKE256: what `DlogGroup::from_label` needs besides the
// per-point map (Arecibo provider macros, reached from public_params -- reference src/proof/nova.rs:196-216).
//   * SVDW constants are computed here from the curve equation (RFC 9380 6.6.1), Z = 1 (halo2curves bn256 / grumpkin SVDW_Z).
//   * iso-Pallas / iso-Vesta coefficients and the 13 isogeny constants are typed in as published by pasta_curves
//     (`ISOGENY_CONSTANTS`; Zcash protocol spec 5.4.9.8); oracle/h2c.py DERIVES the same numbers (Velu) and the CPU suite compares.
//   * SHAKE256 (FIPS 202) is sequential by construction: it stays on the host and is pipelined against the kernel.
// Used by h2c.cu and by the host test build (tests/csrc/h2c_host_test.cc); no CUDA in this header.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>

#include "h2c.cuh"

namespace lurk {

template <class F>
inline F fe_from_hex(const char *hex) {      // 64 hex digits, big-endian, canonical value < p
    F raw = F::zero();
    for (int i = 0; i < 64; i++) {
        char ch = hex[i];
        uint32_t d = ch <= '9' ? ch - '0' : (ch | 0x20) - 'a' + 10;
        int bit = 4 * (63 - i);
        raw.v[bit >> 5] |= d << (bit & 31);
    }
    return F::from_canonical(raw);
}

struct H2cCurveText {
    const char *curve_id, *method;
    const char *iso_a;         // nullptr for SVDW curves
    const char *iso[13];
};
inline const H2cCurveText &h2c_curve_text(int curve_id) {
    static const H2cCurveText t[4] = {
        {"bn256_g1", "SVDW", nullptr, {}},
        {"grumpkin_g1", "SVDW", nullptr, {}},
        {"pallas", "SSWU", "18354a2eb0ea8c9c49be2d7258370742b74134581a27a59f92bb4b0b657a014b",
         {"0e38e38e38e38e38e38e38e38e38e38e4081775473d8375b775f6034aaaaaaab", "3509afd51872d88e267c7ffa51cf412a0f93b82ee4b994958cf863b02814fb76",
          "17329b9ec525375398c7d7ac3d98fd13380af066cfeb6d690eb64faef37ea4f7", "1c71c71c71c71c71c71c71c71c71c71c8102eea8e7b06eb6eebec06955555580",
          "1d572e7ddc099cff5a607fcce0494a799c434ac1c96b6980c47f2ab668bcd71f", "325669becaecd5d11d13bf2a7f22b105b4abf9fb9a1fc81c2aa3af1eae5b6604",
          "1a12f684bda12f684bda12f684bda12f7642b01ad461bad25ad985b5e38e38e4", "1a84d7ea8c396c47133e3ffd28e7a09507c9dc17725cca4ac67c31d8140a7dbb",
          "3fb98ff0d2ddcadd303216cce1db9ff11765e924f745937802e2be87d225b234", "025ed097b425ed097b425ed097b425ed0ac03e8e134eb3e493e53ab371c71c4f",
          "0c02c5bcca0e6b7f0790bfb3506defb65941a3a4a97aa1b35a28279b1d1b42ae", "17033d3c60c68173573b3d7f7d681310d976bbfabbc5661d4d90ab820b12320a",
          "40000000000000000000000000000000224698fc094cf91b992d30ecfffffde5"}},
        {"vesta", "SSWU", "267f9b2ee592271a81639c4d96f787739673928c7d01b212c515ad7242eaa6b1",
         {"38e38e38e38e38e38e38e38e38e38e390205dd51cfa0961a43cd42c800000001", "1d935247b4473d17acecf10f5f7c09a2216b8861ec72bd5d8b95c6aaf703bcc5",
          "18760c7f7a9ad20ded7ee4a9cdf78f8fd59d03d23b39cb11aeac67bbeb586a3d", "31c71c71c71c71c71c71c71c71c71c71e1c521a795ac8356fb539a6f0000002b",
          "0a2de485568125d51454798a5b5c56b2a3ad678129b604d3b7284f7eaf21a2e9", "14735171ee5427780c621de8b91c242a30cd6d53df49d235f169c187d2533465",
          "12f684bda12f684bda12f684bda12f685601f4709a8adcb36bef1642aaaaaaab", "2ec9a923da239e8bd6767887afbe04d121d910aefb03b31d8bee58e5fb81de63",
          "19b0d87e16e2578866d1466e9de10e6497a3ca5c24e9ea634986913ab4443034", "1ed097b425ed097b425ed097b425ed098bc32d36fb21a6a38f64842c55555533",
          "2f44d6c801c1b8bf9e7eb64f890a820c06a767bfc35b5bac58dfecce86b2745e", "3d59f455cafc7668252659ba2b546c7e926847fb9ddd76a1d43d449776f99d2f",
          "40000000000000000000000000000000224698fc0994a8dd8c46eb20fffffde5"}},
    };
    return t[curve_id & 3];
}

// b of y^2 = x^3 + b for the four curves: 3, -17, 5, 5
template <class C> inline typename C::Base curve_b_coeff() {
    using F = typename C::Base;
    switch (C::ID) {
        case 0: return F::from_u64(3);
        case 1: return F::from_u64(17).neg();
        default: return F::from_u64(5);
    }
}

// returns false when domain_prefix / msg_len do not fit the single-block layout of hash_to_field
template <class C>
inline bool h2c_make_params(const char *domain_prefix, size_t msg_len, H2cParams<typename C::Base> &P) {
    using F = typename C::Base;
    const H2cCurveText &txt = h2c_curve_text(C::ID);
    memset(&P, 0, sizeof P);
    // DST' = prefix || "-" || curve_id || "_XMD:BLAKE2b_" || method || "_RO_" || len
    std::string dst = std::string(domain_prefix) + "-" + txt.curve_id + "_XMD:BLAKE2b_" + txt.method + "_RO_";
    if (dst.size() + 1 > (size_t)H2C_MAX_DST || dst.size() > 255) return false;
    if (msg_len > (size_t)H2C_MAX_MSG || msg_len + 3 + dst.size() + 1 > 128 || 65 + dst.size() + 1 > 128) return false;
    memcpy(P.dst, dst.data(), dst.size());
    P.dst[dst.size()] = (uint8_t)dst.size();
    P.dst_len = (uint32_t)dst.size() + 1;
    P.zero_block.init();
    uint64_t zeros[16] = {0};
    P.zero_block.compress(zeros, 128, false);
    // (t - 1) / 2 with p - 1 = 2^s t
    {
        uint32_t t[8];
        for (int i = 0; i < 8; i++) t[i] = F::Params::MOD(i);
        t[0] -= 1;                                                   // p is odd
        const int sh = F::Params::TWO_ADICITY + 1;                   // (p - 1) >> s, then (t - 1) >> 1 with t odd = one more shift
        for (int k = 0; k < sh; k++) { for (int i = 0; i < 7; i++) t[i] = (t[i] >> 1) | (t[i + 1] << 31); t[7] >>= 1; }
        for (int i = 0; i < 8; i++) P.sqrt_exp[i] = t[i];
    }
    P.b = curve_b_coeff<C>();
    if (!txt.iso_a) {
        P.method = 0;
        P.z = F::one();
        const F three = F::from_u64(3), four = F::from_u64(4), two = F::from_u64(2);
        const F gz = P.z.sqr() * P.z + P.b;                          // g(Z), a = 0
        const F h = three * P.z.sqr();                               // 3 Z^2 + 4 a
        P.c1 = gz;
        P.c2 = (P.z * inv_fixed(two)).neg();
        bool sq = false;
        F c3 = sqrt_fixed((gz * h).neg(), P.sqrt_exp, &sq);
        if (!sq) return false;
        if (sgn0(c3)) c3 = c3.neg();
        P.c3 = c3;
        P.c4 = (four * gz * inv_fixed(h)).neg();
    } else {
        P.method = 1;
        P.z = F::from_u64(13).neg();
        P.iso_a = fe_from_hex<F>(txt.iso_a);
        P.iso_b = F::from_u64(1265);
        const F ia = inv_fixed(P.iso_a);
        P.nb_over_a = (P.iso_b * ia).neg();
        P.b_over_za = P.iso_b * ia * inv_fixed(P.z);
        for (int i = 0; i < 13; i++) P.iso[i] = fe_from_hex<F>(txt.iso[i]);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ SHAKE256 (FIPS 202)
struct Shake256 {
    uint64_t st[25];
    uint8_t buf[136];
    size_t pos = 0;
    bool squeezing = false;
    Shake256() { memset(st, 0, sizeof st); }
    static inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
    // Keccak-f[1600], lanes in local variables and every step written out: ~1.7x the table-driven form on one core, which matters
    // because this stream is the one sequential part of from_label (64 MB for a 2^21-point key).
    void permute() {
        static const uint64_t RC[24] = {
            0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull, 0x0000000080000001ull,
            0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
            0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull,
            0x000000000000800aull, 0x800000008000000aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
        uint64_t a00 = st[0], a01 = st[1], a02 = st[2], a03 = st[3], a04 = st[4], a05 = st[5], a06 = st[6], a07 = st[7], a08 = st[8], a09 = st[9],
                 a10 = st[10], a11 = st[11], a12 = st[12], a13 = st[13], a14 = st[14], a15 = st[15], a16 = st[16], a17 = st[17], a18 = st[18],
                 a19 = st[19], a20 = st[20], a21 = st[21], a22 = st[22], a23 = st[23], a24 = st[24];
        for (int r = 0; r < 24; r++) {
            // theta
            const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                           c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
            const uint64_t d0 = c4 ^ rotl(c1, 1), d1 = c0 ^ rotl(c2, 1), d2 = c1 ^ rotl(c3, 1), d3 = c2 ^ rotl(c4, 1), d4 = c3 ^ rotl(c0, 1);
            a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
            a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
            a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
            a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
            a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
            // rho + pi: b[y + 5 ((2x + 3y) mod 5)] = rotl(a[x + 5y], offset[x][y])
            const uint64_t b00 = a00, b10 = rotl(a01, 1), b20 = rotl(a02, 62), b05 = rotl(a03, 28), b15 = rotl(a04, 27);
            const uint64_t b16 = rotl(a05, 36), b01 = rotl(a06, 44), b11 = rotl(a07, 6), b21 = rotl(a08, 55), b06 = rotl(a09, 20);
            const uint64_t b07 = rotl(a10, 3), b17 = rotl(a11, 10), b02 = rotl(a12, 43), b12 = rotl(a13, 25), b22 = rotl(a14, 39);
            const uint64_t b23 = rotl(a15, 41), b08 = rotl(a16, 45), b18 = rotl(a17, 15), b03 = rotl(a18, 21), b13 = rotl(a19, 8);
            const uint64_t b14 = rotl(a20, 18), b24 = rotl(a21, 2), b09 = rotl(a22, 61), b19 = rotl(a23, 56), b04 = rotl(a24, 14);
            // chi, iota
            a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
            a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
            a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
            a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
            a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
            a00 ^= RC[r];
        }
        st[0] = a00; st[1] = a01; st[2] = a02; st[3] = a03; st[4] = a04; st[5] = a05; st[6] = a06; st[7] = a07; st[8] = a08; st[9] = a09;
        st[10] = a10; st[11] = a11; st[12] = a12; st[13] = a13; st[14] = a14; st[15] = a15; st[16] = a16; st[17] = a17; st[18] = a18; st[19] = a19;
        st[20] = a20; st[21] = a21; st[22] = a22; st[23] = a23; st[24] = a24;
    }
    void xor_block(const uint8_t *p) {
        for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, p + 8 * i, 8); st[i] ^= w; }   // little-endian host
    }
    void absorb(const uint8_t *in, size_t n) {
        while (n) {
            size_t k = 136 - pos < n ? 136 - pos : n;
            memcpy(buf + pos, in, k);
            pos += k; in += k; n -= k;
            if (pos == 136) { xor_block(buf); permute(); pos = 0; }
        }
    }
    void squeeze(uint8_t *out, size_t n) {
        if (!squeezing) {
            memset(buf + pos, 0, 136 - pos);
            buf[pos] ^= 0x1f;
            buf[135] ^= 0x80;
            xor_block(buf);
            permute();
            squeezing = true;
            pos = 0;
        }
        while (n) {
            if (pos == 136) { permute(); pos = 0; }
            size_t k = 136 - pos < n ? 136 - pos : n;
            memcpy(out, reinterpret_cast<const uint8_t *>(st) + pos, k);
            pos += k; out += k; n -= k;
        }
    }
};

}  // namespace lurk
