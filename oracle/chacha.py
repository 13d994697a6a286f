"""ChaCha keystream and the Rep3 mask derivation (oracle; test infrastructure only).

Restates `Rep3Rand::masking_field_elements_vec` (mpc-core/src/protocols/rep3/rngs.rs:137-156) on top
of `RngType = rand_chacha::ChaCha12Rng` (mpc-core/src/lib.rs:13; rand_chacha 0.3.1, Cargo.lock:3837,
not vendored): `fill_bytes` emits the ChaCha keystream of the seed as key, 64-bit block counter from 0,
64-bit stream id 0, 12 rounds, words little-endian; element i is
`from_be_bytes_mod_order(a[32i..32i+32]) - from_be_bytes_mod_order(b[32i..32i+32])`.
The block function is checked against the RFC 7539 section 2.3.2 ChaCha20 vector (same core, 20 rounds);
the 12-round stream itself has no vector in the reference tree (parity of the PRF *stream* with
rand_chacha is therefore restated, not pinned).
"""
import struct

MASK = 0xffffffff


def _rotl(x, n):
    return ((x << n) & MASK) | (x >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & MASK; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & MASK; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & MASK; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & MASK; s[b] = _rotl(s[b] ^ s[c], 7)


def block(key_words, counter64, stream64, rounds=12):
    init = [0x61707865, 0x3320646e, 0x79622d32, 0x6b206574] + list(key_words) + [
        counter64 & MASK, (counter64 >> 32) & MASK, stream64 & MASK, (stream64 >> 32) & MASK]
    s = list(init)
    for _ in range(rounds // 2):
        _qr(s, 0, 4, 8, 12); _qr(s, 1, 5, 9, 13); _qr(s, 2, 6, 10, 14); _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15); _qr(s, 1, 6, 11, 12); _qr(s, 2, 7, 8, 13); _qr(s, 3, 4, 9, 14)
    return [(x + y) & MASK for x, y in zip(s, init)]


def keystream_words(seed32, word_pos, nwords, rounds=12):
    key = struct.unpack("<8I", seed32)
    out = []
    w = word_pos
    while len(out) < nwords:
        blk = block(key, w >> 4, 0, rounds)
        take = blk[w & 15:]
        out += take[:nwords - len(out)]
        w += len(take)
    return out


def masking_field_elements_vec(seed1, pos1, seed2, pos2, n, r):
    a = keystream_words(seed1, pos1, 8 * n)
    b = keystream_words(seed2, pos2, 8 * n)

    def be(words):
        return int.from_bytes(struct.pack("<8I", *words), "big")

    return [(be(a[8 * i:8 * i + 8]) - be(b[8 * i:8 * i + 8])) % r for i in range(n)]
