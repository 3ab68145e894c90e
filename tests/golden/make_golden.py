#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

Nothing here uses oracle/ or the CUDA library: expected values come from INDEPENDENT sources so the fixtures
can pin both:
  * keccak256   — the 8 input shapes of the reference's live tests (src/testing/tests/precompiles/keccak256.rs:144-196:
                  [], [123;50], [123;136], [123;200], each at byte offset 0 and 31), digests from the pure-Python
                  keccak below, which is itself checked against hashlib.sha3_256 (same permutation and sponge, pad 0x06)
                  and the universal KAT keccak256("") = c5d2...a470.
  * sha256      — the three inputs of src/testing/tests/precompiles/sha256.rs:119-136 ([], [255;256], [255;10000]),
                  digests from hashlib.sha256.
  * ecrecover   — the two known-answer vectors of src/testing/tests/precompiles/ecrecover.rs:127-143 (copied as DATA:
                  input words hash|v|r|s and the expected address).
  * ecrecover_generated — signatures from the `cryptography` package + failure shapes, expected results from a
                  Python-int secp256k1 (below).
  * u256        — ALU cases (add/sub/mul/div/shl/shr/rol/ror/xor/and/or) with results and flags from Python ints
                  following src/opcodes/execution/{add,sub,mul,div,shift,binop}.rs.
"""
import hashlib
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
M256 = (1 << 256) - 1

RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808a, 0x8000000080008000, 0x000000000000808b,
      0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008a, 0x0000000000000088,
      0x0000000080008009, 0x000000008000000a, 0x000000008000808b, 0x800000000000008b, 0x8000000000008089,
      0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800a, 0x800000008000000a,
      0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
M64 = (1 << 64) - 1


def rol64(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & M64 if n else x


def keccak_f(a):
    """a[x][y] 5x5 lanes; textbook formulation (theta, rho+pi via the (x,y)->(y,2x+3y) walk, chi, iota)"""
    for rnd in range(24):
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ rol64(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        x, y, cur = 1, 0, a[1][0]
        for t in range(24):
            x, y = y, (2 * x + 3 * y) % 5
            cur, a[x][y] = a[x][y], rol64(cur, (t + 1) * (t + 2) // 2)
        a = [[a[x][y] ^ (~a[(x + 1) % 5][y] & M64 & a[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= RC[rnd]
    return a


def sponge256(data: bytes, pad: int) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(pad)
    msg += b"\x00" * (-len(msg) % rate)
    msg[-1] ^= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def keccak256(data: bytes) -> bytes:
    return sponge256(data, 0x01)


def u256_case(op, a, b):
    fl = {}
    if op == "add":
        r = a + b
        of = r > M256
        r &= M256
        out = (r, 0)
        fl = dict(of=of, eq=r == 0, gt=(r != 0) and not of)            # add.rs:35-43
    elif op == "sub":
        of = a < b
        r = (a - b) & M256
        out = (r, 0)
        fl = dict(of=of, eq=r == 0, gt=(r != 0) and not of)            # sub.rs:35-44
    elif op == "mul":
        p = a * b
        lo, hi = p & M256, p >> 256
        out = (lo, hi)
        fl = dict(of=hi != 0, eq=lo == 0, gt=(hi == 0) and (lo != 0))  # mul.rs:41-49
    elif op == "div":
        if b == 0:
            out = (0, 0)
            fl = dict(of=True, eq=False, gt=False)                      # div.rs:36-48
        else:
            q, r = divmod(a, b)
            out = (q, r)
            fl = dict(of=False, eq=q == 0, gt=r == 0)                   # div.rs:50-75
    elif op in ("shl", "shr", "rol", "ror"):
        n = b & 0xFF                                                    # shift.rs:44
        shl = lambda v, k: (v << k) & M256 if k < 256 else 0
        shr = lambda v, k: v >> k if k < 256 else 0
        if op == "shl":
            r = shl(a, n)
        elif op == "shr":
            r = shr(a, n)
        elif op == "rol":
            r = shl(a, n) | shr(a, 256 - n)                             # shift.rs:50-52 (n = 0 => shift by 256 => 0)
        else:
            r = shr(a, n) | shl(a, 256 - n)
        out = (r, 0)
        fl = dict(of=False, eq=r == 0, gt=False)                        # shift.rs:63-67
    else:
        r = {"xor": a ^ b, "and": a & b, "or": a | b}[op]
        out = (r, 0)
        fl = dict(of=False, eq=r == 0, gt=False)                        # binop.rs:47-51
    flags = int(fl["of"]) | int(fl["eq"]) << 1 | int(fl["gt"]) << 2
    return {"op": op, "a": hex(a), "b": hex(b), "out0": hex(out[0]), "out1": hex(out[1]), "flags": flags}


# ---- secp256k1 in Python ints (independent of oracle/secp256k1.hpp and of the CUDA code) ---------------------------
SP = 2**256 - 2**32 - 977
SN = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
SG = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
      0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def ec_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % SP == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, SP) % SP
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, SP) % SP
    x = (lam * lam - a[0] - b[0]) % SP
    return x, (lam * (a[0] - x) - a[1]) % SP


def ec_mul(k, pt):
    acc = None
    while k:
        if k & 1:
            acc = ec_add(acc, pt)
        pt = ec_add(pt, pt)
        k >>= 1
    return acc


def py_ecrecover(z, r, s, v_odd):
    if not (0 < r < SN and 0 < s < SN):
        return None
    rhs = (pow(r, 3, SP) + 7) % SP
    y = pow(rhs, (SP + 1) // 4, SP)
    if y * y % SP != rhs:
        return None
    if (y & 1) != int(v_odd):
        y = SP - y
    rinv = pow(r, -1, SN)
    q = ec_add(ec_mul((-z * rinv) % SN, SG), ec_mul(s * rinv % SN, (r, y)))
    if q is None:
        return None
    return keccak256(q[0].to_bytes(32, "big") + q[1].to_bytes(32, "big"))[12:]


def ecrecover_cases():
    """signatures made with the `cryptography` package (its own secp256k1), recovered with the Python-int code above;
    plus the failure shapes: r = 0, s = 0, r >= n, s >= n, x = r not on the curve, wrong parity (a different key)."""
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import ec
    from cryptography.hazmat.primitives.asymmetric.utils import Prehashed, decode_dss_signature
    rng = random.Random(0xEC)
    out = []
    for i in range(6):
        d = rng.randrange(1, SN)
        key = ec.derive_private_key(d, ec.SECP256K1())
        nums = key.public_key().public_numbers()
        expected = keccak256(nums.x.to_bytes(32, "big") + nums.y.to_bytes(32, "big"))[12:]
        digest = rng.randbytes(32) if i else b"\xff" * 32            # digest >= n exercises the reduction mod n
        r, s = decode_dss_signature(key.sign(digest, ec.ECDSA(Prehashed(hashes.SHA256()))))
        z = int.from_bytes(digest, "big")
        hits = [v for v in (0, 1) if py_ecrecover(z, r, s, v) == expected]
        assert len(hits) == 1
        for v in (0, 1):
            got = py_ecrecover(z, r, s, v)
            out.append({"digest": digest.hex(), "r": hex(r), "s": hex(s), "v": v, "ok": got is not None,
                        "address": got.hex() if got else "", "signer": v == hits[0]})
    z = rng.getrandbits(256)
    good_r = out[0]["r"]
    for r, s in ((0, 5), (5, 0), (SN, 5), (int(good_r, 16), SN), (SN + 3, 7), (2**256 - 1, 2**256 - 1)):
        out.append({"digest": z.to_bytes(32, "big").hex(), "r": hex(r), "s": hex(s), "v": 0, "ok": False, "address": "", "signer": False})
    r = 1
    while py_ecrecover(z, r, 12345, 0) is not None:    # smallest r whose x has no point on the curve
        r += 1
    out.append({"digest": z.to_bytes(32, "big").hex(), "r": hex(r), "s": hex(12345), "v": 0, "ok": False, "address": "", "signer": False})
    return out


def main():
    assert sponge256(b"", 0x06) == hashlib.sha3_256(b"").digest()
    for n in (1, 135, 136, 137, 271, 272, 1000):
        d = bytes((i * 7 + n) & 0xFF for i in range(n))
        assert sponge256(d, 0x06) == hashlib.sha3_256(d).digest(), n
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"

    vec = {"keccak256": [], "sha256": [], "ecrecover": []}
    for n in (0, 50, 136, 200):
        for un in (0, 31):
            data = bytes([123]) * n
            vec["keccak256"].append({"byte": 123, "len": n, "unalignment": un, "digest": keccak256(data).hex(),
                                     "ref": "src/testing/tests/precompiles/keccak256.rs:144-196"})
    for n, un in ((4096, 0), (4096, 31), (135, 5), (137, 17), (272, 1)):      # BASELINE config 3 shapes + block edges
        data = bytes((i * 31 + 7) & 0xFF for i in range(n))
        vec["keccak256"].append({"pattern": "(i*31+7)&0xff", "len": n, "unalignment": un, "digest": keccak256(data).hex(),
                                 "ref": "BASELINE.json configs[2] / rate-boundary cases"})
    for n in (0, 256, 10000):
        vec["sha256"].append({"byte": 255, "len": n, "digest": hashlib.sha256(bytes([255]) * n).hexdigest(),
                              "ref": "src/testing/tests/precompiles/sha256.rs:119-136"})
    vec["ecrecover"] = [
        {"input": "38d18acb67d25c8bb9942764b62f18e17054f66a817bd4295423adf9ed98873e"
                  "000000000000000000000000000000000000000000000000000000000000001b"
                  "38d18acb67d25c8bb9942764b62f18e17054f66a817bd4295423adf9ed98873e"
                  "789d1dd423d25f0772d2748d60f7e4b81bb14d086eba8e8e8efb6dcff8a4ae02",
         "layout": "hash | v (27/28) | r | s", "address": "ceaccac640adf55b2028469bd36ba501f28b699d",
         "ref": "src/testing/tests/precompiles/ecrecover.rs:127-133"},
        {"input": "38d18acb67d25c8bb9942764b62f18e17054f66a817bd4295423adf9ed98873e"
                  "000000000000000000000000000000000000000000000000000000000000001b"
                  "38d18acb67d25c8bb9942764b62f18e17054f66a817bd4295423adf9ed98873e"
                  "7fffffffffffffffffffffffffffffff5d576e7357a4501ddfe92f46681b20a0",
         "layout": "hash | v (27/28) | r | s",
         "address": bytes([88, 198, 174, 93, 17, 93, 119, 163, 216, 169, 239, 54, 214, 164, 45, 35, 105, 43, 170, 127]).hex(),
         "ref": "src/testing/tests/precompiles/ecrecover.rs:135-143"},
    ]
    vec["ecrecover_generated"] = ecrecover_cases()
    with open(os.path.join(HERE, "hash_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)

    rng = random.Random(0x5EED0001)
    edge = [0, 1, 2, 255, 256, 257, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 128, 1 << 255,
            M256 - 1, M256]
    cases = []
    for op in ("add", "sub", "mul", "div", "shl", "shr", "rol", "ror", "xor", "and", "or"):
        pairs = [(a, b) for a in edge for b in edge if rng.random() < 0.35]
        for _ in range(40):
            bits_a, bits_b = rng.choice([8, 33, 64, 100, 128, 200, 256]), rng.choice([8, 33, 64, 100, 128, 200, 256])
            pairs.append((rng.getrandbits(bits_a), rng.getrandbits(bits_b)))
        if op in ("shl", "shr", "rol", "ror"):
            pairs += [(rng.getrandbits(256), n) for n in (0, 1, 31, 32, 33, 63, 64, 127, 128, 200, 255, 256, 257, 511, 1 << 40)]
        for a, b in pairs:
            cases.append(u256_case(op, a, b))
    with open(os.path.join(HERE, "u256_vectors.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print(f"wrote {len(vec['keccak256'])} keccak, {len(vec['sha256'])} sha256, {len(vec['ecrecover'])} ecrecover, {len(cases)} u256 cases")


if __name__ == "__main__":
    main()
