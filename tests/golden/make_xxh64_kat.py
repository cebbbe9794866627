"""Generates tests/golden/xxh64_kat.json.

The reference pins no XXH64 VALUE (third-party github.com/cespare/xxhash/v2 v2.3.0, go.mod:6,
source not vendored; approximateprefix tests only pin counts).  These known-answer vectors pin
our oracle + CUDA hash against two independent XXH64 implementations available in this image:
python-xxhash 3.7.0 and the system libxxhash.so.0.8.2.  The chained vectors follow
approximateprefix/hashing.go:70-94 literally: seed link = XXH64(model||salt),
h_i = XXH64(block_i || LE64(h_{i-1})), trailing partial block included.

Run (in the build container): python tests/golden/make_xxh64_kat.py
"""
import ctypes
import json
import os
import random
import struct

import xxhash

lib = ctypes.CDLL("libxxhash.so.0")
lib.XXH64.restype = ctypes.c_uint64
lib.XXH64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]


def h64(b: bytes) -> int:
    a = xxhash.xxh64(b, seed=0).intdigest()
    c = lib.XXH64(b, len(b), 0)
    assert a == c, (a, c)
    return a


def chain(prompt: bytes, model: bytes, salt: bytes, block_chars: int, max_blocks: int):
    if block_chars <= 0 or len(prompt) < block_chars:
        return []
    if len(prompt) > block_chars * max_blocks:
        prompt = prompt[: block_chars * max_blocks]
    prev = h64(model + salt)
    out = []
    i = 0
    while i + block_chars <= len(prompt):
        prev = h64(prompt[i:i + block_chars] + struct.pack("<Q", prev))
        out.append(prev)
        i += block_chars
    if i < len(prompt):
        out.append(h64(prompt[i:] + struct.pack("<Q", prev)))
    return out


def main():
    rng = random.Random(20260922)
    raw = []
    for n in list(range(0, 100)) + [127, 128, 129, 255, 256, 1000, 2056]:
        b = bytes(rng.randrange(256) for _ in range(n))
        raw.append({"hex": b.hex(), "xxh64": f"{h64(b):016x}"})
    spec = [{"ascii": "", "xxh64": "ef46db3751d8e999"}, {"ascii": "a", "xxh64": "d24ec4f1a98c6e5b"},
            {"ascii": "abc", "xxh64": "44bc2cf5ad770999"}]
    for s in spec:
        assert f"{h64(s['ascii'].encode()):016x}" == s["xxh64"]
    chains = []
    letters = b"abcdefghijklmnopqrstuvwxyz"

    def add(prompt, model, salt, bc, mb):
        chains.append({"prompt_hex": prompt.hex(), "model": model.decode(), "salt": salt.decode(),
                       "block_chars": bc, "max_blocks": mb,
                       "seed": f"{h64(model + salt):016x}",
                       "hashes": [f"{x:016x}" for x in chain(prompt, model, salt, bc, mb)]})

    # the prompts the reference's own tests use (plugin_test.go:61,139,209,455,511)
    add(b"aaaabbbb", b"test-model1", b"", 4, 256)
    add(b"aaaaaa", b"test-model1", b"", 4, 256)
    add(b"aaaabbbbccccdddd", b"test-model", b"", 4, 2)
    add(b"aaaabbbbccccdddd", b"test-model", b"", 4, 3)
    add(b"a" * 128, b"test-model", b"", 64, 256)
    add(b"test1", b"my-model", b"", 64, 256)  # integration prompts: shorter than one block ⇒ no hashes
    # random prompts over the block sizes / alignments the engine has fast and generic paths for
    for bc in (4, 12, 28, 32, 60, 64, 96, 128, 256):
        for n in (bc - 1, bc, bc + 1, 3 * bc, 3 * bc + 5, 33 * bc + 7, 2048, 2049):
            if n < 0:
                continue
            p = bytes(letters[rng.randrange(26)] for _ in range(n))
            add(p, b"model-%d" % bc, b"salt" if bc % 8 == 4 else b"", bc, 256)
    add(bytes(letters[rng.randrange(26)] for _ in range(5000)), b"m", b"", 64, 16)  # truncation at max_blocks
    out = {"source": "python-xxhash 3.7.0 == libxxhash 0.8.2 (asserted equal for every vector)",
           "spec_vectors": spec, "raw": raw, "chains": chains}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "xxh64_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path, len(raw), "raw,", len(chains), "chains")


if __name__ == "__main__":
    main()
