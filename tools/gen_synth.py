"""Seeded synthetic corpora shaped like the BASELINE.json configs.

No file of the reference is read: the word list is generated from a small syllable
inventory with a Zipf-like rank distribution, wrapped in enwik-style XML/wiki markup.

    python tools/gen_synth.py text 100000 out.txt [seed]
    python tools/gen_synth.py binary 100000 out.bin [seed]
"""
import sys
import numpy as np

_ONSETS = ["", "b", "c", "d", "f", "g", "h", "l", "m", "n", "p", "r", "s", "t", "w", "st", "tr", "ch", "sh", "th", "pr", "gr", "pl", "br"]
_VOWELS = ["a", "e", "i", "o", "u", "ea", "ou", "io", "ai", "ee"]
_CODAS = ["", "n", "r", "s", "t", "l", "m", "d", "ng", "nt", "st", "ck", "rd", "ll", "ss"]
_COMMON = ("the of and to in a is that for it as was with be by on not he this are or his from at which but have an had they "
           "you were their one all we can her has there been if more when will would who so no out up into than them only "
           "its time some could these two may first then do any my now such like our over man even most made after also did "
           "many before must through years where much your way well down should because each just those people how too little "
           "state good very make world still own see men work long get here between both life being under never day same "
           "another know while last might us great old year off come since against go came right used take three").split()


def _lexicon(rng, n=6000):
    words = list(_COMMON)
    seen = set(words)
    while len(words) < n:
        k = 1 + min(3, int(rng.geometric(0.55)))
        w = "".join(_ONSETS[rng.integers(len(_ONSETS))] + _VOWELS[rng.integers(len(_VOWELS))] + _CODAS[rng.integers(len(_CODAS))]
                    for _ in range(k))
        if w not in seen and 2 <= len(w) <= 14:
            seen.add(w)
            words.append(w)
    return words


def synth_text(n_bytes, seed=0xE9E80001):
    rng = np.random.default_rng(seed)
    words = _lexicon(rng)
    ranks = np.arange(1, len(words) + 1, dtype=np.float64)
    pz = ranks ** -1.07
    pz /= pz.sum()
    out = []
    size = 0
    page = 0
    while size < n_bytes:
        page += 1
        title = " ".join(words[i].capitalize() for i in rng.choice(len(words), size=1 + rng.integers(3), p=pz))
        head = "  <page>\n    <title>%s</title>\n    <id>%d</id>\n    <revision>\n      <text xml:space=\"preserve\">" % (title, page)
        body = []
        body_len = 0
        while body_len < 3500:
            n_sent = 2 + rng.integers(6)
            para = []
            for _ in range(n_sent):
                n_w = 3 + rng.poisson(14)
                idx = rng.choice(len(words), size=n_w, p=pz)
                toks = []
                for k, i in enumerate(idx):
                    w = words[i]
                    r = rng.random()
                    if r < 0.03:
                        w = "[[" + w + "]]"
                    elif r < 0.05:
                        w = str(int(rng.integers(1, 2100)))
                    elif r < 0.053:
                        w = "&quot;" + w + "&quot;"
                    elif r < 0.06:
                        w = "''" + w + "''"
                    if k == 0:
                        w = w[:1].upper() + w[1:]
                    if k < n_w - 1 and rng.random() < 0.08:
                        w += str(rng.choice([",", ";", ":", " (", ")"]))
                    toks.append(w)
                para.append(" ".join(toks) + str(rng.choice([".", ".", ".", "?", "!"])))
            p = " ".join(para) + "\n\n"
            body.append(p)
            body_len += len(p)
        s = head + "".join(body) + "</text>\n    </revision>\n  </page>\n"
        out.append(s)
        size += len(s)
    return "".join(out).encode("ascii")[:n_bytes]


def synth_binary(n_bytes, seed=0xE9E80003):
    """Alternating 64 KiB blocks: x86-64-ELF-like opcode streams and JPEG-header blocks."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    opcodes = rng.integers(0, 256, size=256, dtype=np.uint8)
    weights = rng.dirichlet(np.full(256, 0.3))
    blk = 0
    while len(out) < n_bytes:
        if blk % 2 == 0:
            b = bytearray(b"\x7fELF\x02\x01\x01" + bytes(9) + b"\x02\x00\x3e\x00\x01\x00\x00\x00" + bytes(40))
            targets = rng.integers(0, 1 << 16, size=32)
            while len(b) < 65536:
                run = rng.choice(opcodes, size=int(rng.integers(8, 40)), p=weights)
                b += bytes(run)
                tgt = int(targets[rng.integers(32)]) - (len(b) & 0xFFFF)
                b += bytes([0xE8 if rng.random() < 0.7 else 0xE9]) + int(tgt & 0xFFFFFFFF).to_bytes(4, "little")
            out += b[:65536]
        else:
            b = bytearray(b"\xff\xd8\xff\xe0\x00\x10JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
            for t in range(2):
                b += b"\xff\xdb\x00\x43" + bytes([t]) + bytes(int(v) for v in np.clip(rng.integers(1, 100, size=64), 1, 255))
            b += b"\xff\xc0\x00\x11\x08\x01\xe0\x02\x80\x03\x01\x22\x00\x02\x11\x01\x03\x11\x01"
            b += b"\xff\xda\x00\x0c\x03\x01\x00\x02\x11\x03\x11\x00\x3f\x00"
            ent = rng.integers(0, 256, size=65536, dtype=np.uint8)
            ent = bytes(ent).replace(b"\xff", b"\xff\x00")
            b += ent
            out += b[:65534] + b"\xff\xd9"
        blk += 1
    return bytes(out[:n_bytes])


if __name__ == "__main__":
    kind, n, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    seed = int(sys.argv[4], 0) if len(sys.argv) > 4 else None
    fn = synth_text if kind == "text" else synth_binary
    data = fn(n, seed) if seed is not None else fn(n)
    open(path, "wb").write(data)
