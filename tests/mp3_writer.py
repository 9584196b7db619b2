"""Test-vector writer for MPEG Layer III streams (there is no MP3 encoder, sample file or codec tool offline): it does
not encode audio, it draws RANDOM BUT WELL-FORMED granules — legal side information, scalefactors, Huffman-coded
spectra in every code book, window switching, mixed blocks, MS stereo, the bit reservoir — and serialises them exactly
as ISO/IEC 11172-3 2.4.1 lays a frame out.  Any conforming decoder must produce the same PCM from such a stream, which
is what tests/test_mp3_cpu.py checks between this repo's decoder and libavcodec's.  `tone_stream` writes the one
stream whose PCM is known in closed form: a single spectral line, i.e. a windowed sinusoid at a known frequency."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_T = {}


def tables():
    if _T:
        return _T
    src = open(os.path.join(ROOT, "whisper-rust-ort_b200", "csrc", "host", "mp3_tables.h")).read()

    def arr(name):
        m = re.search(name + r"(?:\[\d+\])+ = \{(.*?)\};", src, re.S)
        return [int(x) for x in re.findall(r"-?\d+", m.group(1))]

    ids, dim, off = arr("kBookIds"), arr("kBookDim"), arr("kBookOff")
    lens, syms = arr("kHuffLen"), arr("kHuffSym")
    books = {}
    for b, (i, d, o) in enumerate(zip(ids, dim, off)):
        acc, enc = 0, {}
        for k in range(d * d):
            ln = lens[o + k]
            enc[syms[o + k]] = (acc >> (32 - ln), ln)
            acc += 1 << (32 - ln)
        books[i] = (d, enc)
    lin = [1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13]
    sel = {}
    for t in range(32):
        if t in (0, 4, 14):
            continue
        sel[t] = (books[t], 0) if t < 16 else (books[16 if t < 24 else 24], lin[t - 16])
    _T["sel"] = sel
    _T["quad_a"] = list(zip(arr("kQuadACode"), arr("kQuadALen")))
    bl, bs = arr("kBandLong"), arr("kBandShort")
    _T["long"] = [bl[22 * r:22 * r + 22] for r in range(9)]
    _T["short"] = [bs[13 * r:13 * r + 13] for r in range(9)]
    nsf = arr("kLsfPartitions")
    _T["nsf"] = [[nsf[12 * a + 4 * b:12 * a + 4 * b + 4] for b in range(3)] for a in range(6)]
    return _T


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, v, n):
        assert 0 <= v < (1 << n) or n == 0, (v, n)
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def __len__(self):
        return len(self.bits)

    def tobytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(int("".join(map(str, b[i:i + 8])), 2) for i in range(0, len(b), 8))


SR = [[44100, 48000, 32000], [22050, 24000, 16000], [11025, 12000, 8000]]
BR = [[0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320], [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]]
SLEN = [[0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4], [0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3]]


def header(version, sr_idx, br_idx, pad, mode, mode_ext, crc=False):
    vbits = {0: 3, 1: 2, 2: 0}[version]
    w = (0x7FF << 21) | (vbits << 19) | (1 << 17) | ((0 if crc else 1) << 16) | (br_idx << 12) | (sr_idx << 10) | (pad << 9) | (mode << 6) | (mode_ext << 4)
    return w.to_bytes(4, "big")


def frame_size(version, sr_idx, br_idx, pad):
    return (144 if version == 0 else 72) * BR[0 if version == 0 else 1][br_idx] * 1000 // SR[version][sr_idx] + pad


def lsf_slen(sfc):
    if sfc < 400:
        return [(sfc >> 4) // 5, (sfc >> 4) % 5, (sfc % 16) >> 2, sfc % 4], 0
    if sfc < 500:
        sfc -= 400
        return [(sfc >> 2) // 5, (sfc >> 2) % 5, sfc % 4, 0], 1
    sfc -= 500
    return [sfc // 3, sfc % 3, 0, 0], 2


def lsf_slen_intensity(sfc):
    v = sfc >> 1
    if v < 180:
        return [v // 36, (v % 36) // 6, (v % 36) % 6, 0], 3
    if v < 244:
        v -= 180
        return [(v % 64) >> 4, (v % 16) >> 2, v % 4, 0], 4
    v -= 244
    return [v // 3, v % 3, 0, 0], 5


def random_granule(rng, version, row, budget_bits, gr, prev, allow_scfsi, spec=None, force=None, is_right=False):
    """Draws one granule/channel: returns (side-info dict, main-data BitWriter)."""
    T = tables()
    for attempt in range(50):
        g = {"scfsi": [0, 0, 0, 0]}
        g["ws"] = int(rng.random() < 0.5)
        g["block_type"] = int(rng.integers(1, 4)) if g["ws"] else 0
        g["mixed"] = int(g["ws"] and g["block_type"] == 2 and rng.random() < 0.4 and row != 8)
        g["global_gain"] = int(rng.integers(100, 190))
        g["scalefac_scale"] = int(rng.integers(0, 2))
        g["count1table"] = int(rng.integers(0, 2))
        g["preflag"] = int(rng.integers(0, 2)) if version == 0 else 0
        g["subblock_gain"] = [int(rng.integers(0, 8)) for _ in range(3)]
        if spec is not None:                                    # explicit spectrum: plain long blocks, all scalefactors zero
            g.update(ws=0, block_type=0, mixed=0, global_gain=spec["global_gain"], scalefac_scale=0, preflag=0)
        if force is not None:
            g.update(force(gr) if callable(force) else force)
        shortb = g["block_type"] == 2
        w = BitWriter()
        # ---- part 2: scalefactors
        if spec is not None:
            g["sfc"] = 0
        elif version == 0:
            g["sfc"] = int(rng.integers(0, 16))
            s1, s2 = SLEN[0][g["sfc"]], SLEN[1][g["sfc"]]
            if shortb:
                n1 = 17 if g["mixed"] else 18
                for _ in range(n1):
                    w.put(int(rng.integers(0, 1 << s1)), s1)
                for _ in range(18):
                    w.put(int(rng.integers(0, 1 << s2)), s2)
            else:
                if gr == 1 and allow_scfsi and prev is not None and prev["block_type"] != 2:
                    g["scfsi"] = [int(rng.integers(0, 2)) for _ in range(4)]
                for k, (a, b) in enumerate([(0, 6), (6, 11), (11, 16), (16, 21)]):
                    if g["scfsi"][k]:
                        continue
                    for _ in range(a, b):
                        s = s1 if k < 2 else s2
                        w.put(int(rng.integers(0, 1 << s)), s)
        else:
            g["sfc"] = int(rng.integers(0, 512))
            slen, r = lsf_slen_intensity(g["sfc"]) if is_right else lsf_slen(g["sfc"])
            col = (2 if g["mixed"] else 1) if shortb else 0
            for k in range(4):
                for _ in range(T["nsf"][r][col][k]):
                    # right channel of an intensity-stereo pair: positions; the all-ones value is "illegal", which decoders treat differently
                    # (and libavcodec knows only positions 0..15 of the up-to-5-bit LSF field)
                    top = min(16, max(1, (1 << slen[k]) - 1)) if is_right else 1 << slen[k]
                    w.put(int(rng.integers(0, top)), slen[k])
        # ---- part 3: Huffman-coded spectrum
        scale = 0.25 * (0.5 ** attempt)
        bv = int(rng.integers(0, max(1, int(288 * scale)) + 1))
        if g["ws"]:
            g["region0_count"], g["region1_count"] = (8 if shortb and not g["mixed"] else 7), 36
            if shortb:
                r1 = 72 if row == 8 else 36
            else:
                r1 = 36 if version == 0 else (108 if row == 8 else 54)
            r2 = 576
            nreg = 2
        else:
            g["region0_count"] = int(rng.integers(0, 16))
            g["region1_count"] = int(rng.integers(0, 8))
            edge = np.concatenate([[0], np.cumsum(T["long"][row])])
            r1 = int(edge[min(g["region0_count"] + 1, 22)])
            r2 = int(edge[min(g["region0_count"] + g["region1_count"] + 2, 22)])
            nreg = 3
        ids = sorted(T["sel"].keys())
        if is_right:
            # the intensity bound is "right channel all zero": keep to the books without escape values, whose lines libavcodec
            # dequantises in float (its integer path for escaped values rounds very small magnitudes to an exact 0)
            ids = [t for t in ids if t < 16]
        g["table_select"] = [int(rng.choice([0] + ids)) if rng.random() < 0.9 else 0 for _ in range(nreg)]
        if spec is not None:
            bv, g["table_select"] = spec["big_values"], spec["table_select"]
        g["big_values"] = bv
        lim = [min(r1, 2 * bv), min(r2, 2 * bv), 2 * bv]
        pos, peak = 0, 1
        for r in range(3):
            t = g["table_select"][r] if r < nreg else 0
            while pos < lim[r]:
                if t != 0:
                    (d, enc), lb = T["sel"][t]
                    vals = []
                    for _ in range(2):
                        if spec is not None:
                            v = spec["is"][pos + len(vals)]
                        else:
                            v = int(rng.integers(0, d))
                            if v == d - 1 and lb and d == 16:
                                v += int(rng.integers(0, 1 << lb)) if rng.random() < 0.5 else 0
                            if rng.random() < 0.5:
                                v = -v
                        vals.append(v)
                    x, y = abs(vals[0]), abs(vals[1])
                    peak = max(peak, x, y)
                    code, ln = enc[(min(x, 15) << 4) | min(y, 15)]
                    w.put(code, ln)
                    if x >= 15 and lb:
                        w.put(x - 15, lb)
                    if x:
                        w.put(int(vals[0] < 0), 1)
                    if y >= 15 and lb:
                        w.put(y - 15, lb)
                    if y:
                        w.put(int(vals[1] < 0), 1)
                pos += 2
        nq = 0 if spec is not None else int(rng.integers(0, max(1, (576 - 2 * bv) // 4 + 1) * min(1.0, 4 * scale) + 1))
        nq = min(nq, (576 - 2 * bv) // 4)
        for _ in range(nq):
            v = [int(rng.integers(0, 2)) for _ in range(4)]
            sym = (v[0] << 3) | (v[1] << 2) | (v[2] << 1) | v[3]
            if g["count1table"]:
                w.put(15 - sym, 4)
            else:
                w.put(*T["quad_a"][sym])
            for k in range(4):
                if v[k]:
                    w.put(int(rng.integers(0, 2)), 1)
        if spec is None:                                        # keep |xr| = |is|^(4/3) 2^((gain-210)/4) below ~1/4: a real encoder's range
            top = 210 - 8 - int(np.ceil(4 * (4 / 3) * np.log2(peak + 1)))
            g["global_gain"] = int(rng.integers(max(0, top - 60), top + 1))
        if len(w) <= min(budget_bits, 4095):
            g["part2_3_length"] = len(w)
            return g, w
    raise RuntimeError("could not fit a granule into %d bits" % budget_bits)


def side_info(version, channels, main_data_begin, grs):
    w = BitWriter()
    if version == 0:
        w.put(main_data_begin, 9)
        w.put(0, 5 if channels == 1 else 3)
        for c in range(channels):
            for k in range(4):
                w.put(grs[1][c]["scfsi"][k], 1)
    else:
        w.put(main_data_begin, 8)
        w.put(0, 1 if channels == 1 else 2)
    for gr in range(len(grs)):
        for c in range(channels):
            g = grs[gr][c]
            w.put(g["part2_3_length"], 12)
            w.put(g["big_values"], 9)
            w.put(g["global_gain"], 8)
            w.put(g["sfc"], 4 if version == 0 else 9)
            w.put(g["ws"], 1)
            if g["ws"]:
                w.put(g["block_type"], 2)
                w.put(g["mixed"], 1)
                for r in range(2):
                    w.put(g["table_select"][r], 5)
                for k in range(3):
                    w.put(g["subblock_gain"][k], 3)
            else:
                for r in range(3):
                    w.put(g["table_select"][r], 5)
                w.put(g["region0_count"], 4)
                w.put(g["region1_count"], 3)
            if version == 0:
                w.put(g["preflag"], 1)
            w.put(g["scalefac_scale"], 1)
            w.put(g["count1table"], 1)
    out = w.tobytes()
    assert len(out) == ((17 if channels == 1 else 32) if version == 0 else (9 if channels == 1 else 17)), len(out)
    return out


def make_stream(seed, version=0, sr_idx=0, br_idx=9, channels=2, n_frames=6, ms=False, crc=False, reservoir=True, specs=None, force=None, intensity=False):
    """-> (list of frame bytes, sample rate).  `specs`: optional {(frame, gr, ch): spec} with explicit spectra."""
    rng = np.random.default_rng(seed)
    row = version * 3 + sr_idx
    n_gr = 2 if version == 0 else 1
    mode = 3 if channels == 1 else (1 if ms else int(rng.choice([0, 2])))
    mode_ext = (2 if (ms and channels == 2) else 0) | (1 if (intensity and channels == 2) else 0)
    if mode_ext:
        mode = 1
    side = (17 if channels == 1 else 32) if version == 0 else (9 if channels == 1 else 17)
    head = 4 + (2 if crc else 0)
    max_back = 511 if version == 0 else 255
    sizes, pads = [], []
    for _ in range(n_frames):
        pad = int(rng.integers(0, 2))
        pads.append(pad)
        sizes.append(frame_size(version, sr_idx, br_idx, pad))
    areas = [s - head - side for s in sizes]
    starts = np.concatenate([[0], np.cumsum(areas)])            # byte offset of each frame's own main-data area in the pipe
    pipe = bytearray(int(starts[-1]))
    wpos, frames_side = 0, []
    last_mixed = [0] * channels
    last_bt = [0] * channels                                    # window sequence per channel: long -> start -> short... -> stop -> long (2.4.2.7:
    nxt = {0: [0, 0, 1], 1: [2], 2: [2, 2, 3], 3: [0, 1]}       # a short block is only ever entered through a start block and left through a stop block)
    for f in range(n_frames):
        a0, a1 = int(starts[f]), int(starts[f + 1])
        wpos = max(wpos, a0 - (max_back if reservoir else 0))
        if not reservoir:
            wpos = a0
        budget = (a1 - wpos) * 8
        grs = [[None] * channels for _ in range(n_gr)]
        body = BitWriter()
        for gr in range(n_gr):
            for c in range(channels):
                left = budget - len(body)
                share = left // ((n_gr - gr) * channels - c) if rng.random() < 0.5 else left // 2
                spec = specs.get((f, gr, c)) if specs else None
                f_c = force
                if f_c is None and spec is None:
                    bt = int(rng.choice(nxt[last_bt[c]]))
                    mixed = int(bt == 2 and rng.random() < 0.4 and row != 8)
                    if bt == 2 and last_bt[c] == 2:             # a run of short blocks keeps its layout: the long-windowed low
                        mixed = last_mixed[c]                   # subbands of a mixed block are no legal neighbour of short windows
                    f_c = dict(ws=int(bt != 0), block_type=bt, mixed=mixed)
                    if mode_ext and c == 1:                     # joint stereo pairs lines of the two channels: same window layout in both
                        f_c = {k: grs[gr][0][k] for k in ("ws", "block_type", "mixed")}
                g, w = random_granule(rng, version, row, max(share, 80), gr, grs[0][c] if gr else None, allow_scfsi=True, spec=spec, force=f_c, is_right=bool(mode_ext & 1) and c == 1)
                last_bt[c], last_mixed[c] = g["block_type"], g["mixed"]
                grs[gr][c] = g
                body.bits += w.bits
        data = body.tobytes()
        assert wpos + len(data) <= a1, (wpos, len(data), a1)
        pipe[wpos:wpos + len(data)] = data
        frames_side.append(side_info(version, channels, a0 - wpos, grs))
        wpos += len(data)
    frames = []
    for f in range(n_frames):
        fb = header(version, sr_idx, br_idx, pads[f], mode, mode_ext, crc) + (b"\x00\x00" if crc else b"") + frames_side[f] + bytes(pipe[int(starts[f]):int(starts[f + 1])])
        assert len(fb) == sizes[f]
        frames.append(fb)
    return frames, SR[version][sr_idx]
