"""The parallel decomposition behind the device-side GOOGLE encoder (csrc/encode_google.cuh), restated in numpy and pinned on the CPU against
the host encoder's bytes (== the reference encoder's, tests/test_codecs_cpu.py): a block's bytes depend only on its own postings and the
docID before it, so (1) every block is SIZED independently, (2) an exclusive scan of the sizes places it, (3) every block is WRITTEN
independently — here in a shuffled order, to make the independence explicit — and the skiplist entry of a block follows from how many blocks
the encoder session has committed before it (google_codec.cpp:9-176; the countdown carries across terms, google_codec.h:57).  The CUDA
kernels implement exactly these formulas (tests/test_gpu_encoder.py compares their bytes with the reference encoder's on a GPU)."""
import numpy as np
import pytest

import trinity_b200 as tb


def vb_len(x: int) -> int:
    return 1 if x < (1 << 7) else 2 if x < (1 << 14) else 3 if x < (1 << 21) else 4 if x < (1 << 28) else 5


def vb_put(buf: bytearray, at: int, x: int) -> int:
    if x < (1 << 7):
        b = bytes([x])
    elif x < (1 << 14):
        b = bytes([0x80 | (x >> 8), x & 0xFF])
    elif x < (1 << 21):
        b = bytes([0xC0 | (x >> 16), x & 0xFF, (x >> 8) & 0xFF])
    elif x < (1 << 28):
        b = bytes([0xE0 | (x >> 24), (x >> 16) & 0xFF, (x >> 8) & 0xFF, x & 0xFF])
    else:
        b = bytes([0xF0, x & 0xFF, (x >> 8) & 0xFF, (x >> 16) & 0xFF, (x >> 24) & 0xFF])
    buf[at:at + len(b)] = b
    return at + len(b)


def encode_parallel(lists, block_docs, step, countdown, rng):
    """the kernels' plan: sizes -> scan -> chunk offsets -> independent block writes"""
    phase0 = (step - countdown) % step
    # flat block numbering
    blocks = []  # (term, j, first posting, n)
    blk_begin = [0]
    for t, (d, f, p) in enumerate(lists):
        nb = (len(d) + block_docs - 1) // block_docs
        for j in range(nb):
            blocks.append((t, j, j * block_docs, min(block_docs, len(d) - j * block_docs)))
        blk_begin.append(len(blocks))
    hit_begin = [np.concatenate([[0], np.cumsum(f)]).astype(np.int64) for d, f, p in lists]

    def parts(t, j, i0, n):
        d, f, p = lists[t]
        prev_last = int(d[i0 - 1]) if j else 0
        last = int(d[i0 + n - 1])
        deltas = [int(d[i]) - (int(d[i - 1]) if i else 0) for i in range(i0, i0 + n)]
        hits = []
        for i in range(i0, i0 + n):
            pp = 0
            for k in range(int(f[i])):
                pos = int(p[hit_begin[t][i] + k]) if p is not None else k + 1
                hits.append((pos - pp) << 1)
                pp = pos
        body = sum(vb_len(x) for x in deltas[:-1]) + sum(vb_len(int(x)) for x in f[i0:i0 + n]) + sum(vb_len(x) for x in hits)
        return prev_last, last, deltas, hits, body

    # (1) sizes
    size = []
    for t, j, i0, n in blocks:
        prev_last, last, _, _, body = parts(t, j, i0, n)
        size.append(vb_len(last - prev_last) + vb_len(body) + 1 + body)
    # (2) scan + per-term chunk sizes
    boff = np.concatenate([[0], np.cumsum(size)]).astype(np.int64)
    chunk = []
    for t in range(len(lists)):
        b0, nb = blk_begin[t], blk_begin[t + 1] - blk_begin[t]
        phase = (phase0 + b0) % step
        entries = min(65535, (phase + nb) // step)
        chunk.append(2 + int(boff[b0 + nb] - boff[b0]) + 8 * entries)
    toff = np.concatenate([[0], np.cumsum(chunk)]).astype(np.int64)
    out = bytearray(int(toff[-1]))  # zero-filled: a term without documents is its zero u16
    # (3) independent writes, any order
    order = rng.permutation(len(blocks))
    for g in order:
        t, j, i0, n = blocks[g]
        d, f, p = lists[t]
        prev_last, last, deltas, hits, body = parts(t, j, i0, n)
        b0, nb = blk_begin[t], blk_begin[t + 1] - blk_begin[t]
        blk_off = 2 + int(boff[g] - boff[b0])
        at = int(toff[t]) + blk_off
        at = vb_put(out, at, last - prev_last)
        at = vb_put(out, at, body)
        out[at] = n
        at += 1
        for x in deltas[:-1]:
            at = vb_put(out, at, x)
        for x in f[i0:i0 + n]:
            at = vb_put(out, at, int(x))
        for x in hits:
            at = vb_put(out, at, x)
        phase = (phase0 + b0) % step
        c = phase + j + 1
        if c % step == 0 and c // step - 1 < 65535:
            s = int(toff[t]) + 2 + int(boff[b0 + nb] - boff[b0]) + 8 * (c // step - 1)
            out[s:s + 4] = int(prev_last).to_bytes(4, "little")
            out[s + 4:s + 8] = int(blk_off).to_bytes(4, "little")
        if j == 0:
            entries = min(65535, (phase + nb) // step)
            out[int(toff[t]):int(toff[t]) + 2] = entries.to_bytes(2, "little")
    tuples = [(len(lists[t][0]), int(toff[t]), chunk[t]) for t in range(len(lists))]
    return bytes(out), tuples, step - (phase0 + len(blocks)) % step


def _lists(rng):
    def one(n, max_gap, fmax, with_pos=True):
        d = np.cumsum(rng.integers(1, max_gap, n, dtype=np.uint64)).astype(np.uint32) if n else np.zeros(0, np.uint32)
        f = rng.integers(0, fmax + 1, n).astype(np.uint32)
        p = np.concatenate([np.sort(rng.integers(1, 1 << 14, int(x))) for x in f] + [np.zeros(0, np.int64)]).astype(np.uint32)
        return d, f, p
    return [one(1, 9, 2), one(32, 40, 3), one(0, 9, 1), one(33, 40, 1), one(32 * 19 + 7, 500, 4), one(300, 2_000_000, 2), one(0, 9, 1), one(900, 3, 2)]


@pytest.mark.parametrize("block_docs,step,countdown", [(32, 8, 8), (32, 8, 3), (8, 1, 1), (16, 64, 40), (128, 8, 8), (100, 3, 2), (1, 8, 5)])
def test_parallel_plan_reproduces_the_serial_encoder(block_docs, step, countdown):
    rng = np.random.default_rng(block_docs * 7 + step)
    lists = _lists(rng)
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    b.set_google_block(block_docs, step)
    filler = step - countdown  # the host builder starts with a full countdown: put `filler` blocks in front
    if filler:
        b.add_term(np.arange(1, filler * block_docs + 1, dtype=np.uint32), np.ones(filler * block_docs, np.uint32), None)
    skip, nskip = b.index().size, len(b.terms)
    for d, f, p in lists:
        b.add_term(d, f, p)
    want = bytes(b.index()[skip:])
    got, tuples, cd = encode_parallel(lists, block_docs, step, countdown, rng)
    assert len(got) == len(want)
    assert got == want, f"first differing byte at {next(i for i, (x, y) in enumerate(zip(got, want)) if x != y)}"
    assert tuples == [(t[0], t[1] - skip, t[2]) for t in b.terms[nskip:]]
    nblocks = sum((len(d) + block_docs - 1) // block_docs for d, f, p in lists)
    assert cd == step - ((step - countdown) % step + nblocks) % step


def test_positions_omitted_mean_one_to_freq():
    rng = np.random.default_rng(2)
    lists = [(d, f, None) for d, f, p in _lists(rng)]
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    for d, f, _ in lists:
        b.add_term(d, f, None)
    got, tuples, _ = encode_parallel(lists, 32, 8, 8, rng)
    assert got == bytes(b.index()) and tuples == b.terms
