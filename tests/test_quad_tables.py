"""The dealing of a symbol's 48 data bins over the seven decision slots of k_stream_quad / k_mc_quad (csrc/ofdm_stream.cuh,
make_quad_lane): restated here from the frame-build block of the reference (OFDM.c:523-548: data runs and pilots on the centred
grid; :494 the L sequence) and checked without a GPU -- every data bin is owned by exactly one (lane, slot), every slot's
payload word is the one the kernel selects (w1, word_a, w2, word_b, w0, word_c, w1), the shift fields put the bin's bit
pair at bits 31 / 30, and the flip masks carry exactly the bins with L < 0."""
import re
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# centred index c -> data index: runs of OFDM.c:528-547; pilots at c = 11, 25, 39, 53
RUNS = [(6, 10), (12, 24), (26, 31), (33, 38), (40, 52), (54, 58)]
LK = [1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 0,
      1, -1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, 1, 1]          # OFDM.c:494, c = 6..58


def tables():
    by_c = {}
    d = 0
    for lo, hi in RUNS:
        for c in range(lo, hi + 1):
            by_c[c] = d
            d += 1
    assert d == 48
    bin_data, bin_lts = {}, {}
    for p in range(64):                      # natural FFT bin p <-> centred index (p + 32) % 64
        c = (p + 32) & 63
        bin_data[p] = by_c.get(c, -1)
        bin_lts[p] = LK[c - 6] if 6 <= c <= 58 else 0
    return bin_data, bin_lts


def slot_bin(u, t):
    j = t if t < 3 else ((3 if u < 3 else 4) if t == 3 else t + 1)
    return u + 8 * j


def test_every_data_bin_has_one_slot_and_the_kernels_word_selection():
    bin_data, bin_lts = tables()
    owned = {}
    for u in range(8):
        flips = [0, 0, 0]
        for t in range(7):
            p = slot_bin(u, t)
            d = bin_data[p]
            if d < 0:
                continue
            assert bin_lts[p] != 0
            assert p not in owned
            owned[p] = (u, t)
            # the word the kernel reads for this slot (ofdm_stream.cuh: w1, word_a, w2, word_b, w0, word_c, w1)
            word = {0: 1, 1: 1 if u < 2 else 2, 2: 2, 3: 2 if u < 3 else 0, 4: 0, 5: 1 if u == 7 else 0, 6: 1}[t]
            assert word == d >> 4, (u, t, p, d)
            # shift field: w << (30 - 2 (d & 15)) puts bit b = 2d + 1 at bit 31 and bit a = 2d at bit 30
            sh = 30 - 2 * (d & 15)
            assert 0 <= sh <= 30 and (2 * (d & 15) + 1) + sh == 31
            if bin_lts[p] < 0:
                flips[d >> 4] |= 1 << (2 * (d & 15))
        # bit a of exactly the lane's data bins with L < 0
        for w in range(3):
            for bit in range(32):
                if (flips[w] >> bit) & 1:
                    assert bit % 2 == 0
                    dd = 16 * w + bit // 2
                    p = [q for q in range(64) if bin_data[q] == dd][0]
                    assert p % 8 == u and bin_lts[p] < 0
    assert sorted(owned) == sorted(p for p in range(64) if bin_data[p] >= 0) and len(owned) == 48
    # lane use of the decision stage: 48 of 7 x 8 slots
    assert len(owned) / 56 > 0.85


def test_the_header_states_the_same_slot_rule():
    """the comment and code of make_quad_lane are the restatement's source: keep them in step"""
    src = open(os.path.join(ROOT, "ieee-802.11-ofdm-qpsk-simulator_b200", "csrc", "ofdm_stream.cuh")).read()
    assert "const int j = t < 3 ? t : (t == 3 ? (u < 3 ? 3 : 4) : t + 1);" in src
    assert "const uint32_t sh = 30u - 2u * (uint32_t)(d & 15);" in src
    assert re.search(r"word_a = u < 2 \? w1 : w2", src) and re.search(r"word_b = u < 3 \? w2 : w0", src) and re.search(r"word_c = u == 7 \? w1 : w0", src)
