#!/usr/bin/env python
"""Search for the dealing of a frame's 96 data bins to (round, half-warp, lane) that minimises shared-memory bank conflicts of
the item loads of the fused receivers (csrc/ofdm_chain.cuh: c_deal_bin).  Model: 64-bit loads, one half-warp per wavefront,
bank pair = element index mod 16; LTS tiles at element offsets 288 / 360, symbol tiles at 0 / 72 (72 = 8 mod 16)."""
import random
from collections import defaultdict

C = list(range(6, 11)) + list(range(12, 25)) + list(range(26, 32)) + list(range(33, 39)) + list(range(40, 53)) + list(range(54, 59))
DATA_BIN = [(x + 32) % 64 for x in C]          # data index -> natural FFT bin


def wavefronts(elems):
    m = defaultdict(set)
    for e in elems:
        m[e % 16].add(e)
    return max(len(v) for v in m.values())


def group_cost(items):                          # 16 x (bin, symbol)
    a = [288 + b for b, s in items]
    b_ = [360 + b for b, s in items]
    f = [s * 72 + b for b, s in items]
    return wavefronts(a) - 1 + wavefronts(b_) - 1 + wavefronts(f) - 1


def total(groups):
    return sum(group_cost([(b, 0) for b in g] + [(b, 1) for b in g]) for g in groups)


def main():
    in_order = [[(DATA_BIN[(hw * 16 + l + 32 * r) % 48], (hw * 16 + l + 32 * r) // 48) for l in range(16)] for r in range(3) for hw in range(2)]
    print("data indices in order:", sum(group_cost(g) for g in in_order), "extra wavefronts per frame")
    random.seed(1)
    bins = list(DATA_BIN)
    random.shuffle(bins)
    groups = [bins[8 * k:8 * k + 8] for k in range(6)]
    cur = total(groups)
    for _ in range(400000):
        a, b = random.sample(range(6), 2)
        i, j = random.randrange(8), random.randrange(8)
        groups[a][i], groups[b][j] = groups[b][j], groups[a][i]
        t = total(groups)
        if t <= cur:
            cur = t
        else:
            groups[a][i], groups[b][j] = groups[b][j], groups[a][i]
        if cur <= 1:
            break
    print("8 bins x both symbols per half-warp:", cur, "extra wavefronts per frame")
    for g in groups:
        print("   ", ", ".join("%2d" % b for b in sorted(g)))


if __name__ == "__main__":
    main()
