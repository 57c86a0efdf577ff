"""CPU emulation of the IMMA (mma.sync.m16n8k32.u8.u8.s32) fragment ownership used by pm_tiles_imma, to check an
operand layout BEFORE it goes to the GPU: (a) the accumulated tile equals the direct correlation, (b) the number of
shared-memory wavefronts of the A loads (bank conflicts) for a given window pitch.

Fragment ownership (PTX ISA, m16n8k32 with 8-bit operands), lane = 4*g + tig:
  A (16 x 32): a0 = row g,   k 4*tig..4*tig+3 | a1 = row g+8, same k | a2 = row g, k 16+4*tig.. | a3 = row g+8, k 16+4*tig..
  B (32 x 8):  b0 = k 4*tig..4*tig+3, n = g   | b1 = k 16+4*tig.., n = g
  C (16 x 8):  c0,c1 = row g, cols 2*tig, 2*tig+1 | c2,c3 = row g+8, same cols

Two mappings of window columns onto the hardware K slots:
  "perm": slot k of a0 <- window byte 8*tig + (k - 4*tig), a2 <- 8*tig + 4 + ...   (lane words 2*tig, 2*tig+1; shipped)
  "std" : slot k <- window byte k                                                     (lane words tig, tig+4)
"""
import numpy as np


def window_col_of_slot(layout, tig, reg, e):
    """window column (relative to the chunk) held in byte e of A register a0 (reg=0) / a2 (reg=1) of lane tig"""
    if layout == "perm":
        return 8 * tig + 4 * reg + e
    return 4 * tig + 16 * reg + e


def emulate(layout, s=35, R=41, seed=0):
    """C[row][n] = sum_i sum_c sum_slots A[row][slot] * B[slot][n] with the kernel's A / B byte selection."""
    rng = np.random.default_rng(seed)
    W = R + s - 1
    win = rng.integers(0, 256, (W + 16, W + 96), dtype=np.int64)
    tpl = rng.integers(0, 256, (s, s), dtype=np.int64)
    nc = (s + 7 + 31) // 32
    pad = np.zeros((s, 32 * nc + 96), np.int64)
    pad[:, 8:8 + s] = tpl
    direct = np.zeros((R, R), np.int64)
    for y in range(R):
        for x in range(R):
            direct[y, x] = (win[y:y + s, x:x + s] * tpl).sum()
    got = np.zeros((R, R), np.int64)
    slots = [(tig, reg, e) for tig in range(4) for reg in range(2) for e in range(4)]
    for y0 in range(0, R, 16):
        for x0 in range(0, R, 24):
            for b in range(3):
                C = np.zeros((16, 8), np.int64)
                for i in range(s):
                    for c in range(nc):
                        for tig, reg, e in slots:
                            col = window_col_of_slot(layout, tig, reg, e)
                            a_col = win[y0 + i:y0 + i + 16, x0 + 8 * b + 32 * c + col]          # A[:, slot]
                            b_row = np.array([pad[i, 32 * c + 8 + col - n] for n in range(8)])   # B[slot, :]
                            C += np.outer(a_col, b_row)
                ys = slice(y0, min(R, y0 + 16)); xs = slice(x0 + 8 * b, min(R, x0 + 8 * b + 8))
                if xs.start < R:
                    got[ys, xs] = C[:ys.stop - ys.start, :xs.stop - xs.start]
    return np.array_equal(got, direct)


def a_load_wavefronts(layout, wpw):
    """shared-memory wavefronts of one scalar LDS.32 of an A register across a warp (32 banks of 4 bytes)"""
    worst = 0
    for reg in range(2):
        banks = {}
        for lane in range(32):
            g, tig = lane >> 2, lane & 3
            word = g * wpw + (2 * tig + reg if layout == "perm" else tig + 4 * reg)
            banks.setdefault(word % 32, set()).add(word)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


def b_words(layout, tig, g):
    """aligned template words (relative to the chunk) a lane reads for its B fragment, and the byte shift"""
    if layout == "perm":
        ob = 8 + 8 * tig - g
        return [ob >> 2, (ob >> 2) + 1, (ob >> 2) + 2], ob & 3
    ob = 8 + 4 * tig - g
    return [ob >> 2, (ob >> 2) + 1, (ob >> 2) + 4, (ob >> 2) + 5], ob & 3


if __name__ == "__main__":
    for layout in ("perm", "std"):
        ok = all(emulate(layout, s, R, seed) for s, R, seed in ((35, 41, 0), (8, 12, 1), (51, 20, 2), (17, 30, 3)))
        print("%-4s layout: accumulated tiles == direct correlation: %s" % (layout, ok))
        for wpw in (24, 28, 20, 40, 44):
            print("     pitch %2d words: %d wavefront(s) per A load" % (wpw, a_load_wavefronts(layout, wpw)))
        nb = max(len(set(b_words(layout, t, g)[0])) for t in range(4) for g in range(8))
        print("     B fragment: %d aligned words + 2 funnel shifts per lane" % nb)
