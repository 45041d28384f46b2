"""Pick the row-layout padding of each FFT plan by brute force.

Model: shared memory has 32 4-byte banks.  A warp access of `esz`-byte elements is served in
passes of 128 bytes: 8 lanes per pass for 16-byte elements (fp64 complex), 16 lanes for 8-byte
elements (fp32 complex).  Within a pass the lanes must hit distinct (address/esz) mod (128/esz)
slots or the pass is replayed.  Score = total wavefronts / ideal wavefronts over the three
access patterns of microtipi_b200/csrc/wfm_fft.cuh (stage-1 write, stage-2 read/write,
stage-3 read).
"""
import itertools, sys

PLANS = {32: (8, 8, 4, 1), 64: (8, 8, 8, 1), 128: (8, 8, 4, 4), 256: (8, 8, 8, 4),
         512: (8, 8, 8, 8), 1024: (16, 8, 8, 16), 2048: (16, 16, 16, 8)}


def pad(i, pa, pb):
    return i + ((i >> pa) if pa else 0) + ((i >> pb) if pb else 0)


def wavefronts(addrs, esz):
    lanes_per_pass = 128 // esz
    tot = 0
    for p in range(0, len(addrs), lanes_per_pass):
        grp = addrs[p:p + lanes_per_pass]
        slots = {}
        for a in grp:
            slots.setdefault(a % lanes_per_pass, set()).add(a)
        tot += max(len(v) for v in slots.values())
    return tot


def score(N, E, R1, R2, R3, pa, pb, esz, threads=256):
    T = N // E
    S1 = N // R1
    rowlen = pad(N - 1, pa, pb) + 1
    RB = max(1, threads // T)
    nthreads = RB * T
    tot = ideal = 0

    def run(posfn, nb, legs):
        nonlocal tot, ideal
        for u in range(nb):
            for r in range(legs):
                for w0 in range(0, nthreads, 32):
                    addrs = []
                    for tid in range(w0, min(w0 + 32, nthreads)):
                        slot, t = divmod(tid, T)
                        addrs.append(slot * rowlen + pad(posfn(t + T * u, r), pa, pb))
                    tot += wavefronts(addrs, esz)
                    ideal += (len(addrs) * esz + 127) // 128
    run(lambda b, k1: k1 * S1 + b, E // R1, R1)
    if R3 > 1:
        run(lambda b, r: (b // R3) * S1 + r * R3 + (b % R3), E // R2, R2)
        run(lambda b, r: (b % R1) * S1 + (b // R1) * R3 + r, E // R3, R3)
    else:
        run(lambda b, r: b * S1 + r, E // R2, R2)
    return tot / ideal, rowlen


if __name__ == "__main__":
    for esz, name in ((16, "fp64"), (8, "fp32")):
        print(name)
        for N, (E, R1, R2, R3) in PLANS.items():
            best = None
            for pa, pb in itertools.product([0, 1, 2, 3, 4, 5], [0, 4, 5, 6, 7, 8, 9]):
                if pa and pb and pb <= pa:
                    continue
                s, rowlen = score(N, E, R1, R2, R3, pa, pb, esz)
                key = (round(s, 4), rowlen)
                if best is None or key < best[0]:
                    best = (key, pa, pb)
            s0, _ = score(N, E, R1, R2, R3, 0, 0, esz)
            print(f"  N={N:5d} plan={E,R1,R2,R3} unpadded={s0:.2f} best pa={best[1]} pb={best[2]} "
                  f"score={best[0][0]:.3f} rowlen={best[0][1]}")


def score_cols(N, E, R1, R2, R3, C, shift, esz):
    """Column layout: cell = (i + (i >> shift)) * C + c, thread id = t * C + c."""
    T = N // E
    S1 = N // R1
    nthreads = C * T
    tot = ideal = 0

    def padc(i):
        return i + ((i >> shift) if shift else 0)

    def run(posfn, nb, legs):
        nonlocal tot, ideal
        for u in range(nb):
            for r in range(legs):
                for w0 in range(0, nthreads, 32):
                    addrs = []
                    for tid in range(w0, min(w0 + 32, nthreads)):
                        t, c = divmod(tid, C)
                        addrs.append(padc(posfn(t + T * u, r)) * C + c)
                    tot += wavefronts(addrs, esz)
                    ideal += (len(addrs) * esz + 127) // 128
    run(lambda b, k1: k1 * S1 + b, E // R1, R1)
    if R3 > 1:
        run(lambda b, r: (b // R3) * S1 + r * R3 + (b % R3), E // R2, R2)
        run(lambda b, r: (b % R1) * S1 + (b // R1) * R3 + r, E // R3, R3)
    else:
        run(lambda b, r: b * S1 + r, E // R2, R2)
    return tot / ideal


def score_off(N, E, R1, R2, R3, off, esz, vec_last=2, threads=256):
    """Row layout with a per-block offset: cell(i) = i + off[i // S1]; the last stage reads `vec_last` adjacent cells per
    load (wfm_fft.cuh RowOff / VEC_LAST, the fp32 rows).  Returns (total wavefronts / ideal, per-pattern ratios)."""
    T = N // E
    S1 = N // R1

    def padf(i):
        return i + off[i // S1]

    rowlen = (padf(N - 1) + 2) & ~1
    nthreads = max(1, threads // T) * T
    res = {}

    def wave(addrs, width):
        lanes = 128 // width
        tot = 0
        for p in range(0, len(addrs), lanes):
            slots = {}
            for a in addrs[p:p + lanes]:
                slots.setdefault((a * esz // width) % lanes, set()).add(a * esz // width)
            tot += max(len(v) for v in slots.values())
        return tot

    def run(name, posfn, nb, legs, vec=1):
        tot = ideal = 0
        for u in range(nb):
            for r in range(0, legs, vec):
                for w0 in range(0, nthreads, 32):
                    addrs = []
                    for tid in range(w0, min(w0 + 32, nthreads)):
                        slot, t = divmod(tid, T)
                        addrs.append(slot * rowlen + padf(posfn(t + T * u, r)))
                    tot += wave(addrs, esz * vec)
                    ideal += (len(addrs) * esz * vec + 127) // 128
        res[name] = tot / ideal
        return tot, ideal

    a = run("st1", lambda b, k1: k1 * S1 + b, E // R1, R1)
    b2 = run("x2", lambda b, r: (b // R3) * S1 + r * R3 + (b % R3), E // R2, R2)
    c = run("ld3", lambda b, r: (b % R1) * S1 + (b // R1) * R3 + r, E // R3, R3, vec_last)
    return (a[0] + 2 * b2[0] + c[0]) / (a[1] + 2 * b2[1] + c[1]), res


# the fp32 row layouts of wfm_fft.cuh (RowOff<N, 8>): off(k1) = A (k1 & 1) + B ((k1 >> 1) & 1) + C (k1 >> 2)
F32_ROWOFF = {128: ((8, 8, 4, 4), (4, 8, 18)), 256: ((8, 8, 8, 4), (4, 8, 18)), 512: ((8, 8, 8, 8), (8, 18, 36)),
              1024: ((16, 8, 8, 16), (2, 4, 8))}
if __name__ == "__main__":
    print("fp32 rows with per-block offsets + 16-byte last-stage loads")
    for N, ((E, R1, R2, R3), (A, B, Cc)) in F32_ROWOFF.items():
        off = [A * (k & 1) + B * ((k >> 1) & 1) + Cc * (k >> 2) for k in range(R1)]
        s, res = score_off(N, E, R1, R2, R3, off, 8)
        s0, _ = score(N, E, R1, R2, R3, 0, 6 if N >= 512 else 4, 8)
        print(f"  N={N:5d} off={off} score={s:.3f} {res}")
