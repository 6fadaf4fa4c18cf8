"""Index algebra of the batched-affine bucket accumulation (csrc/msm.cu, k_msm_affine_fwd / _back; opt-in with
PB200_MSM_AFFINE=1), modelled with integers in place of curve points.

The entries of every bucket are added as a pairwise tree: round r adds the surviving elements (2i, 2i + 1) of each
bucket.  Layout 0 is the sorted reference list (bucket b at [off[b], off[b] + L)); the output of round r is layout
r + 1, where bucket b keeps ceil(L_r / 2) elements at off_{r+1}[b] = (off_r[b] + b) >> 1.  A thread owns K consecutive
input POSITIONS of the round's layout whatever buckets they belong to.  The model replays that schedule and checks
what the kernels rely on: every read and write stays inside its buffer (cap/2 + nb + 2 and cap/4 + nb + 2 points,
K/2 pair slots per thread), no output slot is written twice, a thread never gets more than K/2 pairs, and every
bucket's last addition yields the sum of its entries."""
import random

K, THREADS = 64, 128


def aff_off(o, b, r):
    for _ in range(r):
        o = (o + b) >> 1
    return o


def aff_len(length, r):
    return (length + (1 << r) - 1) >> r


def positions(cap, nb, r):
    return cap if r == 0 else (cap >> r) + nb + 1


def run(counts, seed=0, slack=0):
    rng = random.Random(seed)
    nb = len(counts)
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    total = off[-1]
    cap = max(total, 1) + slack  # the kernels size everything from cap = n * W >= total
    vals = [rng.randrange(1, 1 << 30) for _ in range(total)]
    cap_a, cap_b = cap // 2 + nb + 2, cap // 4 + nb + 2
    buf_a, buf_b = [None] * cap_a, [None] * cap_b
    sums = [0] * nb
    max_len = max(counts) if counts else 0
    rounds = 0
    while (1 << rounds) < cap:
        rounds += 1
    slots = 0
    for r in range(rounds + 1):
        ctas = -(-(-(-positions(cap, nb, r) // K)) // THREADS)
        slots = max(slots, ctas * THREADS * (K // 2))
    for r in range(rounds):
        if r > 0 and (1 << r) >= max_len:
            continue
        src = vals if r == 0 else (buf_a if r & 1 else buf_b)
        dst = buf_b if r & 1 else buf_a
        in_cap = cap if r == 0 else (cap_a if r & 1 else cap_b)
        out_cap = cap_b if r & 1 else cap_a
        ctas = -(-(-(-positions(cap, nb, r) // K)) // THREADS)
        threads = ctas * THREADS
        end = aff_off(off[nb], nb, r)
        assert end <= in_cap
        written = set()
        for t in range(threads):
            pos0, pos1 = t * K, min(end, t * K + K)
            if pos0 >= end:
                continue
            lo, hi = 0, nb
            while hi - lo > 1:
                mid = (lo + hi) >> 1
                if aff_off(off[mid], mid, r) <= pos0:
                    lo = mid
                else:
                    hi = mid
            n_pairs = 0
            bk = lo
            while bk < nb:
                o = aff_off(off[bk], bk, r)
                if o >= pos1:
                    break
                length = aff_len(off[bk + 1] - off[bk], r)
                if length == 1 and r == 0 and o >= pos0:
                    sums[bk] = src[o]
                if length >= 2:
                    o_next = aff_off(off[bk], bk, r + 1)
                    e = 0 if o >= pos0 else ((pos0 - o + 1) & ~1)
                    while e + 1 < length and o + e < pos1:
                        assert o + e + 1 < in_cap and src[o + e] is not None and src[o + e + 1] is not None
                        assert t + n_pairs * threads < slots
                        if length == 2:
                            sums[bk] = src[o + e] + src[o + e + 1]
                        else:
                            d = o_next + (e >> 1)
                            assert d < out_cap and d not in written
                            written.add(d)
                            dst[d] = src[o + e] + src[o + e + 1]
                        n_pairs += 1
                        e += 2
                    if (length & 1) and pos0 <= o + length - 1 < pos1:
                        d = o_next + ((length - 1) >> 1)
                        assert d < out_cap and d not in written
                        written.add(d)
                        dst[d] = src[o + length - 1]
                bk += 1
            assert n_pairs <= K // 2
    for bk in range(nb):
        assert sums[bk] == sum(vals[off[bk] : off[bk + 1]]), (bk, counts[bk])


def check(trials=120, seed=5):
    rng = random.Random(seed)
    for trial in range(trials):
        nb = rng.choice([1, 2, 8, 64, 512])
        kind = rng.randrange(5)
        if kind == 0:
            counts = [rng.randrange(0, 4) for _ in range(nb)]  # sparse: empty, single and tiny buckets
        elif kind == 1:
            counts = [rng.randrange(20, 50) for _ in range(nb)]  # dense, uniform scalars
        elif kind == 2:
            counts = [0] * nb
            counts[rng.randrange(nb)] = rng.randrange(1, 5000)  # everything in one bucket (equal scalars)
        elif kind == 3:
            counts = [rng.choice([0, 0, 0, 1, 2, 3, 200]) for _ in range(nb)]  # wire values: mostly nothing, a long tail
        else:
            counts = [rng.randrange(0, 130) for _ in range(nb)]
        run(counts, trial, slack=rng.randrange(0, 50))
