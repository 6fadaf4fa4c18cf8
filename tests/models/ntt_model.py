"""Index-math model of the multi-pass NTT schedule implemented in plonk_b200/csrc/ntt.cu.

Pure Python over the oracle's field; used during development to validate the pass decomposition
(tile addressing, inter-pass twiddles, digit-reversed final store) before it is transcribed to CUDA."""
import random
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyref as R

MOD = R.R_MOD


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def root(logm, inverse):
    w = R.ROOT_OF_UNITY
    for _ in range(logm, 32):
        w = w * w % MOD
    return pow(w, -1, MOD) if inverse else w


def table(logm, inverse):
    w = root(logm, inverse)
    out, p = [], 1
    for _ in range(max(1, 1 << (logm - 1))):
        out.append(p)
        p = p * w % MOD
    return out


def sub_dft_dif(S, r, T, W):
    """In-smem DIF radix-2, natural in -> bit-reversed out, layout S[j*T + t]."""
    Rr = 1 << r
    for s in range(r):
        half = Rr >> (s + 1)
        for bidx in range((Rr // 2) * T):
            t = bidx % T
            jj = bidx // T
            grp, pos = divmod(jj, half)
            j0 = grp * 2 * half + pos
            j1 = j0 + half
            w = W[pos << s]
            a, b = S[j0 * T + t], S[j1 * T + t]
            S[j0 * T + t] = (a + b) % MOD
            S[j1 * T + t] = (a - b) * w % MOD


def plan(L, max_r=4):
    """Split L into pass radices (logs)."""
    if L <= max_r:
        return [L]
    p = -(-L // max_r)
    base, extra = divmod(L, p)
    return [base + (1 if i < extra else 0) for i in range(p)]


def ntt_multipass(x, L, inverse, radices, T=2):
    N = 1 << L
    data = list(x) + [0] * (N - len(x))
    p = len(radices)
    out = [None] * N
    for q, r in enumerate(radices):
        Rr = 1 << r
        logLo = sum(radices[q + 1:])
        Lo = 1 << logLo
        H = 1 << sum(radices[:q])
        W = table(r, inverse)
        last = q == p - 1
        if not last:
            m = r + logLo
            Wm = table(m, inverse)
            M = 1 << m
            Tt = min(T, Lo)
            for h in range(H):
                for lo0 in range(0, Lo, Tt):
                    S = [0] * (Rr * Tt)
                    for j in range(Rr):
                        for t in range(Tt):
                            S[j * Tt + t] = data[h * Rr * Lo + j * Lo + lo0 + t]
                    sub_dft_dif(S, r, Tt, W)
                    for k in range(Rr):
                        for t in range(Tt):
                            v = S[bitrev(k, r) * Tt + t]
                            e = k * (lo0 + t)
                            assert e < M
                            tw = Wm[e] if e < M // 2 else (-Wm[e - M // 2]) % MOD
                            data[h * Rr * Lo + k * Lo + lo0 + t] = v * tw % MOD
        else:
            # rows h = k0*(H/R0) + mid ; tile over consecutive k0
            if p == 1:
                R0, Hmid = 1, 1
            else:
                R0 = 1 << radices[0]
                Hmid = H // R0
            Tt = min(T, R0)
            for mid in range(Hmid):
                # digit-reverse mid over radices[1:p-1]
                revmid, tmp, mul = 0, mid, 1
                digs = []
                for rr in reversed(radices[1:p - 1]):
                    digs.append(tmp % (1 << rr)); tmp //= (1 << rr)
                digs.reverse()  # digs[i] = k_{i+1}
                for rr, dgt in zip(radices[1:p - 1], digs):
                    revmid += dgt * mul; mul <<= rr
                for k0b in range(0, R0, Tt):
                    S = [0] * (Rr * Tt)
                    for t in range(Tt):
                        h = (k0b + t) * Hmid + mid
                        for j in range(Rr):
                            S[j * Tt + t] = data[h * Rr + j]
                    sub_dft_dif(S, r, Tt, W)
                    for k in range(Rr):
                        for t in range(Tt):
                            v = S[bitrev(k, r) * Tt + t]
                            out[(k0b + t) + R0 * revmid + H * k] = v
    return out


if __name__ == "__main__":
    rng = random.Random(5)
    for L, radices, T in [(3, [3], 1), (4, [2, 2], 2), (5, [3, 2], 2), (6, [2, 2, 2], 2), (7, [3, 2, 2], 4), (8, [3, 3, 2], 2), (9, [3, 3, 3], 4), (6, [3, 3], 8)]:
        for inverse in (False, True):
            x = [rng.randrange(MOD) for _ in range((1 << L) - 3)]
            got = ntt_multipass(x, L, inverse, radices, T)
            a = list(x) + [0] * ((1 << L) - len(x))
            R.serial_fft(a, root(L, inverse), L)
            assert got == a, (L, radices, inverse)
            print("ok", L, radices, T, inverse)
