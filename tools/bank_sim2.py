"""Bank-conflict simulator for the FFT exchanges with the "local last stage" thread assignment (fft_core.cuh), and a small
search over XOR swizzles.  Mirrors: Stage::base / plan_last_butterfly, smem_put/get (128-bit accesses where a stage with
sigma == 1 meets a contiguous address map), RowAddr (one row = N/E consecutive threads) and ColAddr (thread = c + TC*t).

    python tools/bank_sim2.py row 2048            # prints wavefronts/ideal per stage for the swizzle in use
    python tools/bank_sim2.py col 2048 4 tma128   # column tile with the TMA SWIZZLE_128B address map
"""
import itertools
import sys


def plan_pick(rem, e):
    return (e // 2 if (rem == 2 * e and e >= 8) else e) if rem >= e else rem


def plan(n, e):
    radices, rem = [], n
    while rem > 1:
        r = plan_pick(rem, e)
        radices.append(r)
        rem //= r
    sig, rem = [], n
    for r in radices:
        rem //= r
        sig.append(rem)
    return radices, sig


def reg_pos(n, e, radices, sig, s, t, idx, local_last=True):
    L = len(radices)
    r, sg = radices[s], sig[s]
    g, j = divmod(idx, r)
    tpf = n // e
    if local_last and s == L - 1 and L >= 2:
        rl, rp = radices[-1], radices[-2]
        gp_count = e // rp                    # groups per thread at stage L-2
        h_count = (e // rl) // gp_count       # last-stage butterflies per thread per block
        gp, h = divmod(g, h_count)
        bt = t + gp * tpf
        b = rp * (bt // rl) + (bt % rl) + rl * h
        return b * rl + j
    b = t + g * tpf
    return (b // sg) * (sg * r) + (b % sg) + j * sg


def wavefronts(addrs_bytes, width):
    group = 128 // width
    total = 0
    for g0 in range(0, 32, group):
        banks = {}
        for a in addrs_bytes[g0:g0 + group]:
            for w in range(width // 4):
                banks.setdefault(((a // 4) + w) % 32, set()).add(a // 4 + w)
        total += max(len(v) for v in banks.values())
    return total, 32 // group


def analyse(n, e, mode, tc, elem, addr_fn, local_last=True, quiet=False, contiguous128=True):
    """addr_fn(p, c) -> byte address of position p of column c (rows: c = 0)."""
    radices, sig = plan(n, e)
    tpf = n // e
    L = len(radices)
    res = []
    for s in range(L):
        tot = ideal = 0
        wide = mode == "row" and sig[s] == 1 and elem == 8 and radices[s] % 2 == 0 and contiguous128
        for idx in range(e):
            if wide and idx % 2:
                continue
            for warp in range(0, max(1, (tpf * (tc if mode == "col" else 1)) // 32)):
                addrs = []
                for lane in range(32):
                    th = warp * 32 + lane
                    if mode == "row":
                        t = th % tpf
                        rowbase = (th // tpf) * n * elem
                        addrs.append(rowbase + addr_fn(reg_pos(n, e, radices, sig, s, t, idx, local_last), 0))
                    else:
                        c, t = th % tc, (th // tc) % tpf
                        addrs.append(addr_fn(reg_pos(n, e, radices, sig, s, t, idx, local_last), c))
                w, i = wavefronts(addrs, 16 if wide else min(elem, 16))
                tot += w
                ideal += i
        res.append(tot / ideal)
        if not quiet:
            print(f"  stage {s} (R={radices[s]}, sigma={sig[s]}): wavefronts/ideal = {tot / ideal:.2f}")
    return res


def xor_swizzle(terms):
    """terms: list of (src_shift, mask, dst_shift): p ^= ((p >> src) & mask) << dst."""
    def f(p):
        q = p
        for src, mask, dst in terms:
            q ^= ((p >> src) & mask) << dst
        return q
    return f


def row_addr(swz, elem):
    return lambda p, c: swz(p) * elem


def col_addr(swz, tc, elem):
    return lambda p, c: (swz(p) * tc + c) * elem


def col_addr_tma128(tc, elem):
    """TMA SWIZZLE_128B on the natural tile [row][tc columns]: byte address bits [4:6] ^= bits [7:9]."""
    def f(p, c):
        x = (p * tc + c) * elem
        return x ^ (((x >> 7) & 7) << 4)
    return f


def search_row(n, e=16, elem=8, local_last=True):
    """smallest set of XOR terms (low bits <- higher bits) that makes every stage conflict-free."""
    nb = n.bit_length() - 1
    cands = []
    for dst in range(0, 4):
        for width in (1, 2, 3):
            if dst + width > 4:
                continue
            for src in range(4, nb - width + 1):
                cands.append((src, (1 << width) - 1, dst))
    best = None
    for k in (1, 2, 3):
        for combo in itertools.combinations(cands, k):
            r = analyse(n, e, "row", 1, elem, row_addr(xor_swizzle(combo), elem), local_last, quiet=True)
            if max(r) <= 1.0:
                return combo
            if best is None or max(r) < best[0]:
                best = (max(r), combo)
    return best


def search_col(n, tc, e=16, elem=8, local_last=True):
    nb = n.bit_length() - 1
    w = max(0, 4 - (tc.bit_length() - 1)) if elem == 8 else max(0, 3 - (tc.bit_length() - 1))
    if w == 0:
        return ()
    cands = []
    for dst in range(0, w):
        for width in range(1, w - dst + 1):
            for src in range(max(1, w), nb - width + 1):
                cands.append((src, (1 << width) - 1, dst))
    best = None
    for k in (1, 2, 3):
        for combo in itertools.combinations(cands, k):
            r = analyse(n, e, "col", tc, elem, col_addr(xor_swizzle(combo), tc, elem), local_last, quiet=True)
            if max(r) <= 1.0:
                return combo
            if best is None or max(r) < best[0]:
                best = (max(r), combo)
    return best


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "row"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    if mode == "row":
        combo = search_row(n)
        print("row", n, "swizzle terms (src, mask, dst):", combo)
        if combo and isinstance(combo[0], tuple):
            analyse(n, 16, "row", 1, 8, row_addr(xor_swizzle(combo), 8))
    else:
        tc = int(sys.argv[3]) if len(sys.argv) > 3 else 4
        if len(sys.argv) > 4 and sys.argv[4] == "tma128":
            print("col", n, "tc", tc, "TMA SWIZZLE_128B address map")
            analyse(n, 16, "col", tc, 8, col_addr_tma128(tc, 8))
        else:
            combo = search_col(n, tc)
            print("col", n, "tc", tc, "swizzle terms:", combo)
            if combo == () or (combo and isinstance(combo[0], tuple)):
                analyse(n, 16, "col", tc, 8, col_addr(xor_swizzle(combo), tc, 8))
