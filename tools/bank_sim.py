"""Shared-memory bank-conflict simulator for the in-place FFT exchanges (pyatmosphere_b200/csrc/fft_core.cuh).

For every exchange (write at stage-s positions, read at stage-(s+1) positions, and the reverse for the inverse
transform) and every register index it counts the wavefronts a warp needs: a 64-bit access is served per half
warp, a 128-bit access per quarter warp; within such a group the cost is the largest number of DISTINCT
addresses falling into one 4-byte bank column (32 banks).  Ideal = 1.0 per group.

    python tools/bank_sim.py 2048 16 row     # N, E, row|col [TC] [elem bytes]
"""
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


def reg_pos(n, e, radices, sig, s, t, idx):
    r, sg = radices[s], sig[s]
    g, j = divmod(idx, r)
    b = t + g * (n // e)
    return (b // sg) * (sg * r) + (b % sg) + j * sg


def ident(p):
    return p


def swz_row_for(n):
    """Mirror of swz_row<N, 16> in fft_passes.cuh (complex64 rows)."""
    if n == 2048:
        return lambda p: p ^ (((p >> 4) & 3) << 1) ^ (((p >> 7) & 1) << 3)
    if n == 1024:
        return lambda p: p ^ (((p >> 6) & 3) << 2) ^ (((p >> 4) & 1) << 1)
    if n in (4096, 256):
        return lambda p: p ^ (((p >> 4) & 7) << 1)
    if n in (8192, 512):
        return lambda p: p ^ (((p >> 5) & 3) << 2) ^ (((p >> 4) & 1) << 1)
    return lambda p: p ^ (((p >> 4) & 3) << 1)


def swz_col_for(n, tc):
    """Mirror of swz_col<N, 16, TC>."""
    radices, _ = plan(n, 16)
    rl = radices[-1].bit_length() - 1
    w = 0 if tc >= 16 else 4 - (tc.bit_length() - 1)
    if w == 0:
        return ident
    if n == 8192 and tc == 1:
        return lambda p: p ^ ((p >> 2) & 15) ^ (((p >> 6) & 1) << 2)
    return lambda p: p ^ ((p >> rl) & ((1 << w) - 1))


def wavefronts(addrs_bytes, width):
    """addrs_bytes: byte address per lane (32 lanes); width: access width in bytes (8 or 16)."""
    group = 128 // width            # lanes served together
    total = 0
    for g0 in range(0, 32, group):
        banks = {}
        for a in addrs_bytes[g0:g0 + group]:
            for w in range(width // 4):
                banks.setdefault(((a // 4) + w) % 32, set()).add((a // 4 + w))
        total += max(len(v) for v in banks.values())
    return total, 32 // group


def analyse(n, e, mode, tc=4, elem=8, swz=None):
    radices, sig = plan(n, e)
    tpf = n // e
    swz = swz or (swz_row_for(n) if mode == "row" else swz_col_for(n, tc))
    print(f"N={n} E={e} radices={radices} sigma={sig} mode={mode} tc={tc if mode == 'col' else 1} elem={elem}B")
    L = len(radices)
    worst = 0.0
    for s in range(L):
        tot = ideal = 0
        for idx in range(e):
            for warp in range(2):       # two sample warps
                addrs = []
                for lane in range(32):
                    th = warp * 32 + lane
                    if mode == "row":
                        f, t = divmod(th, tpf) if tpf < 32 * 2 else (0, th)
                        if tpf >= 64:
                            f, t = 0, th % tpf
                        base = f * n
                        p = reg_pos(n, e, radices, sig, s, t % tpf, idx)
                        addrs.append((base + swz(p)) * elem)
                    else:
                        c, t = th % tc, th // tc
                        p = reg_pos(n, e, radices, sig, s, t % tpf, idx)
                        addrs.append((swz(p) * tc + c) * elem)
                wide = mode == "row" and sig[s] == 1 and elem == 8 and radices[s] % 2 == 0
                if wide:            # the kernel moves register pairs (j, j+1) as one 128-bit access
                    if idx % 2:
                        continue
                    w, i = wavefronts(addrs, 16)
                else:
                    w, i = wavefronts(addrs, elem if elem <= 16 else 16)
                tot += w
                ideal += i
        print(f"  stage {s} (R={radices[s]}, sigma={sig[s]}): wavefronts/ideal = {tot / ideal:.2f}")
        worst = max(worst, tot / ideal)
    return worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    e = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    mode = sys.argv[3] if len(sys.argv) > 3 else "row"
    tc = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    elem = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    analyse(n, e, mode, tc, elem)
    print("-- without swizzle --")
    analyse(n, e, mode, tc, elem, swz=ident)
