#!/usr/bin/env python
"""Per-row traffic model of the tiled U(1) matvec (csrc/apply_u1.cu) for a spin-1/2 ring in the Sz=0 sector.

Counts, for a tile width k and an optional cluster of 2^m tiles that exchange their mid-bit bonds on chip, the bytes a
row pulls from the L2 (everything that is not the CTA's own shared memory) and the part of them that cannot be an L2
hit for an ascending tile order (neighbour tiles further away than the L2 holds).  The numbers reproduce the measured
126 B/row and ~49 B/row of DRAM traffic of the shipped k=15, m=0 kernel to ~5 % and are what DESIGN.md section 7 ranks
the next steps by.  Pure host arithmetic: python tools/u1_traffic_model.py [n_sites]"""
import math
import sys


def firing_prob(n, n_dn):
    """probability that a given bond (two distinct sites) is anti-aligned in the fixed-popcount sector"""
    return 2.0 * math.comb(n - 2, n_dn - 1) / math.comb(n, n_dn)


def model(n, k, m, l2_bytes=126e6, l2_usable=0.5, vec=8):
    n_dn = n // 2
    q = firing_prob(n, n_dn)
    hb = n - k
    dim = math.comb(n, n_dn)
    n_tiles = sum(1 for h in range(1 << hb) if 0 <= n_dn - bin(h).count("1") <= k) if hb <= 20 else 2 ** hb
    tile_bytes = dim * vec / n_tiles
    # bonds of the ring by where their two sites live
    low = k - 1                      # both sites inside the tile: shared-memory gathers, index bytes only
    straddle_km1 = 1                 # (k-1, k): contiguous block stream from the neighbour tile
    periodic = 1                     # (n-1, 0): gather from the tile with the top bit flipped
    high = hb - 1                    # (k+j, k+j+1), j = 0 .. hb-2: coalesced streams
    in_cluster = min(max(m - 1, 0), high) + (1 if m >= 1 else 0)   # bonds among the m lowest high bits (+ the k-1|k straddler)
    streams_l2 = (high + straddle_km1 - in_cluster) * q            # 8-byte columns per row that come from the L2
    ell_slots = 2 * math.ceil(0.5 * (low * q + 2.2 * math.sqrt(low * q * (1 - q))))   # padded to the group maximum, in pairs
    bytes_l2 = {
        "own tile": vec, "y": vec, "streams": streams_l2 * vec, "periodic gather": q * 2 * vec + 2,
        "ELL indices": 2 * ell_slots, "codes/lists": 4,
    }
    # DRAM: own tile + y always; a stream on H-bit j misses when the tiles between the two partners exceed the usable L2
    far = [j for j in range(high) if (2 ** j) * tile_bytes * 2 > l2_bytes * l2_usable and j >= m - 1]
    dram = 2 * vec + len(far) * q * vec + q * 2 * vec     # + periodic bond: always far, half-used sectors
    return sum(bytes_l2.values()), bytes_l2, dram, tile_bytes, ell_slots


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dim = math.comb(n, n // 2)
    lts_cap = 6300 * 1.96e9      # B/s the L2 slices deliver chip-wide (B300_MICROARCH.md)
    print(f"ring of {n} sites, Sz=0: {dim:,} rows, P(bond fires) = {firing_prob(n, n // 2):.3f}")
    print(" k  m | max tile KB | L2->SM B/row (streams, ELL idx) | DRAM B/row | ms at the L2 cap | ms at 6.5 TB/s DRAM")
    for k in (14, 15, 16, 17):
        for m in (0, 3, 4):
            tot, parts, dram, tb, slots = model(n, k, m)
            print(f"{k:2d} {m:2d} | {math.comb(k, k // 2) * 8 / 1e3:11.1f} | {tot:6.1f} ({parts['streams']:5.1f}, {parts['ELL indices']:4.1f})"
                  f"          | {dram:6.1f}     | {tot * dim / lts_cap * 1e3:6.2f}           | {dram * dim / 6.5e12 * 1e3:6.2f}")
