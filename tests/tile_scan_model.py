"""Lane-level NumPy/Python model of minhash_tile_kernel (query_recommendation_system_b200/csrc/minhash.cu): the same
shared-memory layouts, index arithmetic and step order, one Python loop iteration per lane.  Test infrastructure
only — it exists so that the ALGORITHM of the tile scan (byte-wise word map, chunk-major head table of byte offsets,
bit-sliced first-touch ranks with static low digits and masked high digits, pairwise un-slicing, head-exhaustion and
short-set fallbacks) is pinned to the reference's golden signatures on the CPU, where the CUDA kernel cannot run."""
import numpy as np

NW, TILE_SETS = 4, 32
M32 = 0xFFFFFFFF


def _ffs(x):
    return (x & -x).bit_length()


def head_pitch(K):
    return (K + 31) // 32 * 32


def order_head(tab, R):
    """qrs_perm_head_invert: entry ((r / 8) * Kp + i) * 8 + r % 8 = 4 * (the element with perm_tab[e, i] == r),
    r < R; the padding hashes K..Kp-1 stay zero."""
    m, K = tab.shape
    Kp = head_pitch(K)
    head = np.zeros(R * Kp, np.int64)
    e_idx, i_idx = np.nonzero(tab < R)
    r = tab[e_idx, i_idx]
    head[((r >> 3) * Kp + i_idx) * 8 + (r & 7)] = e_idx * 4
    return head


def nh_for(R):
    return 6 if R <= (8 << 6) else 9 if R <= (8 << 9) else 12


def tile_scan(rowptr, colidx, m, tab, R, force_tile):
    """Signatures int64[n, K] as minhash_tile_kernel computes them (mode TILE when force_tile, else AUTO)."""
    n, K = rowptr.size - 1, tab.shape[1]
    head = order_head(tab, R)
    Kp = ((K + 31) // 32) * 32 + 2                      # tile_sig_stride
    sig = np.full((n, K), -12345, np.int64)
    for tile in range((n + TILE_SETS - 1) // TILE_SETS):
        s0 = tile * TILE_SETS
        nsets = min(TILE_SETS, n - s0)
        rp = [int(rowptr[s0 + min(t, nsets)]) for t in range(TILE_SETS + 1)]
        bmb = np.zeros(((m + 3) // 4 * 4) * 4, np.uint8)          # the word map, byte view
        sigt = np.full(TILE_SETS * Kp, -777, np.int64)
        scanmask = 0
        for lane in range(32):
            if lane < nsets:
                ln = rp[lane + 1] - rp[lane]
                if ln > 0 and (force_tile or ln * ln >= m):
                    scanmask |= 1 << lane
        for warp in range(NW):
            for jj in range(8):                                     # empty sets: a row of -1
                j = warp * 8 + jj
                if j < nsets and rp[j + 1] == rp[j]:
                    sigt[j * Kp:j * Kp + K] = -1
            todo = (scanmask >> (warp * 8)) & 0xFF                 # build: byte `warp` of every word
            jj = _ffs(todo) - 1 if todo else -1
            t0 = 0
            while jj >= 0:
                lo = rp[warp * 8 + jj]
                ln = rp[warp * 8 + jj + 1] - lo
                for lane in range(32):
                    for u in range(8):
                        t = t0 + lane + 32 * u
                        if t < ln:
                            bmb[int(colidx[lo + t]) * 4 + warp] |= 1 << jj
                t0 += 256
                if t0 >= ln:
                    todo &= todo - 1
                    jj = _ffs(todo) - 1 if todo else -1
                    t0 = 0
        bm = bmb.view(np.uint32)
        Kh, nchunks, NH = head_pitch(K), R >> 3, nh_for(R)
        NP = 3 + NH
        for warp in range(NW):                                      # walk: lane = hash
            for i0 in range(warp * 32, K, NW * 32):
                found = [((~scanmask) & M32) if i0 + l < K else M32 for l in range(32)]
                P = [[0] * NP for _ in range(32)]                   # bit-sliced ranks: P[lane][plane]
                c8, stop = 0, False
                while not stop:
                    before8 = list(found)
                    for cc in range(8):
                        c = c8 + cc
                        if c >= nchunks or all(f == M32 for f in found):
                            stop = True
                            break
                        for lane in range(32):
                            i = i0 + lane
                            q = head[(c * Kh + i) * 8:(c * Kh + i) * 8 + 8]      # two 16-byte loads
                            before = found[lane]
                            for u in range(8):
                                w = int(bm[int(q[u]) >> 2])                          # byte offset into the word map
                                for p in range(3):
                                    if (u >> p) & 1:
                                        P[lane][p] |= w & ~found[lane] & M32
                                found[lane] |= w
                            for p in range(3):                                        # static digits of the chunk
                                if (cc >> p) & 1:
                                    P[lane][3 + p] |= found[lane] & ~before & M32
                    for lane in range(32):                                            # masked high digits
                        new8 = found[lane] & ~before8[lane] & M32
                        for h in range(3, NH):
                            if (c8 >> h) & 1:
                                P[lane][3 + h] |= new8
                    c8 += 8
                for lane in range(32):                              # un-slice: sets j and j + 16 together
                    if i0 + lane >= K:
                        continue
                    for j in range(16):
                        lo_on, hi_on = (scanmask >> j) & 1, (scanmask >> (j + 16)) & 1
                        if not (lo_on or hi_on):
                            continue
                        v2 = 0
                        for p in range(NP):
                            t = (P[lane][p] << (p - j)) & M32 if p >= j else P[lane][p] >> (j - p)
                            v2 |= t & ((0x00010001 << p) & M32)
                        if lo_on:
                            sigt[j * Kp + i0 + lane] = v2 & 0xFFFF
                        if hi_on:
                            sigt[(j + 16) * Kp + i0 + lane] = v2 >> 16
                for src in range(32):                               # still open after the head: gather-min
                    mm = (~found[src]) & M32
                    while mm:
                        j = _ffs(mm) - 1
                        mm &= mm - 1
                        lo, hi = rp[j], rp[j + 1]
                        sigt[j * Kp + i0 + src] = int(tab[colidx[lo:hi], i0 + src].min())
        for j in range(nsets):                                      # short sets (AUTO): direct gather-min
            lo, hi = rp[j], rp[j + 1]
            if hi > lo and not (scanmask >> j) & 1:
                sigt[j * Kp:j * Kp + K] = tab[colidx[lo:hi], :].min(axis=0)
        for j in range(nsets):
            sig[s0 + j] = sigt[j * Kp:j * Kp + K]
    return sig
