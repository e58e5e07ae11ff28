"""Lane-level NumPy/Python model of minhash_tile_kernel (query_recommendation_system_b200/csrc/minhash.cu): the same
shared-memory layouts, index arithmetic and step order, one Python loop iteration per lane.  Test infrastructure
only — it exists so that the ALGORITHM of the tile scan (byte-wise word map, 16-rank staging through the padded
16-bit tile, first-touch recording, head-exhaustion and short-set fallbacks) is pinned to the reference's golden
signatures on the CPU, where the CUDA kernel cannot run."""
import numpy as np

NW, TILE_SETS, CHUNK, PITCH = 4, 32, 16, 24


def _ffs(x):
    return (x & -x).bit_length()


def order_head(tab, R):
    """qrs_perm_head_invert: order_head[i*R + r] = the element with perm_tab[e, i] == r (r < R)."""
    m, K = tab.shape
    head = np.zeros(K * R, np.int64)
    e_idx, i_idx = np.nonzero(tab < R)
    head[i_idx * R + tab[e_idx, i_idx]] = e_idx
    return head


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
        for warp in range(NW):                                      # walk: lane = hash
            for i0 in range(warp * 32, K, NW * 32):
                found = [((~scanmask) & 0xFFFFFFFF) if i0 + l < K else 0xFFFFFFFF for l in range(32)]
                xp = np.zeros(32 * PITCH, np.int64)
                r0 = 0
                while r0 < R and not all(f == 0xFFFFFFFF for f in found):
                    for s8 in range(0, 32, 8):                      # staging: 4 lanes per hash row
                        for lane in range(32):
                            hh, rq4 = s8 + (lane >> 2), (lane & 3) * 4
                            ih = min(i0 + hh, K - 1)
                            xp[hh * PITCH + rq4:hh * PITCH + rq4 + 4] = head[ih * R + r0 + rq4:ih * R + r0 + rq4 + 4]
                    for rq in range(0, CHUNK, 8):
                        for lane in range(32):
                            for u in range(8):
                                w = int(bm[int(xp[lane * PITCH + rq + u])])
                                nw = w & ~found[lane] & 0xFFFFFFFF
                                found[lane] |= w
                                while nw:
                                    j = _ffs(nw) - 1
                                    nw &= nw - 1
                                    sigt[j * Kp + i0 + lane] = r0 + rq + u
                        if all(f == 0xFFFFFFFF for f in found):
                            break
                    r0 += CHUNK
                for src in range(32):                               # still open after the head: gather-min
                    mm = (~found[src]) & 0xFFFFFFFF
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
