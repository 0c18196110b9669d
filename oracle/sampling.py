"""CPU restatement of csrc/sampling.cuh.  TEST INFRASTRUCTURE ONLY.

The sampling rule is the reference's (Load_Data_by_length.py:127-162: uniform draw from [0, n_item), redrawn while the
item occurs in the user's training row -- test variant: training or test row; pad positions get the pad id); the random
stream is the engine's counter-based one (Philox4x32-10, counter = (user, position, attempt // 4, epoch), key = seed,
j = (word * n_item) >> 32), because the reference's sequential Mersenne Twister stream has no parallel equivalent.
Interval binning follows cal_dis / fun_compute_dist_neg (:24-42, :165-180) in float64."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint32 arrays c0..c3; k0, k1 Python ints.  Returns the four output words."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def sample_negatives(rows, forbidden_a, n_item, seed, epoch, forbidden_b=None):
    rows = np.asarray(rows)
    n_user, lrow = rows.shape
    out = np.full_like(rows, n_item)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for u in range(n_user):
        forbid = set(np.asarray(forbidden_a[u]).tolist())
        if forbidden_b is not None:
            forbid |= set(np.asarray(forbidden_b[u]).tolist())
        for t in range(lrow):
            if rows[u, t] == n_item:
                continue
            blk, done = 0, False
            while not done:
                words = philox4x32_10([u], [t], [blk], [epoch & 0xFFFFFFFF], k0, k1)
                for w in words:
                    j = int((int(w[0]) * n_item) >> 32)
                    if j not in forbid:
                        out[u, t] = j; done = True
                        break
                blk += 1
    return out


def neg_intervals(p, q, lens, coords, dd, dist_num):
    """fun_compute_dist_neg with numpy float64 (same formula as cal_dis)."""
    p, q = np.asarray(p), np.asarray(q)
    coords = np.asarray(coords, dtype=np.float64)
    out = np.full_like(p, dist_num)
    d, rad = 12742.0, 0.017453292519943295
    for u in range(p.shape[0]):
        L = int(lens[u])
        for t in range(1, L):
            lat1, lon1 = coords[q[u, t]]
            lat2, lon2 = coords[p[u, t - 1]]
            a, b = (lat1 - lat2) * rad, (lon1 - lon2) * rad
            c = (1.0 - np.cos(a)) / 2 + np.cos(lat1 * rad) * np.cos(lat2 * rad) * (1.0 - np.cos(b)) / 2
            out[u, t] = min(int(d * np.arcsin(np.sqrt(c)) * 1000 / dd), dist_num)
    return out
