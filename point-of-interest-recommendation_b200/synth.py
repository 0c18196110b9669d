"""Synthetic check-in data of the shapes BASELINE.json names.

Everything the reference's loaders produce for the hot path, generated directly as integer
matrices (no CSV round trip): padded POI sequences, prefix masks, uniform rejection negatives
(Load_Data_by_length.py:127-143), POI coordinates in a Singapore-sized box, and the haversine
distance-interval matrices with the reference's alignment (Load_Data_by_length.py:24-42,73-78,
165-180).  Seeded with RandomState(123), the seed the reference gives its Theano RNG (GRU.py:45).
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: n_user, n_item, seq, d, UD(km), dd(m)
    "c1": dict(n_user=2321, n_item=5528, seq=64, d=32, UD=40, dd=200, model="gru"),
    "c2": dict(n_user=10000, n_item=40000, seq=32, d=128, UD=40, dd=200, model="distance2pre"),
    "c5": dict(n_user=1000000, n_item=10000000, seq=256, d=512, UD=40, dd=200, model="distance2pre"),
}


def haversine_km(lat1, lon1, lat2, lon2):
    """Vectorised form of the reference's `cal_dis` distance (Load_Data_by_length.py:24-36)."""
    d = 12742.0
    p = 0.017453292519943295
    a = (lat1 - lat2) * p
    b = (lon1 - lon2) * p
    c = (1.0 - np.cos(a)) / 2 + np.cos(lat1 * p) * np.cos(lat2 * p) * (1.0 - np.cos(b)) / 2
    return d * np.arcsin(np.sqrt(c))


def interval_of(dist_km, dd, dist_num):
    """`min(int(dist * 1000 / dd), dist_num)` (Load_Data_by_length.py:38-39)."""
    return np.minimum((dist_km * 1000.0 / dd).astype(np.int64), dist_num).astype(np.int32)


def sample_negatives(rs, P, M, n_item):
    """Uniform negatives outside the user's own training set, pad positions = n_item
    (Load_Data_by_length.py:127-143), vectorised rejection sampling."""
    U, L = P.shape
    Q = rs.randint(0, n_item, size=(U, L)).astype(np.int32)
    own = np.sort(np.where(M > 0, P, -1), axis=1)
    while True:
        # membership of Q[u, t] in own[u, :]
        pos = np.array([np.searchsorted(own[u], Q[u]) for u in range(U)]) if U <= 64 else _rowwise_searchsorted(own, Q)
        pos = np.minimum(pos, L - 1)
        hit = np.take_along_axis(own, pos, axis=1) == Q
        hit &= M > 0
        n = int(hit.sum())
        if n == 0:
            break
        Q[hit] = rs.randint(0, n_item, size=n)
    Q[M == 0] = n_item
    return Q


def _rowwise_searchsorted(sorted_rows, values):
    # offset trick: make rows disjoint ranges so one global searchsorted does all rows
    U, L = sorted_rows.shape
    big = np.int64(max(int(sorted_rows.max()), int(values.max())) + 2)
    off = (np.arange(U, dtype=np.int64) * big)[:, None]
    flat = (sorted_rows.astype(np.int64) + 1 + off).reshape(-1)
    v = (values.astype(np.int64) + 1 + off).reshape(-1)
    pos = np.searchsorted(flat, v).reshape(U, -1) - (np.arange(U, dtype=np.int64) * L)[:, None]
    return np.clip(pos, 0, L - 1)


def make_dataset(n_user, n_item, seq, UD=40, dd=200, ragged=False, zipf=None, seed=123):
    """Returns a dict with int32 matrices P, Q, M, DP, DQ [n_user x seq], lens, coords [n_item x 2],
    one held-out test POI per user (`tes`), and dist_num."""
    rs = np.random.RandomState(seed)
    dist_num = int(UD * 1000 / dd)
    if zipf:
        ranks = rs.zipf(zipf, size=(n_user, seq)).astype(np.int64)
        perm = rs.permutation(n_item)
        P = perm[(ranks - 1) % n_item].astype(np.int32)
    else:
        P = rs.randint(0, n_item, size=(n_user, seq)).astype(np.int32)
    if ragged:
        lens = rs.randint(max(2, seq // 2), seq + 1, size=n_user).astype(np.int32)
        lens[0] = seq
    else:
        lens = np.full(n_user, seq, dtype=np.int32)
    M = (np.arange(seq)[None, :] < lens[:, None]).astype(np.int32)
    P = np.where(M > 0, P, n_item).astype(np.int32)
    Q = sample_negatives(rs, P, M, n_item)
    coords = np.stack([rs.uniform(1.22, 1.47, n_item), rs.uniform(103.60, 104.04, n_item)], axis=1)
    cpad = np.concatenate([coords, coords[:1]], axis=0)          # pad row: any value, masked below
    prev = cpad[P[:, :-1]]
    curp = cpad[P[:, 1:]]
    curq = cpad[Q[:, 1:]]
    DP = np.full((n_user, seq), dist_num, dtype=np.int32)
    DQ = np.full((n_user, seq), dist_num, dtype=np.int32)
    DP[:, 1:] = interval_of(haversine_km(curp[..., 0], curp[..., 1], prev[..., 0], prev[..., 1]), dd, dist_num)
    DQ[:, 1:] = interval_of(haversine_km(curq[..., 0], curq[..., 1], prev[..., 0], prev[..., 1]), dd, dist_num)
    DP[M == 0] = dist_num
    DQ[M == 0] = dist_num
    tes = rs.randint(0, n_item, size=(n_user, 1)).astype(np.int32)
    return dict(P=P, Q=Q, M=M, DP=DP, DQ=DQ, lens=lens, coords=coords, tes=tes, dist_num=dist_num,
                n_user=n_user, n_item=n_item, seq=seq, dd=dd, UD=UD)


def init_state(n_item, d, H, dist_num=None, seed=123, n_user=None):
    """U(-0.5, 0.5) fp32 tables/weights, zero biases (GRU.py:59-64, GRU_Spatial.py:50-71)."""
    rs = np.random.RandomState(seed + 1)
    u = lambda *shape: rs.uniform(-0.5, 0.5, shape).astype(np.float32)
    st = dict(lt=u(n_item + 1, d), wh=u(3, H, H), bi=np.zeros((3, H), dtype=np.float32))
    if dist_num is None:
        st["ui"] = u(3, H, d)
    else:
        st["ui"] = u(3, H, 2 * d)
        st["di"] = u(dist_num + 1, d)
        st["vs"] = u(dist_num + 1, H)
        st["bs"] = np.zeros((dist_num + 1,), dtype=np.float32)
        st["wd"] = np.float64(rs.uniform(0, 0.5))
        st["loss_weight"] = u(2)
    return st


def write_sequence_file(path, n_user, n_item, min_len=6, max_len=20, seed=123):
    """A synthetic dataset in the reference's on-disk sequence format
    (poidata/extract_whole_user_buys.py:81-90): space separated columns
    check_times pois_different u_id u_pois u_times u_coordinates, fields joined by '/',
    coordinates as 'lat,lon'.  Every POI id occurs at least once so that the loaders' alias table has
    exactly n_item entries."""
    rs = np.random.RandomState(seed)
    lat = rs.uniform(1.22, 1.47, n_item)
    lon = rs.uniform(103.60, 104.04, n_item)
    rows = []
    pool = list(rs.permutation(n_item))
    for u in range(n_user):
        L = int(rs.randint(min_len, max_len + 1))
        seq = [int(pool.pop()) if pool else int(rs.randint(0, n_item)) for _ in range(L)]
        for t in range(1, L):
            if rs.rand() < 0.15:
                seq[t] = seq[rs.randint(0, t)]
        t0 = 1.3e9 / 60.0
        times = np.cumsum(rs.randint(5, 900, size=L)) + t0          # minutes
        rows.append((L, '%0.2f' % (len(set(seq)) / L), 'u%d' % u,
                     '/'.join('p%d' % i for i in seq),
                     '/'.join('%d' % t for t in times),
                     '/'.join('%.6f,%.6f' % (lat[i], lon[i]) for i in seq)))
    used = set(int(x[1:]) for r in rows for x in r[3].split('/'))
    for i in sorted(set(range(n_item)) - used):                    # POIs not visited yet: append to random users
        u = rs.randint(0, n_user)
        L, pd_, uid, ps, ts, cs = rows[u]
        last_t = int(ts.split('/')[-1]) + int(rs.randint(5, 900))
        rows[u] = (L + 1, pd_, uid, ps + '/p%d' % i, ts + '/%d' % last_t, cs + '/%.6f,%.6f' % (lat[i], lon[i]))
    with open(path, 'w') as f:
        f.write('check_times pois_different u_id u_pois u_times u_coordinates\n')
        for r in rows:
            f.write(' '.join(str(x) for x in r) + '\n')
    return path
