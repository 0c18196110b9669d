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
    # c3 / c4: BASELINE.json gives |POI|, d and the negatives per positive; |U| and seq are SURVEY.md 8d's assumptions
    "c3": dict(n_user=10000, n_item=100000, seq=32, d=256, neg=20, UD=40, dd=200, model="prme"),
    "c4": dict(n_user=100000, n_item=1000000, seq=32, d=256, neg=100, UD=40, dd=200, model="geoie"),
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


def prme_km(lat1, lon1, lat2, lon2):
    """Vectorised `cal_dis` of the PRME loader (Load_Data_prme.py:27-36): equatorial radius, 2 asin form."""
    rad = np.pi / 180.0
    a = (lat1 - lat2) * rad
    b = (lon1 - lon2) * rad
    s = 2 * np.arcsin(np.sqrt(np.sin(a / 2) ** 2 + np.cos(lat1 * rad) * np.cos(lat2 * rad) * np.sin(b / 2) ** 2))
    return s * 6378.137


def make_mf_dataset(n_user, n_item, seq, K, seed=123):
    """Check-in sequences for the pairwise models with K negatives per positive (C3 PRME, C4 GeoIE): P [U x seq] uniform
    POIs, Q [U x seq x K] uniform negatives outside the user's own sequence (Load_Data_prme.py:120-140 draws one; here K
    per position), coords [n_item+1 x 2] (pad row 0,0 like Load_Data_prme.py:108-112), PRME's per-position inputs: `dist`
    = km between consecutive POIs (Load_Data_prme.py:60-75) and `gap` = integer minutes U{1..720} (threshold 360)."""
    rs = np.random.RandomState(seed)
    P = rs.randint(0, n_item, size=(n_user, seq)).astype(np.int32)
    Q = rs.randint(0, n_item, size=(n_user, seq, K)).astype(np.int32)
    own = np.sort(P, axis=1)
    for _ in range(8):                                                   # rejection of the user's own POIs
        pos = _rowwise_searchsorted(own, Q.reshape(n_user, -1)).reshape(Q.shape)
        hit = np.take_along_axis(own, pos.reshape(n_user, -1), axis=1).reshape(Q.shape) == Q
        n = int(hit.sum())
        if n == 0:
            break
        Q[hit] = rs.randint(0, n_item, size=n)
    coords = np.zeros((n_item + 1, 2), dtype=np.float64)
    coords[:n_item, 0] = rs.uniform(1.22, 1.47, n_item); coords[:n_item, 1] = rs.uniform(103.60, 104.04, n_item)
    a, b = coords[P[:, 1:]], coords[P[:, :-1]]
    dist = np.zeros((n_user, seq), dtype=np.float32)
    dist[:, 1:] = prme_km(a[..., 0], a[..., 1], b[..., 0], b[..., 1])
    gap = rs.randint(1, 721, size=(n_user, seq)).astype(np.int32)
    return dict(P=P, Q=Q, coords=coords, dist=dist, gap=gap, n_user=n_user, n_item=n_item, seq=seq, K=K)


def init_mf_state(model, n_user, n_item, d, seed=123):
    """U(-0.5, 0.5) fp32 tables (PRME.py:76-83, GeoIE.py:65-72); GeoIE's a, b drawn positive so that a d^b is finite."""
    rs = np.random.RandomState(seed + 1)
    u = lambda *shape: rs.uniform(-0.5, 0.5, shape).astype(np.float32)
    if model == "prme":
        return dict(ds=u(n_item + 1, d), dp=u(n_item + 1, d), du=u(n_user, d))
    return dict(g=u(n_item + 1, d), h=u(n_item + 1, d), t=u(n_user, d), z=u(n_item + 1, d),
                a=np.float64(rs.uniform(0.05, 0.5)), b=np.float64(rs.uniform(0.05, 0.5)))


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


# ----------------------------------------------------------------------------------------------------------------------
# C5 scale (|POI| = 10M, |U| = 1M, seq = 256, d = 512): generated on the DEVICE -- the item table alone is 20.5 GB
# ----------------------------------------------------------------------------------------------------------------------
def hashed_uniform_rows(rows, d, seed, device, scale=1.0, chunk=1 << 20):
    """Table rows whose values depend only on (global row id, column, seed): the same table whatever the sharding.
    value = ((row * 2654435761 + col * 40503 + seed * 97) mod 2^24) / 2^24 - 0.5, times `scale`; fp32 [len(rows) x d]."""
    import torch
    rows = torch.as_tensor(rows, dtype=torch.int64, device=device)
    out = torch.empty((rows.numel(), d), dtype=torch.float32, device=device)
    cols = torch.arange(d, dtype=torch.int64, device=device) * 40503 + int(seed) * 97
    for s in range(0, rows.numel(), chunk):
        r = rows[s:s + chunk, None] * 2654435761 + cols[None, :]
        out[s:s + chunk] = ((r & 0xFFFFFF).to(torch.float32) * (1.0 / (1 << 24)) - 0.5) * scale
    return out


def make_dataset_device(engine, n_user, n_item, seq, users=None, UD=40, dd=200, seed=123):
    """`make_dataset` at C5 scale, every array built on the engine's device: uniform POI sequences (no padding), negatives
    by the engine's device sampler (csrc/sampling.cuh: uniform, rejecting the user's own POIs -- the rule of
    Load_Data_by_length.py:127-143 on a counter-based stream), coordinates in the Singapore-sized box and both interval
    matrices by `poi_neg_intervals` (fp64 haversine of Load_Data_by_length.py:24-42).  `users` (int64 tensor / array of
    global user ids, default all) selects the rows to keep; a user's rows do not depend on which other users are kept.
    Returns int32 device tensors P, Q, DP, DQ [len(users) x seq], lens, float64 coords [n_item+1 x 2], dist_num."""
    import torch
    dev = engine.torch_device
    dist_num = int(UD * 1000 / dd)
    g = torch.Generator(device=dev); g.manual_seed(seed)
    P = torch.randint(0, n_item, (n_user, seq), dtype=torch.int32, device=dev, generator=g)
    coords = torch.zeros((n_item + 1, 2), dtype=torch.float64, device=dev)
    coords[:n_item, 0] = torch.rand(n_item, dtype=torch.float64, device=dev, generator=g) * 0.25 + 1.22
    coords[:n_item, 1] = torch.rand(n_item, dtype=torch.float64, device=dev, generator=g) * 0.44 + 103.60
    if users is not None:
        P = P[torch.as_tensor(users, dtype=torch.int64, device=dev)].contiguous()
    lens = torch.full((P.shape[0],), seq, dtype=torch.int32, device=dev)
    srt = torch.sort(P, dim=1).values.contiguous()
    Q = engine.sample_negatives(P, srt, n_item, seed, 0)
    del srt
    DP = engine.neg_intervals(P, P, lens, coords, float(dd), dist_num)          # interval(p[t-1], p[t]); D at t = 0
    DQ = engine.neg_intervals(P, Q, lens, coords, float(dd), dist_num)          # interval(p[t-1], q[t])
    engine.sync()
    return dict(P=P, Q=Q, DP=DP, DQ=DQ, lens=lens, coords=coords, dist_num=dist_num, n_user=int(P.shape[0]), n_item=n_item, seq=seq)


def init_state_device(n_item, d, dist_num, device, rows=None, seed=123):
    """Distance2Pre parameters at C5 scale on the device.  `rows` = the global item rows this rank holds (default all
    n_item + 1).  Tables U(-0.5, 0.5) like the reference (GRU.py:60, GRU_Spatial.py:57); the GRU / head weights are
    U(-0.5, 0.5) * 4 / sqrt(d): at d = 512 the reference's unscaled init saturates every gate (pre-activations of order
    sqrt(d) * 0.15 ~ 3.4 per term), which makes the gradients vanish and the run numerically meaningless; the scaling keeps
    the recurrence in the regime the d = 128 configuration has with the reference's init (tests use the same scaling)."""
    import torch
    rs = np.random.RandomState(seed + 1)
    sc = 4.0 / np.sqrt(d)
    u = lambda *shape: torch.from_numpy((rs.uniform(-0.5, 0.5, shape) * sc).astype(np.float32)).to(device)
    rows = torch.arange(n_item + 1, device=device) if rows is None else rows
    return dict(lt=hashed_uniform_rows(rows, d, seed, device), ui=u(3, d, 2 * d), wh=u(3, d, d),
                bi=torch.zeros((3, d), dtype=torch.float32, device=device),
                di=torch.from_numpy(rs.uniform(-0.5, 0.5, (dist_num + 1, d)).astype(np.float32)).to(device),
                vs=u(dist_num + 1, d), bs=torch.zeros((dist_num + 1,), dtype=torch.float32, device=device),
                wd=np.float64(rs.uniform(0, 0.5)), loss_weight=rs.uniform(-0.5, 0.5, 2).astype(np.float32))
