"""Design prototype (not product, not oracle): checks on the CPU that the parallel formulation used
by the CUDA path — Boruvka levels == union-by-rank ranks, per-root chains replayed in rank waves —
reproduces the reference's sequential merge sequence exactly (compared with the port's merge trace).
"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from numba import njit

INF = np.uint32(0xFFFFFFFF)

def edge_slots(flow, n8=True):
    H, W = flow.shape[:2]
    N = W * H
    w = np.full((H, W, 4), np.inf)
    f = flow.astype(np.float32)
    def d(a, b):
        dx = (a[..., 0] - b[..., 0]).astype(np.float32).astype(np.float64)
        dy = (a[..., 1] - b[..., 1]).astype(np.float32).astype(np.float64)
        return np.sqrt(dx * dx + dy * dy)
    w[:, 1:, 0] = d(f[:, 1:], f[:, :-1])
    w[1:, :, 1] = d(f[1:, :], f[:-1, :])
    if n8:
        w[1:, 1:, 2] = d(f[1:, 1:], f[:-1, :-1])
        w[:-1, 1:, 3] = d(f[:-1, 1:], f[1:, :-1])
    w = w.reshape(-1)
    order = np.argsort(w, kind='stable').astype(np.uint32)
    E = int(np.isfinite(w).sum())
    rank = np.empty(4 * N, np.uint32)
    rank[order] = np.arange(4 * N, dtype=np.uint32)
    rank[~np.isfinite(w)] = INF
    return w, order[:E], rank, E

@njit(cache=True)
def nbr_of(s, d, W):
    if d == 0: return s - 1
    if d == 1: return s - W
    if d == 2: return s - W - 1
    return s + W - 1

@njit(cache=True)
def boruvka(rank, order, W, H):
    N = W * H
    comp = np.arange(N).astype(np.int32)
    loss_time = np.full(N, INF, np.uint32)
    up = np.arange(N).astype(np.int32)
    lvl = np.zeros(N, np.int32)
    newp = np.arange(N).astype(np.int32)
    level = 0
    while True:
        best = np.full(N, INF, np.uint32)
        for p in range(N):
            x = p % W; y = p // W
            cp = comp[p]
            # own back-edges
            for d in range(4):
                r = rank[4 * p + d]
                if r != INF:
                    q = nbr_of(p, d, W)
                    if comp[q] != cp:
                        if r < best[cp]: best[cp] = r
                        if r < best[comp[q]]: best[comp[q]] = r
        nroots = 0
        for c in range(N):
            if comp[c] != c: continue
            nroots += 1
            t = best[c]
            if t == INF:
                newp[c] = c
                continue
            seq = order[t]
            s = seq >> 2; d = seq & 3
            e = nbr_of(s, d, W)
            cs = comp[s]; ce = comp[e]
            other = ce if cs == c else cs
            mutual = best[other] == t
            if mutual and c == ce:
                newp[c] = c
            else:
                newp[c] = other
                loss_time[c] = t
                lvl[c] = level
        if nroots == 1:
            break
        # pointer jumping
        changed = True
        while changed:
            changed = False
            for c in range(N):
                if comp[c] == c:
                    g = newp[newp[c]]
                    if g != newp[c]:
                        newp[c] = g; changed = True
        for c in range(N):
            if comp[c] == c and newp[c] != c:
                up[c] = newp[c]
        for p in range(N):
            comp[p] = newp[comp[p]]
        level += 1
    for c in range(N):
        if loss_time[c] == INF: lvl[c] = level
    return loss_time, up, lvl, level

@njit(cache=True)
def winners(loss_time, up):
    N = loss_time.shape[0]
    win = np.full(N, -1, np.int32)
    maxhops = 0
    for c in range(N):
        t = loss_time[c]
        if t == INF: continue
        cur = up[c]; hops = 0
        while loss_time[cur] < t:
            cur = up[cur]; hops += 1
        win[c] = cur
        if hops > maxhops: maxhops = hops
    return win, maxhops

@njit(cache=True)
def replay(flow, W, H, loss_time, win, lvl, maxlvl):
    N = W * H
    # events sorted by (winner, time)
    losers = np.nonzero(win >= 0)[0].astype(np.int32)
    key = win[losers].astype(np.int64) * (1 << 32) + loss_time[losers].astype(np.int64)
    o = np.argsort(key)
    ev_loser = losers[o]
    ev_win = win[ev_loser]
    nE = ev_loser.shape[0]
    chain_start = np.full(N + 1, -1, np.int32)
    chain_len = np.zeros(N, np.int32)
    for i in range(nE):
        r = ev_win[i]
        if chain_start[r] < 0: chain_start[r] = i
        chain_len[r] += 1
    size = np.ones(N, np.int32)
    fl = flow.reshape(-1, 2).copy()
    bb = np.empty((N, 4), np.int32)
    for p in range(N):
        bb[p, 0] = p % W; bb[p, 1] = p // W; bb[p, 2] = p % W; bb[p, 3] = p // W
    out_size = np.zeros(nE, np.int32); out_bb = np.zeros((nE, 4), np.int32); out_fl = np.zeros((nE, 2), np.float32)
    for wave in range(1, maxlvl + 1):
        for r in range(N):
            if lvl[r] != wave or chain_len[r] == 0: continue
            s = size[r]; fx = fl[r, 0]; fy = fl[r, 1]
            b0 = bb[r, 0]; b1 = bb[r, 1]; b2 = bb[r, 2]; b3 = bb[r, 3]
            for i in range(chain_start[r], chain_start[r] + chain_len[r]):
                a = ev_loser[i]
                sa = size[a]
                wax = np.float32(fl[a, 0] * np.float32(sa)); way = np.float32(fl[a, 1] * np.float32(sa))
                wbx = np.float32(fx * np.float32(s)); wby = np.float32(fy * np.float32(s))
                sx = np.float32(wax + wbx); sy = np.float32(way + wby)
                inv = 1.0 / np.float64(sa + s)
                fx = np.float32(np.float64(sx) * inv); fy = np.float32(np.float64(sy) * inv)
                s += sa
                b0 = min(b0, bb[a, 0]); b1 = min(b1, bb[a, 1]); b2 = max(b2, bb[a, 2]); b3 = max(b3, bb[a, 3])
                out_size[i] = s; out_fl[i, 0] = fx; out_fl[i, 1] = fy
                out_bb[i, 0] = b0; out_bb[i, 1] = b1; out_bb[i, 2] = b2; out_bb[i, 3] = b3
            size[r] = s; fl[r, 0] = fx; fl[r, 1] = fy
            bb[r, 0] = b0; bb[r, 1] = b1; bb[r, 2] = b2; bb[r, 3] = b3
    return ev_loser, ev_win, out_size, out_bb, out_fl

def run(flow_blurred, n8=True):
    H, W = flow_blurred.shape[:2]
    t0 = time.time()
    w, order, rank, E = edge_slots(flow_blurred, n8)
    t1 = time.time()
    loss_time, up, lvl, maxlvl = boruvka(rank, order, W, H)
    t2 = time.time()
    win, maxhops = winners(loss_time, up)
    ev_loser, ev_win, s, bb, fl = replay(np.ascontiguousarray(flow_blurred, np.float32), W, H, loss_time, win, lvl, maxlvl)
    t3 = time.time()
    print(f"edges {t1-t0:.2f}s boruvka {t2-t1:.2f}s levels={maxlvl} climb maxhops={maxhops} replay {t3-t2:.2f}s")
    return dict(loss_time=loss_time, up=up, lvl=lvl, win=win, ev_loser=ev_loser, ev_win=ev_win, size=s, bbox=bb, flow=fl, E=E)

def check(flow_blurred, n8=True):
    from oracle import cpu
    P = cpu.port()
    persp, inv, upm = P.get_mats()
    ref = P.segment(flow_blurred, persp, inv, upm, neighbors=8 if n8 else 4, trace=True)
    tr = ref['trace']
    r = run(flow_blurred, n8)
    # sort our events by time
    t = r['loss_time'][r['ev_loser']]
    o = np.argsort(t, kind='stable')
    ok = True
    for name, a, b in [("edge_pos", t[o].astype(np.int64), tr['edge_pos'].astype(np.int64)),
                       ("loser", r['ev_loser'][o], tr['loser']), ("winner", r['ev_win'][o], tr['winner']),
                       ("size", r['size'][o], tr['size']), ("bbox", r['bbox'][o], tr['bbox']),
                       ("flow", r['flow'][o].view(np.uint32), tr['flow'].view(np.uint32))]:
        eq = np.array_equal(a, b)
        print(f"  {name}: {'OK' if eq else 'MISMATCH'}")
        if not eq:
            ok = False
            bad = np.nonzero(np.any(np.atleast_2d((a != b).reshape(len(a), -1)), axis=1))[0]
            print("   first bad", bad[:5], a[bad[:3]], b[bad[:3]])
    return ok

if __name__ == "__main__":
    import cv2
    if len(sys.argv) > 1 and sys.argv[1] == "rand":
        rng = np.random.default_rng(1)
        H, W = 90, 160
        f = rng.normal(size=(H, W, 2)).astype(np.float32)
        f[20:50, 30:90] = 0  # exact ties
        fb = cv2.GaussianBlur(f, (0, 0), 1.0); fb[25:45, 40:80] = 0.25
        print(check(fb), check(fb, n8=False))
    else:
        im1 = cv2.imread('/root/reference/data/frame_1052.png'); im2 = cv2.imread('/root/reference/data/frame_1053.png')
        g1 = cv2.cvtColor(im1, cv2.COLOR_BGR2GRAY); g2 = cv2.cvtColor(im2, cv2.COLOR_BGR2GRAY)
        flow = cv2.calcOpticalFlowFarneback(g1, g2, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        fb = cv2.GaussianBlur(flow, (0, 0), 3.0)
        print(check(fb))
