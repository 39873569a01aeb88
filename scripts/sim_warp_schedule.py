"""Offline study of warp scheduling policies for the traversal kernel (no GPU): uses the oracle's per-ray event
strings (N = inner node step, I = instance entry, T = triangle test) and replays them on a 32-lane warp model.
Cost model: issue slots per phase execution (cN, cT, cI) regardless of how many lanes are active."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_binding as ob
from cudatracerlib_b200 import Scene, RAY_DTYPE

cN, cT, cI, cFetch = 45, 36, 90, 30


def bounce_rays(scene, w, h, x0, y0, bw, bh, seed=1):
    v = scene.view
    rays = np.zeros(bw * bh, RAY_DTYPE)
    i = 0
    for y in range(y0, y0 + bh):
        for x in range(x0, x0 + bw):
            o, d = ob.camera_ray(v, x + 0.5, y + 0.5)
            rays["o"][i] = o; rays["d"][i] = d; i += 1
    rays["tmin"] = v.ray_eps; rays["tmax"] = 3e38
    res = ob.trace_rays(v, rays)
    rng = np.random.default_rng(seed)
    hit = res["tri_idx"] != 0xffffffff
    p = rays["o"] + rays["d"] * res["dist"][:, None]
    r = rng.normal(size=(len(rays), 3)); r /= np.linalg.norm(r, axis=1, keepdims=True)
    nd = -rays["d"] + 0.98 * r; nd /= np.linalg.norm(nd, axis=1, keepdims=True)
    b = np.zeros(hit.sum(), RAY_DTYPE)
    b["o"] = (p - rays["d"] * 1e-3)[hit]; b["d"] = nd[hit].astype(np.float32); b["tmin"] = v.ray_eps; b["tmax"] = 3e38
    return rays, b


def events(scene, rays, any_hit=False):
    L = ob.oracle()
    L.orc_trace_events.restype = C.c_longlong
    L.orc_trace_events.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
    off = np.zeros(len(rays) + 1, np.int64)
    tot = L.orc_trace_events(C.byref(scene.view), len(rays), rays.ctypes.data, int(any_hit), None, 0, off.ctypes.data)
    ev = np.zeros(tot, np.uint8)
    L.orc_trace_events(C.byref(scene.view), len(rays), rays.ctypes.data, int(any_hit), ev.ctypes.data, tot, off.ctypes.data)
    return [bytes(ev[off[i]:off[i + 1]]) for i in range(len(rays))]


def ideal(evs):
    n = sum(e.count(b"N") for e in evs); t = sum(e.count(b"T") for e in evs); i = sum(e.count(b"I") for e in evs)
    return (n * cN + t * cT + i * cI) / 32.0, n, t, i


def sim(evs, policy, theta=8, theta_lo=1, batch=False):
    """returns issue-slot cost for one warp stream processing all rays in order (warps are statistically alike)."""
    nxt = 0; cost = 0
    lane = [None] * 32; pos = [0] * 32
    n_rays = len(evs)
    lanes_N = lanes_T = execs_N = execs_T = 0
    while True:
        # refill
        idle = [l for l in range(32) if lane[l] is None]
        if idle and nxt < n_rays and (not batch or len(idle) == 32):
            for l in idle:
                if nxt < n_rays:
                    lane[l] = evs[nxt]; pos[l] = 0; nxt += 1
                    if len(lane[l]) == 0: lane[l] = None
            cost += cFetch
        act = [l for l in range(32) if lane[l] is not None]
        if not act:
            if nxt >= n_rays: break
            continue
        def state(l): return lane[l][pos[l]]
        def adv(l):
            pos[l] += 1
            if pos[l] >= len(lane[l]): lane[l] = None
        def count():
            cN_ = [l for l in range(32) if lane[l] is not None and state(l) == 78]
            cT_ = [l for l in range(32) if lane[l] is not None and state(l) == 84]
            cI_ = [l for l in range(32) if lane[l] is not None and state(l) == 73]
            return cN_, cT_, cI_
        if policy == "whilewhile":
            # N phase until no active lane is in N state
            while True:
                a, b, c = count()
                if not a: break
                for l in a: adv(l)
                cost += cN; lanes_N += len(a); execs_N += 1
            while True:
                a, b, c = count()
                if not b and not c: break
                if c:
                    for l in c: adv(l)
                    cost += cI
                if b:
                    for l in b: adv(l)
                    cost += cT; lanes_T += len(b); execs_T += 1
        elif policy == "threshold":
            a, b, c = count()
            # node phase while leaf-waiters below theta
            while a and (len(b) + len(c)) < theta:
                for l in a: adv(l)
                cost += cN; lanes_N += len(a); execs_N += 1
                a, b, c = count()
                if any(lane[l] is None for l in range(32)) and nxt < n_rays: break  # go refill
            if (len(b) + len(c)) >= theta or not a:
                while b or c:
                    if c and (len(c) >= 4 or not b):
                        for l in c: adv(l)
                        cost += cI
                    elif b:
                        for l in b: adv(l)
                        cost += cT; lanes_T += len(b); execs_T += 1
                    a, b, c = count()
                    if len(b) + len(c) < theta_lo: break
        elif policy == "majority":
            a, b, c = count()
            if len(a) >= max(len(b), len(c)) and a:
                for l in a: adv(l)
                cost += cN; lanes_N += len(a); execs_N += 1
            elif len(b) >= len(c) and b:
                for l in b: adv(l)
                cost += cT; lanes_T += len(b); execs_T += 1
            elif c:
                for l in c: adv(l)
                cost += cI
    return cost, lanes_N / max(1, execs_N), lanes_T / max(1, execs_T)


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "c2"
    s = Scene(kind, 1920, 1080)
    prim, bnc = bounce_rays(s, 1920, 1080, 900, 500, 128, 48)
    for name, rays in (("primary", prim), ("bounce", bnc)):
        evs = events(s, rays)
        idl, n, t, i = ideal(evs)
        print(f"{kind} {name}: rays {len(rays)} N/ray {n/len(rays):.1f} T/ray {t/len(rays):.1f} I/ray {i/len(rays):.2f} ideal slots/ray {idl/len(rays):.1f}")
        for pol, kw in (("whilewhile", dict(batch=True)), ("whilewhile", {}), ("majority", {}), ("threshold", dict(theta=4)), ("threshold", dict(theta=8)), ("threshold", dict(theta=12)),
                        ("threshold", dict(theta=8, theta_lo=4)), ("threshold", dict(theta=12, theta_lo=6)), ("threshold", dict(theta=16, theta_lo=8))):
            c, ln, lt = sim(evs, pol, **kw)
            print(f"   {pol:11s} {str(kw):32s} slots/ray {c/len(rays):7.1f}  efficiency {idl/c:.3f}  lanes/N-exec {ln:.1f} lanes/T-exec {lt:.1f}")


def sim_rule(evs, thT, thI, thR, merge_first_I=True, costs=(45, 36, 90, 30)):
    """Per-step rule: refill if idle >= thR; T-step if nT >= thT; I-step if nI >= thI; else N-step; if no N lane, run the fullest other phase.
    merge_first_I: a ray whose first event is 'I' (single-instance scene) does the instance entry inside the refill step."""
    kN, kT, kI, kF = costs
    nxt = 0; cost = 0; n_rays = len(evs)
    lane = [None] * 32; pos = [0] * 32
    stat = {"N": [0, 0], "T": [0, 0], "I": [0, 0], "F": [0, 0]}
    while True:
        sN = []; sT = []; sI = []; idle = []
        for l in range(32):
            if lane[l] is None: idle.append(l)
            else:
                c = lane[l][pos[l]]
                (sN if c == 78 else sT if c == 84 else sI).append(l)
        pool = nxt < n_rays
        if not pool and len(idle) == 32: break
        def run(which):
            nonlocal cost, nxt
            if which == "F":
                k = 0
                for l in idle:
                    if nxt < n_rays:
                        e = evs[nxt]; nxt += 1; k += 1
                        if merge_first_I and e[:1] == b"I": e = e[1:]
                        lane[l] = e if len(e) else None; pos[l] = 0
                cost += kF + (kI if merge_first_I else 0); stat["F"][0] += k; stat["F"][1] += 1
                return
            ls, k = {"N": (sN, kN), "T": (sT, kT), "I": (sI, kI)}[which]
            for l in ls:
                pos[l] += 1
                if pos[l] >= len(lane[l]): lane[l] = None
            cost += k; stat[which][0] += len(ls); stat[which][1] += 1
        if pool and len(idle) >= thR: run("F")
        elif len(sT) >= thT: run("T")
        elif len(sI) >= thI: run("I")
        elif sN: run("N")
        else:
            cand = [(len(sT), "T"), (len(sI), "I"), (len(idle) if pool else 0, "F")]
            cand.sort(reverse=True)
            run(cand[0][1])
    return cost, {k: (v[0] / max(1, v[1])) for k, v in stat.items()}


def study(kind):
    s = Scene(kind, 1920, 1080)
    prim, bnc = bounce_rays(s, 1920, 1080, 900, 500, 128, 48)
    for name, rays, anyhit in (("primary", prim, False), ("bounce", bnc, False), ("bounce-anyhit", bnc, True)):
        evs = events(s, rays, anyhit)
        idl, n, t, i = ideal(evs)
        idl += len(rays) * 30 / 32
        print(f"{kind} {name}: rays {len(rays)} N/ray {n/len(rays):.1f} T/ray {t/len(rays):.1f} I/ray {i/len(rays):.2f} ideal slots/ray {idl/len(rays):.1f}")
        best = []
        for thT in (4, 6, 8, 12, 16):
            for thI in (4, 8, 16):
                for thR in (4, 8, 12, 16):
                    c, st = sim_rule(evs, thT, thI, thR)
                    best.append((c, thT, thI, thR, st))
        best.sort(key=lambda x: x[0])
        for c, thT, thI, thR, st in best[:4] + best[-1:]:
            print(f"   thT {thT:2d} thI {thI:2d} thR {thR:2d} slots/ray {c/len(rays):7.1f} eff {idl/c:.3f} lanes/exec " + " ".join(f"{k}:{v:.1f}" for k, v in st.items()))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "rule":
    study(sys.argv[1])
