"""Peephole folding, scheduler, slot allocator, encoder and big-integer emulator for the Fq2-granular VM.

Operation (after folding):  dst = [xi]^pxi (A*B | A^2 | 0) + sum(addends),  A, B = signed sums of <= 3 slot
values, every term may carry xi (multiply by 1+u) and conj flags.  Steps are warp-synchronous and UNIFORM
(same kind and same operand/addend counts on every lane; short lanes are padded with the zero slot).
Kinds: MUL (888 wide MACs per lane), SQR (600), LIN (0), IO.
"""
import heapq
import struct

from .trace import Q

R = 1 << 384
RINV = pow(R, -1, Q)
NOP = 0xFFFF
ZERO_SLOT = 0
MAX_ADD = 8


# ---- folding ---------------------------------------------------------------------------------------------
def _reads(op):
    if op["kind"] == "FMA":
        return [t[0] for t in op["a"]] + [t[0] for t in op["b"]] + [t[0] for t in op.get("g1", [])] + [t[0] for t in op["add"]]
    return [op["src"]] if op["op"] == "store" else []


def fold(ops):
    """a LIN whose terms contain a single-use product (coefficient +/-1, not conjugated) becomes the addend
    list of that product; the product moves to the LIN's position (its other terms are defined by then)"""
    ops = [dict(op) for op in ops]
    for op in ops:
        if op["kind"] == "FMA":
            op.setdefault("pxi", 0)
    uses = {}
    for op in ops:
        for v in _reads(op):
            uses[v] = uses.get(v, 0) + 1
    producer = {op["dst"]: i for i, op in enumerate(ops) if op["kind"] == "FMA"}
    dead = set()
    for xi_, x in enumerate(ops):
        if x["kind"] != "FMA" or x["mode"] != "lin":
            continue
        counts = {}
        for t in x["add"]:
            counts[t[0]] = counts.get(t[0], 0) + 1
        best = None
        for t in x["add"]:
            v, sg, fxi, fcj = t
            pi = producer.get(v)
            if pi is None or pi in dead or counts[v] not in (1, 3) or uses.get(v, 0) != counts[v] or fcj:
                continue
            if counts[v] == 3 and fxi:
                continue
            m = ops[pi]
            if m["mode"] == "lin" or m["add"] or m["pxi"] or m.get("m3"):
                continue
            if sg < 0 and m["mode"] == "sqr":
                continue                                # -(A^2) has no operand to negate
            if best is None or pi > best[0]:
                best = (pi, t)
        if best is None:
            continue
        pi, t = best
        m = ops[pi]
        b = [(w, -s, fx, fc) for w, s, fx, fc in m["b"]] if t[1] < 0 else list(m["b"])
        others = {}
        for u in x["add"]:
            if u[0] != t[0]:
                others[u] = others.get(u, 0) + 1
        g1, rest, m3 = [], [], 0
        if counts[t[0]] == 3:                           # dst = 3 * (P + g1) + rest
            m3 = 1
            for u, c in others.items():
                g1 += [u] * (c // 3)
                rest += [u] * (c % 3)
        else:
            for u, c in others.items():
                rest += [u] * c
        if len(g1) > 4 or len(g1) + len(rest) > MAX_ADD:
            continue
        ops[xi_] = {"kind": "FMA", "mode": m["mode"], "dst": x["dst"], "a": list(m["a"]), "b": b, "add": rest, "pxi": t[2],
                    "g1": g1, "m3": m3}
        dead.add(pi)
    ops = [op for i, op in enumerate(ops) if i not in dead]
    for op in ops:
        if op["kind"] == "FMA":
            _group(op)
    return ops


def _group(op):
    """dst = m * (P + g1) + add with m in {1, 3}: terms whose multiplicity is a multiple of three move into the
    tripled group once per three copies (the 3t +/- 2z shape of the cyclotomic squaring loads 4 slots instead of 8)"""
    if "g1" in op:
        return            # grouped by the fold
    op["g1"], op["m3"] = [], 0
    if op["mode"] != "lin":
        return            # a product in the group would be tripled too: only pure sums are regrouped here
    counts = {}
    for t in op["add"]:
        counts[t] = counts.get(t, 0) + 1
    trip = {t: c // 3 for t, c in counts.items() if c >= 3}
    if sum(trip.values()) < 2 or sum(trip.values()) > 4:
        return
    g1, rest = [], []
    for t, c in counts.items():
        k = trip.get(t, 0)
        g1 += [t] * k
        rest += [t] * (c - 3 * k)
    op["g1"], op["add"], op["m3"] = g1, rest, 1


# ---- scheduling ------------------------------------------------------------------------------------------
def _deps(op):
    return _reads(op)


def _kind(op):
    return {"mul": "MUL", "sqr": "SQR", "lin": "LIN"}[op["mode"]] if op["kind"] == "FMA" else "IO"


MACS = {"MUL": 888, "SQR": 600, "LIN": 0, "IO": 0}


def schedule(ops, L, window=0):
    n = len(ops)
    producer, store_of = {}, {}
    for i, op in enumerate(ops):
        if "dst" in op:
            producer[op["dst"]] = i
        if op["kind"] == "IO" and op["op"] == "store":
            store_of[op["dst_handle"]] = i
    preds = []
    for op in ops:
        ps = {producer[v] for v in _deps(op)}
        if op["kind"] == "IO" and op["op"] == "load" and op["src_handle"] in store_of:
            ps.add(store_of[op["src_handle"]])
        preds.append(sorted(ps))
    succs = [[] for _ in range(n)]
    for i, ps in enumerate(preds):
        for p in ps:
            assert p < i, "ops must be in topological order"
            succs[p].append(i)
    cost = [300 + 4 * MACS[_kind(op)] + 40 * len(_deps(op)) for op in ops]
    prio = [0] * n
    for i in range(n - 1, -1, -1):
        prio[i] = cost[i] + max((prio[s] for s in succs[i]), default=0)
    indeg = [len(ps) for ps in preds]
    ready = {"MUL": [], "SQR": [], "LIN": [], "IO": []}
    parked, scheduled, low = [], [False] * n, 0

    def push(i):
        if window and i >= low + window:
            heapq.heappush(parked, i)
        else:
            heapq.heappush(ready[_kind(ops[i])], (-prio[i], i))
    for i in range(n):
        if indeg[i] == 0:
            push(i)
    steps, done = [], 0
    while done < n:
        while low < n and scheduled[low]:
            low += 1
        while parked and parked[0] < low + window:
            i = heapq.heappop(parked)
            heapq.heappush(ready[_kind(ops[i])], (-prio[i], i))
        if len(ready["MUL"]) >= L:
            kind = "MUL"
        elif len(ready["SQR"]) >= L:
            kind = "SQR"
        elif ready["LIN"]:
            kind = "LIN"
        elif ready["IO"]:
            kind = "IO"
        else:      # nothing fills a step: take the partial step that wastes the least multiplier time
            kind = "MUL" if len(ready["MUL"]) * 888 >= len(ready["SQR"]) * 600 else "SQR"
            if not ready[kind]:
                kind = "SQR" if kind == "MUL" else "MUL"
        if kind == "IO":
            chosen = [heapq.heappop(ready[kind])[1] for _ in range(min(L, len(ready[kind])))]
        else:      # lanes of a step share one shape: max tripled-group size + max addend count <= 8 fields
            chosen, deferred, m1, m2 = [], [], 0, 0
            while ready[kind] and len(chosen) < L:
                item = heapq.heappop(ready[kind])
                op = ops[item[1]]
                a1, a2 = len(op.get("g1", [])), len(op["add"])
                if max(m1, a1) + max(m2, a2) <= 8:
                    chosen.append(item[1]); m1, m2 = max(m1, a1), max(m2, a2)
                else:
                    deferred.append(item)
            for item in deferred:
                heapq.heappush(ready[kind], item)
        assert chosen, "deadlock"
        steps.append((kind, chosen))
        done += len(chosen)
        for i in chosen:
            scheduled[i] = True
            for s in succs[i]:
                indeg[s] -= 1
                if indeg[s] == 0:
                    push(s)
    return steps


def allocate(ops, steps):
    last_use = {}
    for t, (_, chosen) in enumerate(steps):
        for i in chosen:
            for v in _deps(ops[i]):
                last_use[v] = t
    slot, free, nslots, dying = {}, [], 1, {}
    for v, t in last_use.items():
        dying.setdefault(t, []).append(v)
    for t, (_, chosen) in enumerate(steps):
        for v in dying.get(t, []):
            if v in slot:
                heapq.heappush(free, slot[v])
        for i in chosen:
            if "dst" not in ops[i]:
                continue
            d = ops[i]["dst"]
            if d not in last_use:
                dying.setdefault(t + 1, []).append(d)
            if free:
                slot[d] = heapq.heappop(free)
            else:
                slot[d] = nslots; nslots += 1
    return slot, nslots


# lane instruction, 64 bytes: dst, a[3], b[3], add[8] (u16, 6 * slot), flags a[3] b[3] add[8] (u8: 1 neg, 2 xi, 4 conj),
# lane bytes (pxi | io op, io seg, io width, pad), header[16]
_FMT = "<H3H3H8H14B4B16B"
KIND_CODE = {"LIN": 0, "MUL": 1, "SQR": 2, "IO": 3}


def _flag(t):
    return (1 if t[1] < 0 else 0) | (2 if t[2] else 0) | (4 if t[3] else 0)


def encode(ops, steps, slot, L):
    assert struct.calcsize(_FMT) == 64
    out = bytearray()
    o6 = lambda v: 6 * slot[v]
    Z = 6 * ZERO_SLOT
    for kind, chosen in steps:
        lanes = [ops[i] for i in chosen] + [None] * (L - len(chosen))
        live = [op for op in lanes if op]
        if kind == "IO":
            hdr = [3] + [0] * 15
            for op in lanes:
                if op is None:
                    out += struct.pack(_FMT, NOP, *([Z] * 14), *([0] * 14), 0, 0, 0, 0, *hdr)
                    continue
                sl = o6(op["dst"]) if op["op"] == "load" else o6(op["src"])
                out += struct.pack(_FMT, sl, op["idx"], *([Z] * 13), *([0] * 14), 1 if op["op"] == "load" else 2, op["seg"], op["width"], 0, *hdr)
            continue
        na = max(len(op["a"]) for op in live)
        nb = max(len(op["b"]) for op in live)
        n1 = max(len(op["g1"]) for op in live)
        nadd = max(len(op["add"]) for op in live)
        assert na <= 3 and nb <= 3 and n1 <= 4 and n1 + nadd <= 8
        # union of the flags used at each position: lets the kernel skip mask work on plain steps
        fa = fb = fadd = 0
        for op in live:
            for t in op["a"]:
                fa |= _flag(t)
            for t in op["b"]:
                fb |= _flag(t)
            for t in op["add"] + op["g1"]:
                fadd |= _flag(t)
        anypxi = (1 if any(op["pxi"] for op in live) else 0) | (2 if any(op["m3"] for op in live) else 0)
        hdr = [KIND_CODE[kind], na, nb, nadd, fa, fb, fadd, anypxi, n1] + [0] * 7
        for op in lanes:
            if op is None:
                out += struct.pack(_FMT, NOP, *([Z] * 14), *([0] * 14), 0, 0, 0, 0, *hdr)
                continue
            o = [Z] * 14
            fl = [0] * 14
            cls = lambda t: (t[3], t[2], t[1] < 0)
            for base, terms in ((0, sorted(op["a"], key=cls)), (3, sorted(op["b"], key=cls)), (6, sorted(op["g1"], key=cls)),
                                (6 + n1, sorted(op["add"], key=cls))):
                for j, t in enumerate(terms):
                    o[base + j] = o6(t[0]); fl[base + j] = _flag(t)
            out += struct.pack(_FMT, o6(op["dst"]), *o, *fl, (1 if op["pxi"] else 0) | (2 if op["m3"] else 0), 0, 0, 0, *hdr)
    return bytes(out)


# ---- emulator --------------------------------------------------------------------------------------------
def _term(v, fl):
    t0, t1 = v
    if fl & 4:
        t1 = -t1
    if fl & 2:
        t0, t1 = t0 - t1, t0 + t1
    if fl & 1:
        t0, t1 = -t0, -t1
    return t0, t1


class Emulator:
    def __init__(self, code, L, nslots, consts):
        self.code, self.L, self.nslots = code, L, nslots
        self.consts = []
        for c0, c1 in consts:
            self.consts += [c0 * R % Q, c1 * R % Q]

    def run(self, segs):
        slots = [None] * self.nslots
        slots[ZERO_SLOT] = (0, 0)
        segs = dict(segs)
        segs[4] = self.consts
        L = self.L
        nsteps = len(self.code) // (64 * L)

        def rd(o):
            assert o % 6 == 0
            v = slots[o // 6]
            assert v is not None, "read of an unwritten slot %d" % (o // 6)
            return v

        for s in range(nsteps):
            pending = []
            for l in range(L):
                ins = struct.unpack_from(_FMT, self.code, (s * L + l) * 64)
                dst, o, fl, lane, hdr = ins[0], ins[1:15], ins[15:29], ins[29:33], ins[33:49]
                kind = hdr[0]
                if kind == 3:
                    if dst == NOP:
                        continue
                    op_, seg, width = lane[0], lane[1], lane[2]
                    if op_ == 1:
                        g = segs[seg]
                        pending.append((dst, (g[o[0]], g[o[0] + 1] if width == 2 else 0)))
                    elif op_ == 2:
                        g = segs[seg]
                        while len(g) <= o[0] + width - 1:
                            g.append(None)
                        v = rd(dst)
                        g[o[0]] = v[0]
                        if width == 2:
                            g[o[0] + 1] = v[1]
                    continue
                na, nb, nadd = hdr[1], hdr[2], hdr[3]

                def gather(base, n, U):
                    a0 = a1 = 0
                    for j in range(n):
                        f = fl[base + j]
                        t0, t1 = rd(o[base + j])
                        if U & 6:
                            t0, t1 = _term((t0, t1), f)
                        elif f & 1:                      # device: Q - t, so the unreduced sum stays in [0, n Q]
                            t0, t1 = Q - t0, Q - t1
                        a0 += t0; a1 += t1
                    if U & 6:
                        return a0 % Q, a1 % Q
                    assert 0 <= a0 <= n * Q and 0 <= a1 <= n * Q
                    return a0, a1
                p0 = p1 = 0
                if kind == 1:
                    a0, a1 = gather(0, na, hdr[4]); b0, b1 = gather(3, nb, hdr[5])
                    nb1 = nb * Q - b1
                    assert 0 <= nb1 and a0 * b0 + a1 * nb1 < Q * R and a0 * b1 + a1 * b0 < Q * R
                    p0 = (a0 * b0 + a1 * nb1) * RINV % Q
                    p1 = (a0 * b1 + a1 * b0) * RINV % Q
                elif kind == 2:
                    a0, a1 = gather(0, na, hdr[4])
                    assert a0 < Q and a1 < Q
                    p0 = (a0 + a1) * (a0 + Q - a1) * RINV % Q
                    p1 = a0 * (2 * a1) * RINV % Q
                n1 = hdr[8]
                if lane[0] & 1:
                    p0, p1 = p0 - p1, p0 + p1
                for j in range(n1):
                    t0, t1 = _term(rd(o[6 + j]), fl[6 + j])
                    p0 += t0; p1 += t1
                if lane[0] & 2:
                    p0, p1 = 3 * p0, 3 * p1
                for j in range(nadd):
                    t0, t1 = _term(rd(o[6 + n1 + j]), fl[6 + n1 + j])
                    p0 += t0; p1 += t1
                if dst != NOP:
                    pending.append((dst, (p0 % Q, p1 % Q)))
            for dst, v in pending:
                slots[dst // 6] = v
        return segs


def stats(ops, steps, L):
    cnt = {k: sum(1 for op in ops if _kind(op) == k) for k in ("MUL", "SQR", "LIN", "IO")}
    st = {k: sum(1 for kd, _ in steps if kd == k) for k in ("MUL", "SQR", "LIN", "IO")}
    macs = 888 * cnt["MUL"] + 600 * cnt["SQR"]
    pipe = (888 * st["MUL"] + 600 * st["SQR"]) * L
    terms = sum(len(_deps(op)) for op in ops if op["kind"] == "FMA")
    return {"ops": cnt, "steps": st, "wide_macs": macs, "mac_fill": macs / pipe if pipe else 0.0, "terms": terms,
            "accesses_per_kmac": 1000.0 * (terms + cnt["MUL"] + cnt["SQR"] + cnt["LIN"]) / max(macs, 1)}
