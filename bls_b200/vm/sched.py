"""Scheduler, slot allocator, encoder and big-integer emulator for the Fq virtual machine.

A program is a list of warp-synchronous steps; in every step each of the L lanes of a lane group
executes at most one operation and all operations of a step have the same kind:
  MUL  dst = (a0 [+/- a1]) * (b0 [+/- b1])      one 384-bit Montgomery multiplication per lane
  LIN  dst = sum of positive terms - sum of negative terms, reduced to the canonical residue
  IO   slot <- global element (inputs, constants, spilled values) or global element <- slot
A MUL step costs the IMAD.WIDE pipe 300 wide MACs per lane whether the lane is busy or not, so the
scheduler's first objective is the number of MUL steps (lower bound ceil(#MUL / L)).  Steps are
UNIFORM: the shape of a step (operand pattern of a MUL step; positive / negative term counts and
reduction depth of a LIN step) is the same for every lane -- short lanes are padded with the zero
slot -- so the kernel executes them without divergence.
"""
import heapq
import struct

from .trace import Q

R = 1 << 384
RINV = pow(R, -1, Q)
NOP = 0xFFFF
ZERO_SLOT = 0            # slot 0 of every lane group holds 0 (written once by the kernel)


def _deps(op):
    if op["kind"] == "MUL":
        return [v for o in op["operands"] for v, _ in o]
    if op["kind"] == "LIN":
        return [v for v, _ in op["terms"]]
    return [op["src"]] if op["op"] == "store" else []


def schedule(prog, L, window=0, shape_slack=9):
    ops = prog.ops
    n = len(ops)
    producer = {}
    store_of = {}
    for i, op in enumerate(ops):
        if "dst" in op:
            producer[op["dst"]] = i
        if op["kind"] == "IO" and op["op"] == "store":
            store_of[op["dst_handle"]] = i
    preds = []
    for op in ops:
        ps = {producer[v] for v in _deps(op)}
        if op["kind"] == "IO" and op["op"] == "load" and op["src_handle"] in store_of:
            ps.add(store_of[op["src_handle"]])          # a spilled value is re-read after its store
        preds.append(sorted(ps))
    succs = [[] for _ in range(n)]
    for i, ps in enumerate(preds):
        for p in ps:
            succs[p].append(i)
    def kind_of(op):
        return ("MUL%d" % op["k"]) if op["kind"] == "MUL" else op["kind"]
    cost = [650 * (op["k"] + 1) if op["kind"] == "MUL" else (60 + 16 * len(op["terms"]) if op["kind"] == "LIN" else 200) for op in ops]
    prio = [0] * n
    for i in range(n - 1, -1, -1):
        prio[i] = cost[i] + max((prio[s] for s in succs[i]), default=0)
    indeg = [len(ps) for ps in preds]
    ready = {"MUL1": [], "MUL2": [], "LIN": [], "IO": []}
    parked = []                  # ready but beyond the look-ahead window (bounds live ranges / slots)
    scheduled = [False] * n
    low = 0

    def push(i):
        if window and i >= low + window:
            heapq.heappush(parked, i)
        else:
            heapq.heappush(ready[kind_of(ops[i])], (-prio[i], i))
    for i in range(n):
        if indeg[i] == 0:
            push(i)
    steps = []
    done = 0
    while done < n:
        while low < n and scheduled[low]:
            low += 1
        while parked and parked[0] < low + window:
            i = heapq.heappop(parked)
            heapq.heappush(ready[kind_of(ops[i])], (-prio[i], i))
        if len(ready["MUL2"]) >= L:
            kind = "MUL2"
        elif len(ready["MUL1"]) >= L:
            kind = "MUL1"
        elif ready["LIN"]:
            kind = "LIN"
        elif ready["IO"]:
            kind = "IO"
        else:
            kind = "MUL2" if len(ready["MUL2"]) >= len(ready["MUL1"]) else "MUL1"
        if kind == "LIN":
            # lanes of a LIN step share one shape (max positive + max negative term count <= 9 operand fields)
            chosen, deferred, mp, mn = [], [], 0, 0
            while ready["LIN"] and len(chosen) < L:
                item = heapq.heappop(ready["LIN"])
                terms = ops[item[1]]["terms"]
                p_ = sum(1 for _, sg in terms if sg > 0); n_ = len(terms) - p_
                if max(mp, p_) + max(mn, n_) <= 9 and (not chosen or max(mp, p_) + max(mn, n_) <= shape_slack + mp + mn):
                    chosen.append(item[1]); mp, mn = max(mp, p_), max(mn, n_)
                else:
                    deferred.append(item)
            for item in deferred:
                heapq.heappush(ready["LIN"], item)
        else:
            chosen = [heapq.heappop(ready[kind])[1] for _ in range(min(L, len(ready[kind])))]
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


def allocate(prog, steps):
    """slot for every value; a value dies after the step of its last use; slot 0 is the zero slot"""
    ops = prog.ops
    last_use = {}
    for t, (_, chosen) in enumerate(steps):
        for i in chosen:
            for v in _deps(ops[i]):
                last_use[v] = t
    slot = {}
    free, nslots = [], 1
    dying = {}
    for v, t in last_use.items():
        dying.setdefault(t, []).append(v)
    for t, (_, chosen) in enumerate(steps):
        # operands read in this step are released before its results are placed: the kernel
        # synchronises the warp between the loads and the stores of a step
        for v in dying.get(t, []):
            if v in slot:
                heapq.heappush(free, slot[v])
        for i in chosen:
            if "dst" not in ops[i]:
                continue
            d = ops[i]["dst"]
            if d not in last_use:             # dead result: still needs a landing slot for this step
                dying.setdefault(t + 1, []).append(d)
            if free:
                slot[d] = heapq.heappop(free)
            else:
                slot[d] = nslots; nslots += 1
    return slot, nslots


_FMT = "<H9HBBH8B"     # dst, o[9], lane byte, lane byte 2, pad, header[8]  = 32 bytes


def encode(prog, steps, slot, L):
    """-> bytes: nsteps * L lane instructions of 32 bytes (layout in vm.cuh).  Slot operands are
    stored as 3 * slot (16-byte units)."""
    assert struct.calcsize(_FMT) == 32
    out = bytearray()
    padded_terms = 0
    o3 = lambda v: 3 * slot[v]
    Z = 3 * ZERO_SLOT
    for kind, chosen in steps:
        lanes = [prog.ops[i] for i in chosen] + [None] * (L - len(chosen))
        if kind in ("MUL1", "MUL2"):
            live = [op for op in lanes if op]
            present = anyneg = 0
            for op in live:
                for oi, o in enumerate(op["operands"]):
                    for ti, (v, sg) in enumerate(o):
                        present |= 1 << (2 * oi + ti)
                        if sg < 0:
                            anyneg |= 1 << (2 * oi + ti)
            hdr = [1 if kind == "MUL1" else 4, present, anyneg, 0, 0, 0, 0, 0]
            for op in lanes:
                if op is None:
                    out += struct.pack(_FMT, NOP, *([Z] * 9), 0, 0, 0, *hdr)
                    continue
                o = [Z] * 9
                signs = 0
                for oi, opd in enumerate(op["operands"]):
                    assert 1 <= len(opd) <= 2
                    for ti, (v, sg) in enumerate(opd):
                        o[2 * oi + ti] = o3(v)
                        if sg < 0:
                            signs |= 1 << (2 * oi + ti)
                out += struct.pack(_FMT, o3(op["dst"]), *o, signs, 0, 0, *hdr)
        elif kind == "LIN":
            live = [op for op in lanes if op]
            npos = max(sum(1 for _, sg in op["terms"] if sg > 0) for op in live)
            nneg = max(sum(1 for _, sg in op["terms"] if sg < 0) for op in live)
            weight = max(len(op["terms"]) for op in live)
            assert npos + nneg <= 9 and weight <= 8
            mode = 0 if weight <= 1 and nneg == 0 else (1 if weight <= 2 else 2)
            hdr = [2, npos, nneg, mode, 0, 0, 0, 0]
            padded_terms += (npos + nneg) * L
            for op in lanes:
                if op is None:
                    out += struct.pack(_FMT, NOP, *([Z] * 9), 0, 0, 0, *hdr)
                    continue
                pos = [o3(v) for v, sg in op["terms"] if sg > 0]
                neg = [o3(v) for v, sg in op["terms"] if sg < 0]
                o = pos + [Z] * (9 - len(pos) - len(neg)) + neg[::-1]      # negative terms occupy o[8], o[7], ...
                out += struct.pack(_FMT, o3(op["dst"]), *o, len(neg), 0, 0, *hdr)
        else:
            hdr = [3, 0, 0, 0, 0, 0, 0, 0]
            for op in lanes:
                if op is None:
                    out += struct.pack(_FMT, NOP, *([Z] * 9), 0, 0, 0, *hdr)
                    continue
                assert 0 <= op["seg"] < 8 and 0 <= op["idx"] < 65536
                sl = o3(op["dst"]) if op["op"] == "load" else o3(op["src"])
                out += struct.pack(_FMT, sl, op["idx"], *([Z] * 8), 1 if op["op"] == "load" else 2, op["seg"], 0, *hdr)
    return bytes(out), padded_terms


class Emulator:
    """executes an encoded program on Python integers (Montgomery domain), one lane group"""

    def __init__(self, code, L, nslots, consts):
        self.code, self.L, self.nslots = code, L, nslots
        self.consts = [c * R % Q for c in consts]

    def run(self, segs):
        """segs: dict seg -> list of Montgomery-form integers (mutated for outputs)"""
        slots = [None] * self.nslots
        slots[ZERO_SLOT] = 0
        segs = dict(segs)
        segs[4] = self.consts
        L = self.L
        nsteps = len(self.code) // (32 * L)

        def rd(o):
            assert o % 3 == 0
            v = slots[o // 3]
            assert v is not None, "read of an unwritten slot %d" % (o // 3)
            return v

        for s in range(nsteps):
            pending = []
            for l in range(L):
                ins = struct.unpack_from(_FMT, self.code, (s * L + l) * 32)
                dst, o, lb, lb2, hdr = ins[0], ins[1:10], ins[10], ins[11], ins[13:21]
                kind = hdr[0]
                if kind in (1, 4):
                    vals = []
                    for oi in range(2 if kind == 1 else 4):
                        acc = 0
                        for ti in range(2):
                            bit = 2 * oi + ti
                            if hdr[1] >> bit & 1:
                                t = rd(o[bit])
                                acc += (Q - t) if lb >> bit & 1 else t
                            else:
                                assert o[bit] == 3 * ZERO_SLOT and not (lb >> bit & 1)
                        assert acc <= 2 * Q
                        vals.append(acc)
                    t = vals[0] * vals[1] + (vals[2] * vals[3] if kind == 4 else 0)
                    assert t < Q * R
                    if dst != NOP:
                        pending.append((dst, t * RINV % Q))
                elif kind == 2:
                    npos, nneg = hdr[1], hdr[2]
                    acc = lb * Q
                    for j in range(npos):
                        acc += rd(o[j])
                    for j in range(nneg):
                        acc -= rd(o[8 - j])
                    assert 0 <= acc < 8 * Q
                    if hdr[3] == 0:
                        assert acc < Q
                    elif hdr[3] == 1:
                        assert acc < 2 * Q
                    if dst != NOP:
                        pending.append((dst, acc % Q))
                elif kind == 3 and dst != NOP:
                    if lb == 1:
                        pending.append((dst, segs[lb2][o[0]]))
                    elif lb == 2:
                        seg = segs[lb2]
                        while len(seg) <= o[0]:
                            seg.append(None)
                        seg[o[0]] = rd(dst)
            for dst, v in pending:        # all loads of a step happen before its stores
                assert 0 <= v < Q
                slots[dst // 3] = v
        return segs


def stats(prog, steps, L, padded_terms=0):
    nm = sum(1 for op in prog.ops if op["kind"] == "MUL")
    nm1 = sum(1 for op in prog.ops if op["kind"] == "MUL" and op["k"] == 1)
    nm2 = nm - nm1
    s1 = sum(1 for k, _ in steps if k == "MUL1"); s2 = sum(1 for k, _ in steps if k == "MUL2")
    nl = sum(1 for op in prog.ops if op["kind"] == "LIN")
    ms = s1 + s2
    ls = sum(1 for k, _ in steps if k == "LIN")
    ios = sum(1 for k, _ in steps if k == "IO")
    terms = sum(len(op["terms"]) for op in prog.ops if op["kind"] == "LIN")
    macs = 300 * nm1 + 444 * nm2
    return {"mul1_ops": nm1, "mul2_ops": nm2, "mul1_steps": s1, "mul2_steps": s2, "wide_macs": macs,
            "mac_fill": macs / ((300 * s1 + 444 * s2) * L) if ms else 0.0,
            "mul_ops": nm, "lin_ops": nl, "lin_terms": terms, "mul_steps": ms, "lin_steps": ls, "io_steps": ios,
            "mul_fill": nm / (ms * L) if ms else 0.0, "lin_fill": nl / (ls * L) if ls else 0.0,
            "lin_padded_terms_per_step": padded_terms / (ls * L) if ls else 0.0}
