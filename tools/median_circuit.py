#!/usr/bin/env python
"""Designs, verifies and emits the min/max circuits of the shared-sort 5x5 and 3x3 medians (dmc_front8u.cu).

A thread of the median kernel produces R vertically adjacent outputs for one pixel pair (two fp16 lanes).  Rather than
running a 25-input selection network per output, the work common to neighbouring outputs is done once:

  S[i]   = the 5 horizontal neighbours of input row i, sorted                        (each row serves 5 outputs)
  P[i]   = merge(S[i], S[i+1])                        sorted 10                       (each pair serves 2 output pairs)
  M      = ranks 7..12 of merge(P[o+1], P[o+3])       the only ranks of the 4 shared rows that can be the median
  out(o)   = rank 5 of S[o]   u M      (11 values)
  out(o+1) = rank 5 of S[o+5] u M

Everything is a min/max circuit, so the zero-one principle applies: the circuit computes the median of every input iff
it does so for the 2^25 binary inputs, which verify() enumerates bit-parallel.

    python tools/median_circuit.py            # verify + op counts
    python tools/median_circuit.py --emit     # also rewrite depthmapcompression_b200/csrc/dmc_median_gen.inc
"""
import argparse
import os
import sys

import numpy as np

INF, NINF = "+inf", "-inf"


class Circuit:
    """Min/max DAG with constant folding and common-subexpression elimination."""

    def __init__(self):
        self.ops = []          # (kind, a, b) with kind in {"in", "min", "max"}
        self.memo = {}

    def inp(self, name):
        self.ops.append(("in", name, None)); return len(self.ops) - 1

    def _op(self, kind, a, b):
        if a == b: return a
        if kind == "min":
            if a == INF: return b
            if b == INF: return a
            if a == NINF or b == NINF: return NINF
        else:
            if a == NINF: return b
            if b == NINF: return a
            if a == INF or b == INF: return INF
        key = (kind, min(a, b), max(a, b))
        if key not in self.memo:
            self.ops.append((kind, key[1], key[2])); self.memo[key] = len(self.ops) - 1
        return self.memo[key]

    def mn(self, a, b): return self._op("min", a, b)
    def mx(self, a, b): return self._op("max", a, b)
    def xchg(self, a, b): return self.mn(a, b), self.mx(a, b)

    def live(self, outs):
        need = set(); stack = [o for o in outs]
        while stack:
            n = stack.pop()
            if n in need or isinstance(n, str): continue
            need.add(n); k, a, b = self.ops[n]
            if k != "in": stack += [a, b]
        return need

    def count(self, outs):
        return sum(1 for n in self.live(outs) if self.ops[n][0] != "in")


SORT5 = [(0, 1), (3, 4), (2, 4), (2, 3), (0, 3), (0, 2), (1, 4), (1, 3), (1, 2)]      # 9 exchanges
SORT3 = [(0, 1), (1, 2), (0, 1)]


def sort_net(c, v, net):
    v = list(v)
    for i, j in net: v[i], v[j] = c.xchg(v[i], v[j])
    return v


def oe_merge(c, a, b):
    """Batcher's odd-even merge of two sorted lists of arbitrary lengths (Knuth 5.3.4)."""
    if not a: return list(b)
    if not b: return list(a)
    if len(a) == 1 and len(b) == 1: return list(c.xchg(a[0], b[0]))
    v = oe_merge(c, a[0::2], b[0::2]); w = oe_merge(c, a[1::2], b[1::2])
    out = [v[0]]; i = 0
    while i < len(w) or i + 1 < len(v):
        wi = w[i] if i < len(w) else None; vi = v[i + 1] if i + 1 < len(v) else None
        if wi is not None and vi is not None: out += list(c.xchg(wi, vi))
        elif wi is not None: out.append(wi)
        else: out.append(vi)
        i += 1
    return out


def rank_formula(c, x, y, k):
    """k-th smallest (1-based) of the union of sorted x and y:  min over i+j=k of max(x_i, y_j), x_0 = y_0 = -inf."""
    best = INF
    for i in range(0, k + 1):
        j = k - i
        if i > len(x) or j > len(y): continue
        xi = x[i - 1] if i else NINF; yj = y[j - 1] if j else NINF
        best = c.mn(best, c.mx(xi, yj))
    return best


def mid_of_merge(c, p, q, lo, hi, how):
    if how == "oe": return oe_merge(c, p, q)[lo:hi + 1]
    return [rank_formula(c, p, q, k + 1) for k in range(lo, hi + 1)]


def build25(how_mid="oe", how_sel="formula"):
    """Two vertically adjacent outputs from six rows of five (rows 0..4 -> out0, rows 1..5 -> out1)."""
    c = Circuit()
    rows = [[c.inp("r%dc%d" % (r, k)) for k in range(5)] for r in range(6)]
    S = [sort_net(c, r, SORT5) for r in rows]
    P1 = oe_merge(c, S[1], S[2]); P3 = oe_merge(c, S[3], S[4])
    M = mid_of_merge(c, P1, P3, 7, 12, how_mid)
    def sel(s):
        if how_sel == "formula": return rank_formula(c, s, M, 6)
        return oe_merge(c, s, M)[5]
    return c, rows, S, P1, P3, M, sel(S[0]), sel(S[5])


def eval01(c, outs, inputs):
    """Evaluates the circuit on every binary assignment of `inputs` (<= 25 of them), bit-parallel."""
    n = len(inputs); words = max(1, (1 << n) // 64)
    val = {}
    for k, node in enumerate(inputs):
        if k < 6:
            pat = 0
            for b in range(64):
                if (b >> k) & 1: pat |= 1 << b
            val[node] = np.full(words, pat, dtype=np.uint64)
        else:
            idx = np.arange(words, dtype=np.uint64)
            val[node] = np.where((idx >> np.uint64(k - 6)) & np.uint64(1), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0))
    need = c.live(outs)
    for nidx in sorted(need):
        k, a, b = c.ops[nidx]
        if k == "in":
            if nidx not in val: raise ValueError("circuit depends on an input outside the window")
            continue
        val[nidx] = (val[a] & val[b]) if k == "min" else (val[a] | val[b])
    return [val[o] for o in outs], val


def popcount_ge(inputs_vals, thresh):
    """Bit-parallel 'at least thresh of the inputs are one' via a ripple counter."""
    nbits = 5
    cnt = [np.zeros_like(inputs_vals[0]) for _ in range(nbits)]
    for v in inputs_vals:
        carry = v
        for b in range(nbits):
            cnt[b], carry = cnt[b] ^ carry, cnt[b] & carry
    ones = np.uint64(0xFFFFFFFFFFFFFFFF)
    # value >= thresh, compare bit-serially from the top
    gt = np.zeros_like(cnt[0]); eq = np.full_like(cnt[0], ones)
    for b in reversed(range(nbits)):
        tb = ones if (thresh >> b) & 1 else np.uint64(0)
        gt |= eq & cnt[b] & ~tb
        eq &= ~(cnt[b] ^ tb)
    return gt | eq


def verify25(how_mid, how_sel):
    c, rows, S, P1, P3, M, o0, o1 = build25(how_mid, how_sel)
    for out, rr in ((o0, rows[0:5]), (o1, rows[1:6])):
        inputs = [n for r in rr for n in r]
        (res,), val = eval01(c, [out], inputs)
        want = popcount_ge([val[n] for n in inputs], 13)      # median of 25 binary values is 1 iff >= 13 ones
        if not np.array_equal(res, want): return False
    return True


def build9():
    """One 3x3 output from three sorted rows of three."""
    c = Circuit()
    rows = [[c.inp("r%dc%d" % (r, k)) for k in range(3)] for r in range(3)]
    S = [sort_net(c, r, SORT3) for r in rows]
    lo = c.mx(c.mx(S[0][0], S[1][0]), S[2][0])
    hi = c.mn(c.mn(S[0][2], S[1][2]), S[2][2])
    a, b, d = S[0][1], S[1][1], S[2][1]
    mid = c.mx(c.mn(a, b), c.mn(c.mx(a, b), d))
    out = c.mx(c.mn(lo, mid), c.mn(c.mx(lo, mid), hi))
    return c, rows, S, out


def verify9():
    c, rows, S, out = build9()
    inputs = [n for r in rows for n in r]
    (res,), val = eval01(c, [out], inputs)
    return np.array_equal(res, popcount_ge([val[n] for n in inputs], 5))


# ---- code emission ---------------------------------------------------------------------------------------------------
def emit_function(c, name, in_lists, out_nodes, out_names, lines, fmod_note=True):
    """Straight-line code for the sub-circuit that computes out_nodes from the nodes of in_lists (name -> node list)."""
    sym = {}
    for arr, nodes in in_lists:
        for k, n in enumerate(nodes): sym[n] = "%s[%d]" % (arr, k)
    need = set(); stack = list(out_nodes)
    while stack:
        n = stack.pop()
        if n in need or n in sym: continue
        need.add(n); k, a, b = c.ops[n]
        if k == "in": raise ValueError("%s: reaches a raw input" % name)
        stack += [a, b]
    order = sorted(need)
    pair = {}
    for n in order:
        k, a, b = c.ops[n]
        other = c.memo.get(("max" if k == "min" else "min", a, b))
        if other in need: pair[n] = other
    done = set(); nx = 0; nsingle = 0
    for n in order:
        if n in done: continue
        k, a, b = c.ops[n]
        if n in pair:
            lo, hi = (n, pair[n]) if k == "min" else (pair[n], n)
            sym[lo], sym[hi] = "t%d" % lo, "t%d" % hi
            lines.append("    DMC_MED_XCHG(%d, %s, %s, t%d, t%d);" % (nx, sym[a], sym[b], lo, hi)); nx += 1
            done |= {lo, hi}
        else:
            sym[n] = "t%d" % n
            lines.append("    const DMC_MED_T t%d = %s(%s, %s);" % (n, "DMC_MED_MIN" if k == "min" else "DMC_MED_MAX", sym[a], sym[b])); nsingle += 1
            done.add(n)
    for nm, n in zip(out_names, out_nodes): lines.append("    %s = %s;" % (nm, sym[n]))
    return nx, nsingle


def emit(path):
    c, rows, S, P1, P3, M, o0, o1 = build25("oe", "formula")
    L = ["// dmc_median_gen.inc -- GENERATED by tools/median_circuit.py (do not edit): the min/max circuits of the shared-sort",
         "// medians.  The includer defines DMC_MED_T (lane type), DMC_MED_FN (function qualifiers), DMC_MED_MIN / DMC_MED_MAX and",
         "// DMC_MED_XCHG(i, a, b, lo, hi), which declares lo = min(a, b) and hi = max(a, b); i numbers the exchanges of one",
         "// function so that every n-th one can be routed to the FMA pipe.  All circuits verified on every binary input.", ""]
    stats = {}
    L.append("DMC_MED_FN void med_sort5(DMC_MED_T (&v)[5]) {")
    for k, (i, j) in enumerate(SORT5): L.append("    { DMC_MED_XCHG(%d, v[%d], v[%d], lo, hi); v[%d] = lo; v[%d] = hi; }" % (k, i, j, i, j))
    L.append("}"); L.append("")
    L.append("DMC_MED_FN void med_sort3(DMC_MED_T (&v)[3]) {")
    for k, (i, j) in enumerate(SORT3): L.append("    { DMC_MED_XCHG(%d, v[%d], v[%d], lo, hi); v[%d] = lo; v[%d] = hi; }" % (k, i, j, i, j))
    L.append("}"); L.append("")
    L.append("// sorted a[5], sorted b[5] -> sorted p[10]")
    L.append("DMC_MED_FN void med_merge55(const DMC_MED_T (&a)[5], const DMC_MED_T (&b)[5], DMC_MED_T (&p)[10]) {")
    stats["merge55"] = emit_function(c, "merge55", [("a", S[1]), ("b", S[2])], P1, ["p[%d]" % k for k in range(10)], L)
    L.append("}"); L.append("")
    L.append("// sorted p[10], sorted q[10] -> m[0..5] = ranks 7..12 of their union")
    L.append("DMC_MED_FN void med_mid6(const DMC_MED_T (&p)[10], const DMC_MED_T (&q)[10], DMC_MED_T (&m)[6]) {")
    stats["mid6"] = emit_function(c, "mid6", [("p", P1), ("q", P3)], M, ["m[%d]" % k for k in range(6)], L)
    L.append("}"); L.append("")
    L.append("// sorted s[5] and m[6] (ranks 7..12 of the other 20 window values) -> median of the 25")
    L.append("DMC_MED_FN DMC_MED_T med_select(const DMC_MED_T (&s)[5], const DMC_MED_T (&m)[6]) {")
    L.append("    DMC_MED_T r;")
    stats["select"] = emit_function(c, "select", [("s", S[0]), ("m", M)], [o0], ["r"], L)
    L.append("    return r;"); L.append("}"); L.append("")
    c9, rows9, S9, out9 = build9()
    L.append("// three sorted rows of three -> median of the 9")
    L.append("DMC_MED_FN DMC_MED_T med_select9(const DMC_MED_T (&a)[3], const DMC_MED_T (&b)[3], const DMC_MED_T (&c)[3]) {")
    L.append("    DMC_MED_T r;")
    stats["select9"] = emit_function(c9, "select9", [("a", S9[0]), ("b", S9[1]), ("c", S9[2])], [out9], ["r"], L)
    L.append("    return r;"); L.append("}")
    open(path, "w").write("\n".join(L) + "\n")
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emit", action="store_true")
    args = ap.parse_args()
    print("3x3:", "ok" if verify9() else "WRONG", "ops per output (excluding the shared row sort):", build9()[0].count([build9()[3]]) - 9 * 2)
    for how_mid in ("oe", "formula"):
        for how_sel in ("formula", "oe"):
            c, rows, S, P1, P3, M, o0, o1 = build25(how_mid, how_sel)
            tot = c.count([o0, o1])
            print("5x5 mid=%-7s sel=%-7s: %s  ops for two outputs incl. 6 row sorts = %d  (row sorts %d, P merges %d, mid6 %d, selects %d)" % (
                how_mid, how_sel, "ok" if verify25(how_mid, how_sel) else "WRONG", tot,
                c.count([n for s in S for n in s]), c.count(P1 + P3) - c.count([n for s in S[1:5] for n in s]),
                c.count(M) - c.count(P1 + P3), tot - c.count(M) - c.count(S[0] + S[5]) + 0))
    if args.emit:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "depthmapcompression_b200", "csrc", "dmc_median_gen.inc")
        stats = emit(os.path.normpath(path))
        print("emitted", os.path.normpath(path), stats)
    return 0


if __name__ == "__main__":
    sys.exit(main())
